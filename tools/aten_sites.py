"""Which Python lines issue the small ATen ops of one eager training step?  A TorchDispatchMode records, for every ATen
op that touches a CUDA tensor, the innermost stack frame inside this repo.
    python tools/aten_sites.py [workload]"""
import collections
import sys
import traceback
import torch
from torch.utils._python_dispatch import TorchDispatchMode
sys.path.insert(0, ".")
import bench
import pggan_b200 as pg

wl = sys.argv[1] if len(sys.argv) > 1 else "c4_256x256_b16"
pg.set_compute_dtype(torch.float16)
scale, batch, depths, _ = bench.WORKLOADS[wl]
tr = pg.ProgressiveGAN(pg.make_args(latent_dim=512, depths=depths), 0, variant="W")
tr.set_scale(scale, 0.5 if scale else 0.0)
res = 4 * 2 ** scale
img = torch.rand(batch, 3, res, res, device="cuda") * 2 - 1
zd, zg, eps = torch.randn(batch, 512, device="cuda"), torch.randn(batch, 512, device="cuda"), torch.rand(batch, 1, device="cuda")
for _ in range(3):
    tr.train_step(img, latent_D=zd, latent_G=zg, eps=eps)
torch.cuda.synchronize()
SKIP = ("view", "reshape", "as_strided", "detach", "alias", "expand", "permute", "transpose", "t.", "select", "slice",
        "unsqueeze", "squeeze", "empty", "_local_scalar", "is_", "size", "stride", "numel", "storage", "set_", "resize_",
        "unbind", "split", "chunk", "contiguous", "_unsafe_view", "lift_fresh", "clone_no")
sites = collections.Counter()


class Rec(TorchDispatchMode):
    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        name = str(func).replace("aten.", "")
        if not any(name.startswith(s) for s in SKIP):
            flat = [a for a in list(args) + list((kwargs or {}).values()) if isinstance(a, torch.Tensor)]
            if any(t.is_cuda for t in flat) or "device" in (kwargs or {}):
                fr = "?"
                for f in reversed(traceback.extract_stack()):
                    if "/pggan_b200/" in f.filename:
                        fr = f"{f.filename.split('/')[-1]}:{f.lineno} {f.name}"
                        break
                numel = max([t.numel() for t in flat], default=0)
                sites[(name, fr, "big" if numel > 1 << 16 else "small")] += 1
        return func(*args, **(kwargs or {}))


with Rec():
    tr.train_step(img, latent_D=zd, latent_G=zg, eps=eps)
    torch.cuda.synchronize()
print(f"{wl}: {sum(sites.values())} recorded ATen ops on CUDA tensors in one step (views / metadata ops skipped)")
for (name, frame, sz), n in sites.most_common(70):
    print(f"{n:4d}  {name:30s} {sz:5s} {frame}")
