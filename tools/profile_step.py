"""Kernel timeline of the training step with torch.profiler (CUPTI): true warm durations + idle gaps.
    python tools/profile_step.py [graph|eager] [workload]"""
import collections
import re
import sys

import torch

sys.path.insert(0, ".")
import pggan_b200 as pg  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402

from bench import WORKLOADS  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "graph"
workload = sys.argv[2] if len(sys.argv) > 2 else "c2_32x32_b64_fade"
scale, B, depths, _ = WORKLOADS[workload]
pg.set_compute_dtype(torch.float16)
torch.manual_seed(0)
tr = pg.ProgressiveGAN(pg.make_args(latent_dim=512, depths=depths), 0, variant="W")
tr.set_scale(scale, 0.5 if scale else 0.0)
res = 4 * 2 ** scale
img = torch.rand(B, 3, res, res, device="cuda") * 2 - 1
zd, zg, eps = torch.randn(B, 512, device="cuda"), torch.randn(B, 512, device="cuda"), torch.rand(B, 1, device="cuda")


def eager():
    tr.train_step(img, latent_D=zd, latent_G=zg, eps=eps)


for _ in range(3):
    eager()
step = eager
if mode == "graph":
    step = tr.capture(B)
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Memcpy" not in e.name
       and "Memset" not in e.name]
kern = sorted([(e.time_range.start, e.time_range.end, e.name) for e in evs], key=lambda t: t[0])
t0, t1 = kern[0][0], max(k[1] for k in kern)
busy = sum(k[1] - k[0] for k in kern)
print(f"{workload} mode {mode}: 3 steps span {(t1-t0)/1e3:.2f} ms, kernel-busy {busy/1e3:.2f} ms, launches {len(kern)}, "
      f"idle {(t1-t0-busy)/1e3:.2f} ms")
agg = collections.defaultdict(lambda: [0, 0.0])
for s, e, n in kern:
    n = re.sub(r"pg::|void |\(.*", "", n)[:60]
    agg[n][0] += 1
    agg[n][1] += e - s
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{t/busy*100:5.1f}% {t/3e3:8.3f} ms/step n/step={c/3:6.1f} avg {t/c:8.1f} us  {n}")
gaps = [kern[i + 1][0] - kern[i][1] for i in range(len(kern) - 1)]
gaps = [g for g in gaps if g > 0]
print("gaps: total %.2f ms, mean %.2f us, >5us: %d, >20us: %d" % (sum(gaps) / 1e3, sum(gaps) / max(1, len(gaps)),
                                                                  sum(g > 5 for g in gaps), sum(g > 20 for g in gaps)))
