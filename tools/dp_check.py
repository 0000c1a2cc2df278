"""Data-parallel correctness check on real GPUs (torchrun, >= 2 ranks): after a few training steps with a DIFFERENT
batch per rank, every replica must hold bit-identical parameters (same all-reduced gradients -> same Adam update).
A parameter whose gradient missed its all-reduce (e.g. a broken bucket-ready notification) diverges immediately.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import pggan_b200 as pg  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pg.set_compute_dtype(torch.bfloat16)
torch.manual_seed(0)
width, scale, B = 64, 3, 8
tr = pg.ProgressiveGAN(pg.make_args(latent_dim=width, depths=[width] * (scale + 1)), local, variant="W")
tr.set_scale(scale, 0.5)
assert tr.data_parallel
torch.manual_seed(1000 + rank)                       # different data per rank
for step in range(3):
    img = torch.rand(B, 3, 32, 32, device="cuda") * 2 - 1
    tr.train_step(img)
torch.cuda.synchronize()
worst = 0.0
for net in (tr.G, tr.D):
    for p in net.parameters():
        ref = p.detach().clone()
        dist.broadcast(ref, src=0)
        worst = max(worst, float((p.detach() - ref).abs().max()))
w = torch.tensor([worst], device="cuda")
dist.all_reduce(w, op=dist.ReduceOp.MAX)
moved = float(sum((p.detach().float() ** 2).sum() for p in tr.D.parameters()))
if rank == 0:
    print(f"dp_check: world {world}, max |param - rank0 param| over all ranks = {float(w):.3e} (must be 0); "
          f"|D|^2 = {moved:.4f}", flush=True)
    assert float(w) == 0.0
dist.barrier()
os._exit(0)
