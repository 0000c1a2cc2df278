"""Data-parallel correctness check on real GPUs (torchrun, >= 2 ranks), real G and D at 32x32:

 (1) GRADIENT MEAN.  With a different batch per rank, the flat gradient buffers after the bucketed all-reduce must equal
     the SUM over ranks of the per-shard gradients (the Adam kernel folds the 1/world): every rank recomputes all
     shards on a non-parallel twin trainer (same weights, rank-local minibatch-stddev groups) and compares.  This
     exercises the counted bucket readiness (in-place gradient sink ON under DP) in its learning step and in its
     steady state, eagerly and inside a captured CUDA graph.
 (2) REPLICAS STAY IDENTICAL.  After a few real training steps every replica must hold bit-identical parameters.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import pggan_b200 as pg  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pg.set_compute_dtype(torch.float16)
torch.manual_seed(0)
width, scale, B = 64, 3, 8
res = 4 * 2 ** scale


def make(dp):
    t = pg.ProgressiveGAN(pg.make_args(latent_dim=width, depths=[width] * (scale + 1)), local, variant="W",
                          data_parallel=dp, bucket_bytes=256 << 10)       # small buckets: several per network
    t.set_scale(scale, 0.5)
    return t


tr = make(True)
assert tr.data_parallel and len(tr.sync_D.buckets) > 2
twin = make(False)
twin.G.load_state_dict(tr.G.state_dict())
twin.D.load_state_dict(tr.D.state_dict())
for t in (tr, twin):                                   # keep the gradients: no parameter update in part (1)
    t.opt_G.step = lambda: None
    t.opt_D.step = lambda: None


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def gather(t):
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return out


gen = torch.Generator(device="cuda").manual_seed(1000 + rank)       # different data per rank
worst = 0.0
modes = ["eager (learning pass)", "eager (counted launches)", "eager (counted launches)", "cuda graph replay"]
replay = None
for it, mode in enumerate(modes):
    img = torch.rand(B, 3, res, res, device="cuda", generator=gen) * 2 - 1
    zd = torch.randn(B, width, device="cuda", generator=gen)
    zg = torch.randn(B, width, device="cuda", generator=gen)
    eps = torch.rand(B, 1, device="cuda", generator=gen)
    if mode.startswith("cuda graph"):
        replay = tr.capture(B)
        replay(img, zd, zg, eps)
    else:
        tr.train_step(img, latent_D=zd, latent_G=zg, eps=eps)
    torch.cuda.synchronize()
    got_G, got_D = tr.opt_G.grads.flat.clone(), tr.opt_D.grads.flat.clone()
    sum_G, sum_D = torch.zeros_like(got_G), torch.zeros_like(got_D)
    for r, (i_r, zd_r, zg_r, e_r) in enumerate(zip(gather(img), gather(zd), gather(zg), gather(eps))):
        twin.train_step(i_r, latent_D=zd_r, latent_G=zg_r, eps=e_r)
        sum_G += twin.opt_G.grads.flat
        sum_D += twin.opt_D.grads.flat
    eG, eD = rel(got_G, sum_G), rel(got_D, sum_D)
    early = (sum(tr.sync_D._launched), len(tr.sync_D.buckets))
    worst = max(worst, eG, eD)
    if rank == 0:
        print(f"dp_check (1) {mode:26s}: all-reduced grads vs sum of {world} per-shard grads: G {eG:.2e}  D {eD:.2e}   "
              f"(D buckets {early[1]})", flush=True)
w = torch.tensor([worst], device="cuda")
dist.all_reduce(w, op=dist.ReduceOp.MAX)
# fp32 atomics (bias column sums, RGB weight gradients) make two evaluations of the same shard differ in the last bits
assert float(w) < 1e-4, float(w)

# ---- (2) real updates: replicas stay bit-identical
del tr, twin, replay
tr = make(True)
torch.manual_seed(2000 + rank)
for step in range(3):
    tr.train_step(torch.rand(B, 3, res, res, device="cuda") * 2 - 1)
torch.cuda.synchronize()
worst = 0.0
for net in (tr.G, tr.D):
    for p in net.parameters():
        ref = p.detach().clone()
        dist.broadcast(ref, src=0)
        worst = max(worst, float((p.detach() - ref).abs().max()))
w2 = torch.tensor([worst], device="cuda")
dist.all_reduce(w2, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"dp_check (2) world {world}: max |param - rank0 param| over all ranks after 3 steps = {float(w2):.3e} (must be 0)",
          flush=True)
    assert float(w2) == 0.0
    print("dp_check OK", flush=True)
dist.barrier()
os._exit(0)
