"""Data-parallel host logic on CPU: world_size 2, gloo.  GradSync's bucketed all-reduce of the flat
gradient buffer must give the SUM over ranks (so grad_scale = 1/world yields the DDP mean), fire once
per step even with two backward passes, and cope with parameters that receive no gradient."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pggan_b200.dist import GradSync, broadcast_parameters
    from pggan_b200.optim import FlatGrads
    torch.manual_seed(100 + rank)                      # different init per rank, like a fresh add_block
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 8),
                              torch.nn.Linear(8, 4))
    stale = torch.nn.Linear(3, 3)                      # registered but never used: no gradient arrives
    broadcast_parameters(net)
    params = list(reversed(list(net.parameters()))) + list(stale.parameters())
    grads = FlatGrads(params)
    sync = GradSync(grads, bucket_bytes=1024)          # several small buckets
    assert len(sync.buckets) > 2
    torch.manual_seed(7 + rank)
    x1, x2 = torch.randn(5, 16), torch.randn(5, 16)
    # three identical steps: the first armed backward LEARNS the contribution counts (every bucket is reduced in
    # finish()), the later ones launch each bucket from the contribution that completes it
    early = []
    for step in range(3):
        grads.zero()
        net(x1).pow(2).sum().backward()                # first backward of the step: NOT reduced yet
        local_after_first = grads.flat.clone()
        sync.arm()
        net(x2).sum().backward()                       # last backward: buckets go out as they complete
        early.append(sum(sync._launched))              # launched before finish()
        sync.finish()
        if step == 0:
            first = grads.flat.clone()
        else:
            assert torch.equal(first, grads.flat), "counted launch changed the reduced gradients"
    assert early[0] == 0 and early[1] >= len(sync.buckets) - 1 and early[2] == early[1], early
    # an in-place contribution reported through the sink listener counts like an autograd accumulation
    from pggan_b200 import ops
    sync.relearn()
    for step in range(2):
        grads.zero()
        sync.arm()
        net(x2).sum().backward()
        with torch.no_grad():
            params[0].grad.add_(1.0)
        ops._sunk(params[0])
        sync.finish()
    assert sync._expected[sync._param_bucket[id(params[0])]] >= 2
    grads.zero()
    net(x1).pow(2).sum().backward()
    local_after_first = grads.flat.clone()
    sync.relearn()
    sync.arm()
    net(x2).sum().backward()
    sync.finish()
    q.put((rank, grads.flat.tolist(), local_after_first.tolist(), [p.detach().tolist() for p in net.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


def test_gradsync_sums_over_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, f0, l0, p0), (_, f1, l1, p1) = res
    f0, f1 = torch.tensor(f0), torch.tensor(f1)
    p0, p1 = [torch.tensor(t) for t in p0], [torch.tensor(t) for t in p1]
    assert torch.equal(f0, f1)                         # every rank holds the same reduced buffer
    for a, b in zip(p0, p1):
        assert torch.equal(a, b)                       # broadcast made the replicas identical
    # single-process recomputation of both ranks' full-step gradients
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 8), torch.nn.Linear(8, 4))
    with torch.no_grad():
        for p, v in zip(net.parameters(), p0):
            p.copy_(v)
    total = None
    for rank in range(world):
        torch.manual_seed(7 + rank)
        x1, x2 = torch.randn(5, 16), torch.randn(5, 16)
        net.zero_grad()
        (net(x1).pow(2).sum() + net(x2).sum()).backward()
        g = torch.cat([p.grad.reshape(-1) for p in reversed(list(net.parameters()))])
        total = g if total is None else total + g
    n = total.numel()
    # flat layout pads each parameter to a multiple of 4 floats: compare parameter by parameter
    from pggan_b200.optim import _pad4
    off_flat, off_ref = 0, 0
    for p in reversed(list(net.parameters())):
        k = p.numel()
        assert torch.allclose(f0[off_flat:off_flat + k], total[off_ref:off_ref + k], rtol=1e-5, atol=1e-6)
        off_flat += _pad4(k)
        off_ref += k
    assert float(f0[off_flat:].abs().max()) == 0.0     # the stale layer stayed zero
