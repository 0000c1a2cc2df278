"""Mirror of the PGGAN half of reference lib/utils.py: update_net, setup_ddp, upscale2d,
downscale2d, num_flat_features."""
from __future__ import annotations

import torch

from . import ops


def update_net(optimizer, loss):
    """reference lib/utils.py:72-75"""
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()


def setup_ddp(gpu, ngpus_per_node, init_method="tcp://127.0.0.1:3456"):
    """reference lib/utils.py:78-83 (NCCL, single node)"""
    torch.distributed.init_process_group(backend="nccl", init_method=init_method, world_size=ngpus_per_node,
                                         rank=gpu)


def _is_image(x):
    return x.dim() == 4 and x.shape[1] <= 4 and x.dtype == torch.float32


def upscale2d(x, factor=2):
    """nearest-neighbour x`factor` — reference lib/utils.py:106-118 (exact indexing)"""
    assert isinstance(factor, int) and factor >= 1
    if factor == 1:
        return x
    assert factor == 2, "the PGGAN path only up-samples by 2"
    if _is_image(x):                       # fp32 RGB stays fp32 NCHW: plain replication
        return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    return ops.export_act(ops.Upsample2x.apply(x, 1.0))


def downscale2d(x, factor=2):
    """avg_pool2d — reference lib/utils.py:120-124"""
    assert isinstance(factor, int) and factor >= 1
    if factor == 1:
        return x
    assert factor == 2, "the PGGAN path only down-samples by 2"
    if _is_image(x):
        return torch.nn.functional.avg_pool2d(x, (2, 2))
    return ops.export_act(ops.Pool2x2.apply(x, 0.25))


def num_flat_features(x):
    """reference lib/utils.py:126-131"""
    size = x.size()[1:]
    num_features = 1
    for s in size:
        num_features *= s
    return num_features
