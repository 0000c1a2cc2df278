"""Data-parallel gradient synchronisation: bucketed all-reduce overlapped with backward.

What the reference intends with --use_mGPU (DistributedDataParallel, lib/model.py:74-79) but never
actually runs (it keeps only `.module`): every rank holds a full replica and its own
batch, gradients are summed across ranks once per optimizer step and the 1/world_size is folded into
the Adam kernel.  Buckets are contiguous slices of the flat gradient buffer (optim.FlatGrads) in
gradient-ready order; a bucket's all-reduce is launched (async, on NCCL's stream over NVLink/NVSwitch)
by the gradient contribution that completes it, so communication overlaps the rest of the backward
pass.  Only live parameters are registered.

The D step of the WGAN-GP variant has two backward calls (GP's own, then loss_D's): arm() is called
before the LAST one only, so each bucket is reduced once per step.  Backend-agnostic (gloo on CPU in
the tests, NCCL on the GPUs).
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist

from .optim import FlatGrads


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """rank `src` weights to everyone — at start and after every add_block (new blocks are randomly
    initialised per rank; the reference never re-broadcasts)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for p in module.parameters():
        dist.broadcast(p.data, src=src, group=group)


class GradSync:
    """Bucket readiness is COUNTED, not hooked: the training step is a fixed program, so every parameter receives the
    same number of gradient contributions in every armed backward.  A contribution is either an in-place accumulation
    by a wgrad / bias kernel (ops._GRAD_SINK: reported through ops.add_sink_listener) or autograd's own accumulation
    (post-accumulate-grad hook: the RGB layers and whatever else is not sunk).  The first armed backward LEARNS the
    counts (its buckets are all reduced in finish()); from then on a bucket's all-reduce is launched by the
    contribution that completes it, so the in-place gradient sink stays on under data parallelism and NCCL overlaps
    the rest of backward."""

    def __init__(self, grads: FlatGrads, bucket_bytes: int = 32 << 20, group=None):
        self.grads = grads
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.buckets: List[dict] = []
        self._param_bucket = {}
        cur = {"lo": 0, "hi": 0, "params": []}
        for p, off in zip(grads.params, grads.offsets):
            n_pad = (p.numel() + 3) // 4 * 4
            if cur["params"] and (off + n_pad - cur["lo"]) * 4 > bucket_bytes:
                self.buckets.append(cur)
                cur = {"lo": off, "hi": off, "params": []}
            cur["params"].append(p)
            cur["hi"] = off + n_pad
        self.buckets.append(cur)
        for bi, b in enumerate(self.buckets):
            for p in b["params"]:
                self._param_bucket[id(p)] = bi
        self._armed = False
        self._expected = None            # learned: contributions per bucket in one armed backward
        self._seen: List[int] = []
        self._launched: List[bool] = []
        self._handles = []
        self._hooks = [p.register_post_accumulate_grad_hook(self._notify) for p in grads.params]
        from . import ops
        ops.add_sink_listener(self._notify)

    # ------------------------------------------------------------------------------------
    def arm(self):
        """call right before the last backward of the optimizer step"""
        self._armed = self.world > 1
        self._seen = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)
        self._handles = []

    def _launch(self, bi: int):
        b = self.buckets[bi]
        self._launched[bi] = True
        view = self.grads.flat[b["lo"]:b["hi"]]
        self._handles.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _notify(self, p):
        if not self._armed:
            return
        bi = self._param_bucket.get(id(p))
        if bi is None:
            return
        if self._launched[bi]:
            raise RuntimeError("GradSync: a gradient contribution arrived after its bucket was reduced — the step's "
                               "structure changed; create a new GradSync (set_optimizers) or call relearn()")
        self._seen[bi] += 1
        if self._expected is None or self._seen[bi] < self._expected[bi]:
            return
        # identical launch order on every rank: buckets go out strictly in index order
        for j in range(len(self.buckets)):
            if self._launched[j]:
                continue
            if self._seen[j] >= self._expected[j]:          # (a bucket that expects nothing is complete as it is)
                self._launch(j)
            else:
                break

    def relearn(self):
        self._expected = None

    def finish(self):
        """after the backward: launch whatever did not fire (unused parameters, the learning pass), wait for all.
        Gradients then hold the SUM over ranks; pass grad_scale = 1/world to the optimizer."""
        if not self._armed:
            return
        if self._expected is None:
            self._expected = list(self._seen)
        elif self._expected != self._seen:
            # fewer contributions than learned (more would have raised above): learn again, reduce the rest now
            self._expected = list(self._seen)
        for bi in range(len(self.buckets)):
            if not self._launched[bi]:
                self._launch(bi)
        for h in self._handles:
            h.wait()
        self._armed = False

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
        from . import ops
        ops.remove_sink_listener(self._notify)
