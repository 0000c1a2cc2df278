"""Fused multi-tensor Adam over the live parameters of one network.

torch.optim.Adam semantics as the reference configures it (lib/model.py:95-97: lr_G 1e-4 / lr_D 1e-5,
betas (beta1, beta2), eps 1e-8, no weight decay, no amsgrad); parameters whose gradient is None are
skipped exactly like torch skips them (the stale RGB layers of SURVEY.md D16 are simply not registered).

Gradients live in ONE flat fp32 buffer (`p.grad` is a view into it): zero_grad is one memset, the
data-parallel all-reduce runs on contiguous bucket slices (dist.GradSync), and the update is one
kernel launch (pggan_adam_step) that also folds the 1/world_size of the gradient mean.
The step counter lives on the device, so the whole step is CUDA-graph capturable.
"""
from __future__ import annotations

import ctypes
from typing import Iterable, List

import torch

from . import _lib as L
from . import ops


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


class FlatGrads:
    """one flat gradient buffer for an ordered list of parameters (order = gradient-ready order)"""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        if not self.params:
            raise ValueError("no parameters")
        dev = self.params[0].device
        self.offsets = []
        off = 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise TypeError("master parameters must be fp32")
            self.offsets.append(off)
            off += _pad4(p.numel())
        self.total = off
        self.flat = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.attach()
        if dev.type == "cuda":
            ops.register_grad_sink(self.params)      # kernels may accumulate straight into these views

    @staticmethod
    def view_like(buf, p, off):
        """the slice of a flat per-element buffer (gradients, Adam moments) as a tensor of the parameter's logical shape
        WITH THE PARAMETER'S STRIDES (channels_last conv weights are physically [O][kh][kw][I]: element k of the slice
        belongs to element k of the parameter's memory)"""
        flat = buf[off:off + p.numel()]
        if ops.is_ohwi(p):
            o, i, kh, kw = p.shape
            return flat.view(o, kh, kw, i).permute(0, 3, 1, 2)
        return flat.view_as(p)

    def _view(self, p, off):
        return self.view_like(self.flat, p, off)

    def attach(self):
        for p, off in zip(self.params, self.offsets):
            p.grad = self._view(p, off)

    def zero(self):
        self.flat.zero_()
        # a backward with create_graph=True (or set_to_none) may have replaced the views: re-attach
        for p, off in zip(self.params, self.offsets):
            g = p.grad
            if g is None or g.data_ptr() != self.flat.data_ptr() + 4 * off or g.stride() != p.stride():
                p.grad = self._view(p, off)


class FusedAdam:
    """drop-in for the two `torch.optim.Adam` objects of reference lib/model.py:95-97"""

    def __init__(self, params, lr, betas=(0.0, 0.99), eps=1e-8, grad_scale=1.0, all_params=None):
        """all_params: every parameter of the network in `module.parameters()` order (the order in which the
        reference's torch.optim.Adam indexes them) — gives state_dict() / load_state_dict() the reference's indices;
        parameters that are not registered here (stale RGB layers) simply have no state, as in torch."""
        self.grads = params if isinstance(params, FlatGrads) else FlatGrads(params)
        self.all_params = list(all_params) if all_params is not None else list(self.grads.params)
        self.lr, self.betas, self.eps, self.grad_scale = float(lr), (float(betas[0]), float(betas[1])), float(eps), \
            float(grad_scale)
        g = self.grads
        dev = g.flat.device
        if dev.type != "cuda":
            raise L.PgganLibraryError("FusedAdam needs CUDA parameters (no CPU fallback)")
        self.exp_avg = torch.zeros_like(g.flat)
        self.exp_avg_sq = torch.zeros_like(g.flat)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
        chunk = L.load().pggan_adam_chunk_size()
        entries = (L.AdamEntry * len(g.params))()
        chunk_tensor, chunk_offset = [], []
        # 16-bit packed fprop operands of the channels_last conv weights, rewritten by the Adam kernel itself
        self._shadowed = [p for p in g.params if ops.is_ohwi(p)] if ops._ACT_FMT != L.PG_F32 else []
        self.shadow = None
        shadow_off = {}
        if self._shadowed:
            tot = 0
            for p in self._shadowed:
                shadow_off[id(p)] = tot
                tot += (p.numel() + 7) // 8 * 8
            self.shadow = torch.zeros(tot, dtype=ops.get_compute_dtype(), device=dev)
            for p in self._shadowed:
                o, i, kh, kw = p.shape
                t = self.shadow[shadow_off[id(p)]:shadow_off[id(p)] + p.numel()].view(o, kh * kw * i)
                ops.register_shadow(p, "f", ops._ACT_FMT, None, t)
        for i, (p, off) in enumerate(zip(g.params, g.offsets)):
            if not (p.is_contiguous() or ops.is_ohwi(p)):
                raise ValueError("parameters must be dense (contiguous or channels_last)")
            if id(p) in shadow_off:
                entries[i].shadow = self.shadow.data_ptr() + 2 * shadow_off[id(p)]
                entries[i].shadow_dtype = ops._ACT_FMT
            entries[i].param = p.data_ptr()
            entries[i].grad = g.flat.data_ptr() + 4 * off
            entries[i].exp_avg = self.exp_avg.data_ptr() + 4 * off
            entries[i].exp_avg_sq = self.exp_avg_sq.data_ptr() + 4 * off
            entries[i].numel = p.numel()
            for c0 in range(0, p.numel(), chunk):
                chunk_tensor.append(i)
                chunk_offset.append(c0)
        raw = bytes(entries)
        self._table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        self._chunk_tensor = torch.tensor(chunk_tensor, dtype=torch.int32, device=dev)
        self._chunk_offset = torch.tensor(chunk_offset, dtype=torch.int64, device=dev)
        self._n_tensors, self._n_chunks = len(g.params), len(chunk_tensor)
        self._param_ptrs = [p.data_ptr() for p in g.params]
        self.sync_shadows()

    def sync_shadows(self):
        """(re)write the packed 16-bit fprop copies from the current parameter values and mark them valid — at
        construction, and after anything but this optimizer's own step changed the parameters (load_state_dict,
        the capture warm-up's restore)"""
        for p in self._shadowed:
            key = (id(p), "f", ops._ACT_FMT, None)
            ent = ops._SHADOWS.get(key)
            if ent is None or ent[0] is not p:
                continue
            o, i, kh, kw = p.shape
            src = p.detach().permute(0, 2, 3, 1).reshape(-1)           # the parameter's memory, in order
            L.check(L.load().pggan_convert(ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(ent[1].data_ptr()),
                                           p.numel(), L.PG_F32, ops._ACT_FMT, ops._s()), "pggan_convert")
        ops.shadows_synced(self._shadowed)

    # torch.optim.Optimizer-like surface used by update_net (reference lib/utils.py:72-75)
    def zero_grad(self, set_to_none: bool = False):
        self.grads.zero()

    def step(self):
        for p, ptr in zip(self.grads.params, self._param_ptrs):
            if p.data_ptr() != ptr:
                raise RuntimeError("a parameter was re-allocated after the optimizer was built "
                                   "(re-create the optimizers after add_block/.to(), as the reference does)")
        b1, b2 = self.betas
        L.check(L.load().pggan_adam_step(ctypes.c_void_p(self._table.data_ptr()), self._n_tensors, self._n_chunks,
                                         ctypes.c_void_p(self._chunk_tensor.data_ptr()),
                                         ctypes.c_void_p(self._chunk_offset.data_ptr()),
                                         ctypes.c_void_p(self.step_count.data_ptr()), self.lr, b1, b2, self.eps,
                                         self.grad_scale, ops._s()), "pggan_adam_step")
        ops.clear_pack_cache(self.grads.params)     # packed bf16/tf32 copies of THESE weights are stale now ...
        ops.shadows_synced(self._shadowed)          # ... except the fprop operands this kernel has just rewritten

    # ---- torch.optim.Adam wire format (reference lib/checkpoint.py:24-25, pggan/model.py:97-101) ----------------
    def state_dict(self):
        """the dict torch.optim.Adam(module.parameters(), ...).state_dict() would hold: per-parameter
        {step, exp_avg, exp_avg_sq} keyed by the parameter's index in module.parameters(), plus one param_group"""
        index = {id(p): i for i, p in enumerate(self.all_params)}
        step = int(self.step_count.item())
        state = {}
        if step > 0:
            for p, off in zip(self.grads.params, self.grads.offsets):
                n = p.numel()
                state[index[id(p)]] = {"step": torch.tensor(float(step)),
                                       "exp_avg": FlatGrads.view_like(self.exp_avg, p, off).clone(),
                                       "exp_avg_sq": FlatGrads.view_like(self.exp_avg_sq, p, off).clone()}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(len(self.all_params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        """accepts the format above (written by this class or by torch.optim.Adam over the same module)"""
        if "state" not in sd or "param_groups" not in sd:
            raise ValueError("FusedAdam.load_state_dict: not a torch.optim.Adam state dict (no 'state'/'param_groups')")
        group = sd["param_groups"][0]
        if len(group["params"]) != len(self.all_params):
            raise ValueError(f"FusedAdam.load_state_dict: the checkpoint's optimizer covers {len(group['params'])} "
                             f"parameters, this network has {len(self.all_params)} (grow it to the checkpoint's "
                             "scale_index first)")
        self.lr = float(group["lr"])
        self.betas = (float(group["betas"][0]), float(group["betas"][1]))
        self.eps = float(group["eps"])
        pos = {pid: i for i, pid in enumerate(group["params"])}        # saved id -> position in module.parameters()
        by_pos = {pos[k]: v for k, v in sd["state"].items()}
        index = {id(p): i for i, p in enumerate(self.all_params)}
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        steps = set()
        for p, off in zip(self.grads.params, self.grads.offsets):
            st = by_pos.get(index[id(p)])
            if st is None:
                continue
            n = p.numel()
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError("FusedAdam.load_state_dict: moment shape mismatch for a parameter of shape "
                                 f"{tuple(p.shape)}")
            FlatGrads.view_like(self.exp_avg, p, off).copy_(st["exp_avg"])
            FlatGrads.view_like(self.exp_avg_sq, p, off).copy_(st["exp_avg_sq"])
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"FusedAdam.load_state_dict: parameters carry different step counts {sorted(steps)}")
        self.step_count.fill_(steps.pop() if steps else 0)
        self.sync_shadows()
