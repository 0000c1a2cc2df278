"""Kernel wrappers and double-differentiable autograd Functions over libpggan_b200.so.

Every Function's backward is itself built from Functions of this file, so
``torch.autograd.grad(..., create_graph=True)`` (R1 — reference lib/loss.py:125-135; WGAN-GP —
reference pggan/loss.py:78-90) can be differentiated again: {fprop, dgrad, wgrad} are closed under
differentiation, LeakyReLU contributes only its mask, pooling / up-sampling /
blends are linear, and minibatch-stddev has a true second-order kernel.

Feature maps are channels_last tensors ([B,C,H,W] logical, NHWC physical); images/logits/parameters are fp32.

Precision modes (set_compute_dtype):
  torch.float16 (default, the benchmarked mode)  activations, packed weights AND gradient tensors are IEEE half
          (11-bit significand: meets north_star's 1e-2 image/logit and 1e-3 gradient-penalty bars with margin), fp32
          accumulate and epilogues, conversions saturate at +-65504 (never inf).  Range: the trainer applies a STATIC
          LOSS SCALE (trainer.ProgressiveGAN.loss_scale = 1024) that the fused Adam kernel divides out; the penalty's
          first-order chain is seeded with 1 and is kept near unit norm by the penalty itself.
  "mixed"  ACTIVATIONS and packed weights are IEEE half (11-bit significand: meets
          north_star's 1e-2 / 1e-3 bars), GRADIENT tensors are bf16 (fp32's exponent range: no loss scaling), fp32
          accumulate.  tcgen05.mma kind::f16 wants BOTH operands in one format (a bf16 x fp16 instruction descriptor
          faults: measured, tools/probe_mixed.py), so every GEMM runs in the format of its activation/gradient operand:
          forward fp16 x fp16; dgrad and the double backward's fprop bf16 x bf16 (the weights have a second, bf16
          packed copy); wgrad bf16 x bf16 with the saved activation re-rounded to bf16 on the way in (pggan_convert).
          No loss scaling needed, but the bf16 rounding of the penalty's first-order gradient chain leaves the GP value
          at 1e-3..2e-3 where ||grad|| is near 1 (measured), and the conversions cost ~4 % of a 256x256 step.
  torch.bfloat16                           everything bf16
  torch.float32                            3xTF32 operands, fp32 accumulate: the parity mode

Precision follows the TENSOR: every Function computes in the family of its input feature tensor — an fp32 tensor runs
the fp32 (3xTF32) path and yields fp32, a 16-bit tensor the 16-bit path — in forward, backward and double backward
alike (the backward reads the family off its saved tensors, autograd's own dtype checks cast at a family boundary).
The precision MODE only decides what `enter()` converts external fp32 inputs to; the networks call it once at their
entry.  In the 16-bit modes they stay in fp32 while the image is 4x4 (scale 0): there the image gradient has 48
elements per sample and ||grad|| ~ 0.8, where GP = 10*(||g||-1)^2 amplifies a norm error 8x (measured 2.6e-3..6e-3 GP
error at 16 bits against 1e-3 asked), and a 2 ms latency-bound step gains nothing from 16 bits anyway.

Two tensor CLASSES therefore exist: "a" (activation-like: produced by a forward pass) and "g" (gradient-like: produced
by a backward pass, including everything the double backward derives from gradients).  Every 16-bit feature tensor
carries the torch dtype of the GRADIENT class (`get_compute_dtype()`), so autograd's own gradient accumulation and
dtype checks act on honestly-labelled tensors; in mixed mode an activation tensor is therefore a bf16-LABELLED
container of fp16 bits.  Only kernels of this library read activations; `to_float()` / the networks' fp32 outputs are
the way out, and the per-layer public `forward()` of layers.py exports honest fp16.  The class of every operand is
static at each call site and passed down as `xc` / `cls` arguments.
"""
from __future__ import annotations

import contextlib
import ctypes
import os
from typing import Optional

import torch

from . import _lib as L

A, G = "a", "g"                # tensor classes, see the module docstring
MIXED = "mixed"
_MODE = "fp16"
_COMPUTE_DTYPE = torch.float16           # torch label of 16-bit feature tensors == real format of the gradient class
_ACT_FMT, _GRAD_FMT = L.PG_F16, L.PG_F16
_REAL = {L.PG_BF16: torch.bfloat16, L.PG_F16: torch.float16, L.PG_F32: torch.float32}
_SKIP_PARAM_GRADS = False      # see input_grad_only()
PROFILE = None                 # bench.py sets a dict: (kind, shape) -> [(start_event, end_event), ...]


class _Timed:
    """CUDA-event bracket on the launch stream around one GEMM-class launch (only when PROFILE is set)"""

    def __init__(self, kind, shape):
        self.key = (kind, shape) if PROFILE is not None else None

    def __enter__(self):
        if self.key is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *a):
        if self.key is not None:
            self.e1.record()
            PROFILE.setdefault(self.key, []).append((self.e0, self.e1))


def set_compute_dtype(dtype) -> None:
    """torch.float16 (the default), "mixed" (fp16 activations/weights, bf16 gradients), torch.bfloat16 (same kernels
    and speed, 8-bit significand) or torch.float32 (3xTF32 operands: the parity mode)."""
    global _COMPUTE_DTYPE, _ACT_FMT, _GRAD_FMT, _MODE
    if dtype == MIXED:
        _MODE, _COMPUTE_DTYPE, _ACT_FMT, _GRAD_FMT = MIXED, torch.bfloat16, L.PG_F16, L.PG_BF16
    elif dtype == torch.bfloat16:
        _MODE, _COMPUTE_DTYPE, _ACT_FMT, _GRAD_FMT = "bf16", torch.bfloat16, L.PG_BF16, L.PG_BF16
    elif dtype == torch.float16:
        _MODE, _COMPUTE_DTYPE, _ACT_FMT, _GRAD_FMT = "fp16", torch.float16, L.PG_F16, L.PG_F16
    elif dtype == torch.float32:
        _MODE, _COMPUTE_DTYPE, _ACT_FMT, _GRAD_FMT = "fp32", torch.float32, L.PG_F32, L.PG_F32
    else:
        raise ValueError('compute dtype must be "mixed", torch.bfloat16, torch.float16 or torch.float32 (3xTF32)')
    _KIND_CACHE.clear()
    _PACK_CACHE.clear()


def get_compute_dtype() -> torch.dtype:
    """torch dtype that 16-bit feature tensors carry (the real format of the gradient class)"""
    return _COMPUTE_DTYPE


def get_precision_mode() -> str:
    """one of "mixed", "bf16", "fp16", "fp32" """
    return _MODE


def _fmt(cls: str) -> int:
    return _ACT_FMT if cls == A else _GRAD_FMT


@contextlib.contextmanager
def input_grad_only():
    """Inside this context the backward of the GEMM-class Functions skips weight/bias gradients.
    Used around ``autograd.grad(d_out, x_in, create_graph=True)`` in R1/GP, where only the input
    gradient is requested (a Python autograd.Function cannot see that from ctx.needs_input_grad)."""
    global _SKIP_PARAM_GRADS
    prev = _SKIP_PARAM_GRADS
    _SKIP_PARAM_GRADS = True
    try:
        yield
    finally:
        _SKIP_PARAM_GRADS = prev


# ------------------------------------------------------------------------------- gradient sinks
# Parameters registered here (by optim.FlatGrads) own a persistent fp32 `.grad` view into a flat buffer.  Inside a
# `grad_sink()` context (the trainer opens one around its own backward() calls) a plain backward (no create_graph)
# lets the GEMM-class Functions ACCUMULATE weight / bias gradients straight into that view (pggan_unpack_wgrad
# accumulate=1, atomics of the bias column sums) and return None to autograd: no temporary gradient tensor, no
# AccumulateGrad add kernel, no zero-fill.  Outside the context the Functions behave like any autograd Function
# (gradients are returned), so torch.autograd.grad / hooks / backward(inputs=...) see what they expect.
# Every in-place accumulation is reported to the listeners (dist.GradSync counts them for bucket readiness).
_GRAD_SINK: dict = {}
_SINK_ENABLED = False
_SINK_LISTENERS: list = []


def register_grad_sink(params):
    import weakref
    for p in params:
        _GRAD_SINK[id(p)] = weakref.ref(p)


def unregister_grad_sink(params):
    for p in params:
        _GRAD_SINK.pop(id(p), None)


@contextlib.contextmanager
def grad_sink(enabled: bool = True):
    global _SINK_ENABLED
    prev = _SINK_ENABLED
    _SINK_ENABLED = enabled
    try:
        yield
    finally:
        _SINK_ENABLED = prev


def add_sink_listener(fn):
    _SINK_LISTENERS.append(fn)


def remove_sink_listener(fn):
    if fn in _SINK_LISTENERS:
        _SINK_LISTENERS.remove(fn)


def _sink(p):
    """the parameter's .grad if gradients may be accumulated into it in place, else None"""
    if p is None or not _SINK_ENABLED:
        return None
    ref = _GRAD_SINK.get(id(p))
    if ref is None or ref() is not p:
        return None
    g = p.grad
    if g is None or g.dtype != torch.float32 or g.shape != p.shape or g.stride() != p.stride() or \
            not (g.is_contiguous() or is_ohwi(g)):
        return None
    return g


def _sunk(p):
    """a gradient contribution of `p` has just been accumulated in place"""
    for fn in _SINK_LISTENERS:
        fn(p)


# ------------------------------------------------------------------------------------ helpers
def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(t: torch.Tensor):
    if not t.is_cuda:
        raise L.PgganLibraryError("pggan_b200 ops run on CUDA tensors only (no CPU fallback)")


def cl(x: torch.Tensor) -> torch.Tensor:
    """Feature tensor in the kernels' physical layout (NHWC for 4-D, row-major for 2-D)."""
    if x.dim() == 4:
        return x.contiguous(memory_format=torch.channels_last)
    return x.contiguous()


class _Import(torch.autograd.Function):
    """external tensor -> internal feature tensor of class `cls` whose real format differs from the label
    (mixed mode activations); the gradient goes back honestly labelled"""

    @staticmethod
    def forward(ctx, x, cls):
        ctx.dtype = x.dtype
        return x.to(_REAL[_fmt(cls)]).view(_COMPUTE_DTYPE)

    @staticmethod
    def backward(ctx, g):
        return g.to(ctx.dtype), None


def _small(x: torch.Tensor) -> bool:
    """the fp32 family: a feature tensor that already is fp32 stays fp32 (see the module docstring)"""
    return x.dtype == torch.float32


def _fmt_of(x: torch.Tensor, cls: str) -> int:
    """element format of a feature tensor: fp32 tensors are fp32 (an island, or the fp32 mode), 16-bit ones carry the
    format of their class"""
    return L.PG_F32 if x.dtype == torch.float32 else _fmt(cls)


class _ToF32(torch.autograd.Function):
    """internal 16-bit feature tensor -> fp32 (entering an fp32 island outside a Function, mixed mode)"""

    @staticmethod
    def forward(ctx, x, cls):
        return to_float(x, cls)

    @staticmethod
    def backward(ctx, g):
        return g.to(_COMPUTE_DTYPE), None


def feat(x: torch.Tensor, cls: str = A, small=None) -> torch.Tensor:
    """feature tensor of class `cls` in the kernels' layout.  Precision family: fp32 stays fp32 (and the fp32 mode is
    all fp32); `small` overrides — True forces fp32, False forces the mode's 16-bit format (a gradient arriving in
    fp32 for a 16-bit layer, the partner of a 16-bit tensor).  A tensor that already carries the 16-bit feature dtype
    is trusted to BE of class `cls`."""
    if small is None:
        small = _small(x)
    if small or _COMPUTE_DTYPE == torch.float32:
        if x.dtype != torch.float32:
            if x.dtype == _COMPUTE_DTYPE and _REAL[_fmt(cls)] != _COMPUTE_DTYPE:      # mixed mode: relabelled bits
                x = _ToF32.apply(x, cls) if (torch.is_grad_enabled() and x.requires_grad) else to_float(x, cls)
            else:
                x = x.float()
        return cl(x)
    if x.dtype != _COMPUTE_DTYPE:
        real = _REAL[_fmt(cls)]
        if real == _COMPUTE_DTYPE:
            x = x.to(real)
        elif torch.is_grad_enabled() and x.requires_grad:
            x = _Import.apply(x, cls)
        else:
            x = x.to(real).view(_COMPUTE_DTYPE)
    return cl(x)


def enter(x: torch.Tensor, cls: str = A, f32: bool = False) -> torch.Tensor:
    """an external tensor (latents, test inputs) as a feature tensor of the current precision mode — or of the fp32
    family if `f32`.  This is the ONE place where fp32 is turned into 16 bits; differentiable."""
    return feat(x, cls, small=bool(f32) or _COMPUTE_DTYPE == torch.float32)


def to_float(x: torch.Tensor, cls: str = A) -> torch.Tensor:
    """fp32 values of an internal feature tensor (no autograd; see ExportFloat)"""
    real = _REAL[_fmt(cls)]
    if x.dtype == _COMPUTE_DTYPE and real != _COMPUTE_DTYPE:
        x = x.view(real)
    return x.float()


def to_honest(x: torch.Tensor, cls: str = A) -> torch.Tensor:
    """the same bits under their real torch dtype"""
    real = _REAL[_fmt(cls)]
    if x.dtype == _COMPUTE_DTYPE and real != _COMPUTE_DTYPE:
        return x.view(real)
    return x


class ExportFloat(torch.autograd.Function):
    """internal activation -> fp32 tensor for the caller (D's `feature` output); differentiable"""

    @staticmethod
    def forward(ctx, x):
        ctx.f32 = x.dtype == torch.float32
        return to_float(x, A)

    @staticmethod
    def backward(ctx, g):
        return feat(g, G, small=ctx.f32)


class ExportAct(torch.autograd.Function):
    """internal activation -> honestly-labelled 16-bit tensor (the public per-layer forward() in mixed mode)"""

    @staticmethod
    def forward(ctx, x):
        return to_honest(x, A)

    @staticmethod
    def backward(ctx, g):
        return feat(g, G, small=False)


def export_act(x: torch.Tensor) -> torch.Tensor:
    if _REAL[_ACT_FMT] == _COMPUTE_DTYPE or x.dtype != _COMPUTE_DTYPE:
        return x
    return ExportAct.apply(x)


# fade-in alpha: a float, optionally mirrored in a 1-element fp32 device tensor the kernels read at run time
class AlphaRef:
    __slots__ = ("value", "dev")

    def __init__(self, value, dev=None):
        self.value, self.dev = float(value), dev

    def __float__(self):
        return float(self.value)


def _coef(c0, c1=0.0, alpha=None) -> L.Coef:
    """pg_coef for  c0 + c1*alpha;  alpha: None, a float or an AlphaRef"""
    if isinstance(c0, tuple):
        c0, c1, alpha = c0
    if alpha is None or c1 == 0.0:
        return L.Coef(float(c0), 0.0, None)
    if isinstance(alpha, AlphaRef) and alpha.dev is not None:
        return L.Coef(float(c0), float(c1), alpha.dev.data_ptr())
    return L.Coef(float(c0) + float(c1) * float(alpha), 0.0, None)


def _affine(c0, c1, alpha):
    """coefficient (c0 + c1*alpha) as the tuple the kernel wrappers accept in place of a float"""
    return (float(c0), float(c1), alpha)


def _dims(x: torch.Tensor):
    if x.dim() == 4:
        b, c, h, w = x.shape
        return b, c, h, w
    if x.dim() == 2:
        return x.shape[0], x.shape[1], 1, 1
    raise ValueError(f"expected a 2-D or 4-D feature tensor, got shape {tuple(x.shape)}")


def _new_like_layout(x: torch.Tensor, channels: int, dtype: torch.dtype, h=None, w=None) -> torch.Tensor:
    b, _, hh, ww = _dims(x)
    h = hh if h is None else h
    w = ww if w is None else w
    if x.dim() == 2:
        return torch.empty((b, channels), dtype=dtype, device=x.device)
    return torch.empty((b, channels, h, w), dtype=dtype, device=x.device, memory_format=torch.channels_last)


# ------------------------------------------------------------------------------ weight packing
# A "layout" says how the logical GEMM operand [N][taps][C] of a layer maps onto its fp32 parameter:
#   None / ("oihw",)        OIHW conv weight or [O,I] linear weight as they are
#   ("pad_c", Cp)           OIHW conv whose input has Cp >= I channels (zero padded: the 513->576 mbstd conv)
#   ("lin_out_perm", C0)    [C0*16, K] linear whose OUTPUT rows are consumed in (h,w,c) order (G's format layer)
#   ("lin_in_perm", C)      [O, C*16] linear whose INPUT columns arrive in (h,w,c) order (D's 8192->512 layer)
_PACK_CACHE: dict = {}


def is_ohwi(W: torch.Tensor) -> bool:
    """a 4-D weight kept in torch.channels_last memory format: physically [O][kh][kw][I] (layers.py stores the 3x3 conv
    weights that way; the logical shape, state_dict values and checkpoints stay OIHW)"""
    return W.dim() == 4 and W.shape[2] * W.shape[3] > 1 and not W.is_contiguous() and \
        W.is_contiguous(memory_format=torch.channels_last)


def _maps(wshape, layout, ohwi=False):
    """-> (N, C, taps, fmap, tmap): logical dims of the fprop GEMM and the pg_pack_map argument tuples
    (n_total, n_valid, n2, taps, c_total, c_valid, c2, base, s_n1, s_n2, s_t, s_c1, s_c2) for the
    fprop operand [N][taps][C] and the dgrad operand [C][taps][N] (taps reversed).
    ohwi: the parameter memory is [O][taps][I] (channels_last) instead of [O][I][taps]."""
    O, I = int(wshape[0]), int(wshape[1])
    T = 1
    for d in wshape[2:]:
        T *= int(d)
    kind = layout[0] if layout else "oihw"
    if kind in ("oihw", "pad_c"):
        Cp = int(layout[1]) if kind == "pad_c" else I
        if ohwi:      # off(o, t, i) = o*T*I + t*I + i
            return O, Cp, T, (O, O, 1, T, Cp, I, 1, 0, I * T, 0, I, 1, 0), \
                (Cp, I, 1, T, O, O, 1, (T - 1) * I, 1, 0, -I, I * T, 0)
        return O, Cp, T, (O, O, 1, T, Cp, I, 1, 0, I * T, 0, 1, T, 0), (Cp, I, 1, T, O, O, 1, T - 1, T, 0, -1, I * T, 0)
    if kind == "lin_out_perm":
        C0 = int(layout[1])
        S = O // C0          # 16 spatial positions; parameter row r = c*S + hw, logical n = hw*C0 + c
        return O, I, 1, (O, O, C0, 1, I, I, 1, 0, I, S * I, 0, 1, 0), (I, I, 1, 1, O, O, C0, 0, 1, 0, 0, I, S * I)
    if kind == "lin_in_perm":
        C = int(layout[1])
        S = I // C           # parameter column j = c*S + hw, logical k = hw*C + c
        return O, I, 1, (O, O, 1, 1, I, I, C, 0, I, 0, 0, 1, S), (I, I, C, 1, O, O, 1, 0, 1, S, 0, I, 0)
    raise ValueError(f"unknown weight layout {layout}")


def _packmap(args):
    m = L.PackMap()
    (m.n_total, m.n_valid, m.n2, m.taps, m.c_total, m.c_valid, m.c2, m.base, m.s_n1, m.s_n2, m.s_t, m.s_c1,
     m.s_c2) = args
    return m


# Packed copies that somebody else keeps in sync with the parameter: optim.FusedAdam writes the fprop operand of a
# channels_last conv weight in the same pass that updates it.  (id(W), kind, fmt, layout) -> [W, tensor, version];
# the entry is valid while W._version equals the version recorded when the copy was last written.
_SHADOWS: dict = {}


def register_shadow(W, kind, fmt, layout, tensor):
    _SHADOWS[(id(W), kind, fmt, layout)] = [W, tensor, None]


def shadows_synced(params):
    """the owner has just rewritten the shadow copies of `params` from their current values"""
    ids = {id(p): p for p in params}
    for key, ent in _SHADOWS.items():
        if key[0] in ids and ent[0] is ids[key[0]]:
            ent[2] = ent[0]._version


def drop_shadows(params=None):
    if params is None:
        _SHADOWS.clear()
        return
    ids = {id(p) for p in params}
    for k in [k for k in _SHADOWS if k[0] in ids]:
        del _SHADOWS[k]


def _pack(W: torch.Tensor, kind: str, dtype=None, layout=None, fmt=None) -> torch.Tensor:
    """fp32 parameter -> K-major packed operand in element format `fmt` (default: the activation format; a dgrad /
    double-backward GEMM asks for the gradient format — both operands of a tensor-core GEMM share one format);
    kind 'f' = fprop [N][taps*C], 't' = dgrad [C][taps*N] with the taps reversed.  Cached per (parameter, version) for
    nn.Parameters.  (`dtype` is accepted for compatibility and ignored: the precision mode decides.)
    A channels_last ([O][kh][kw][I]) conv weight IS its fprop operand: fp32 mode uses the parameter memory itself, the
    16-bit modes the copy the fused Adam kernel maintains (register_shadow), else a cast."""
    _need_cuda(W)
    fmt = _ACT_FMT if fmt is None else fmt
    cacheable = isinstance(W, torch.nn.Parameter)
    if cacheable:
        ent = _SHADOWS.get((id(W), kind, fmt, layout))
        if ent is not None and ent[0] is W and ent[2] == W._version:
            return ent[1]
    key = (id(W), W._version, kind, fmt, layout) if cacheable else None
    if cacheable:
        hit = _PACK_CACHE.get(key)
        if hit is not None and hit[0] is W:
            return hit[1]
    Wc = W.detach()
    ohwi = is_ohwi(Wc) and Wc.dtype == torch.float32
    if not ohwi and (Wc.dtype != torch.float32 or not Wc.is_contiguous()):
        Wc = Wc.float().contiguous()
    N, C, T, fmap, tmap = _maps(Wc.shape, layout, ohwi)
    args = fmap if kind == "f" else tmap
    if ohwi and kind == "f" and fmt == L.PG_F32 and args[4] == args[5]:
        return Wc.permute(0, 2, 3, 1).reshape(args[0], args[3] * args[4])      # a view: no copy at all
    out = torch.empty((args[0], args[3] * args[4]), dtype=torch.float32 if fmt == L.PG_F32 else _COMPUTE_DTYPE,
                      device=W.device)
    m = _packmap(args)
    L.check(L.load().pggan_pack_weight(_p(Wc), _p(out), fmt, ctypes.byref(m), _s()), "pggan_pack_weight")
    if cacheable:
        if len(_PACK_CACHE) > 1024:
            _PACK_CACHE.clear()
        _PACK_CACHE[key] = (W, out)
    return out


def k_convert(x: torch.Tensor, from_fmt: int, to_fmt: int) -> torch.Tensor:
    """the same values in another element format (same torch label for the 16-bit formats)"""
    if from_fmt == to_fmt:
        return x
    out_dtype = torch.float32 if to_fmt == L.PG_F32 else (_COMPUTE_DTYPE if _COMPUTE_DTYPE != torch.float32 else _REAL[to_fmt])
    out = torch.empty_like(x, dtype=out_dtype)
    L.check(L.load().pggan_convert(_p(x), _p(out), x.numel(), from_fmt, to_fmt, _s()), "pggan_convert")
    return out


def clear_pack_cache(params=None):
    """drop packed copies (all, or only those of `params`) — called by the optimizer after it has
    updated the parameters in place through the Adam kernel"""
    if params is None:
        _PACK_CACHE.clear()
        return
    ids = {id(p) for p in params}
    for k in [k for k in _PACK_CACHE if k[0] in ids]:
        del _PACK_CACHE[k]


# ------------------------------------------------------------------------------- mask hand-shake (double backward)
FUSE_DOUBLE_BACKWARD_MASKS = os.environ.get("PGGAN_NO_MASK_FUSE", "0") in ("", "0")   # tests switch it off to compare against the unfused composition


class _MaskLink:
    """Hand-shake between a Function that multiplies its output by a LeakyReLU mask in the FIRST backward (the producer:
    a ConvT with mask_src, UpsampleMask, MaskMul) and the ConvT that consumes that output.  In the double backward the
    roles swap: the producer's backward starts by masking the incoming gradient, and that gradient is the output of the
    consumer's backward conv — so the consumer applies the mask in its GEMM epilogue and leaves its output's address in
    `token`; the producer then skips its own pass (one read + one write of the feature map saved per layer).
    Only offered when the consumer is the producer output's only user (`sole`), so autograd cannot have summed anything
    into the gradient in between; a token that does not match the incoming gradient raises."""
    __slots__ = ("token",)

    def __init__(self):
        self.token = None


def _tag_masked(out, link, y, slope):
    out._pg_mask = (link, y, slope)
    return out


def _premasked(link, gg) -> bool:
    """did the consumer's backward already apply this producer's mask to gg?"""
    if link is None or link.token is None:
        return False
    tok, link.token = link.token, None
    if tok != (gg.data_ptr(), tuple(gg.shape)):
        raise RuntimeError("pggan_b200: a pre-masked gradient was replaced before it reached the masking Function "
                           "(the dgrad's output has more than one user?)")
    return True


# ------------------------------------------------------------------------------- raw kernels
_KIND_CACHE: dict = {}


def conv_pn_fusable(x, N, taps, up2=False) -> bool:
    """may k_conv(..., pn_eps=...) fuse the pixel norm into this layer's epilogue? (one epilogue thread must own all
    N channels of a pixel: the rows kernel, or the plain kernel when one N tile covers the layer)"""
    b, c, h, w = _dims(x)
    if up2:
        h, w = 2 * h, 2 * w
    fmt = _fmt_of(x, A)
    key = ("pn", b, h, w, c, N, taps, fmt, up2)
    hit = _KIND_CACHE.get(key)
    if hit is None:
        p = L.ConvParams()
        p.B, p.H, p.W, p.C, p.N, p.taps = b, h, w, c, N, taps
        p.dtype = p.out_dtype = fmt
        p.up2 = 1 if up2 else 0
        hit = bool(L.load().pggan_conv_pn_fusable(ctypes.byref(p)))
        _KIND_CACHE[key] = hit
    return hit


def conv_up2_fusable(x_low, N, taps=9) -> bool:
    """may conv3x3(upscale2d(x_low)) run as ONE kernel (k_conv(..., up2=True))?  Offered where the plain tensor-core
    kernel would take the unfused layer; the 16-bit modes only (the fp32 island rule looks at the OUTPUT size)."""
    b, c, h, w = _dims(x_low)
    fmt = _fmt_of(x_low, A)
    key = ("up2", b, h, w, c, N, taps, fmt)
    hit = _KIND_CACHE.get(key)
    if hit is None:
        p = L.ConvParams()
        p.B, p.H, p.W, p.C, p.N, p.taps = b, 2 * h, 2 * w, c, N, taps
        p.dtype = p.out_dtype = fmt
        hit = x_low.dim() == 4 and bool(L.load().pggan_conv_up2_fusable(ctypes.byref(p)))
        _KIND_CACHE[key] = hit
    return hit


_SPLITK_WS: dict = {}
SPLITK_WS_BYTES = 4096 + (20 << 20)


def _splitk_ws(device) -> torch.Tensor:
    """the split-K workspace of pg_conv_params.ws_k for the current stream of `device`: one buffer per (device, stream) —
    launches that share it are stream-ordered —, zeroed once (the library leaves it zero)"""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _SPLITK_WS.get(key)
    if ws is None:
        ws = torch.zeros(SPLITK_WS_BYTES, dtype=torch.uint8, device=device)
        _SPLITK_WS[key] = ws
    return ws


def k_conv(x, wp, bias, N, taps, scale, slope=1.0, mask_src=None, mask_slope=1.0, raw_out=False, pn_eps=None, xc=A,
           up2=False):
    """pn_eps: also apply the pixel norm in the epilogue; returns (y, rnorm[pixels] fp32) then.
    xc: class of x (and of y); wp must be packed in the same element format (_pack(..., fmt=_fmt(xc))).
    up2: x is the low-resolution tensor; the convolution reads its nearest-2x up-sampling (never materialised)."""
    _need_cuda(x)
    b, c, h, w = _dims(x)
    if up2:
        h, w = 2 * h, 2 * w
    xf = _fmt_of(x, xc)
    out_dtype = torch.float32 if raw_out else x.dtype
    y = _new_like_layout(x, N, out_dtype, h, w)
    p = L.ConvParams()
    rn = None
    if pn_eps is not None:
        rn = torch.empty((b * h * w,), dtype=torch.float32, device=x.device)
        p.pn_rnorm, p.pn_eps = rn.data_ptr(), float(pn_eps)
    p.x, p.wp, p.bias, p.mask_src, p.y = x.data_ptr(), wp.data_ptr(), (bias.data_ptr() if bias is not None else None), \
        (mask_src.data_ptr() if mask_src is not None else None), y.data_ptr()
    p.B, p.H, p.W, p.C, p.N, p.taps = b, h, w, c, N, taps
    p.dtype = xf
    p.out_dtype = L.PG_F32_RAW if raw_out else xf
    p.scale, p.slope, p.mask_slope = float(scale), float(slope), float(mask_slope)
    p.up2 = 1 if up2 else 0
    wk = _splitk_ws(x.device)
    p.ws_k, p.ws_k_bytes = wk.data_ptr(), wk.numel()
    ws = None
    if xf == L.PG_F32:          # fp32 mode = 3xTF32: the library splits operands into these
        ws = (torch.empty(3 * x.numel(), dtype=torch.float32, device=x.device),
              torch.empty(3 * wp.numel(), dtype=torch.float32, device=x.device))
        p.ws_x, p.ws_w = ws[0].data_ptr(), ws[1].data_ptr()
    with _Timed("conv_gemm", (b, h, w, c, N, taps)):
        L.check(L.load().pggan_conv_fprop(ctypes.byref(p), _s()), "pggan_conv_fprop")
    return y if pn_eps is None else (y, rn)


def dgrad_mask_fusable(batch, h, w, fwd_in_channels, fwd_out_channels, taps=9, dtype=None, f32=False) -> bool:
    """True if the dgrad of a (fwd_in_channels -> fwd_out_channels) conv at this size runs on a kernel whose epilogue
    applies a LeakyReLU-backward mask for free (see pggan_conv_kernel_kind); the blocks use it to decide whether a
    producer may defer its mask to this layer's dgrad."""
    gfmt = L.PG_F32 if f32 else _GRAD_FMT
    key = (batch, h, w, fwd_in_channels, fwd_out_channels, taps, gfmt)
    hit = _KIND_CACHE.get(key)
    if hit is None:
        p = L.ConvParams()
        p.B, p.H, p.W, p.C, p.N, p.taps = batch, h, w, fwd_out_channels, fwd_in_channels, taps     # dgrad: roles swapped
        p.dtype = gfmt
        p.out_dtype = gfmt
        hit = L.load().pggan_conv_kernel_kind(ctypes.byref(p)) in (L.PG_CONV_KIND_PLAIN, L.PG_CONV_KIND_ROWS)
        _KIND_CACHE[key] = hit
    return hit


DIRECT_WGRAD = os.environ.get("PGGAN_DETERMINISTIC", "0") in ("", "0")     # see set_deterministic()


def set_deterministic(flag: bool = True) -> None:
    """True: keep the summation order of every weight gradient fixed — the wgrad GEMMs write per-split slices that an
    un-pack pass adds in order (+0.4 ms per C4 step) instead of adding their partial sums to .grad with fp32 reductions.
    Conv / split-K results are order-independent in both modes; what remains unordered in either mode are the fp32
    atomics of the bias and RGB-layer gradients (column sums), so two runs still differ in the last bit of those.
    Also settable with PGGAN_DETERMINISTIC=1."""
    global DIRECT_WGRAD
    DIRECT_WGRAD = not flag


def _dense_map(fmap) -> bool:
    """does the logical operand [N][taps][C] coincide with the parameter memory (no permutation, no padding)?"""
    n_total, n_valid, n2, taps, c_total, c_valid, c2, base, s_n1, s_n2, s_t, s_c1, s_c2 = fmap
    return n2 == 1 and c2 == 1 and base == 0 and s_c1 == 1 and (taps == 1 or s_t == c_total) and \
        s_n1 == taps * c_total and n_valid == n_total and c_valid == c_total


def k_wgrad(g, x, scale, wshape, layout=None, accumulate_into=None, xc=A):
    """weight gradient in the parameter's own fp32 layout: a new tensor, or added into `accumulate_into`.
    g is a gradient-class tensor, x of class `xc` (converted to the gradient format when that differs)."""
    _need_cuda(g)
    b, n, h, w = _dims(g)
    c = _dims(x)[1]
    ohwi = accumulate_into is not None and is_ohwi(accumulate_into)       # the gradient view has the parameter's strides
    N, C, taps, fmap, _ = _maps(wshape, layout, ohwi)
    if (n, c) != (N, C):
        raise ValueError(f"wgrad operand channels ({n},{c}) do not match the layer ({N},{C})")
    p = L.WgradParams()
    p.g, p.x = g.data_ptr(), x.data_ptr()
    p.B, p.H, p.W, p.C, p.N, p.taps = b, h, w, c, n, taps
    gfmt = _fmt_of(g, G)
    p.dtype = gfmt
    if _fmt_of(x, xc) != gfmt:
        # both GEMM operands must share one element format: the saved fp16 activation is re-rounded to bf16 (mixed mode)
        x = k_convert(x, _fmt_of(x, xc), gfmt)
    p.x = x.data_ptr()
    ws = None
    if gfmt == L.PG_F32:
        ws = (torch.empty(3 * g.numel(), dtype=torch.float32, device=g.device),
              torch.empty(3 * x.numel(), dtype=torch.float32, device=g.device))
        p.ws_g, p.ws_x = ws[0].data_ptr(), ws[1].data_ptr()
    if accumulate_into is not None and DIRECT_WGRAD and _dense_map(fmap):
        p.direct, p.direct_scale = accumulate_into.data_ptr(), float(scale)
        if not L.load().pggan_conv_wgrad_direct_ok(ctypes.byref(p)):
            p.direct, p.direct_scale = None, 0.0
    if p.direct:
        # the gradient has the operand's own layout (channels_last conv weight, plain linear weight): the GEMM's epilogue
        # adds its scaled partial sums straight into .grad — no slice buffer, no un-pack pass
        with _Timed("wgrad_gemm", (b, h, w, c, n, taps)):
            L.check(L.load().pggan_conv_wgrad(ctypes.byref(p), _s()), "pggan_conv_wgrad")
        return None
    slices = L.load().pggan_conv_wgrad_slices(ctypes.byref(p))
    acc = torch.empty((slices, n, taps, c), dtype=torch.float32, device=g.device)
    p.acc = acc.data_ptr()
    with _Timed("wgrad_gemm", (b, h, w, c, n, taps)):
        L.check(L.load().pggan_conv_wgrad(ctypes.byref(p), _s()), "pggan_conv_wgrad")
    m = _packmap(fmap)
    if accumulate_into is not None:
        L.check(L.load().pggan_unpack_wgrad(_p(acc), _p(accumulate_into), ctypes.byref(m), slices, float(scale), 1,
                                            _s()), "pggan_unpack_wgrad")
        return None
    pad = fmap[1] < fmap[0] or fmap[5] < fmap[4]
    out = (torch.zeros if pad else torch.empty)(tuple(wshape), dtype=torch.float32, device=g.device)
    L.check(L.load().pggan_unpack_wgrad(_p(acc), _p(out), ctypes.byref(m), slices, float(scale), 0, _s()),
            "pggan_unpack_wgrad")
    return out


def _vn(fmt: int) -> int:
    return 4 if fmt == L.PG_F32 else 8


def k_mask_colsum(g, y, slope, want_out, colsum_scale, colsum_into=None, cls=G):
    """fused LeakyReLU backward (+ bias gradient): returns (g*mask(y) or None, colsum or None).
    colsum_into: an fp32 [C] tensor the column sums are ADDED to (then the second result is None)."""
    b, c, h, w = _dims(g)
    vn = _vn(_fmt_of(g, cls))
    if c % vn != 0 or (c // vn) > 256 or 256 % (c // vn) != 0:          # odd widths: two plain passes
        out = k_mask_mul(g, y, slope, cls) if (want_out and y is not None) else (g if want_out else None)
        src = out if out is not None else (k_mask_mul(g, y, slope, cls) if y is not None else g)
        cs = k_colsum(src, colsum_scale, cls) if colsum_scale is not None else None
        if cs is not None and colsum_into is not None:
            colsum_into.add_(cs)
            cs = None
        return out, cs
    out = torch.empty_like(g) if (want_out and y is not None) else None
    cs = None
    if colsum_scale is not None:
        cs = colsum_into if colsum_into is not None else torch.zeros((c,), dtype=torch.float32, device=g.device)
    if out is None and cs is None:
        return (g if want_out else None), None
    L.check(L.load().pggan_mask_colsum(_p(g), _p(y), _p(out), _p(cs), b * h * w, c, float(slope),
                                       float(colsum_scale or 0.0), _fmt_of(g, cls), _s()), "pggan_mask_colsum")
    if want_out and out is None:
        out = g
    return out, (None if colsum_into is not None else cs)


def k_mask_mul(g, y, slope, cls=G):
    out = torch.empty_like(g)
    L.check(L.load().pggan_lrelu_mask_mul(_p(g), _p(y), _p(out), g.numel(), float(slope), _fmt_of(g, cls), _s()),
            "pggan_lrelu_mask_mul")
    return out


def k_colsum(g, scale, cls=G):
    b, c, h, w = _dims(g)
    if c % _vn(_fmt_of(g, cls)) != 0:      # e.g. the 1-logit decision layer: a [B,1] column
        gf = to_float(g, cls)
        g2 = gf.reshape(b, c) if g.dim() == 2 else gf.permute(0, 2, 3, 1).reshape(-1, c)
        return g2.sum(0) * scale
    out = torch.zeros((c,), dtype=torch.float32, device=g.device)
    L.check(L.load().pggan_colsum(_p(g), _p(out), b * h * w, c, float(scale), _fmt_of(g, cls), _s()), "pggan_colsum")
    return out


def k_upsample(x, scale, mask_src=None, slope=1.0, cls=A, colsum=None, colsum_scale=1.0):
    """scale: a float or an _affine() coefficient.  colsum (with mask_src): an fp32 [C] tensor that receives
    += colsum_scale * (column sums of the result) in the same pass."""
    b, c, h, w = _dims(x)
    out = _new_like_layout(x, c, x.dtype, 2 * h, 2 * w)
    if mask_src is not None and colsum is not None:
        L.check(L.load().pggan_upsample2x_mask_colsum(_p(x), _p(mask_src), _p(out), _p(colsum), b, h, w, c, _coef(scale),
                                                      float(slope), float(colsum_scale), _fmt_of(x, cls), _s()),
                "pggan_upsample2x_mask_colsum")
        return out
    if mask_src is not None:
        L.check(L.load().pggan_upsample2x_mask(_p(x), _p(mask_src), _p(out), b, h, w, c, _coef(scale), float(slope),
                                               _fmt_of(x, cls), _s()), "pggan_upsample2x_mask")
        return out
    L.check(L.load().pggan_upsample2x(_p(x), _p(out), b, h, w, c, _coef(scale), _fmt_of(x, cls), _s()),
            "pggan_upsample2x")
    return out


def upsample_colsum_fusable(x, cls=G) -> bool:
    if os.environ.get("PGGAN_NO_POOL_COLSUM", "0") not in ("", "0"):
        return False
    c = _dims(x)[1]
    vn = _vn(_fmt_of(x, cls))
    return x.dim() == 4 and c % vn == 0 and (c // vn) <= 256 and 256 % (c // vn) == 0


def k_pool(x, scale, blend_with=None, alpha=1.0, cls=A):
    """scale: a float or an _affine() coefficient; alpha: a float or an AlphaRef"""
    b, c, h, w = _dims(x)
    out = _new_like_layout(x, c, x.dtype, h // 2, w // 2)
    L.check(L.load().pggan_pool2x2(_p(x), _p(blend_with), _p(out), b, h, w, c, _coef(scale), _coef(0.0, 1.0, alpha),
                                   _fmt_of(x, cls), _s()), "pggan_pool2x2")
    return out


def k_axpby(a, b, sa, sb, cls=A):
    """sa / sb: floats or _affine() coefficients; fp32 tensors run as fp32 whatever the class"""
    out = torch.empty_like(a)
    fmt = _fmt_of(a, cls)
    L.check(L.load().pggan_axpby(_p(a), _p(b), _p(out), a.numel(), _coef(sa), _coef(sb), fmt, _s()), "pggan_axpby")
    return out


# ----------------------------------------------------------------------- GEMM-class Functions
def _perm_bias(b, layout, inverse=False):
    """bias of a ("lin_out_perm", C0) layer in logical (h,w,c) order, or a logical gradient back"""
    if b is None or not layout or layout[0] != "lin_out_perm":
        return b
    C0 = int(layout[1])
    S = b.numel() // C0
    return b.view(S, C0).t().reshape(-1) if inverse else b.view(C0, S).t().reshape(-1)


class ConvF(torch.autograd.Function):
    """y = lrelu(scale * (conv(x, W) + b), slope)   [fprop; slope = 1 -> linear]

    LeakyReLU-backward fusion (the mask is sign(stored activation), so whoever holds the activation can apply it):
      x_act_slope : x is the LeakyReLU output of a producer that was run with defer_mask=True; this layer's dgrad
                    epilogue then applies that producer's mask to gx (no separate leaky_relu_backward pass).
      defer_mask  : the consumer of y applies THIS layer's mask (a ConvF with x_act_slope, or a pooling Function
                    with act_slope): the incoming gradient is already masked and y need not be saved.
      xc          : tensor class of x (and of y): "a" in a forward pass, "g" when the double backward runs an fprop
                    on a gradient-like tensor (ConvT.backward)."""

    @staticmethod
    def forward(ctx, x, W, b, scale, slope, raw_out, layout=None, x_act_slope=None, defer_mask=False, xc=A):
        x = feat(x, xc)
        N, C, taps, _, _ = _maps(W.shape, layout)
        bias = None if b is None else _perm_bias(b.detach().float(), layout).contiguous()
        y = k_conv(x, _pack(W, "f", None, layout, _fmt_of(x, xc)), bias, N, taps, scale, slope, raw_out=raw_out, xc=xc)
        ctx.mask_here = slope != 1.0 and not defer_mask
        ctx.save_for_backward(x, W, y if ctx.mask_here else None)
        ctx.scale, ctx.slope, ctx.has_bias, ctx.layout = scale, slope, b is not None, layout
        ctx.x_act_slope = x_act_slope
        ctx.xc = xc
        ctx.w_param = W if isinstance(W, torch.nn.Parameter) else None
        ctx.bias_param = b if isinstance(b, torch.nn.Parameter) else None
        if defer_mask and ctx.bias_param is not None and not (layout and layout[0] == "lin_out_perm"):
            # whoever applies this layer's mask in the backward (the pool's adjoint) may take the bias gradient along
            y._pg_bias = (ctx.bias_param, scale)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, W, y = ctx.saved_tensors
        g = feat(gy, G, small=(x.dtype == torch.float32))      # (the fp32 logits' gradient of a 16-bit layer: converted)
        need_w = ctx.needs_input_grad[1] and not _SKIP_PARAM_GRADS
        need_b = ctx.has_bias and ctx.needs_input_grad[2] and not _SKIP_PARAM_GRADS
        gx = gW = gb = None
        xm, xs = (x, ctx.x_act_slope) if ctx.x_act_slope is not None else (None, 1.0)
        if ctx.w_param is not None:
            W = ctx.w_param              # the Parameter object itself (gradient sink / packed-copy cache are keyed by it)
        if torch.is_grad_enabled():
            # create_graph=True: stay composable (every piece is a differentiable Function)
            if ctx.mask_here:
                link = _MaskLink()
                g = _tag_masked(MaskMul.apply(g, y, ctx.slope, link), link, y, ctx.slope)
            if need_b:
                gb = ColSum.apply(g, ctx.scale)
        else:
            # plain backward: LeakyReLU mask and bias gradient in one pass over g; gradients of registered
            # parameters are accumulated in place (see _GRAD_SINK)
            b_sink = _sink(ctx.bias_param) if (need_b and not (ctx.layout and ctx.layout[0] == "lin_out_perm")) else None
            pre = getattr(gy, "_pg_colsum", None) if need_b else None
            if pre is not None and not isinstance(pre, str) and (ctx.mask_here or g is not gy):
                pre = None                  # (a buffer of column sums of a tensor that is not what this layer sums)
            if isinstance(pre, str):
                # "sunk": the pool backward that produced gy added this layer's bias gradient to .grad itself (and told the
                # sink listeners) — never add it a second time
                b_sink = None
                if ctx.mask_here:
                    g, _ = k_mask_colsum(g, y, ctx.slope, True, None)
            elif pre is not None:
                # ... or summed the columns into a buffer of its own (no gradient sink in this context)
                if b_sink is not None:
                    b_sink.add_(pre, alpha=ctx.scale)
                else:
                    gb = pre * ctx.scale
            else:
                g, gb = k_mask_colsum(g, y if ctx.mask_here else None, ctx.slope, True,
                                      ctx.scale if need_b else None, colsum_into=b_sink)
            if b_sink is not None:
                _sunk(ctx.bias_param)
            if gb is not None:
                gb = _perm_bias(gb, ctx.layout, inverse=True)
            if ctx.needs_input_grad[0]:
                gx = ConvT.apply(g, W, ctx.scale, ctx.layout, xm, xs)
            if need_w:
                w_sink = _sink(ctx.w_param)
                gW = k_wgrad(g, x, ctx.scale, tuple(W.shape), ctx.layout, accumulate_into=w_sink, xc=ctx.xc)
                if w_sink is not None:
                    _sunk(ctx.w_param)
            return gx, gW, gb, None, None, None, None, None, None, None
        if gb is not None:
            gb = _perm_bias(gb, ctx.layout, inverse=True)
        if ctx.needs_input_grad[0]:
            link = _MaskLink() if xm is not None else None
            gx = ConvT.apply(g, W, ctx.scale, ctx.layout, xm, xs, link, not (need_w or need_b))
            if link is not None:
                _tag_masked(gx, link, xm, xs)
        if need_w:
            gW = ConvW.apply(g, x, ctx.scale, tuple(W.shape), ctx.layout, ctx.xc)
        return gx, gW, gb, None, None, None, None, None, None, None


class ConvActPN(torch.autograd.Function):
    """G's  [upscale2d ->] conv3x3 -> c*(.+b) -> LeakyReLU -> pixel norm  (reference lib/blocks.py:119-151 with
    lib/layers.py:13-14 and lib/utils.py:106-118).  Forward: ONE kernel where the epilogue thread owns the pixel (GEMM
    with bias/He/LReLU/pixel-norm epilogue), two otherwise; with up2 the nearest-2x up-sampling is the GEMM's PROLOGUE
    (parity tiles read shifted boxes of the low-resolution tensor: the up-sampled tensor is never written).  Backward:
    first pass fuses pixel-norm backward, the LeakyReLU mask (sign(y_pn) == sign(pre)) and the bias gradient; the
    up-sampled input is re-materialised only for the weight gradient, its gradient is pooled back.
    The generator is never differentiated twice."""

    @staticmethod
    def forward(ctx, x, W, b, scale, slope, eps, layout=None, up2=False):
        if up2:
            x = feat(x, A)
            if not conv_up2_fusable(x, int(W.shape[0])):       # the fused prologue where the plain kernel takes the layer
                x = k_upsample(x, 1.0)
                up2 = False
                ctx.up_input = True
            else:
                ctx.up_input = False
        else:
            x = feat(x, A)
            ctx.up_input = False
        ctx.up2 = up2
        N, C, taps, _, _ = _maps(W.shape, layout)
        bias = None if b is None else _perm_bias(b.detach().float(), layout).contiguous()
        if x.dim() == 4 and conv_pn_fusable(x, N, taps, up2):
            # conv + He scale + bias + LeakyReLU + pixel norm in ONE kernel (the epilogue thread owns the whole pixel)
            y, r = k_conv(x, _pack(W, "f", None, layout, _fmt_of(x, A)), bias, N, taps, scale, slope, pn_eps=eps, up2=up2)
        else:
            a = k_conv(x, _pack(W, "f", None, layout, _fmt_of(x, A)), bias, N, taps, scale, slope, up2=up2)
            bb, c, h, w = _dims(a)
            y = torch.empty_like(a)
            r = torch.empty((bb * h * w,), dtype=torch.float32, device=a.device)
            L.check(L.load().pggan_pixelnorm_fwd(_p(a), _p(y), _p(r), bb * h * w, c, float(eps), _fmt_of(a, A), _s()),
                    "pggan_pixelnorm_fwd")
        ctx.save_for_backward(x, W, y, r)
        ctx.scale, ctx.slope, ctx.has_bias, ctx.layout = scale, slope, b is not None, layout
        ctx.w_param = W if isinstance(W, torch.nn.Parameter) else None
        ctx.bias_param = b if isinstance(b, torch.nn.Parameter) else None
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, W, y, r = ctx.saved_tensors
        gy = feat(gy, G, small=(y.dtype == torch.float32))
        bb, c, h, w = _dims(y)
        need_b = ctx.has_bias and ctx.needs_input_grad[2] and not _SKIP_PARAM_GRADS
        g = torch.empty_like(gy)
        b_sink = _sink(ctx.bias_param) if (need_b and not (ctx.layout and ctx.layout[0] == "lin_out_perm")) else None
        gb = None
        if need_b:
            gb = b_sink if b_sink is not None else torch.zeros((c,), dtype=torch.float32, device=y.device)
        L.check(L.load().pggan_pixelnorm_bwd(_p(gy), _p(y), _p(r), _p(g), bb * h * w, c, float(ctx.slope), _p(gb),
                                             float(ctx.scale), _fmt_of(gy, G), _fmt_of(y, A), _s()), "pggan_pixelnorm_bwd")
        if b_sink is not None:
            _sunk(ctx.bias_param)
        gx = gW = None
        upsampled = ctx.up2 or ctx.up_input           # the conv's input was upscale2d(x_low)
        if ctx.needs_input_grad[0]:
            N, C, taps, _, _ = _maps(W.shape, ctx.layout)
            gx = k_conv(g, _pack(W, "t", None, ctx.layout, _fmt_of(g, G)), None, C, taps, ctx.scale, xc=G)
            if upsampled:
                gx = k_pool(gx, 1.0, cls=G)           # adjoint of the nearest up-sampling: 2x2 window sum
        if ctx.needs_input_grad[1] and not _SKIP_PARAM_GRADS:
            w_sink = _sink(ctx.w_param)
            xw = k_upsample(x, 1.0) if ctx.up2 else x          # fused forward: re-materialise up2(x) for this GEMM only
            gW = k_wgrad(g, xw, ctx.scale, tuple(W.shape), ctx.layout, accumulate_into=w_sink, xc=A)
            if w_sink is not None:
                _sunk(ctx.w_param)
        if b_sink is not None:
            gb = None
        elif gb is not None:
            gb = _perm_bias(gb, ctx.layout, inverse=True)
        return gx, gW, gb, None, None, None, None, None


class ConvT(torch.autograd.Function):
    """gx = scale * conv_transpose(g, W)   [dgrad: the same kernel on flipped/transposed weights]"""

    @staticmethod
    def forward(ctx, g, W, scale, layout=None, mask_src=None, mask_slope=1.0, link=None, sole=False):
        """mask_src: gx is additionally multiplied by (mask_src > 0 ? 1 : mask_slope) in the GEMM epilogue
        (the LeakyReLU backward of the layer that produced this conv's input); link: this Function's _MaskLink then.
        sole: this Function is the only user of g — if g's producer masked it (tag left by _tag_masked), the double
        backward applies that mask in the epilogue of the conv below instead of in a pass of its own."""
        ctx.out_mask = getattr(g, "_pg_mask", None) if (sole and FUSE_DOUBLE_BACKWARD_MASKS) else None
        g = feat(g, G)
        if ctx.out_mask is not None and ctx.out_mask[1].element_size() != g.element_size():
            ctx.out_mask = None
        N, C, taps, _, _ = _maps(W.shape, layout)
        gx = k_conv(g, _pack(W, "t", None, layout, _fmt_of(g, G)), None, C, taps, scale, mask_src=mask_src,
                    mask_slope=mask_slope, xc=G)
        ctx.save_for_backward(g, W, mask_src)
        ctx.scale, ctx.layout, ctx.mask_slope, ctx.link = scale, layout, mask_slope, link
        ctx.w_param = W if isinstance(W, torch.nn.Parameter) else None
        return gx

    @staticmethod
    def backward(ctx, ggx):
        g, W, mask_src = ctx.saved_tensors
        pre = mask_src is not None and _premasked(ctx.link, ggx)
        ggx = feat(ggx, G, small=(g.dtype == torch.float32))
        if mask_src is not None and not pre:
            ggx = MaskMul.apply(ggx, mask_src, ctx.mask_slope)
        dg = dW = None
        if ctx.w_param is not None:
            W = ctx.w_param
        if ctx.needs_input_grad[0]:
            if ctx.out_mask is not None and not torch.is_grad_enabled() and ggx.dtype == g.dtype:
                link, ym, sl = ctx.out_mask
                N, C, taps, _, _ = _maps(W.shape, ctx.layout)
                dg = k_conv(ggx, _pack(W, "f", None, ctx.layout, _fmt_of(ggx, G)), None, N, taps, ctx.scale, mask_src=ym,
                            mask_slope=sl, xc=G)
                link.token = (dg.data_ptr(), tuple(dg.shape))
            else:
                dg = ConvF.apply(ggx, W, None, ctx.scale, 1.0, False, ctx.layout, None, False, G)
        if ctx.needs_input_grad[1] and not _SKIP_PARAM_GRADS:
            w_sink = _sink(ctx.w_param) if not torch.is_grad_enabled() else None
            if w_sink is not None:
                # the penalty's own backward (plain, no higher order): accumulate d/dW of the dgrad straight into .grad
                k_wgrad(g, ggx, ctx.scale, tuple(W.shape), ctx.layout, accumulate_into=w_sink, xc=G)
                _sunk(ctx.w_param)
            else:
                dW = ConvW.apply(g, ggx, ctx.scale, tuple(W.shape), ctx.layout, G)
        return dg, dW, None, None, None, None, None, None


class ConvW(torch.autograd.Function):
    """gW = scale * wgrad(g, x)   in the parameter's layout; g gradient-class, x of class xc"""

    @staticmethod
    def forward(ctx, g, x, scale, wshape, layout=None, xc=A):
        g, x = feat(g, G), feat(x, xc)
        ctx.save_for_backward(g, x)
        ctx.scale, ctx.layout, ctx.xc = scale, layout, xc
        return k_wgrad(g, x, scale, wshape, layout, xc=xc)

    @staticmethod
    def backward(ctx, ggW):
        g, x = ctx.saved_tensors
        dg = dx = None
        ggW = ggW.contiguous()
        if ctx.needs_input_grad[0]:
            dg = ConvF.apply(x, ggW, None, ctx.scale, 1.0, False, ctx.layout, None, False, ctx.xc)
            if ctx.xc != G:
                # conv(x, ggW) came out in x's class; it is the gradient of g
                dg = feat(to_float(dg, ctx.xc), G)
        if ctx.needs_input_grad[1]:
            dx = ConvT.apply(g, ggW, ctx.scale, ctx.layout)
        return dg, dx, None, None, None, None


class ColSum(torch.autograd.Function):
    """gb[c] = scale * sum over pixels of g[:, c]"""

    @staticmethod
    def forward(ctx, g, scale):
        g = feat(g, G)
        ctx.shape, ctx.scale, ctx.dtype = tuple(g.shape), scale, g.dtype
        return k_colsum(g, scale, G)

    @staticmethod
    def backward(ctx, ggb):
        shape = ctx.shape
        view = (1, -1, 1, 1) if len(shape) == 4 else (1, -1)
        return cl((ggb * ctx.scale).to(ctx.dtype).view(view).expand(shape)), None


class MaskMul(torch.autograd.Function):
    """g * (y > 0 ? 1 : slope): LeakyReLU backward from the saved OUTPUT (sign(y) == sign(pre)).  link: see _MaskLink."""

    @staticmethod
    def forward(ctx, g, y, slope, link=None):
        g = feat(g, G, small=(y.dtype == torch.float32))
        ctx.save_for_backward(y)
        ctx.slope, ctx.link = slope, link
        return k_mask_mul(g, y, slope, G)

    @staticmethod
    def backward(ctx, gg):
        (y,) = ctx.saved_tensors
        if _premasked(ctx.link, gg):
            return gg, None, None, None
        return MaskMul.apply(feat(gg, G, small=(y.dtype == torch.float32)), y, ctx.slope), None, None, None


# ------------------------------------------------------------------ resampling / blend Functions
class Upsample2x(torch.autograd.Function):
    """nearest 2x (reference lib/utils.py:106-118), times `scale` (a float or an _affine() coefficient)"""

    @staticmethod
    def forward(ctx, x, scale, cls=A):
        ctx.scale = scale
        x = feat(x, cls)
        ctx.f32 = x.dtype == torch.float32
        return k_upsample(x, scale, cls=cls)

    @staticmethod
    def backward(ctx, g):
        return Pool2x2.apply(feat(g, G, small=ctx.f32), ctx.scale, None, G), None, None


class UpsampleMask(torch.autograd.Function):
    """scale * up2(g) * (y > 0 ? 1 : slope): avg-pool backward fused with the LeakyReLU backward of the pooled
    activation y (one pass instead of two).  link: see _MaskLink (use _upsample_mask())."""

    @staticmethod
    def forward(ctx, g, y, scale, slope, link=None):
        ctx.save_for_backward(y)
        ctx.scale, ctx.slope, ctx.link = scale, slope, link
        return k_upsample(feat(g, G, small=(y.dtype == torch.float32)), scale, mask_src=y, slope=slope, cls=G)

    @staticmethod
    def backward(ctx, gg):
        (y,) = ctx.saved_tensors
        f32 = y.dtype == torch.float32
        if not _premasked(ctx.link, gg):
            gg = MaskMul.apply(feat(gg, G, small=f32), y, ctx.slope)
        return Pool2x2.apply(gg, ctx.scale, None, G), None, None, None, None


def _upsample_mask(g, y, scale, slope, bias_info=None):
    """bias_info: (bias parameter, He scale) of the conv that produced y (tag left by ConvF.forward), or None"""
    if not torch.is_grad_enabled():
        g = feat(g, G, small=(y.dtype == torch.float32))
        if _SKIP_PARAM_GRADS or bias_info is None or not bias_info[0].requires_grad or not upsample_colsum_fusable(g):
            return k_upsample(g, scale, mask_src=y, slope=slope, cls=G)
        # plain backward: y is the output of a conv whose bias gradient is the column sum of this result — taken in the
        # same pass: straight into .grad inside a gradient-sink context, else into a buffer handed to that conv's backward
        # through the tensor (ConvF.backward picks up `_pg_colsum`)
        bparam, bscale = bias_info
        sink = _sink(bparam)
        if sink is not None:
            out = k_upsample(g, scale, mask_src=y, slope=slope, cls=G, colsum=sink, colsum_scale=bscale)
            _sunk(bparam)
            out._pg_colsum = "sunk"
            return out
        cs = torch.zeros((g.shape[1],), dtype=torch.float32, device=g.device)
        out = k_upsample(g, scale, mask_src=y, slope=slope, cls=G, colsum=cs)
        out._pg_colsum = cs
        return out
    link = _MaskLink()
    return _tag_masked(UpsampleMask.apply(g, y, scale, slope, link), link, y, slope)


class Pool2x2(torch.autograd.Function):
    """scale * (2x2 window sum); scale = 0.25 is nn.AvgPool2d(2) (reference lib/blocks.py:193); scale may be an
    _affine() coefficient.  act_slope: x is the LeakyReLU output of a ConvF run with defer_mask=True — its mask is
    applied here, in backward."""

    @staticmethod
    def forward(ctx, x, scale, act_slope=None, cls=A):
        ctx.bias_info = getattr(x, "_pg_bias", None) if act_slope is not None else None
        x = feat(x, cls)
        ctx.scale, ctx.act_slope, ctx.f32 = scale, act_slope, x.dtype == torch.float32
        if act_slope is not None:
            ctx.save_for_backward(x)
        return k_pool(x, scale, cls=cls)

    @staticmethod
    def backward(ctx, g):
        g = feat(g, G, small=ctx.f32)
        if ctx.act_slope is not None:
            (x,) = ctx.saved_tensors
            return _upsample_mask(g, x, ctx.scale, ctx.act_slope, ctx.bias_info), None, None, None
        return Upsample2x.apply(g, ctx.scale, G), None, None, None


class PoolBlend(torch.autograd.Function):
    """(1-alpha) * low + alpha * avgpool2(x): D's feature-space fade fused with the pooling
    (reference lib/blocks.py:193 + pggan/nets.py:262-265).  act_slope as in Pool2x2; alpha a float or an AlphaRef."""

    @staticmethod
    def forward(ctx, x, low, alpha, act_slope=None):
        ctx.bias_info = getattr(x, "_pg_bias", None) if act_slope is not None else None
        x = feat(x, A)
        ctx.alpha, ctx.act_slope, ctx.f32 = alpha, act_slope, x.dtype == torch.float32
        if act_slope is not None:
            ctx.save_for_backward(x)
        return k_pool(x, 0.25, blend_with=feat(low, A, small=ctx.f32), alpha=alpha)

    @staticmethod
    def backward(ctx, g):
        g = feat(g, G, small=ctx.f32)
        gx = None
        if ctx.needs_input_grad[0]:
            if ctx.act_slope is not None:
                (x,) = ctx.saved_tensors
                gx = _upsample_mask(g, x, _affine(0.0, 0.25, ctx.alpha), ctx.act_slope, ctx.bias_info)
            else:
                gx = Upsample2x.apply(g, _affine(0.0, 0.25, ctx.alpha), G)
        gl = Scale.apply(g, _affine(1.0, -1.0, ctx.alpha), G) if ctx.needs_input_grad[1] else None
        return gx, gl, None, None


class Scale(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, s, cls=A):
        ctx.s, ctx.cls = s, cls
        # a tensor that is already in a feature dtype keeps its precision family
        x = feat(x, cls, small=(x.dtype == torch.float32) if x.dtype in (torch.float32, _COMPUTE_DTYPE) else None)
        ctx.f32 = x.dtype == torch.float32
        return k_axpby(x, None, s, 0.0, cls)

    @staticmethod
    def backward(ctx, g):
        return Scale.apply(feat(g, G, small=ctx.f32), ctx.s, G), None, None


class Lerp(torch.autograd.Function):
    """(1-alpha) * a + alpha * b on feature maps"""

    @staticmethod
    def forward(ctx, a, b, alpha, cls=A):
        ctx.alpha = alpha
        return k_axpby(feat(a, cls), feat(b, cls), _affine(1.0, -1.0, alpha), _affine(0.0, 1.0, alpha), cls)

    @staticmethod
    def backward(ctx, g):
        g = feat(g, G)
        ga = Scale.apply(g, _affine(1.0, -1.0, ctx.alpha), G) if ctx.needs_input_grad[0] else None
        gb = Scale.apply(g, _affine(0.0, 1.0, ctx.alpha), G) if ctx.needs_input_grad[1] else None
        return ga, gb, None, None


# ---------------------------------------------------------------------------------- pixel norm
class PixelNorm(torch.autograd.Function):
    """x * rsqrt(mean_c(x^2) + eps)  (reference lib/layers.py:13-14)."""

    @staticmethod
    def forward(ctx, x, eps):
        x = feat(x, A)
        b, c, h, w = _dims(x)
        y = torch.empty_like(x)
        r = torch.empty((b * h * w,), dtype=torch.float32, device=x.device)
        L.check(L.load().pggan_pixelnorm_fwd(_p(x), _p(y), _p(r), b * h * w, c, float(eps), _fmt_of(x, A), _s()),
                "pggan_pixelnorm_fwd")
        ctx.save_for_backward(y, r)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        y, r = ctx.saved_tensors
        gy = feat(gy, G, small=(y.dtype == torch.float32))
        b, c, h, w = _dims(y)
        gx = torch.empty_like(gy)
        L.check(L.load().pggan_pixelnorm_bwd(_p(gy), _p(y), _p(r), _p(gx), b * h * w, c, 1.0, None, 0.0,
                                             _fmt_of(gy, G), _fmt_of(y, A), _s()), "pggan_pixelnorm_bwd")
        return gx, None


# ------------------------------------------------------------------------------------ RGB 1x1
def _k_rgb_expand(img, wc, bias, C, pool, scale, slope, cls, f32=False):
    b, cimg, hi, wi = img.shape
    h, w = (hi // 2, wi // 2) if pool else (hi, wi)
    f32 = f32 or _COMPUTE_DTYPE == torch.float32
    out = torch.empty((b, C, h, w), dtype=torch.float32 if f32 else _COMPUTE_DTYPE, device=img.device,
                      memory_format=torch.channels_last)
    L.check(L.load().pggan_rgb_expand(_p(img), _p(wc), _p(bias), _p(out), b, h, w, C, cimg, int(pool), float(scale),
                                      float(slope), _fmt_of(out, cls), _s()), "pggan_rgb_expand")
    return out


def _k_rgb_reduce(x, wr, bias, cimg, up, scale, up_scale, cls):
    b, c, h, w = _dims(x)
    ho, wo = (2 * h, 2 * w) if up else (h, w)
    out = torch.empty((b, cimg, ho, wo), dtype=torch.float32, device=x.device)
    L.check(L.load().pggan_rgb_reduce(_p(x), _p(wr), _p(bias), _p(out), b, h, w, c, cimg, int(up), float(scale),
                                      float(up_scale), _fmt_of(x, cls), _s()), "pggan_rgb_reduce")
    return out


def _k_rgb_wgrad(f, img, pool, scale, cls, want_bias=False):
    b, c, h, w = _dims(f)
    cimg = img.shape[1]
    acc = torch.zeros((c, cimg), dtype=torch.float32, device=f.device)
    bs = torch.zeros((c,), dtype=torch.float32, device=f.device) if want_bias else None
    L.check(L.load().pggan_rgb_wgrad(_p(f), _p(img), _p(acc), _p(bs), b, h, w, c, cimg, int(pool), float(scale),
                                     _fmt_of(f, cls), _s()), "pggan_rgb_wgrad")
    return (acc, bs) if want_bias else acc


def _img(t):
    return t.float().contiguous()


class RgbExpand(torch.autograd.Function):
    """NHWC[b,h,w,c] = lrelu(scale*(sum_j Wc[c,j]*p_j + bias[c]), slope); p = image pixel or its 2x2 mean.
    Wc is [C, cimg] fp32.  (fromRGBBlock, reference lib/blocks.py:285-292)
    out_cls: class of the produced feature map ("g" when this is the dgrad of a toRGB layer)."""

    @staticmethod
    def forward(ctx, img, Wc, bias, pool, scale, slope, defer_mask=False, out_cls=A, f32=False):
        """defer_mask: the consumer (a ConvF with x_act_slope) applies this layer's LeakyReLU mask in its dgrad
        f32: produce an fp32 feature map (the fp32 family) instead of the mode's 16-bit format"""
        img = _img(img)
        _need_cuda(img)
        wc = Wc.detach().float().contiguous()
        bs = None if bias is None else bias.detach().float().contiguous()
        y = _k_rgb_expand(img, wc, bs, wc.shape[0], pool, scale, slope, out_cls, f32)
        ctx.mask_here = slope != 1.0 and not defer_mask
        ctx.save_for_backward(img, Wc, y if ctx.mask_here else None)
        ctx.pool, ctx.scale, ctx.slope, ctx.has_bias = pool, scale, slope, bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        img, Wc, y = ctx.saved_tensors
        g = feat(gy, G)
        need_w = ctx.needs_input_grad[1] and not _SKIP_PARAM_GRADS
        need_b = ctx.has_bias and ctx.needs_input_grad[2] and not _SKIP_PARAM_GRADS
        gimg = gW = gb = None
        if torch.is_grad_enabled():
            if ctx.mask_here:
                g = MaskMul.apply(g, y, ctx.slope)
            if ctx.needs_input_grad[0]:
                gimg = RgbReduce.apply(g, Wc, None, ctx.pool, ctx.scale, 0.25, G)
            if need_w:
                gW = RgbWgrad.apply(g, img, ctx.pool, ctx.scale, G)
            if need_b:
                gb = ColSum.apply(g, ctx.scale)
            return gimg, gW, gb, None, None, None, None, None, None
        if ctx.mask_here:
            g = k_mask_mul(g, y, ctx.slope, G)
        if ctx.needs_input_grad[0]:
            wr = Wc.detach().float().t().contiguous()
            gimg = _k_rgb_reduce(g, wr, None, wr.shape[0], ctx.pool, ctx.scale, 0.25, G)
        if need_w or need_b:                   # weight and bias gradients share one pass over g
            gW, gb = _k_rgb_wgrad(g, img, ctx.pool, ctx.scale, G, want_bias=True)
            if not need_w:
                gW = None
            if not need_b:
                gb = None
        return gimg, gW, gb, None, None, None, None, None, None


class RgbReduce(torch.autograd.Function):
    """img[b,j,h,w] = scale*(sum_c Wc[c,j]*x[b,h,w,c] + bias[j]); if up, written (times up_scale) to
    the 2x2 block of a 2x larger image.  (toRGBBlock, reference lib/blocks.py:164-170); xc: class of x"""

    @staticmethod
    def forward(ctx, x, Wc, bias, up, scale, up_scale, xc=A):
        x = feat(x, xc)
        wr = Wc.detach().float().t().contiguous()          # [cimg, C]
        bs = None if bias is None else bias.detach().float().contiguous()
        out = _k_rgb_reduce(x, wr, bs, wr.shape[0], up, scale, up_scale, xc)
        ctx.save_for_backward(x, Wc)
        ctx.up, ctx.scale, ctx.up_scale, ctx.has_bias, ctx.xc = up, scale, up_scale, bias is not None, xc
        return out

    @staticmethod
    def backward(ctx, gout):
        x, Wc = ctx.saved_tensors
        gout = _img(gout)
        s = ctx.scale * (4.0 * ctx.up_scale if ctx.up else 1.0)    # adjoint of "write 4 copies * up_scale"
        gx = gW = gb = None
        if ctx.needs_input_grad[0]:
            gx = RgbExpand.apply(gout, Wc, None, ctx.up, s, 1.0, False, G, x.dtype == torch.float32)
        if not _SKIP_PARAM_GRADS:
            if ctx.needs_input_grad[1]:
                gW = RgbWgrad.apply(x, gout, ctx.up, s, ctx.xc)
            if ctx.has_bias and ctx.needs_input_grad[2]:
                gb = gout.sum(dim=(0, 2, 3)) * (ctx.scale * (ctx.up_scale if ctx.up else 1.0))
        return gx, gW, gb, None, None, None, None


class RgbWgrad(torch.autograd.Function):
    """gWc[c,j] = scale * sum_pixels f[.,c] * p_j;  fc: class of f"""

    @staticmethod
    def forward(ctx, f, img, pool, scale, fc=G):
        f, img = feat(f, fc), _img(img)
        ctx.save_for_backward(f, img)
        ctx.pool, ctx.scale, ctx.fc = pool, scale, fc
        return _k_rgb_wgrad(f, img, pool, scale, fc)

    @staticmethod
    def backward(ctx, ggW):
        f, img = ctx.saved_tensors
        df = dimg = None
        if ctx.needs_input_grad[0]:
            df = RgbExpand.apply(img, ggW, None, ctx.pool, ctx.scale, 1.0, False, G, f.dtype == torch.float32)
        if ctx.needs_input_grad[1]:
            dimg = RgbReduce.apply(f, ggW, None, ctx.pool, ctx.scale, 0.25, ctx.fc)
        return df, dimg, None, None, None


class ImageFade(torch.autograd.Function):
    """(1-alpha) * up2(low) + alpha * x on fp32 NCHW images; low=None -> low = avgpool2(x)
    (G's RGB fade, reference pggan/nets.py:155-156; real-image fade, pggan/model.py:217-221).
    alpha: a float or an AlphaRef (device scalar: a captured graph follows the fade-in)."""

    @staticmethod
    def forward(ctx, x, low, alpha):
        x = _img(x)
        _need_cuda(x)
        lo = None if low is None else _img(low)
        b, c, h, w = x.shape
        out = torch.empty_like(x)
        L.check(L.load().pggan_image_fade(_p(x), _p(lo), _p(out), b, c, h, w, _coef(0.0, 1.0, alpha), _s()),
                "pggan_image_fade")
        ctx.alpha, ctx.self_low = alpha, low is None
        return out

    @staticmethod
    def backward(ctx, g):
        g = _img(g)
        b, c, h, w = g.shape
        if ctx.self_low:
            # F(x) = (1-a)*up(avgpool(x)) + a*x is self-adjoint (up o avgpool = 0.25 * up o up^T)
            return ImageFade.apply(g, None, ctx.alpha), None, None
        gx = glow = None
        if ctx.needs_input_grad[0]:
            gx = k_axpby(g, None, _affine(0.0, 1.0, ctx.alpha), 0.0)
        if ctx.needs_input_grad[1]:
            glow = torch.empty((b, c, h // 2, w // 2), dtype=torch.float32, device=g.device)
            L.check(L.load().pggan_image_pool_sum(_p(g), _p(glow), b, c, h, w, _coef(1.0, -1.0, ctx.alpha), _s()),
                    "pggan_image_pool_sum")
        return gx, glow, None


# -------------------------------------------------------------------------- minibatch stddev
class MbstdStat(torch.autograd.Function):
    """per-group scalar of reference lib/blocks.py:216-224 (x: [B, ...] features, any layout)"""

    @staticmethod
    def forward(ctx, x, group):
        x = feat(x, A)
        B = x.shape[0]
        K = x.numel() // B
        stat = torch.empty((B // group,), dtype=torch.float32, device=x.device)
        L.check(L.load().pggan_mbstd_fwd(_p(x), _p(stat), B, K, group, _fmt_of(x, A), _s()), "pggan_mbstd_fwd")
        ctx.save_for_backward(x)
        ctx.group = group
        return stat

    @staticmethod
    def backward(ctx, gs):
        (x,) = ctx.saved_tensors
        return MbstdBwd.apply(x, gs.float().contiguous(), ctx.group), None


class MbstdBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gs, group):
        B = x.shape[0]
        K = x.numel() // B
        gx = torch.empty_like(x)
        L.check(L.load().pggan_mbstd_bwd(_p(x), _p(gs), _p(gx), B, K, group, _fmt_of(x, A), _fmt_of(gx, G), _s()),
                "pggan_mbstd_bwd")
        ctx.save_for_backward(x, gs)
        ctx.group = group
        return gx

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, u):
        x, gs = ctx.saved_tensors
        u = feat(u, G, small=(x.dtype == torch.float32))
        if u.stride() != x.stride():
            u = u.contiguous(memory_format=torch.channels_last) if x.dim() == 4 else u.contiguous()
        B = x.shape[0]
        K = x.numel() // B
        ggs = torch.empty_like(gs)
        gx2 = torch.empty_like(u)
        L.check(L.load().pggan_mbstd_bwd2(_p(x), _p(u), _p(gs), _p(ggs), _p(gx2), B, K, ctx.group, _fmt_of(x, A), _fmt_of(u, G),
                                          _s()), "pggan_mbstd_bwd2")
        return gx2, ggs, None


class StatChannel(torch.autograd.Function):
    """per-group statistic [B/G] fp32 -> the constant extra channel [B,1,H,W] of the activation class
    (the expand/cat half of reference lib/blocks.py:226-231)"""

    @staticmethod
    def forward(ctx, stat, group, batch, h, w, f32=False):
        ctx.group = group
        y = stat.repeat_interleave(group).view(batch, 1, 1, 1).expand(batch, 1, h, w)
        return enter(y.contiguous(), A, f32)

    @staticmethod
    def backward(ctx, g):
        gs = to_float(g, G).sum(dim=(1, 2, 3)).view(-1, ctx.group).sum(dim=1)
        return gs, None, None, None, None, None


# ------------------------------------------------------------------------- loss reductions
class RowsSumSq(torch.autograd.Function):
    """[B, ...] fp32 -> [B] sum of squares (R1: reference lib/loss.py:132-134)"""

    @staticmethod
    def forward(ctx, g):
        g = _img(g)
        _need_cuda(g)
        B = g.shape[0]
        out = torch.zeros((B,), dtype=torch.float32, device=g.device)
        L.check(L.load().pggan_rows_sumsq(_p(g), _p(out), B, g.numel() // B, _s()), "pggan_rows_sumsq")
        ctx.save_for_backward(g)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        (g,) = ctx.saved_tensors
        B = g.shape[0]
        coef = (2.0 * gout).float().contiguous()
        out = torch.empty_like(g)
        L.check(L.load().pggan_rows_scale(_p(g), _p(coef), _p(out), B, g.numel() // B, _s()), "pggan_rows_scale")
        return out


class GradientPenalty(torch.autograd.Function):
    """w_gp * sum_b (||g_b||_2 - 1)^2 and, fused, w_drift * sum_b pred_b^2 (value only)
    (reference pggan/loss.py:85-87, 98-100).  Returns a 2-vector [gp, drift].
    seed_scale: g is seed_scale times the true gradient (the first-order chain was seeded with seed_scale instead of 1
    to keep its half-precision tensors in the normal range); the norm is taken of g / seed_scale."""

    @staticmethod
    def forward(ctx, g, pred, w_gp, w_drift, seed_scale=1.0):
        g = _img(g)
        _need_cuda(g)
        B = g.shape[0]
        n = g.numel() // B
        sumsq = torch.zeros((B,), dtype=torch.float32, device=g.device)
        coef = torch.empty((B,), dtype=torch.float32, device=g.device)
        res = torch.empty((2,), dtype=torch.float32, device=g.device)
        pr = None if pred is None else pred.detach().float().contiguous()
        L.check(L.load().pggan_rows_sumsq(_p(g), _p(sumsq), B, n, _s()), "pggan_rows_sumsq")
        inv2 = 1.0 / (float(seed_scale) * float(seed_scale))
        if inv2 != 1.0:
            sumsq.mul_(inv2)
        L.check(L.load().pggan_gp_finalize(_p(sumsq), _p(pr), _p(res), _p(coef), B, float(w_gp), float(w_drift), _s()),
                "pggan_gp_finalize")
        ctx.save_for_backward(g, coef)
        ctx.inv2 = inv2
        return res

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gres):
        g, coef = ctx.saved_tensors
        B = g.shape[0]
        c = (coef * gres[0] * ctx.inv2).contiguous()          # d/dg of f(||g/s||): coef(||g/s||) * (g/s) / s
        out = torch.empty_like(g)
        L.check(L.load().pggan_rows_scale(_p(g), _p(c), _p(out), B, g.numel() // B, _s()), "pggan_rows_scale")
        return out, None, None, None, None
