"""Kernel wrappers and double-differentiable autograd Functions over libpggan_b200.so.

Every Function's backward is itself built from Functions of this file, so
``torch.autograd.grad(..., create_graph=True)`` (R1 — reference lib/loss.py:125-135; WGAN-GP —
reference pggan/loss.py:78-90) can be differentiated again: {fprop, dgrad, wgrad} are closed under
differentiation, LeakyReLU contributes only its mask, pooling / up-sampling /
blends are linear, and minibatch-stddev has a true second-order kernel.

Feature maps are channels_last tensors ([B,C,H,W] logical, NHWC physical) in the compute dtype
(torch.bfloat16, or torch.float32 which runs on kind::tf32 MMAs); images/logits/parameters are fp32.
"""
from __future__ import annotations

import contextlib
import ctypes
from typing import Optional

import torch

from . import _lib as L

_COMPUTE_DTYPE = torch.bfloat16
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


def set_compute_dtype(dtype: torch.dtype) -> None:
    """torch.bfloat16 (bf16 operands, fp32 accumulate — the default), torch.float16 (same kernels and speed, 11-bit
    significand: meets the 1e-2 forward bar bf16 misses beyond 16x16; 5-bit exponent, so activations / gradients
    must stay inside +-65504 / above ~6e-8) or torch.float32 (3xTF32 operands, fp32 accumulate: the parity mode)."""
    global _COMPUTE_DTYPE
    if dtype not in (torch.bfloat16, torch.float16, torch.float32):
        raise ValueError("compute dtype must be torch.bfloat16, torch.float16 or torch.float32 (3xTF32)")
    _COMPUTE_DTYPE = dtype


def get_compute_dtype() -> torch.dtype:
    return _COMPUTE_DTYPE


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
# Parameters registered here (by optim.FlatGrads) own a persistent fp32 `.grad` view into a flat buffer.
# In a plain backward (no create_graph) the GEMM-class Functions then ACCUMULATE weight / bias gradients
# straight into that view (pggan_unpack_wgrad accumulate=1, atomics of the bias column sums) and return
# None to autograd: no temporary gradient tensor, no AccumulateGrad add kernel, no zero-fill.
# Not used under data parallelism (GradSync needs autograd's post-accumulate hooks for bucket readiness).
_GRAD_SINK: dict = {}


def register_grad_sink(params):
    import weakref
    for p in params:
        _GRAD_SINK[id(p)] = weakref.ref(p)


def unregister_grad_sink(params):
    for p in params:
        _GRAD_SINK.pop(id(p), None)


def _sink(p):
    """the parameter's .grad if gradients may be accumulated into it in place, else None"""
    if p is None:
        return None
    ref = _GRAD_SINK.get(id(p))
    if ref is None or ref() is not p:
        return None
    g = p.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.shape != p.shape:
        return None
    return g


# ------------------------------------------------------------------------------------ helpers
def _enum(dtype: torch.dtype) -> int:
    if dtype == torch.bfloat16:
        return L.PG_BF16
    if dtype == torch.float16:
        return L.PG_F16
    if dtype == torch.float32:
        return L.PG_F32
    raise TypeError(f"unsupported feature dtype {dtype}")


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


def feat(x: torch.Tensor) -> torch.Tensor:
    """cl() + cast to the compute dtype."""
    if x.dtype != _COMPUTE_DTYPE:
        x = x.to(_COMPUTE_DTYPE)
    return cl(x)


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


def _maps(wshape, layout):
    """-> (N, C, taps, fmap, tmap): logical dims of the fprop GEMM and the pg_pack_map argument tuples
    (n_total, n_valid, n2, taps, c_total, c_valid, c2, base, s_n1, s_n2, s_t, s_c1, s_c2) for the
    fprop operand [N][taps][C] and the dgrad operand [C][taps][N] (taps reversed)."""
    O, I = int(wshape[0]), int(wshape[1])
    T = 1
    for d in wshape[2:]:
        T *= int(d)
    kind = layout[0] if layout else "oihw"
    if kind == "oihw":
        return O, I, T, (O, O, 1, T, I, I, 1, 0, I * T, 0, 1, T, 0), (I, I, 1, T, O, O, 1, T - 1, T, 0, -1, I * T, 0)
    if kind == "pad_c":
        Cp = int(layout[1])
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


def _pack(W: torch.Tensor, kind: str, dtype: torch.dtype, layout=None) -> torch.Tensor:
    """fp32 parameter -> K-major packed operand in `dtype`; kind 'f' = fprop [N][taps*C], 't' = dgrad
    [C][taps*N] with the taps reversed.  Cached per (parameter, version) for nn.Parameters."""
    _need_cuda(W)
    cacheable = isinstance(W, torch.nn.Parameter)
    key = (id(W), W._version, kind, dtype, layout) if cacheable else None
    if cacheable:
        hit = _PACK_CACHE.get(key)
        if hit is not None and hit[0] is W:
            return hit[1]
    Wc = W.detach()
    if Wc.dtype != torch.float32 or not Wc.is_contiguous():
        Wc = Wc.float().contiguous()
    N, C, T, fmap, tmap = _maps(Wc.shape, layout)
    args = fmap if kind == "f" else tmap
    out = torch.empty((args[0], args[3] * args[4]), dtype=dtype, device=W.device)
    m = _packmap(args)
    L.check(L.load().pggan_pack_weight(_p(Wc), _p(out), _enum(dtype), ctypes.byref(m), _s()), "pggan_pack_weight")
    if cacheable:
        if len(_PACK_CACHE) > 1024:
            _PACK_CACHE.clear()
        _PACK_CACHE[key] = (W, out)
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


# ------------------------------------------------------------------------------- raw kernels
_KIND_CACHE: dict = {}


def conv_pn_fusable(x, N, taps) -> bool:
    """may k_conv(..., pn_eps=...) fuse the pixel norm into this layer's epilogue? (one epilogue thread must own all
    N channels of a pixel: the rows kernel, or the plain kernel when one N tile covers the layer)"""
    b, c, h, w = _dims(x)
    key = ("pn", b, h, w, c, N, taps, x.dtype)
    hit = _KIND_CACHE.get(key)
    if hit is None:
        p = L.ConvParams()
        p.B, p.H, p.W, p.C, p.N, p.taps = b, h, w, c, N, taps
        p.dtype = p.out_dtype = _enum(x.dtype)
        hit = bool(L.load().pggan_conv_pn_fusable(ctypes.byref(p)))
        _KIND_CACHE[key] = hit
    return hit


def k_conv(x, wp, bias, N, taps, scale, slope=1.0, mask_src=None, mask_slope=1.0, raw_out=False, pn_eps=None):
    """pn_eps: also apply the pixel norm in the epilogue; returns (y, rnorm[pixels] fp32) then."""
    _need_cuda(x)
    b, c, h, w = _dims(x)
    out_dtype = torch.float32 if raw_out else x.dtype
    y = _new_like_layout(x, N, out_dtype)
    p = L.ConvParams()
    rn = None
    if pn_eps is not None:
        rn = torch.empty((b * h * w,), dtype=torch.float32, device=x.device)
        p.pn_rnorm, p.pn_eps = rn.data_ptr(), float(pn_eps)
    p.x, p.wp, p.bias, p.mask_src, p.y = x.data_ptr(), wp.data_ptr(), (bias.data_ptr() if bias is not None else None), \
        (mask_src.data_ptr() if mask_src is not None else None), y.data_ptr()
    p.B, p.H, p.W, p.C, p.N, p.taps = b, h, w, c, N, taps
    p.dtype = _enum(x.dtype)
    p.out_dtype = L.PG_F32_RAW if raw_out else _enum(x.dtype)
    p.scale, p.slope, p.mask_slope = float(scale), float(slope), float(mask_slope)
    ws = None
    if x.dtype == torch.float32:          # fp32 mode = 3xTF32: the library splits operands into these
        ws = (torch.empty(3 * x.numel(), dtype=torch.float32, device=x.device),
              torch.empty(3 * wp.numel(), dtype=torch.float32, device=x.device))
        p.ws_x, p.ws_w = ws[0].data_ptr(), ws[1].data_ptr()
    with _Timed("conv_gemm", (b, h, w, c, N, taps)):
        L.check(L.load().pggan_conv_fprop(ctypes.byref(p), _s()), "pggan_conv_fprop")
    return y if pn_eps is None else (y, rn)


def dgrad_mask_fusable(batch, h, w, fwd_in_channels, fwd_out_channels, taps=9, dtype=None) -> bool:
    """True if the dgrad of a (fwd_in_channels -> fwd_out_channels) conv at this size runs on a kernel whose epilogue
    applies a LeakyReLU-backward mask for free (see pggan_conv_kernel_kind); the blocks use it to decide whether a
    producer may defer its mask to this layer's dgrad."""
    dtype = dtype or _COMPUTE_DTYPE
    key = (batch, h, w, fwd_in_channels, fwd_out_channels, taps, dtype)
    hit = _KIND_CACHE.get(key)
    if hit is None:
        p = L.ConvParams()
        p.B, p.H, p.W, p.C, p.N, p.taps = batch, h, w, fwd_out_channels, fwd_in_channels, taps     # dgrad: roles swapped
        p.dtype = _enum(dtype)
        p.out_dtype = _enum(dtype)
        hit = L.load().pggan_conv_kernel_kind(ctypes.byref(p)) in (L.PG_CONV_KIND_PLAIN, L.PG_CONV_KIND_ROWS)
        _KIND_CACHE[key] = hit
    return hit


def k_wgrad(g, x, scale, wshape, layout=None, accumulate_into=None):
    """weight gradient in the parameter's own fp32 layout: a new tensor, or added into `accumulate_into`"""
    _need_cuda(g)
    b, n, h, w = _dims(g)
    c = _dims(x)[1]
    N, C, taps, fmap, _ = _maps(wshape, layout)
    if (n, c) != (N, C):
        raise ValueError(f"wgrad operand channels ({n},{c}) do not match the layer ({N},{C})")
    p = L.WgradParams()
    p.g, p.x = g.data_ptr(), x.data_ptr()
    p.B, p.H, p.W, p.C, p.N, p.taps = b, h, w, c, n, taps
    p.dtype = _enum(g.dtype)
    slices = L.load().pggan_conv_wgrad_slices(ctypes.byref(p))
    acc = torch.empty((slices, n, taps, c), dtype=torch.float32, device=g.device)
    p.acc = acc.data_ptr()
    ws = None
    if g.dtype == torch.float32:
        ws = (torch.empty(3 * g.numel(), dtype=torch.float32, device=g.device),
              torch.empty(3 * x.numel(), dtype=torch.float32, device=g.device))
        p.ws_g, p.ws_x = ws[0].data_ptr(), ws[1].data_ptr()
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


def k_mask_colsum(g, y, slope, want_out, colsum_scale, colsum_into=None):
    """fused LeakyReLU backward (+ bias gradient): returns (g*mask(y) or None, colsum or None).
    colsum_into: an fp32 [C] tensor the column sums are ADDED to (then the second result is None)."""
    b, c, h, w = _dims(g)
    vn = 16 // g.element_size()
    if c % vn != 0 or (c // vn) > 256 or 256 % (c // vn) != 0:          # odd widths: two plain passes
        out = k_mask_mul(g, y, slope) if (want_out and y is not None) else (g if want_out else None)
        src = out if out is not None else (k_mask_mul(g, y, slope) if y is not None else g)
        cs = k_colsum(src, colsum_scale) if colsum_scale is not None else None
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
                                       float(colsum_scale or 0.0), _enum(g.dtype), _s()), "pggan_mask_colsum")
    if want_out and out is None:
        out = g
    return out, (None if colsum_into is not None else cs)


def k_mask_mul(g, y, slope):
    out = torch.empty_like(g)
    L.check(L.load().pggan_lrelu_mask_mul(_p(g), _p(y), _p(out), g.numel(), float(slope), _enum(g.dtype), _s()),
            "pggan_lrelu_mask_mul")
    return out


def k_colsum(g, scale):
    b, c, h, w = _dims(g)
    if c % (16 // g.element_size()) != 0:      # e.g. the 1-logit decision layer: a [B,1] column
        g2 = g.float().reshape(b, c) if g.dim() == 2 else g.float().permute(0, 2, 3, 1).reshape(-1, c)
        return g2.sum(0) * scale
    out = torch.zeros((c,), dtype=torch.float32, device=g.device)
    L.check(L.load().pggan_colsum(_p(g), _p(out), b * h * w, c, float(scale), _enum(g.dtype), _s()), "pggan_colsum")
    return out


def k_upsample(x, scale, mask_src=None, slope=1.0):
    b, c, h, w = _dims(x)
    out = _new_like_layout(x, c, x.dtype, 2 * h, 2 * w)
    if mask_src is not None:
        L.check(L.load().pggan_upsample2x_mask(_p(x), _p(mask_src), _p(out), b, h, w, c, float(scale), float(slope),
                                               _enum(x.dtype), _s()), "pggan_upsample2x_mask")
        return out
    L.check(L.load().pggan_upsample2x(_p(x), _p(out), b, h, w, c, float(scale), _enum(x.dtype), _s()),
            "pggan_upsample2x")
    return out


def k_pool(x, scale, blend_with=None, alpha=1.0):
    b, c, h, w = _dims(x)
    out = _new_like_layout(x, c, x.dtype, h // 2, w // 2)
    L.check(L.load().pggan_pool2x2(_p(x), _p(blend_with), _p(out), b, h, w, c, float(scale), float(alpha),
                                   _enum(x.dtype), _s()), "pggan_pool2x2")
    return out


def k_axpby(a, b, sa, sb):
    out = torch.empty_like(a)
    L.check(L.load().pggan_axpby(_p(a), _p(b), _p(out), a.numel(), float(sa), float(sb), _enum(a.dtype), _s()),
            "pggan_axpby")
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
                    with act_slope): the incoming gradient is already masked and y need not be saved."""

    @staticmethod
    def forward(ctx, x, W, b, scale, slope, raw_out, layout=None, x_act_slope=None, defer_mask=False):
        x = feat(x)
        N, C, taps, _, _ = _maps(W.shape, layout)
        bias = None if b is None else _perm_bias(b.detach().float(), layout).contiguous()
        y = k_conv(x, _pack(W, "f", x.dtype, layout), bias, N, taps, scale, slope, raw_out=raw_out)
        ctx.mask_here = slope != 1.0 and not defer_mask
        ctx.save_for_backward(x, W, y if ctx.mask_here else None)
        ctx.scale, ctx.slope, ctx.has_bias, ctx.layout = scale, slope, b is not None, layout
        ctx.x_act_slope = x_act_slope
        ctx.bias_param = b if isinstance(b, torch.nn.Parameter) else None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, W, y = ctx.saved_tensors
        g = feat(gy)
        need_w = ctx.needs_input_grad[1] and not _SKIP_PARAM_GRADS
        need_b = ctx.has_bias and ctx.needs_input_grad[2] and not _SKIP_PARAM_GRADS
        gx = gW = gb = None
        xm, xs = (x, ctx.x_act_slope) if ctx.x_act_slope is not None else (None, 1.0)
        if torch.is_grad_enabled():
            # create_graph=True: stay composable (every piece is a differentiable Function)
            if ctx.mask_here:
                g = MaskMul.apply(g, y, ctx.slope)
            if need_b:
                gb = ColSum.apply(g, ctx.scale)
        else:
            # plain backward: LeakyReLU mask and bias gradient in one pass over g; gradients of registered
            # parameters are accumulated in place (see _GRAD_SINK)
            b_sink = _sink(ctx.bias_param) if (need_b and not (ctx.layout and ctx.layout[0] == "lin_out_perm")) else None
            g, gb = k_mask_colsum(g, y if ctx.mask_here else None, ctx.slope, True,
                                  ctx.scale if need_b else None, colsum_into=b_sink)
            if gb is not None:
                gb = _perm_bias(gb, ctx.layout, inverse=True)
            if ctx.needs_input_grad[0]:
                gx = ConvT.apply(g, W, ctx.scale, ctx.layout, xm, xs)
            if need_w:
                gW = k_wgrad(g, x, ctx.scale, tuple(W.shape), ctx.layout, accumulate_into=_sink(W))
            return gx, gW, gb, None, None, None, None, None, None
        if gb is not None:
            gb = _perm_bias(gb, ctx.layout, inverse=True)
        if ctx.needs_input_grad[0]:
            gx = ConvT.apply(g, W, ctx.scale, ctx.layout, xm, xs)
        if need_w:
            gW = ConvW.apply(g, x, ctx.scale, tuple(W.shape), ctx.layout)
        return gx, gW, gb, None, None, None, None, None, None


class ConvActPN(torch.autograd.Function):
    """G's  conv3x3 -> c*(.+b) -> LeakyReLU -> pixel norm  (reference lib/blocks.py:119-139 with
    lib/layers.py:13-14): two kernels forward (GEMM with fused bias/He/LReLU epilogue, pixel norm) and a
    backward whose first pass fuses pixel-norm backward, the LeakyReLU mask (sign(y_pn) == sign(pre))
    and the bias gradient.  The generator is never differentiated twice."""

    @staticmethod
    def forward(ctx, x, W, b, scale, slope, eps, layout=None):
        x = feat(x)
        N, C, taps, _, _ = _maps(W.shape, layout)
        bias = None if b is None else _perm_bias(b.detach().float(), layout).contiguous()
        if x.dim() == 4 and conv_pn_fusable(x, N, taps):
            # conv + He scale + bias + LeakyReLU + pixel norm in ONE kernel (the epilogue thread owns the whole pixel)
            y, r = k_conv(x, _pack(W, "f", x.dtype, layout), bias, N, taps, scale, slope, pn_eps=eps)
        else:
            a = k_conv(x, _pack(W, "f", x.dtype, layout), bias, N, taps, scale, slope)
            bb, c, h, w = _dims(a)
            y = torch.empty_like(a)
            r = torch.empty((bb * h * w,), dtype=torch.float32, device=a.device)
            L.check(L.load().pggan_pixelnorm_fwd(_p(a), _p(y), _p(r), bb * h * w, c, float(eps), _enum(a.dtype), _s()),
                    "pggan_pixelnorm_fwd")
        ctx.save_for_backward(x, W, y, r)
        ctx.scale, ctx.slope, ctx.has_bias, ctx.layout = scale, slope, b is not None, layout
        ctx.bias_param = b if isinstance(b, torch.nn.Parameter) else None
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, W, y, r = ctx.saved_tensors
        gy = feat(gy)
        bb, c, h, w = _dims(y)
        need_b = ctx.has_bias and ctx.needs_input_grad[2] and not _SKIP_PARAM_GRADS
        g = torch.empty_like(y)
        b_sink = _sink(ctx.bias_param) if (need_b and not (ctx.layout and ctx.layout[0] == "lin_out_perm")) else None
        gb = None
        if need_b:
            gb = b_sink if b_sink is not None else torch.zeros((c,), dtype=torch.float32, device=y.device)
        L.check(L.load().pggan_pixelnorm_bwd(_p(gy), _p(y), _p(r), _p(g), bb * h * w, c, float(ctx.slope), _p(gb),
                                             float(ctx.scale), _enum(y.dtype), _s()), "pggan_pixelnorm_bwd")
        gx = gW = None
        if ctx.needs_input_grad[0]:
            N, C, taps, _, _ = _maps(W.shape, ctx.layout)
            gx = k_conv(g, _pack(W, "t", g.dtype, ctx.layout), None, C, taps, ctx.scale)
        if ctx.needs_input_grad[1] and not _SKIP_PARAM_GRADS:
            gW = k_wgrad(g, x, ctx.scale, tuple(W.shape), ctx.layout, accumulate_into=_sink(W))
        if b_sink is not None:
            gb = None
        elif gb is not None:
            gb = _perm_bias(gb, ctx.layout, inverse=True)
        return gx, gW, gb, None, None, None, None


class ConvT(torch.autograd.Function):
    """gx = scale * conv_transpose(g, W)   [dgrad: the same kernel on flipped/transposed weights]"""

    @staticmethod
    def forward(ctx, g, W, scale, layout=None, mask_src=None, mask_slope=1.0):
        """mask_src: gx is additionally multiplied by (mask_src > 0 ? 1 : mask_slope) in the GEMM epilogue
        (the LeakyReLU backward of the layer that produced this conv's input)."""
        g = feat(g)
        N, C, taps, _, _ = _maps(W.shape, layout)
        gx = k_conv(g, _pack(W, "t", g.dtype, layout), None, C, taps, scale, mask_src=mask_src, mask_slope=mask_slope)
        ctx.save_for_backward(g, W, mask_src)
        ctx.scale, ctx.layout, ctx.mask_slope = scale, layout, mask_slope
        return gx

    @staticmethod
    def backward(ctx, ggx):
        g, W, mask_src = ctx.saved_tensors
        ggx = feat(ggx)
        if mask_src is not None:
            ggx = MaskMul.apply(ggx, mask_src, ctx.mask_slope)
        dg = dW = None
        if ctx.needs_input_grad[0]:
            dg = ConvF.apply(ggx, W, None, ctx.scale, 1.0, False, ctx.layout)
        if ctx.needs_input_grad[1] and not _SKIP_PARAM_GRADS:
            dW = ConvW.apply(g, ggx, ctx.scale, tuple(W.shape), ctx.layout)
        return dg, dW, None, None, None, None


class ConvW(torch.autograd.Function):
    """gW = scale * wgrad(g, x)   in the parameter's layout"""

    @staticmethod
    def forward(ctx, g, x, scale, wshape, layout=None):
        g, x = feat(g), feat(x)
        ctx.save_for_backward(g, x)
        ctx.scale, ctx.layout = scale, layout
        return k_wgrad(g, x, scale, wshape, layout)

    @staticmethod
    def backward(ctx, ggW):
        g, x = ctx.saved_tensors
        dg = dx = None
        ggW = ggW.contiguous()
        if ctx.needs_input_grad[0]:
            dg = ConvF.apply(x, ggW, None, ctx.scale, 1.0, False, ctx.layout)
        if ctx.needs_input_grad[1]:
            dx = ConvT.apply(g, ggW, ctx.scale, ctx.layout)
        return dg, dx, None, None, None


class ColSum(torch.autograd.Function):
    """gb[c] = scale * sum over pixels of g[:, c]"""

    @staticmethod
    def forward(ctx, g, scale):
        g = feat(g)
        ctx.shape, ctx.scale, ctx.dtype = tuple(g.shape), scale, g.dtype
        return k_colsum(g, scale)

    @staticmethod
    def backward(ctx, ggb):
        shape = ctx.shape
        view = (1, -1, 1, 1) if len(shape) == 4 else (1, -1)
        return cl((ggb * ctx.scale).to(ctx.dtype).view(view).expand(shape)), None


class MaskMul(torch.autograd.Function):
    """g * (y > 0 ? 1 : slope): LeakyReLU backward from the saved OUTPUT (sign(y) == sign(pre))."""

    @staticmethod
    def forward(ctx, g, y, slope):
        g = feat(g)
        ctx.save_for_backward(y)
        ctx.slope = slope
        return k_mask_mul(g, y, slope)

    @staticmethod
    def backward(ctx, gg):
        (y,) = ctx.saved_tensors
        return MaskMul.apply(feat(gg), y, ctx.slope), None, None


# ------------------------------------------------------------------ resampling / blend Functions
class Upsample2x(torch.autograd.Function):
    """nearest 2x (reference lib/utils.py:106-118), times `scale`"""

    @staticmethod
    def forward(ctx, x, scale):
        ctx.scale = scale
        return k_upsample(feat(x), scale)

    @staticmethod
    def backward(ctx, g):
        return Pool2x2.apply(feat(g), ctx.scale), None


class UpsampleMask(torch.autograd.Function):
    """scale * up2(g) * (y > 0 ? 1 : slope): avg-pool backward fused with the LeakyReLU backward of the pooled
    activation y (one pass instead of two)"""

    @staticmethod
    def forward(ctx, g, y, scale, slope):
        ctx.save_for_backward(y)
        ctx.scale, ctx.slope = scale, slope
        return k_upsample(feat(g), scale, mask_src=y, slope=slope)

    @staticmethod
    def backward(ctx, gg):
        (y,) = ctx.saved_tensors
        return Pool2x2.apply(MaskMul.apply(feat(gg), y, ctx.slope), ctx.scale), None, None, None


class Pool2x2(torch.autograd.Function):
    """scale * (2x2 window sum); scale = 0.25 is nn.AvgPool2d(2) (reference lib/blocks.py:193).
    act_slope: x is the LeakyReLU output of a ConvF run with defer_mask=True — its mask is applied here, in backward."""

    @staticmethod
    def forward(ctx, x, scale, act_slope=None):
        x = feat(x)
        ctx.scale, ctx.act_slope = scale, act_slope
        if act_slope is not None:
            ctx.save_for_backward(x)
        return k_pool(x, scale)

    @staticmethod
    def backward(ctx, g):
        if ctx.act_slope is not None:
            (x,) = ctx.saved_tensors
            return UpsampleMask.apply(feat(g), x, ctx.scale, ctx.act_slope), None, None
        return Upsample2x.apply(feat(g), ctx.scale), None, None


class PoolBlend(torch.autograd.Function):
    """(1-alpha) * low + alpha * avgpool2(x): D's feature-space fade fused with the pooling
    (reference lib/blocks.py:193 + pggan/nets.py:262-265).  act_slope as in Pool2x2."""

    @staticmethod
    def forward(ctx, x, low, alpha, act_slope=None):
        x = feat(x)
        ctx.alpha, ctx.act_slope = alpha, act_slope
        if act_slope is not None:
            ctx.save_for_backward(x)
        return k_pool(x, 0.25, blend_with=feat(low), alpha=alpha)

    @staticmethod
    def backward(ctx, g):
        g = feat(g)
        gx = None
        if ctx.needs_input_grad[0]:
            if ctx.act_slope is not None:
                (x,) = ctx.saved_tensors
                gx = UpsampleMask.apply(g, x, 0.25 * ctx.alpha, ctx.act_slope)
            else:
                gx = Upsample2x.apply(g, 0.25 * ctx.alpha)
        gl = Scale.apply(g, 1.0 - ctx.alpha) if ctx.needs_input_grad[1] else None
        return gx, gl, None, None


class Scale(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, s):
        ctx.s = s
        return k_axpby(feat(x), None, s, 0.0)

    @staticmethod
    def backward(ctx, g):
        return Scale.apply(feat(g), ctx.s), None


class Lerp(torch.autograd.Function):
    """(1-alpha) * a + alpha * b on feature maps"""

    @staticmethod
    def forward(ctx, a, b, alpha):
        ctx.alpha = alpha
        return k_axpby(feat(a), feat(b), 1.0 - alpha, alpha)

    @staticmethod
    def backward(ctx, g):
        g = feat(g)
        ga = Scale.apply(g, 1.0 - ctx.alpha) if ctx.needs_input_grad[0] else None
        gb = Scale.apply(g, ctx.alpha) if ctx.needs_input_grad[1] else None
        return ga, gb, None


# ---------------------------------------------------------------------------------- pixel norm
class PixelNorm(torch.autograd.Function):
    """x * rsqrt(mean_c(x^2) + eps)  (reference lib/layers.py:13-14).  `act_slope` != 1 declares that
    x is a LeakyReLU output whose mask should be folded into THIS backward (the producing ConvF was
    then run with its own mask disabled) — sign(y) == sign(x) == sign(pre-activation)."""

    @staticmethod
    def forward(ctx, x, eps):
        x = feat(x)
        b, c, h, w = _dims(x)
        y = torch.empty_like(x)
        r = torch.empty((b * h * w,), dtype=torch.float32, device=x.device)
        L.check(L.load().pggan_pixelnorm_fwd(_p(x), _p(y), _p(r), b * h * w, c, float(eps), _enum(x.dtype), _s()),
                "pggan_pixelnorm_fwd")
        ctx.save_for_backward(y, r)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        y, r = ctx.saved_tensors
        gy = feat(gy)
        b, c, h, w = _dims(y)
        gx = torch.empty_like(y)
        L.check(L.load().pggan_pixelnorm_bwd(_p(gy), _p(y), _p(r), _p(gx), b * h * w, c, 1.0, None, 0.0,
                                             _enum(y.dtype), _s()), "pggan_pixelnorm_bwd")
        return gx, None


# ------------------------------------------------------------------------------------ RGB 1x1
def _k_rgb_expand(img, wc, bias, C, pool, scale, slope, dtype):
    b, cimg, hi, wi = img.shape
    h, w = (hi // 2, wi // 2) if pool else (hi, wi)
    out = torch.empty((b, C, h, w), dtype=dtype, device=img.device, memory_format=torch.channels_last)
    L.check(L.load().pggan_rgb_expand(_p(img), _p(wc), _p(bias), _p(out), b, h, w, C, cimg, int(pool), float(scale),
                                      float(slope), _enum(dtype), _s()), "pggan_rgb_expand")
    return out


def _k_rgb_reduce(x, wr, bias, cimg, up, scale, up_scale):
    b, c, h, w = _dims(x)
    ho, wo = (2 * h, 2 * w) if up else (h, w)
    out = torch.empty((b, cimg, ho, wo), dtype=torch.float32, device=x.device)
    L.check(L.load().pggan_rgb_reduce(_p(x), _p(wr), _p(bias), _p(out), b, h, w, c, cimg, int(up), float(scale),
                                      float(up_scale), _enum(x.dtype), _s()), "pggan_rgb_reduce")
    return out


def _k_rgb_wgrad(f, img, pool, scale, want_bias=False):
    b, c, h, w = _dims(f)
    cimg = img.shape[1]
    acc = torch.zeros((c, cimg), dtype=torch.float32, device=f.device)
    bs = torch.zeros((c,), dtype=torch.float32, device=f.device) if want_bias else None
    L.check(L.load().pggan_rgb_wgrad(_p(f), _p(img), _p(acc), _p(bs), b, h, w, c, cimg, int(pool), float(scale),
                                     _enum(f.dtype), _s()), "pggan_rgb_wgrad")
    return (acc, bs) if want_bias else acc


def _img(t):
    return t.float().contiguous()


class RgbExpand(torch.autograd.Function):
    """NHWC[b,h,w,c] = lrelu(scale*(sum_j Wc[c,j]*p_j + bias[c]), slope); p = image pixel or its 2x2 mean.
    Wc is [C, cimg] fp32.  (fromRGBBlock, reference lib/blocks.py:285-292)"""

    @staticmethod
    def forward(ctx, img, Wc, bias, pool, scale, slope, defer_mask=False):
        """defer_mask: the consumer (a ConvF with x_act_slope) applies this layer's LeakyReLU mask in its dgrad"""
        img = _img(img)
        _need_cuda(img)
        wc = Wc.detach().float().contiguous()
        bs = None if bias is None else bias.detach().float().contiguous()
        y = _k_rgb_expand(img, wc, bs, wc.shape[0], pool, scale, slope, _COMPUTE_DTYPE)
        ctx.mask_here = slope != 1.0 and not defer_mask
        ctx.save_for_backward(img, Wc, y if ctx.mask_here else None)
        ctx.pool, ctx.scale, ctx.slope, ctx.has_bias = pool, scale, slope, bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        img, Wc, y = ctx.saved_tensors
        g = feat(gy)
        need_w = ctx.needs_input_grad[1] and not _SKIP_PARAM_GRADS
        need_b = ctx.has_bias and ctx.needs_input_grad[2] and not _SKIP_PARAM_GRADS
        gimg = gW = gb = None
        if torch.is_grad_enabled():
            if ctx.mask_here:
                g = MaskMul.apply(g, y, ctx.slope)
            if ctx.needs_input_grad[0]:
                gimg = RgbReduce.apply(g, Wc, None, ctx.pool, ctx.scale, 0.25)
            if need_w:
                gW = RgbWgrad.apply(g, img, ctx.pool, ctx.scale)
            if need_b:
                gb = ColSum.apply(g, ctx.scale)
            return gimg, gW, gb, None, None, None, None
        if ctx.mask_here:
            g = k_mask_mul(g, y, ctx.slope)
        if ctx.needs_input_grad[0]:
            wr = Wc.detach().float().t().contiguous()
            gimg = _k_rgb_reduce(g, wr, None, wr.shape[0], ctx.pool, ctx.scale, 0.25)
        if need_w or need_b:                   # weight and bias gradients share one pass over g
            gW, gb = _k_rgb_wgrad(g, img, ctx.pool, ctx.scale, want_bias=True)
            if not need_w:
                gW = None
            if not need_b:
                gb = None
        return gimg, gW, gb, None, None, None, None


class RgbReduce(torch.autograd.Function):
    """img[b,j,h,w] = scale*(sum_c Wc[c,j]*x[b,h,w,c] + bias[j]); if up, written (times up_scale) to
    the 2x2 block of a 2x larger image.  (toRGBBlock, reference lib/blocks.py:164-170)"""

    @staticmethod
    def forward(ctx, x, Wc, bias, up, scale, up_scale):
        x = feat(x)
        wr = Wc.detach().float().t().contiguous()          # [cimg, C]
        bs = None if bias is None else bias.detach().float().contiguous()
        out = _k_rgb_reduce(x, wr, bs, wr.shape[0], up, scale, up_scale)
        ctx.save_for_backward(x, Wc)
        ctx.up, ctx.scale, ctx.up_scale, ctx.has_bias = up, scale, up_scale, bias is not None
        return out

    @staticmethod
    def backward(ctx, gout):
        x, Wc = ctx.saved_tensors
        gout = _img(gout)
        s = ctx.scale * (4.0 * ctx.up_scale if ctx.up else 1.0)    # adjoint of "write 4 copies * up_scale"
        gx = gW = gb = None
        if ctx.needs_input_grad[0]:
            gx = RgbExpand.apply(gout, Wc, None, ctx.up, s, 1.0)
        if not _SKIP_PARAM_GRADS:
            if ctx.needs_input_grad[1]:
                gW = RgbWgrad.apply(x, gout, ctx.up, s)
            if ctx.has_bias and ctx.needs_input_grad[2]:
                gb = gout.sum(dim=(0, 2, 3)) * (ctx.scale * (ctx.up_scale if ctx.up else 1.0))
        return gx, gW, gb, None, None, None


class RgbWgrad(torch.autograd.Function):
    """gWc[c,j] = scale * sum_pixels f[.,c] * p_j"""

    @staticmethod
    def forward(ctx, f, img, pool, scale):
        f, img = feat(f), _img(img)
        ctx.save_for_backward(f, img)
        ctx.pool, ctx.scale = pool, scale
        return _k_rgb_wgrad(f, img, pool, scale)

    @staticmethod
    def backward(ctx, ggW):
        f, img = ctx.saved_tensors
        df = dimg = None
        if ctx.needs_input_grad[0]:
            df = RgbExpand.apply(img, ggW, None, ctx.pool, ctx.scale, 1.0)
        if ctx.needs_input_grad[1]:
            dimg = RgbReduce.apply(f, ggW, None, ctx.pool, ctx.scale, 0.25)
        return df, dimg, None, None


class ImageFade(torch.autograd.Function):
    """(1-alpha) * up2(low) + alpha * x on fp32 NCHW images; low=None -> low = avgpool2(x)
    (G's RGB fade, reference pggan/nets.py:155-156; real-image fade, pggan/model.py:217-221)."""

    @staticmethod
    def forward(ctx, x, low, alpha):
        x = _img(x)
        _need_cuda(x)
        lo = None if low is None else _img(low)
        b, c, h, w = x.shape
        out = torch.empty_like(x)
        L.check(L.load().pggan_image_fade(_p(x), _p(lo), _p(out), b, c, h, w, float(alpha), _s()), "pggan_image_fade")
        ctx.alpha, ctx.self_low = alpha, low is None
        return out

    @staticmethod
    def backward(ctx, g):
        g = _img(g)
        b, c, h, w = g.shape
        pooled = torch.empty((b, c, h // 2, w // 2), dtype=torch.float32, device=g.device)
        L.check(L.load().pggan_image_pool_sum(_p(g), _p(pooled), b, c, h, w, 1.0 - ctx.alpha, _s()),
                "pggan_image_pool_sum")
        if ctx.self_low:
            # d/dx of (1-a)*up(pool(x)): up(pooled)/4
            gx = ctx.alpha * g + pooled.repeat_interleave(2, 2).repeat_interleave(2, 3) * 0.25
            return gx, None, None
        return ctx.alpha * g, pooled, None


# -------------------------------------------------------------------------- minibatch stddev
class MbstdStat(torch.autograd.Function):
    """per-group scalar of reference lib/blocks.py:216-224 (x: [B, ...] features, any layout)"""

    @staticmethod
    def forward(ctx, x, group):
        x = feat(x)
        B = x.shape[0]
        K = x.numel() // B
        stat = torch.empty((B // group,), dtype=torch.float32, device=x.device)
        L.check(L.load().pggan_mbstd_fwd(_p(x), _p(stat), B, K, group, _enum(x.dtype), _s()), "pggan_mbstd_fwd")
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
        L.check(L.load().pggan_mbstd_bwd(_p(x), _p(gs), _p(gx), B, K, group, _enum(x.dtype), _s()), "pggan_mbstd_bwd")
        ctx.save_for_backward(x, gs)
        ctx.group = group
        return gx

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, u):
        x, gs = ctx.saved_tensors
        u = feat(u)
        B = x.shape[0]
        K = x.numel() // B
        ggs = torch.empty_like(gs)
        gx2 = torch.empty_like(x)
        L.check(L.load().pggan_mbstd_bwd2(_p(x), _p(u), _p(gs), _p(ggs), _p(gx2), B, K, ctx.group, _enum(x.dtype),
                                          _s()), "pggan_mbstd_bwd2")
        return gx2, ggs, None


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
    (reference pggan/loss.py:85-87, 98-100).  Returns a 2-vector [gp, drift]."""

    @staticmethod
    def forward(ctx, g, pred, w_gp, w_drift):
        g = _img(g)
        _need_cuda(g)
        B = g.shape[0]
        n = g.numel() // B
        sumsq = torch.zeros((B,), dtype=torch.float32, device=g.device)
        coef = torch.empty((B,), dtype=torch.float32, device=g.device)
        res = torch.empty((2,), dtype=torch.float32, device=g.device)
        pr = None if pred is None else pred.detach().float().contiguous()
        L.check(L.load().pggan_rows_sumsq(_p(g), _p(sumsq), B, n, _s()), "pggan_rows_sumsq")
        L.check(L.load().pggan_gp_finalize(_p(sumsq), _p(pr), _p(res), _p(coef), B, float(w_gp), float(w_drift), _s()),
                "pggan_gp_finalize")
        ctx.save_for_backward(g, coef)
        return res

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gres):
        g, coef = ctx.saved_tensors
        B = g.shape[0]
        c = (coef * gres[0]).contiguous()
        out = torch.empty_like(g)
        L.check(L.load().pggan_rows_scale(_p(g), _p(c), _p(out), B, g.numel() // B, _s()), "pggan_rows_scale")
        return out, None, None, None
