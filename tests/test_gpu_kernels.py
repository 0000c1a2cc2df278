"""Kernel-level parity on the GPU: every C-ABI entry point against plain PyTorch fp32 on the same
inputs (inputs pre-rounded to the compute dtype so only accumulation order / output rounding differ)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu

DT = [torch.bfloat16, torch.float32, torch.float16, "mixed"]
# bf16: outputs are rounded to 8 significant bits (rms rel err ~2e-3), fp16: 11 bits (~2.5e-4); fp32 mode is 3xTF32
# (fp32-class; tensor-core fp32 accumulation truncates: ~2^-24 per MMA step, so long K drifts to ~1e-4).
# "mixed": activation-class tensors ("a") are fp16, gradient-class tensors ("g") bf16.
_TOL = {torch.bfloat16: 4e-3, torch.float32: 2e-4, torch.float16: 6e-4}


def real(dtype, cls="a"):
    """element format of a tensor of class `cls` in precision mode `dtype`"""
    if dtype == "mixed":
        return torch.float16 if cls == "a" else torch.bfloat16
    return dtype


def tol(dtype, cls="a"):
    return _TOL[real(dtype, cls)]


class _Tol:
    """TOL[dtype] = tolerance of a gradient-class result (the looser one in mixed mode)"""

    def __getitem__(self, dtype):
        return tol(dtype, "g")


TOL = _Tol()


@pytest.fixture(autouse=True)
def _setup():
    import pggan_b200  # noqa: F401
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    yield
    from pggan_b200 import _lib
    _lib.load().pggan_set_conv_backend(0)
    import pggan_b200 as pg
    pg.set_compute_dtype(torch.float16)


def rnd(t, dtype, cls="a"):
    """round to the operand precision the tensor core sees, keep fp32 storage"""
    r = real(dtype, cls)
    if r in (torch.bfloat16, torch.float16):
        return t.to(r).float()
    return t          # fp32 mode takes full fp32 operands (3xTF32)


def to_feat(t, dtype, cls="a"):
    """external tensor -> internal feature tensor of class `cls` in the current precision mode's 16-bit format
    (Functions handed a plain fp32 tensor compute in the fp32 family: precision follows the tensor)"""
    from pggan_b200 import ops
    return ops.enter(t, cls)


def f32(t, cls="a"):
    from pggan_b200 import ops
    return ops.to_float(t, cls)


CONV_SHAPES = [
    # B, H, W, Cin, Cout, taps
    (2, 8, 8, 64, 64, 9),
    (3, 4, 4, 512, 512, 9),        # ragged batch tile (BB=8 > B)
    (2, 32, 32, 128, 256, 9),
    (1, 64, 64, 64, 128, 9),
    (1, 8, 128, 64, 64, 9),        # W == tile width
    (2, 16, 16, 576, 512, 9),      # K chunk count not a power of two (mbstd-padded layer shape)
    (5, 1, 1, 512, 8192, 1),       # latent linear
    (64, 1, 1, 8192, 512, 1),      # D linear
    (2, 16, 16, 256, 32, 9),       # N tile 32
    (4, 32, 32, 128, 64, 9),       # swapped orientation (N <= 128): 64 output channels, wgrad stacks 2 taps
    (3, 16, 16, 64, 128, 9),       # swapped, ragged batch; wgrad stacks 4 taps of the 64-channel input
    (2, 64, 64, 64, 64, 9),        # the 256^2-stage layer shape (64 -> 64)
    (9, 8, 8, 64, 128, 9),         # swapped with a partially filled 256-pixel box in the batch dimension
    (2, 64, 64, 32, 32, 9),        # 64-byte K span (32 bf16 channels): the 512^2-stage shapes
    (2, 64, 64, 16, 16, 9),        # 32-byte K span, N tile 16: the 1024^2-stage shapes
    (1, 128, 128, 32, 16, 9),
    (2, 32, 32, 16, 32, 9),
    (2, 32, 32, 32, 64, 9),
    # conv_rows kernel (W % 128 == 0, C*2 bytes = one swizzle span): smem-resident rows, taps as shifted views
    (2, 32, 128, 64, 64, 9),       # 128-byte span, one segment per row
    (3, 64, 256, 64, 64, 9),       # two 128-pixel segments per row, strips cut at image ends
    (1, 128, 128, 64, 128, 9),     # N = 128 (ring of 5 rows)
    (2, 64, 256, 32, 32, 9),       # 64-byte span
    (1, 256, 512, 32, 64, 9),
    (1, 64, 1024, 16, 16, 9),      # 32-byte span, N tile 16
    (2, 128, 256, 16, 32, 9),
    (1, 4, 128, 64, 16, 9),        # shortest strip
    (2, 32, 128, 32, 16, 9),       # wgrad_rows: remaining (C, N) span combinations
    (1, 16, 128, 16, 64, 9),
    (1, 8, 256, 64, 32, 9),
    (5, 64, 128, 16, 16, 9),       # more strips than one per CTA range boundary
    (1, 8, 512, 32, 32, 9),        # conv_rows pixel groups: J = 2, N' = 64
    (2, 16, 256, 32, 16, 9),       # J = 2, N' = 32
    (1, 32, 512, 16, 16, 9),       # J = 4, N' = 64
]


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("backend", [0, 1])
@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv_fprop(shape, backend, dtype):
    from pggan_b200 import ops, _lib
    import pggan_b200 as pg
    B, H, W, Ci, Co, taps = shape
    if backend == 1 and B * H * W * Co * Ci * taps > 3e9:
        pytest.skip("SIMT cross-check kept to small shapes")
    pg.set_compute_dtype(dtype)
    _lib.load().pggan_set_conv_backend(backend)
    before = _lib.launch_counts()
    dev = "cuda"
    k = 3 if taps == 9 else 1
    x = rnd(torch.randn(B, Ci, H, W, device=dev), dtype)
    w = rnd(torch.randn(Co, Ci, k, k, device=dev), dtype)
    b = torch.randn(Co, device=dev)
    scale = (2.0 / (Ci * taps)) ** 0.5
    ref = F.leaky_relu((F.conv2d(x, w, b, padding=k // 2)) * scale, 0.2)
    if taps == 1:
        y = ops.ConvF.apply(to_feat(x.view(B, Ci), dtype), w.view(Co, Ci), b, scale, 0.2, False)
        ref = ref.view(B, Co)
    else:
        y = ops.ConvF.apply(to_feat(x, dtype), w, b, scale, 0.2, False)
    torch.cuda.synchronize()
    after = _lib.launch_counts()
    if backend == 0:
        assert after["tcgen05"] > before["tcgen05"], "tensor-core path did not run"
    else:
        assert after["simt"] > before["simt"]
    assert y.dtype == ops.get_compute_dtype()
    assert rel_l2(f32(y), ref) < tol(dtype), rel_l2(f32(y), ref)
    if taps == 9 and dtype != torch.float32 and backend == 0 and B * H * W * Co * Ci < 2e8:
        # precision follows the tensor: the same call on an fp32 tensor runs the 3xTF32 path and returns fp32
        y32 = ops.ConvF.apply(x.contiguous(memory_format=torch.channels_last), w, b, scale, 0.2, False)
        assert y32.dtype == torch.float32 and rel_l2(y32, ref) < _TOL[torch.float32]


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
@pytest.mark.parametrize("shape,force", [((16, 4, 4, 512, 512), None), ((16, 8, 8, 512, 512), None), ((4, 16, 16, 512, 256), None),
                                         ((3, 8, 8, 256, 128), "128,9"), ((16, 4, 4, 512, 512), "256,3"),
                                         ((5, 8, 8, 128, 64), "64,3"), ((32, 4, 4, 256, 256), "64,12")])
def test_conv_split_k(shape, force, dtype, monkeypatch):
    """few output tiles, long K: the K extent is split over otherwise idle SMs (pg_conv_params.ws_k), partial
    accumulators are summed in the workspace and the last CTA of a tile runs the (mask / bias / LeakyReLU / pixel-norm)
    epilogue — same result as the unsplit kernel, reproducible bits, workspace left at zero"""
    from pggan_b200 import ops, _lib
    import pggan_b200 as pg
    pg.set_compute_dtype(dtype)
    _lib.load().pggan_set_conv_backend(0)
    B, H, W, Ci, Co = shape
    if force and dtype == torch.float32 and (9 * 3 * Ci // 32) % int(force.split(",")[1]) != 0:
        pytest.skip("split count does not divide the 3xTF32 K extent")
    dev = "cuda"
    x = rnd(torch.randn(B, Ci, H, W, device=dev), dtype)
    w = rnd(torch.randn(Co, Ci, 3, 3, device=dev), dtype)
    b = torch.randn(Co, device=dev)
    m = rnd(torch.randn(B, Co, H, W, device=dev), dtype)
    scale = (2.0 / (Ci * 9)) ** 0.5
    xf, mf = to_feat(x, dtype), to_feat(m, dtype)
    wp = ops._pack(w, "f", None, None, ops._fmt_of(xf, "a"))

    def run(env):
        if env is None:
            monkeypatch.delenv("PGGAN_FORCE_SPLITK", raising=False)
        else:
            monkeypatch.setenv("PGGAN_FORCE_SPLITK", env)
        y = ops.k_conv(xf, wp, b, Co, 9, scale, 0.2, mask_src=mf, mask_slope=0.2)
        torch.cuda.synchronize()
        return y

    y1 = run(force)
    y2 = run(force)                                               # the workspace is back at zero, the sum order fixed
    assert torch.equal(y1, y2)
    y0 = run("0,1")                                               # one K range: the unsplit kernel
    ref = F.leaky_relu(F.conv2d(x, w, b, padding=1) * scale, 0.2)
    ref = torch.where(m > 0, ref, ref * 0.2)
    assert rel_l2(f32(y1), ref) < tol(dtype), rel_l2(f32(y1), ref)
    assert rel_l2(f32(y1), f32(y0)) < (1e-4 if dtype == torch.float32 else 1e-3)
    ws = ops._splitk_ws(xf.device)
    assert int(ws.view(torch.int32).abs().max()) == 0
    if Co <= 256 and dtype != torch.float32:
        # fused pixel norm on a split layer (two passes over the summed partials)
        monkeypatch.delenv("PGGAN_FORCE_SPLITK", raising=False)
        yp, r = ops.k_conv(xf, wp, b, Co, 9, scale, 0.2, pn_eps=1e-8) if ops.conv_pn_fusable(xf, Co, 9) else (None, None)
        if yp is not None:
            a = F.leaky_relu(F.conv2d(x, w, b, padding=1) * scale, 0.2)
            refp = a * torch.rsqrt((a ** 2).mean(dim=1, keepdim=True) + 1e-8)
            assert rel_l2(f32(yp), refp) < tol(dtype), rel_l2(f32(yp), refp)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("shape", [(2, 16, 16, 64, 128), (1, 32, 128, 64, 64), (1, 16, 256, 16, 32), (2, 32, 32, 128, 64)])
def test_conv_mask_and_raw_out(shape, dtype):
    """epilogue mask (dgrad fused with a LeakyReLU backward) and fp32 raw output on the plain, rows and swapped kernels"""
    from pggan_b200 import ops
    import pggan_b200 as pg
    pg.set_compute_dtype(dtype)
    B, H, W, Ci, Co = shape
    x = rnd(torch.randn(B, Ci, H, W, device="cuda"), dtype)
    w = rnd(torch.randn(Co, Ci, 3, 3, device="cuda"), dtype)
    ysrc = torch.randn(B, Co, H, W, device="cuda")
    wp = ops._pack(w, "f", dtype)
    y = ops.k_conv(to_feat(x, dtype), wp, None, Co, 9, 0.05, 1.0, mask_src=to_feat(ysrc, dtype), mask_slope=0.2)
    ref = F.conv2d(x, w, None, padding=1) * 0.05
    ref = torch.where(f32(to_feat(ysrc, dtype)) > 0, ref, ref * 0.2)
    assert rel_l2(f32(y), ref) < tol(dtype)
    yr = ops.k_conv(to_feat(x, dtype), wp, None, Co, 9, 0.05, 1.0, raw_out=True)
    assert yr.dtype == torch.float32
    assert rel_l2(yr, F.conv2d(x, w, None, padding=1) * 0.05) < 1e-4


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("backend", [0, 1])
@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv_dgrad_wgrad(shape, backend, dtype):
    """first-order backward of ConvF = ConvT (dgrad) + ConvW (wgrad) + ColSum (bias)"""
    from pggan_b200 import ops, _lib
    import pggan_b200 as pg
    B, H, W, Ci, Co, taps = shape
    if backend == 1 and B * H * W * Co * Ci * taps > 1e9:
        pytest.skip("SIMT cross-check kept to small shapes")
    pg.set_compute_dtype(dtype)
    _lib.load().pggan_set_conv_backend(backend)
    dev = "cuda"
    k = 3 if taps == 9 else 1
    x = rnd(torch.randn(B, Ci, H, W, device=dev), dtype)
    w = rnd(torch.randn(Co, Ci, k, k, device=dev), dtype)
    b = torch.randn(Co, device=dev)
    gy = rnd(torch.randn(B, Co, H, W, device=dev), dtype, "g")
    scale = (2.0 / (Ci * taps)) ** 0.5
    xr, wr, br = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    (F.conv2d(xr, wr, br, padding=k // 2) * scale).backward(gy)
    xm, wm, bm = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    if taps == 1:
        y = ops.ConvF.apply(to_feat(xm.view(B, Ci), dtype), wm.view(Co, Ci), bm, scale, 1.0, False)
        y.backward(to_feat(gy.view(B, Co), dtype, "g"))
    else:
        y = ops.ConvF.apply(to_feat(xm, dtype), wm, bm, scale, 1.0, False)
        y.backward(to_feat(gy, dtype, "g"))
    assert rel_l2(xm.grad, xr.grad) < TOL[dtype], ("dgrad", rel_l2(xm.grad, xr.grad))
    # (mixed mode: the fp16 activation operand of the weight-gradient GEMM is re-rounded to bf16 — both tensor-core
    # operands must share one format — so the weight gradient carries bf16's input rounding)
    assert rel_l2(wm.grad, wr.grad) < (4e-3 if dtype == "mixed" else 2e-4), ("wgrad", rel_l2(wm.grad, wr.grad))
    assert rel_l2(bm.grad, br.grad) < 1e-4, ("bias", rel_l2(bm.grad, br.grad))


@pytest.mark.parametrize("dtype", DT)
def test_conv_double_backward(dtype):
    """grad-of-grad through ConvF -> MaskMul -> ConvT, as R1 / WGAN-GP use it"""
    from pggan_b200 import ops
    import pggan_b200 as pg
    pg.set_compute_dtype(dtype)
    B, H, W, Ci, Co = 2, 8, 8, 64, 64
    x = rnd(torch.randn(B, Ci, H, W, device="cuda"), dtype)
    w = rnd(torch.randn(Co, Ci, 3, 3, device="cuda") * 0.1, dtype)
    b = torch.randn(Co, device="cuda") * 0.1

    def run(fn):
        xx, ww, bb = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
        y = fn(xx, ww, bb)
        (g,) = torch.autograd.grad(y.sum(), xx, create_graph=True)
        (g.float() ** 2).sum().backward()
        return g.detach().float(), ww.grad, xx.grad

    g_ref, gw_ref, _ = run(lambda xx, ww, bb: F.leaky_relu(F.conv2d(xx, ww, bb, padding=1) * 0.1, 0.2))
    g_my, gw_my, _ = run(lambda xx, ww, bb: ops.ExportFloat.apply(ops.ConvF.apply(to_feat(xx, dtype), ww, bb, 0.1, 0.2, False)))
    assert rel_l2(g_my, g_ref) < 2 * TOL[dtype]
    assert rel_l2(gw_my, gw_ref) < 4 * TOL[dtype], rel_l2(gw_my, gw_ref)


@pytest.mark.parametrize("dtype", DT)
def test_elementwise(dtype):
    from pggan_b200 import ops
    import pggan_b200 as pg
    pg.set_compute_dtype(dtype)
    B, C, H, W = 3, 64, 8, 16
    x = to_feat(rnd(torch.randn(B, C, H, W, device="cuda"), dtype), dtype)            # activation class
    y = to_feat(rnd(torch.randn(B, C, H, W, device="cuda"), dtype), dtype)
    g = to_feat(rnd(torch.randn(B, C, H, W, device="cuda"), dtype, "g"), dtype, "g")  # gradient class
    # mask mul: gradient masked by the sign of an activation
    out = ops.k_mask_mul(g, y, 0.2)
    ref = torch.where(f32(y) > 0, f32(g, "g"), f32(g, "g") * 0.2)
    assert rel_l2(f32(out, "g"), ref) < tol(dtype, "g")
    # fused mask + column sum
    out2, cs2 = ops.k_mask_colsum(g, y, 0.2, True, 0.5)
    assert torch.equal(out2, out)
    assert rel_l2(cs2, 0.5 * f32(out, "g").sum(dim=(0, 2, 3))) < 1e-4
    # upsample: exact indexing (bit copies, either class)
    up = ops.k_upsample(x, 1.0)
    assert torch.equal(up, x.repeat_interleave(2, 2).repeat_interleave(2, 3))
    assert up.is_contiguous(memory_format=torch.channels_last)
    upm = ops.k_upsample(g, 0.25, mask_src=up, slope=0.2, cls="g")
    gm = f32(g, "g").repeat_interleave(2, 2).repeat_interleave(2, 3) * 0.25
    assert rel_l2(f32(upm, "g"), torch.where(f32(up) > 0, gm, gm * 0.2)) < tol(dtype, "g")
    # ... and with the column sums of the result taken in the same pass (bias gradient of the conv above the pool)
    cs3 = torch.zeros(C, device="cuda")
    upm2 = ops.k_upsample(g, 0.25, mask_src=up, slope=0.2, cls="g", colsum=cs3, colsum_scale=0.5)
    assert torch.equal(upm2, upm)
    assert rel_l2(cs3, 0.5 * f32(upm, "g").sum(dim=(0, 2, 3))) < 1e-5
    # pool
    pl = ops.k_pool(x, 0.25)
    assert rel_l2(f32(pl), F.avg_pool2d(f32(x), 2)) < tol(dtype)
    low = to_feat(rnd(torch.randn(B, C, H // 2, W // 2, device="cuda"), dtype), dtype)
    pb = ops.k_pool(x, 0.25, blend_with=low, alpha=0.3)
    assert rel_l2(f32(pb), 0.7 * f32(low) + 0.3 * F.avg_pool2d(f32(x), 2)) < tol(dtype)
    # the same with alpha as a device scalar (what a captured CUDA graph reads)
    adev = torch.full((1,), 0.3, device="cuda")
    pb2 = ops.k_pool(x, 0.25, blend_with=low, alpha=ops.AlphaRef(0.3, adev))
    assert torch.equal(pb2, pb)
    adev.fill_(0.9)
    pb3 = ops.k_pool(x, 0.25, blend_with=low, alpha=ops.AlphaRef(0.3, adev))      # the DEVICE value wins
    assert rel_l2(f32(pb3), 0.1 * f32(low) + 0.9 * F.avg_pool2d(f32(x), 2)) < tol(dtype)
    ups = ops.k_upsample(g, ops._affine(0.0, 0.25, ops.AlphaRef(0.3, adev)), cls="g")
    assert rel_l2(f32(ups, "g"), f32(g, "g").repeat_interleave(2, 2).repeat_interleave(2, 3) * 0.225) < tol(dtype, "g")
    # axpby
    ab = ops.k_axpby(x, y, 0.25, 0.75)
    assert rel_l2(f32(ab), 0.25 * f32(x) + 0.75 * f32(y)) < tol(dtype)
    ab2 = ops.k_axpby(g, None, ops._affine(1.0, -1.0, ops.AlphaRef(0.0, adev)), 0.0, cls="g")
    assert rel_l2(f32(ab2, "g"), 0.1 * f32(g, "g")) < tol(dtype, "g")
    # colsum
    cs = ops.k_colsum(g, 0.5)
    assert rel_l2(cs, 0.5 * f32(g, "g").sum(dim=(0, 2, 3))) < 1e-5
    # Upsample/Pool autograd pair
    xr = f32(x).clone().requires_grad_()
    (ops.ExportFloat.apply(ops.Upsample2x.apply(to_feat(xr, dtype), 1.0)) * f32(up)).sum().backward()
    assert rel_l2(xr.grad, 4 * F.avg_pool2d(f32(up), 2)) < 2 * tol(dtype, "g")


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
@pytest.mark.parametrize("shape", [(4, 16, 16, 128, 256, 9), (16, 4, 4, 512, 512, 9), (2, 32, 128, 64, 64, 9),
                                   (1, 16, 256, 32, 16, 9), (8, 1, 1, 8192, 512, 1), (3, 8, 8, 64, 128, 9)])
def test_wgrad_accumulates_straight_into_the_gradient(shape, dtype, monkeypatch):
    """a weight gradient whose layout is the GEMM's own ([N][taps][C]: channels_last conv weight, plain linear weight)
    is accumulated by the kernel's epilogue (pg_wgrad_params.direct): same numbers as slices + un-pack, added on top
    of what the gradient held"""
    from pggan_b200 import ops
    import pggan_b200 as pg
    pg.set_compute_dtype(dtype)
    B, H, W, Ci, Co, taps = shape
    if dtype == torch.float32 and B * H * W * Ci * Co > 3e8:
        pytest.skip("fp32 cross-check kept to small shapes")
    dev = "cuda"
    k = 3 if taps == 9 else 1
    torch.manual_seed(5)
    x = rnd(torch.randn(B, Ci, H, W, device=dev), dtype)
    g = rnd(torch.randn(B, Co, H, W, device=dev), dtype, "g")
    if taps == 1:
        xf, gf = to_feat(x.view(B, Ci), dtype), to_feat(g.view(B, Co), dtype, "g")
        wshape = (Co, Ci)
        grad0 = torch.randn(Co, Ci, device=dev)
        ref = g.view(B, Co).t() @ x.view(B, Ci)
    else:
        xf, gf = to_feat(x, dtype), to_feat(g, dtype, "g")
        wshape = (Co, Ci, k, k)
        grad0 = torch.randn(Co, Ci, k, k, device=dev).contiguous(memory_format=torch.channels_last)
        ref = torch.nn.grad.conv2d_weight(x, wshape, g, padding=1)
    outs = []
    for direct in (True, False):
        monkeypatch.setattr(ops, "DIRECT_WGRAD", direct)
        acc = grad0.clone(memory_format=torch.preserve_format)
        before = _lib_other_launches()
        ops.k_wgrad(gf, xf, 0.25, wshape, None, accumulate_into=acc)
        torch.cuda.synchronize()
        if direct and dtype != torch.float32:            # (fp32: the 3xTF32 operand splits are counted there too)
            assert _lib_other_launches() == before, "no un-pack launch on the direct path"
        outs.append(acc)
    want = grad0 + 0.25 * ref
    assert rel_l2(outs[0], want) < (2e-4 if dtype == torch.float32 else 3e-3), rel_l2(outs[0], want)
    assert rel_l2(outs[0], outs[1]) < 1e-5, rel_l2(outs[0], outs[1])


def _emulate_map(n_elems, args, device):
    """index tensor [n_total, taps, c_total] into the flat parameter (-1 = zero padding) of a pg_pack_map tuple"""
    n_total, n_valid, n2, taps, c_total, c_valid, c2, base, s_n1, s_n2, s_t, s_c1, s_c2 = args
    n = torch.arange(n_total, device=device).view(-1, 1, 1)
    t = torch.arange(taps, device=device).view(1, -1, 1)
    c = torch.arange(c_total, device=device).view(1, 1, -1)
    off = base + (n // n2) * s_n1 + (n % n2) * s_n2 + t * s_t + (c // c2) * s_c1 + (c % c2) * s_c2
    valid = ((n < n_valid) & (c < c_valid)).expand(n_total, taps, c_total)
    return torch.where(valid, off.expand(n_total, taps, c_total), torch.full_like(off, -1).expand(n_total, taps, c_total))


@pytest.mark.parametrize("shape,layout", [((512, 8192), ("lin_in_perm", 512)), ((8192, 512), ("lin_out_perm", 512)),
                                          ((512, 2048), ("lin_in_perm", 128)), ((256, 384), None),
                                          ((96, 33, 3, 3), ("pad_c", 64)), ((1, 512), None)])
def test_pack_and_unpack_through_the_layout_maps(shape, layout):
    """operand packing / gradient un-packing of the permuting layouts (row-span kernels for the big linear layers, the
    generic kernel for the rest) against the index map evaluated in torch: bit-exact copies"""
    from pggan_b200 import ops, _lib as L
    import pggan_b200 as pg
    pg.set_compute_dtype(torch.float16)
    torch.manual_seed(3)
    W = torch.nn.Parameter(torch.randn(*shape, device="cuda"))
    N, C, T, fm, tm = ops._maps(W.shape, layout)
    flat = W.detach().reshape(-1)
    for kind, args in (("f", fm), ("t", tm)):
        idx = _emulate_map(flat.numel(), args, "cuda")
        want = torch.where(idx >= 0, flat[idx.clamp(min=0)], torch.zeros((), device="cuda")).to(torch.float16)
        got = ops._pack(W, kind, None, layout, L.PG_F16)
        assert torch.equal(got.reshape(want.shape), want), (kind, layout)
    # gradient slices [S][N][taps][C] -> the parameter's layout, overwrite and accumulate
    idx = _emulate_map(flat.numel(), fm, "cuda")
    slices = 3
    acc = torch.randn(slices, *idx.shape, device="cuda")
    want = torch.zeros_like(flat)
    want[idx[idx >= 0]] = (acc.sum(0) * 0.5)[idx >= 0]
    m = ops._packmap(fm)
    import ctypes
    out = torch.full_like(flat, 7.0) if not (fm[1] < fm[0] or fm[5] < fm[4]) else torch.zeros_like(flat)
    L.check(L.load().pggan_unpack_wgrad(acc.data_ptr(), out.data_ptr(), ctypes.byref(m), slices, 0.5, 0,
                                        torch.cuda.current_stream().cuda_stream), "unpack")
    assert rel_l2(out, want) < 1e-6
    L.check(L.load().pggan_unpack_wgrad(acc.data_ptr(), out.data_ptr(), ctypes.byref(m), slices, 0.5, 1,
                                        torch.cuda.current_stream().cuda_stream), "unpack")
    assert rel_l2(out, 2 * want) < 1e-6


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("C", [16, 64, 512])
def test_pixelnorm(C, dtype):
    from pggan_b200 import ops
    import pggan_b200 as pg
    pg.set_compute_dtype(dtype)
    x = rnd(torch.randn(4, C, 8, 8, device="cuda"), dtype)
    xr = x.clone().requires_grad_()
    ref = xr * torch.rsqrt((xr ** 2).mean(dim=1, keepdim=True) + 1e-8)
    g = rnd(torch.randn_like(x), dtype, "g")
    ref.backward(g)
    xm = x.clone().requires_grad_()
    y = ops.PixelNorm.apply(to_feat(xm, dtype), 1e-8)
    y.backward(to_feat(g, dtype, "g"))
    assert rel_l2(f32(y), ref) < tol(dtype)
    assert rel_l2(xm.grad, xr.grad) < 3 * tol(dtype, "g"), rel_l2(xm.grad, xr.grad)
    # 2-D latent path
    z = torch.randn(6, 512, device="cuda")
    assert rel_l2(f32(ops.PixelNorm.apply(z, 1e-8)), z * torch.rsqrt((z ** 2).mean(1, keepdim=True) + 1e-8)) < tol(dtype)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("C", [16, 128])
def test_rgb_ops(C, dtype):
    from pggan_b200 import ops
    import pggan_b200 as pg
    pg.set_compute_dtype(dtype)
    B, R = 3, 16
    img = torch.rand(B, 3, R, R, device="cuda") * 2 - 1
    Wf = torch.randn(C, 3, device="cuda")
    bf = torch.randn(C, device="cuda")
    for pool in (False, True):
        ir, wr, br = img.clone().requires_grad_(), Wf.clone().requires_grad_(), bf.clone().requires_grad_()
        src = F.avg_pool2d(ir, 2) if pool else ir
        ref = F.leaky_relu(F.conv2d(src, wr.view(C, 3, 1, 1), br) * 0.7, 0.2)
        g = rnd(torch.randn_like(ref), dtype, "g")
        ref.backward(g)
        im, wm, bm = img.clone().requires_grad_(), Wf.clone().requires_grad_(), bf.clone().requires_grad_()
        y = ops.RgbExpand.apply(im, wm, bm, pool, 0.7, 0.2)
        y.backward(to_feat(g, dtype, "g"))
        assert rel_l2(f32(y), ref) < tol(dtype)
        assert rel_l2(im.grad, ir.grad) < 3 * tol(dtype, "g"), ("expand gimg", pool, rel_l2(im.grad, ir.grad))
        assert rel_l2(wm.grad, wr.grad) < 3 * tol(dtype, "g"), ("expand gW", pool, rel_l2(wm.grad, wr.grad))
        assert rel_l2(bm.grad, br.grad) < 3 * tol(dtype, "g")
    x = rnd(torch.randn(B, C, R, R, device="cuda"), dtype)
    Wt = torch.randn(3, C, device="cuda")
    bt = torch.randn(3, device="cuda")
    for up in (False, True):
        xr, wr, br = x.clone().requires_grad_(), Wt.clone().requires_grad_(), bt.clone().requires_grad_()
        ref = F.conv2d(xr, wr.view(3, C, 1, 1), br) * 0.3
        if up:
            ref = ref.repeat_interleave(2, 2).repeat_interleave(2, 3)
        g = torch.randn_like(ref)
        ref.backward(g)
        xm, wm, bm = x.clone().requires_grad_(), Wt.clone().requires_grad_(), bt.clone().requires_grad_()
        y = ops.RgbReduce.apply(to_feat(xm, dtype), wm.t(), bm, up, 0.3, 1.0)
        y.backward(g)
        assert y.dtype == torch.float32 and rel_l2(y, ref) < 1e-5
        assert rel_l2(xm.grad, xr.grad) < 2 * tol(dtype, "g"), ("reduce gx", up)
        assert rel_l2(wm.grad, wr.grad) < 2 * tol(dtype, "g"), ("reduce gW", up, rel_l2(wm.grad, wr.grad))
        assert rel_l2(bm.grad, br.grad) < 1e-4
    # fades: exact fp32 formulas, alpha as a float and as a device scalar; the backward through the fade
    low = torch.randn(B, 3, R // 2, R // 2, device="cuda")
    out = ops.ImageFade.apply(img, low, 0.3)
    assert rel_l2(out, 0.7 * low.repeat_interleave(2, 2).repeat_interleave(2, 3) + 0.3 * img) < 1e-6
    adev = torch.full((1,), 0.3, device="cuda")
    assert torch.equal(ops.ImageFade.apply(img, low, ops.AlphaRef(0.3, adev)), out)
    out = ops.ImageFade.apply(img, None, 0.3)
    ref = 0.7 * F.interpolate(F.avg_pool2d(img, 2), scale_factor=2, mode="nearest") + 0.3 * img
    assert rel_l2(out, ref) < 1e-6
    for with_low in (True, False):
        ir, lr = img.clone().requires_grad_(), low.clone().requires_grad_()
        lo_ref = lr if with_low else F.avg_pool2d(ir, 2)
        ref = 0.7 * F.interpolate(lo_ref, scale_factor=2, mode="nearest") + 0.3 * ir
        gg = torch.randn_like(ref)
        ref.backward(gg)
        im, lm = img.clone().requires_grad_(), low.clone().requires_grad_()
        ops.ImageFade.apply(im, lm if with_low else None, ops.AlphaRef(0.3, adev)).backward(gg)
        assert rel_l2(im.grad, ir.grad) < 1e-6
        if with_low:
            assert rel_l2(lm.grad, lr.grad) < 1e-6


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("B,group", [(8, 4), (4, 4), (6, 6), (3, 3)])
def test_mbstd(B, group, dtype):
    from pggan_b200 import ops
    import pggan_b200 as pg
    pg.set_compute_dtype(dtype)
    C = 64
    x = rnd(torch.randn(B, C, 4, 4, device="cuda"), dtype)

    def ref_stat(t):
        y = t.reshape(B // group, group, -1)
        return torch.sqrt(y.var(dim=1, unbiased=True) + 1e-8).mean(dim=1)

    xr = x.clone().requires_grad_()
    wgt = torch.randn(B // group, device="cuda")
    s_ref = ref_stat(xr)
    (g_ref,) = torch.autograd.grad((s_ref * wgt).sum(), xr, create_graph=True)
    u = torch.randn_like(x)
    (g_ref * u).sum().backward()
    xm = x.clone().requires_grad_()
    s = ops.MbstdStat.apply(to_feat(xm, dtype), group)
    (g_my,) = torch.autograd.grad((s * wgt).sum(), xm, create_graph=True)
    (g_my.float() * u).sum().backward()
    assert rel_l2(s, s_ref) < 1e-4
    assert rel_l2(g_my.float(), g_ref) < 3 * tol(dtype, "g")
    assert rel_l2(xm.grad, xr.grad) < 6 * tol(dtype, "g"), rel_l2(xm.grad, xr.grad)


def test_mbstd_single_sample_and_channel():
    """B == 1: the extra channel is all zeros (reference lib/blocks.py:232-233); B % 4 != 0: one group of B"""
    from pggan_b200.blocks import concatenate_stddev_channel
    from pggan_b200 import ops
    import pggan_b200 as pg
    pg.set_compute_dtype(torch.float32)
    x = torch.randn(1, 64, 4, 4, device="cuda")
    y = concatenate_stddev_channel(x)
    assert y.shape == (1, 65, 4, 4) and float(y[:, 64].abs().max()) == 0.0 and torch.equal(y[:, :64], x)
    x = torch.randn(6, 64, 4, 4, device="cuda")
    y = concatenate_stddev_channel(x)
    ref = torch.sqrt(x.reshape(1, 6, -1).var(dim=1, unbiased=True) + 1e-8).mean()
    assert rel_l2(y[:, 64], ref.expand(6, 4, 4)) < 1e-5
    pg.set_compute_dtype("mixed")
    y = concatenate_stddev_channel(ops.enter(x))     # exported as an honest fp16 tensor in mixed mode
    assert y.dtype == torch.float16 and rel_l2(y[:, 64].float(), ref.expand(6, 4, 4)) < 1e-3


def test_gp_reduce_and_r1():
    from pggan_b200 import ops
    B, n = 5, 3 * 32 * 32
    g = torch.randn(B, 3, 32, 32, device="cuda") * 0.02
    pred = torch.randn(B, device="cuda")
    gr = g.clone().requires_grad_()
    norms = (gr.view(B, -1) ** 2).sum(1).sqrt()
    ref = ((norms - 1) ** 2).sum() * 10.0
    ref.backward()
    gm = g.clone().requires_grad_()
    res = ops.GradientPenalty.apply(gm, pred, 10.0, 0.001)
    res[0].backward()
    assert rel_l2(res[0], ref) < 1e-5
    assert rel_l2(res[1], (pred ** 2).sum() * 0.001) < 1e-5
    assert rel_l2(gm.grad, gr.grad) < 1e-5
    gm2 = g.clone().requires_grad_()
    (0.5 * ops.RowsSumSq.apply(gm2).mean(0)).backward()
    assert rel_l2(gm2.grad, g / B) < 1e-5
    # odd row length (scalar path)
    h = torch.randn(3, 1001, device="cuda")
    assert rel_l2(ops.RowsSumSq.apply(h), (h ** 2).sum(1)) < 1e-5


def _lib_other_launches():
    from pggan_b200 import _lib
    return _lib.launch_counts()["other"]


def test_fused_adam_matches_torch():
    from pggan_b200 import FusedAdam
    torch.manual_seed(1)
    from pggan_b200 import ops
    import pggan_b200 as pg
    pg.set_compute_dtype(torch.float16)
    shapes = [(512, 512, 3, 3), (512,), (3, 64, 1, 1), (1, 512), (1,), (70000,), (64, 32, 3, 3)]
    mine = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    # 3x3 conv weights live in channels_last memory ([O][kh][kw][I]) as layers.py keeps them
    mine[0].data = mine[0].data.contiguous(memory_format=torch.channels_last)
    mine[6].data = mine[6].data.contiguous(memory_format=torch.channels_last)
    assert ops.is_ohwi(mine[0]) and not ops.is_ohwi(mine[2])
    ref = [torch.nn.Parameter(p.detach().clone().contiguous()) for p in mine]
    opt_m = FusedAdam(mine, lr=1e-3, betas=(0.0, 0.99))
    opt_r = torch.optim.Adam(ref, lr=1e-3, betas=(0.0, 0.99))
    for step in range(4):
        opt_m.zero_grad()
        opt_r.zero_grad()
        for pm, pr in zip(mine, ref):
            g = torch.randn_like(pm)
            pm.grad.add_(g)
            pr.grad = g.clone()
        opt_m.step()
        opt_r.step()
    for pm, pr in zip(mine, ref):
        assert rel_l2(pm, pr) < 1e-6
    assert int(opt_m.step_count.item()) == 4
    # the Adam kernel maintains the packed fp16 fprop operand of the channels_last conv weights (no pack kernel)
    for k in (0, 6):
        before = _lib_other_launches()
        wp = ops._pack(mine[k], "f")
        assert _lib_other_launches() == before, "the shadow copy must be served without a pack launch"
        O_, I_ = mine[k].shape[:2]
        assert torch.equal(wp, mine[k].detach().permute(0, 2, 3, 1).reshape(O_, 9 * I_).to(torch.float16))
        # ... the dgrad operand [i][8-t][o] comes from the tiled transpose kernel
        wt = ops._pack(mine[k], "t", fmt=ops._fmt("g"))
        assert torch.equal(wt, mine[k].detach().flip(2, 3).permute(1, 2, 3, 0).reshape(I_, 9 * O_).to(torch.float16))
    mine[6].data.mul_(1.0)                              # a version bump (e.g. load_state_dict) invalidates the shadow
    assert torch.equal(ops._pack(mine[6], "f"), mine[6].detach().permute(0, 2, 3, 1).reshape(64, 9 * 32).to(torch.float16))
    # optimizer state in torch.optim.Adam's wire format, both directions (reference pggan/model.py:97-101)
    sd_m, sd_r = opt_m.state_dict(), opt_r.state_dict()
    assert sorted(sd_m["state"].keys()) == sorted(sd_r["state"].keys())
    for k in sd_r["state"]:
        assert float(sd_m["state"][k]["step"]) == float(sd_r["state"][k]["step"])
        assert rel_l2(sd_m["state"][k]["exp_avg_sq"], sd_r["state"][k]["exp_avg_sq"]) < 1e-6
    opt_r2 = torch.optim.Adam(ref, lr=1e-3, betas=(0.0, 0.99))
    opt_r2.load_state_dict(sd_m)                                   # torch reads what FusedAdam wrote
    opt_m2 = FusedAdam(mine, lr=5e-4, betas=(0.5, 0.9))
    opt_m2.load_state_dict(sd_r)                                   # FusedAdam reads what torch wrote
    assert opt_m2.lr == 1e-3 and opt_m2.betas == (0.0, 0.99) and int(opt_m2.step_count.item()) == 4
    for o_m, o_r in ((opt_m2, opt_r2),):
        o_m.zero_grad()
        o_r.zero_grad()
        for pm, pr in zip(mine, ref):
            g = torch.randn_like(pm)
            pm.grad.add_(g)
            pr.grad = g.clone()
        o_m.step()
        o_r.step()
    for pm, pr in zip(mine, ref):
        assert rel_l2(pm, pr) < 1e-6
    with pytest.raises(ValueError):
        opt_m2.load_state_dict({"exp_avg": None})


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("shape", [(1, 32, 128, 64, 64), (1, 16, 512, 16, 16), (1, 16, 256, 32, 32), (1, 8, 256, 32, 16),
                                   (2, 16, 16, 128, 256), (2, 32, 32, 64, 128), (1, 64, 128, 16, 32)])
def test_conv_fused_pixelnorm(shape, dtype):
    """conv + He scale + bias + LeakyReLU + pixel norm in one kernel (epilogue thread owns the pixel): rows kernel with
    J = 1 / 2 / 4 pixel groups and the plain kernel with one N tile, against PyTorch; ConvActPN uses it where
    pggan_conv_pn_fusable says so and the two-kernel path elsewhere (e.g. the swapped-orientation layers)."""
    from pggan_b200 import ops
    import pggan_b200 as pg
    pg.set_compute_dtype(dtype)
    B, H, W, Ci, Co = shape
    x = rnd(torch.randn(B, Ci, H, W, device="cuda"), dtype)
    w = rnd(torch.randn(Co, Ci, 3, 3, device="cuda"), dtype)
    b = torch.randn(Co, device="cuda")
    scale = (2.0 / (Ci * 9)) ** 0.5
    a = F.leaky_relu(F.conv2d(x, w, b, padding=1) * scale, 0.2)
    ref = a * torch.rsqrt((a ** 2).mean(dim=1, keepdim=True) + 1e-8)
    xf = to_feat(x, dtype)
    if dtype != torch.float32:          # fp32 (3xTF32) triples K: the 16-/32-channel shapes leave the rows envelope
        assert ops.conv_pn_fusable(xf, Co, 9)
    if ops.conv_pn_fusable(xf, Co, 9):
        y, r = ops.k_conv(xf, ops._pack(w, "f", dtype), b, Co, 9, scale, 0.2, pn_eps=1e-8)
        r_ref = torch.rsqrt((a ** 2).mean(dim=1) + 1e-8).reshape(-1)
        assert rel_l2(f32(y), ref) < tol(dtype), rel_l2(f32(y), ref)
        assert rel_l2(r, r_ref) < 1e-3
    # the autograd Function (either path) against PyTorch, forward and input gradient
    xm = x.clone().requires_grad_()
    wm = torch.nn.Parameter(w.clone())
    ym = ops.ConvActPN.apply(to_feat(xm, dtype), wm, b, scale, 0.2, 1e-8)
    assert rel_l2(f32(ym), ref) < tol(dtype)
    assert not ops.conv_pn_fusable(to_feat(rnd(torch.randn(2, 256, 32, 32, device="cuda"), dtype), dtype), 128, 9) \
        or dtype == torch.float32, "256->128 (K = 2304) runs on the swapped kernel: must not be reported fusable"


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("shape", [(2, 8, 8, 64, 64), (3, 4, 4, 512, 512), (2, 16, 16, 512, 256), (1, 64, 64, 256, 128),
                                   (2, 64, 64, 128, 64), (5, 32, 32, 128, 128), (1, 128, 128, 128, 64)])
def test_conv_fused_upsample(shape, dtype, monkeypatch):
    """conv3x3(upscale2d(x)) as ONE kernel (the up-sampling is the GEMM's prologue: parity tiles reading shifted boxes
    of the low-resolution tensor) against F.conv2d on the materialised up-sampled tensor; then G's whole
    upscale -> conv -> LeakyReLU -> pixel norm chain (ops.ConvActPN with up2) with its gradients."""
    from pggan_b200 import ops
    import pggan_b200 as pg
    pg.set_compute_dtype(dtype)
    B, H, W, Ci, Co = shape                      # H, W: LOW resolution
    x = rnd(torch.randn(B, Ci, H, W, device="cuda"), dtype)
    w = rnd(torch.randn(Co, Ci, 3, 3, device="cuda"), dtype)
    b = torch.randn(Co, device="cuda")
    scale = (2.0 / (Ci * 9)) ** 0.5
    xf = to_feat(x, dtype)
    up = x.repeat_interleave(2, 2).repeat_interleave(2, 3)
    ref = F.leaky_relu(F.conv2d(up, w, b, padding=1) * scale, 0.2)
    # the fused call is offered where the plain tensor-core kernel takes the unfused layer (small N with long K goes to
    # the swapped-orientation kernel, which has no up-sampling prologue: ConvActPN then up-samples separately)
    if dtype not in (torch.float32,) and shape in [(2, 8, 8, 64, 64), (3, 4, 4, 512, 512), (2, 16, 16, 512, 256), (5, 32, 32, 128, 128)]:
        assert ops.conv_up2_fusable(xf, Co)
    if ops.conv_up2_fusable(xf, Co):
        y = ops.k_conv(xf, ops._pack(w, "f", fmt=ops._fmt_of(xf, "a")), b, Co, 9, scale, 0.2, up2=True)
        assert y.shape == (B, Co, 2 * H, 2 * W) and y.is_contiguous(memory_format=torch.channels_last)
        assert rel_l2(f32(y), ref) < tol(dtype), rel_l2(f32(y), ref)
        # the unfused pair gives the same bits (same MMAs in the same order on the same operands; K kept in one range)
        monkeypatch.setenv("PGGAN_FORCE_SPLITK", "0,1")
        y2 = ops.k_conv(ops.k_upsample(xf, 1.0), ops._pack(w, "f", fmt=ops._fmt_of(xf, "a")), b, Co, 9, scale, 0.2)
        assert torch.equal(y, y2)
    # the whole G-block chain and its backward
    xr, wr, br = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    a = F.leaky_relu(F.conv2d(xr.repeat_interleave(2, 2).repeat_interleave(2, 3), wr, br, padding=1) * scale, 0.2)
    r = a * torch.rsqrt((a ** 2).mean(dim=1, keepdim=True) + 1e-8)
    gy = rnd(torch.randn_like(r), dtype, "g")
    r.backward(gy)
    xm, wm, bm = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    ym = ops.ConvActPN.apply(to_feat(xm, dtype), wm, bm, scale, 0.2, 1e-8, None, True)
    ym.backward(to_feat(gy, dtype, "g"))
    assert rel_l2(f32(ym), r) < tol(dtype)
    # (gradients pass a LeakyReLU mask: a forward difference of 1e-5 flips a few signs, which costs ~3e-3 in relative L2)
    gt = max(4 * tol(dtype, "g"), 5e-3)
    assert rel_l2(xm.grad, xr.grad) < gt, rel_l2(xm.grad, xr.grad)
    assert rel_l2(wm.grad, wr.grad) < gt, rel_l2(wm.grad, wr.grad)
    assert rel_l2(bm.grad, br.grad) < gt
