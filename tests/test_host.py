"""Host-side logic that needs no GPU: module tree / state_dict keys, the C-ABI library loads and
exports every declared symbol, schedule bookkeeping."""
import os
import re

import pytest
import torch

from oracle import pggan_oracle as O


def test_state_dict_keys_match_reference_layout():
    """Appendix C of SURVEY.md: key names/shapes of the reference at scale 3"""
    import pggan_b200 as pg
    G = pg.Generator(512, 512)
    D = pg.Discriminator(512, apply_minibatch_norm=True)
    for d in (512, 512, 256):
        G.add_block(d)
        D.add_block(d)
    gk, dk = G.state_dict(), D.state_dict()
    assert gk["latent_format_layer.module.weight"].shape == (8192, 512)
    assert gk["first_block.block.0.module.weight"].shape == (512, 512, 3, 3)
    assert gk["blocks.2.block.0.module.weight"].shape == (256, 512, 3, 3)
    assert gk["blocks.2.block.3.module.bias"].shape == (256,)
    assert gk["toRGB_blocks.3.toRGB.module.weight"].shape == (3, 256, 1, 1)
    assert dk["fromRGB_blocks.3.fromRGB.module.weight"].shape == (256, 3, 1, 1)
    assert dk["blocks.2.block.0.module.weight"].shape == (256, 256, 3, 3)
    assert dk["blocks.2.block.2.module.weight"].shape == (512, 256, 3, 3)
    assert dk["minibatch_normalization_block.conv.module.weight"].shape == (512, 513, 3, 3)
    assert dk["minibatch_normalization_block.linear.module.weight"].shape == (512, 8192)
    assert dk["decision_layer.module.weight"].shape == (1, 512)
    assert set(gk) == set(O.init_generator_sd(512, [512, 512, 512, 256], 3))
    assert set(dk) == set(O.init_discriminator_sd([512, 512, 512, 256], 3))


def test_library_exports_every_declared_symbol():
    from pggan_b200 import _lib
    lib = _lib.load()                      # raises if the .so is missing: there is no fallback
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "pggan_b200.h")).read()
    declared = set(re.findall(r"\b(pggan_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pggan_version() >= 100


def test_ops_refuse_cpu_tensors():
    import pytest
    import pggan_b200 as pg
    conv = pg.EqualizedConv2d(64, 64, 3, padding=1)
    with pytest.raises(Exception):
        conv(torch.randn(1, 64, 4, 4))


def test_schedule_matches_reference_bookkeeping():
    """check_jump / change_alpha arithmetic of reference pggan/model.py:176-204, without kernels"""
    import pggan_b200 as pg
    tr = pg.ProgressiveGAN.__new__(pg.ProgressiveGAN)
    tr.args = pg.make_args()
    tr.scale_index = 1
    tr.alpha_index = 0
    tr.G = type("X", (), {"alpha": 0})()
    tr.D = type("X", (), {"alpha": 0})()
    tr.alpha_jump_value = 1 / tr.args.alpha_jump_Ntimes[1]
    tr.next_alpha_jump_step = 2000
    tr.next_scale_jump_step = 10 ** 9
    for step in range(2000, 2000 + 100 * 100 + 1):
        tr.check_jump(step)
    assert tr.alpha_index == 100 and abs(tr.G.alpha - 1.0) < 1e-9 and tr.next_alpha_jump_step == 0


def _emulate_pack(W, args):
    """numpy-free emulation of pack_weight_kernel's index map (pggan_b200/csrc/elementwise.cu)"""
    n_total, n_valid, n2, taps, c_total, c_valid, c2, base, s_n1, s_n2, s_t, s_c1, s_c2 = args
    flat = W.reshape(-1)
    n = torch.arange(n_total).view(-1, 1, 1)
    t = torch.arange(taps).view(1, -1, 1)
    c = torch.arange(c_total).view(1, 1, -1)
    off = base + (n // n2) * s_n1 + (n % n2) * s_n2 + t * s_t + (c // c2) * s_c1 + (c % c2) * s_c2
    valid = (n < n_valid) & (c < c_valid) & (t >= 0)
    out = torch.zeros(n_total, taps, c_total)
    out[valid.expand_as(out)] = flat[off[valid.expand_as(off)]]
    return out


def test_weight_layout_maps():
    """the pg_pack_map tuples of ops._maps against explicit torch permutations"""
    from pggan_b200 import ops
    torch.manual_seed(0)
    # OIHW 3x3: fprop [O][kh*3+kw][I]; dgrad [I][flipped tap][O]
    W = torch.randn(6, 5, 3, 3)
    N, C, T, fm, tm = ops._maps(W.shape, None)
    assert (N, C, T) == (6, 5, 9)
    assert torch.equal(_emulate_pack(W, fm), W.permute(0, 2, 3, 1).reshape(6, 9, 5))
    assert torch.equal(_emulate_pack(W, tm), W.flip(2, 3).permute(1, 2, 3, 0).reshape(5, 9, 6))
    # padded input channels (the 513 -> 576 mbstd conv)
    N, C, T, fm, tm = ops._maps(W.shape, ("pad_c", 8))
    f = _emulate_pack(W, fm)
    assert f.shape == (6, 9, 8) and torch.equal(f[:, :, :5], W.permute(0, 2, 3, 1).reshape(6, 9, 5)) and f[:, :, 5:].abs().sum() == 0
    t = _emulate_pack(W, tm)
    assert t.shape == (8, 9, 6) and torch.equal(t[:5], W.flip(2, 3).permute(1, 2, 3, 0).reshape(5, 9, 6)) and t[5:].abs().sum() == 0
    # G's format layer: rows (c,hw) consumed as (hw,c)
    C0, S, K = 4, 16, 7
    W = torch.randn(C0 * S, K)
    N, C, T, fm, tm = ops._maps(W.shape, ("lin_out_perm", C0))
    Wp = W.view(C0, S, K).transpose(0, 1).reshape(S * C0, K)
    assert torch.equal(_emulate_pack(W, fm)[:, 0, :], Wp)
    assert torch.equal(_emulate_pack(W, tm)[:, 0, :], Wp.t())
    b = torch.randn(C0 * S)
    assert torch.equal(ops._perm_bias(b, ("lin_out_perm", C0)), b.view(C0, S).t().reshape(-1))
    assert torch.equal(ops._perm_bias(ops._perm_bias(b, ("lin_out_perm", C0)), ("lin_out_perm", C0), inverse=True), b)
    # D's linear: columns (c,hw) read as (hw,c)
    O, Cc = 3, 5
    W = torch.randn(O, Cc * S)
    N, C, T, fm, tm = ops._maps(W.shape, ("lin_in_perm", Cc))
    Wl = W.view(O, Cc, S).transpose(1, 2).reshape(O, S * Cc)
    assert torch.equal(_emulate_pack(W, fm)[:, 0, :], Wl)
    assert torch.equal(_emulate_pack(W, tm)[:, 0, :], Wl.t())


def test_reference_import_paths_resolve_to_this_package():
    import importlib
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "pggan_b200", "compat"))
    try:
        for m in ("lib", "lib.layers", "pggan", "pggan.nets"):
            sys.modules.pop(m, None)
        layers = importlib.import_module("lib.layers")
        nets = importlib.import_module("pggan.nets")
        import pggan_b200
        assert layers.EqualizedConv2d is pggan_b200.EqualizedConv2d
        assert nets.Generator is pggan_b200.Generator
        assert importlib.import_module("lib.custom_layers").EqualizedLinear is not None
        assert importlib.import_module("pggan.loss").WGANGPLoss is pggan_b200.WGANGPLoss
    finally:
        sys.path.pop(0)
        for m in [k for k in sys.modules if k == "lib" or k.startswith("lib.") or k == "pggan" or k.startswith("pggan.")]:
            sys.modules.pop(m, None)


def test_checkpoint_roundtrip_reference_format(tmp_path):
    """lib/checkpoint.py:5-34 signatures: save_checkpoint(model, optimizer, name, ckpt_dict), load_checkpoint(args, name)
    with args.ckpt_step (None -> 'latest'); the file is the reference's dict (a torch.optim.Adam optimizer state)"""
    import types
    import inspect
    import pytest
    import pggan_b200 as pg
    from pggan_b200 import checkpoint
    assert list(inspect.signature(checkpoint.load_checkpoint).parameters) == ["args", "name"]
    assert list(inspect.signature(checkpoint.save_checkpoint).parameters) == ["model", "optimizer", "name", "ckpt_dict"]
    G = pg.Generator(32, 32)
    G.add_block(32)
    args = types.SimpleNamespace(save_root=str(tmp_path), run_id="run0", ckpt_id="run0", ckpt_step=None, isMaster=True)
    opt = torch.optim.Adam(G.parameters(), lr=1e-4, betas=(0.0, 0.99))
    ckpt_dict = {"args": vars(args), "global_step": 7, "scale_index": 1, "alpha_G": 0.25}
    checkpoint.save_checkpoint(G, opt, name="G", ckpt_dict=ckpt_dict)
    assert os.path.exists(tmp_path / "run0" / "ckpt" / "G_7.pt") and os.path.exists(tmp_path / "run0" / "ckpt" / "G_latest.pt")
    G2 = pg.Generator(32, 32)
    D2 = pg.Discriminator(32, apply_minibatch_norm=True)
    checkpoint.grow_to_checkpoint(G2, D2, [32, 32], 1)
    for step in (None, 7):
        args.ckpt_step = step
        ck = checkpoint.load_checkpoint(args, "G")
        assert ck["scale_index"] == 1 and ck["global_step"] == 7 and "param_groups" in ck["optimizer"]
        G2.load_state_dict(ck["model"], strict=False)
        for (k1, v1), (k2, v2) in zip(G.state_dict().items(), G2.state_dict().items()):
            assert k1 == k2 and torch.equal(v1.cpu(), v2.cpu())
    args.ckpt_step = 99
    with pytest.raises(FileNotFoundError):
        checkpoint.load_checkpoint(args, "G")


def test_conv_params_struct_layout_matches_the_library():
    """INTEGRATION.md's ctypes structs are the ones of pggan_b200/_lib.py; the library reports its own sizeof()"""
    import ctypes
    from pggan_b200 import _lib
    lib = _lib.load()
    assert ctypes.sizeof(_lib.ConvParams) == lib.pggan_sizeof_conv_params()
    assert ctypes.sizeof(_lib.WgradParams) == lib.pggan_sizeof_wgrad_params()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    for name, _ in _lib.ConvParams._fields_ + _lib.WgradParams._fields_:
        assert f'"{name}"' in doc, f"INTEGRATION.md's struct listing lacks field {name}"


def test_precision_modes_and_tensor_classes():
    """the mode table of ops.py: which element format each tensor class has, and the label dtype"""
    import pytest
    import pggan_b200 as pg
    from pggan_b200 import ops, _lib as L
    try:
        assert ops.get_precision_mode() == "fp16"                   # the default and benchmarked mode
        assert ops.get_compute_dtype() == torch.float16
        assert ops._fmt("a") == L.PG_F16 and ops._fmt("g") == L.PG_F16
        pg.set_compute_dtype("mixed")
        assert ops.get_compute_dtype() == torch.bfloat16 and ops._fmt("a") == L.PG_F16 and ops._fmt("g") == L.PG_BF16
        for dt, fa in ((torch.bfloat16, L.PG_BF16), (torch.float16, L.PG_F16), (torch.float32, L.PG_F32)):
            pg.set_compute_dtype(dt)
            assert ops.get_compute_dtype() == dt and ops._fmt("a") == fa and ops._fmt("g") == fa
        with pytest.raises(ValueError):
            pg.set_compute_dtype(torch.float64)
        pg.set_compute_dtype("mixed")
        # an fp32 tensor imported as an activation becomes fp16 BITS under the bf16 label; to_float reads them back
        x = torch.tensor([1.0, 0.333251953125, -2.5, 70000.0])
        xi = ops.enter(x, "a")
        assert xi.dtype == torch.bfloat16 and torch.equal(xi.view(torch.float16), x.to(torch.float16))
        assert torch.equal(ops.to_float(xi, "a")[:3], x.to(torch.float16).float()[:3])
        gi = ops.enter(x, "g")                                      # gradient class: honest bf16
        assert ops.feat(x, "a").dtype == torch.float32              # precision follows the tensor: fp32 stays fp32
        assert gi.dtype == torch.bfloat16 and torch.equal(gi, x.to(torch.bfloat16))
        # alpha coefficients
        c = ops._coef(ops._affine(1.0, -1.0, 0.25))
        assert abs(c.c0 - 0.75) < 1e-7 and c.c1 == 0.0 and not c.alpha
    finally:
        pg.set_compute_dtype(torch.float16)


def test_dense_map_and_switches():
    """which weight gradients may be accumulated by the GEMM epilogue (ops._dense_map), the deterministic switch, and the
    mask hand-shake's token (ops._premasked) — host logic only"""
    from pggan_b200 import ops
    import pggan_b200 as pg
    # channels_last 3x3 conv weight and plain [out,in] linear weight: the operand IS the parameter memory
    assert ops._dense_map(ops._maps((64, 32, 3, 3), None, ohwi=True)[3])
    assert ops._dense_map(ops._maps((512, 128), None)[3])
    # OIHW 3x3, the permuted linear layers, the padded conv: not dense
    assert not ops._dense_map(ops._maps((64, 32, 3, 3), None, ohwi=False)[3])
    assert not ops._dense_map(ops._maps((512, 8192), ("lin_in_perm", 512))[3])
    assert not ops._dense_map(ops._maps((8192, 512), ("lin_out_perm", 512))[3])
    assert not ops._dense_map(ops._maps((512, 513, 3, 3), ("pad_c", 576), ohwi=True)[3])
    before = ops.DIRECT_WGRAD
    pg.set_deterministic(True)
    assert ops.DIRECT_WGRAD is False
    pg.set_deterministic(False)
    assert ops.DIRECT_WGRAD is True
    ops.DIRECT_WGRAD = before
    # the token protocol: no token -> mask as usual; matching token -> skip once; foreign token -> loud failure
    link = ops._MaskLink()
    t = torch.zeros(2, 3)
    assert ops._premasked(None, t) is False and ops._premasked(link, t) is False
    link.token = (t.data_ptr(), tuple(t.shape))
    assert ops._premasked(link, t) is True and link.token is None
    link.token = (t.data_ptr() + 64, tuple(t.shape))
    with pytest.raises(RuntimeError):
        ops._premasked(link, t)
