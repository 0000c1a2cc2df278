"""Network / loss / step parity on the GPU: pggan_b200 modules (CUDA kernels through the C ABI)
against the oracle (oracle/pggan_oracle.py, fp32 PyTorch) on identical weights and inputs.

Tolerances are BASELINE.json's (north_star): G images and D logits <= 1e-3 relative L2 in fp32/TF32 and <= 1e-2 at
16-bit precision; gradient-penalty value <= 1e-3 relative — relative L2 over the whole tensor.  They are asserted in
the fp32 (3xTF32) parity mode and in torch.float16, the default and BENCHMARKED mode (IEEE-half activations, weights
and gradient tensors, fp32 accumulate, static loss scale in the trainer), at the real widths of BASELINE configs 1-5.
The optional "mixed" mode (fp16 activations, bf16 gradients) is covered at kernel level and by one network test."""
import pytest
import torch

from conftest import rel_l2
from gpu_util import build_pair, req, grads_by_name
from oracle import pggan_oracle as O

pytestmark = pytest.mark.gpu

BENCH = torch.float16        # the benchmarked precision mode (bench.py default)
MIXED = "mixed"
# Forward outputs vs the plain fp32 oracle: north_star's bars.
OUT_TOL = {torch.float32: 1e-3, BENCH: 1e-2, MIXED: 1e-2}
GP_TOL = 1e-3          # gradient-penalty value, every mode
# Gradients are not covered by north_star.  LeakyReLU masks come from the sign of pre-activations, so a forward
# difference of relative size e flips ~e of them and moves the gradient by ~0.8*sqrt(e) in relative L2 (1e-4 -> 0.8e-2
# in fp32 mode, 1e-3 -> 2.5e-2 at 16 bits) on top of the rounding of the gradient tensors themselves.
GRAD_TOL = {torch.float32: 2e-2, BENCH: 8e-2, MIXED: 8e-2}


@pytest.fixture(autouse=True)
def _setup():
    import pggan_b200 as pg
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    pg.set_compute_dtype(BENCH)


def _check_param_grads(mine, ref_params, tol, what, bias_factor=1.0):
    worst = 0.0
    for k, g in mine.items():
        gr = ref_params[k].grad
        if gr is None:
            assert g is None or float(g.abs().max()) == 0.0, (what, k)
            continue
        assert g is not None, (what, k)
        den = float(gr.norm())
        if den < 1e-12:
            continue
        e = rel_l2(g, gr)
        worst = max(worst, e)
        assert e < tol * (bias_factor if k.endswith(".bias") else 1.0), (what, k, e)
    return worst


@pytest.mark.parametrize("dtype", [torch.float32, BENCH])
@pytest.mark.parametrize("scale,alpha,width,batch", [(0, 0.0, 512, 8), (1, 0.5, 512, 8), (3, 0.5, 512, 8),
                                                     (2, 1.0, 64, 6), (5, 0.3, 64, 4)])
def test_forward_and_first_order(scale, alpha, width, batch, dtype):
    import pggan_b200 as pg
    pg.set_compute_dtype(dtype)
    depths = [width] * (scale + 1)
    latent = width
    G, D, g_sd, d_sd = build_pair(scale, depths, latent, seed=scale)
    G.alpha = D.alpha = alpha
    res = 4 * 2 ** scale
    torch.manual_seed(100 + scale)
    z = torch.randn(batch, latent, device="cuda")
    x = torch.rand(batch, 3, res, res, device="cuda") * 2 - 1

    gp_, dp_ = req(g_sd), req(d_sd)
    xr = x.clone().requires_grad_()
    img_ref = O.generator_forward(gp_, z, alpha)
    logit_ref, feat_ref = O.discriminator_forward(dp_, xr, alpha, True)

    img = G(z)
    xm = x.clone().requires_grad_()
    logit, feat = D(xm, True)
    assert img.dtype == torch.float32 and img.shape == img_ref.shape
    assert logit.dtype == torch.float32 and logit.shape == (batch, 1)
    assert feat.dtype == torch.float32 and feat.shape == feat_ref.shape
    e_img, e_logit, e_feat = rel_l2(img, img_ref), rel_l2(logit, logit_ref), rel_l2(feat, feat_ref)
    print(f"\n[{dtype}] scale {scale} width {width}: vs fp32 oracle: img {e_img:.2e} logit {e_logit:.2e} feat {e_feat:.2e}")
    assert e_img < OUT_TOL[dtype], e_img
    assert e_logit < OUT_TOL[dtype], e_logit
    assert e_feat < OUT_TOL[dtype], e_feat

    # first-order gradients of both networks
    wimg = torch.randn_like(img_ref)
    (img_ref * wimg).sum().backward()
    (img * wimg).sum().backward()
    wg = _check_param_grads(grads_by_name(G), gp_, GRAD_TOL[dtype], "G")
    wl = torch.randn_like(logit_ref)
    (logit_ref * wl).sum().backward()
    (logit * wl).sum().backward()
    wd = _check_param_grads(grads_by_name(D), dp_, GRAD_TOL[dtype], "D")
    e_gx = rel_l2(xm.grad, xr.grad)
    print(f"    worst param-grad err G {wg:.2e} D {wd:.2e}; dD/dx {e_gx:.2e}")
    assert e_gx < GRAD_TOL[dtype]


def _gp_tol(dtype, bias_std):
    """north_star: GP value <= 1e-3 relative — asserted for every mode on the reference initialisation (zero biases,
    lib/layers.py:51-53: ||grad|| ~ 0.1, where GP = 10*sum(||g||-1)^2 damps a norm error 4x).  The stress variant with
    random biases puts ||grad|| at 0.8..1.6, where the same formula AMPLIFIES a norm error 5-8x; there a 16-bit forward
    pass (1e-3 off, i.e. ~1e-3 of the LeakyReLU masks of a 4x4x512 map flipped) cannot hold 1e-3 and 1e-2 is asserted;
    the fp32 parity mode holds 1e-3 there too."""
    return GP_TOL if (dtype == torch.float32 or bias_std == 0.0) else 1e-2


@pytest.mark.parametrize("bias_std", [0.0, 0.3])
@pytest.mark.parametrize("dtype", [torch.float32, BENCH])
@pytest.mark.parametrize("scale,width,batch", [(0, 512, 8), (2, 512, 8), (3, 64, 4)])
def test_r1_and_gradient_penalty(scale, width, batch, dtype, bias_std):
    """the double backward through D: value and accumulated D parameter gradients"""
    import pggan_b200 as pg
    pg.set_compute_dtype(dtype)
    alpha = 0.5 if scale else 0.0
    depths = [width] * (scale + 1)
    G, D, g_sd, d_sd = build_pair(scale, depths, width, seed=10 + scale, bias_std=bias_std)
    G.alpha = D.alpha = alpha
    res = 4 * 2 ** scale
    torch.manual_seed(7)
    x = torch.rand(batch, 3, res, res, device="cuda") * 2 - 1
    fake = torch.rand(batch, 3, res, res, device="cuda") * 2 - 1
    eps = torch.rand(batch, 1)
    args = pg.make_args()
    lc = pg.WGANGPLoss(args)

    # ---- WGAN-GP (reference pggan/loss.py:54-92)
    dp_ = req(d_sd)
    xr = x.clone().requires_grad_()
    gp_ref = O.gradient_penalty(lambda t: O.discriminator_forward(dp_, t, alpha), xr, fake, 10.0, eps=eps)
    xm = x.clone().requires_grad_()
    D.zero_grad()
    gp = lc.get_gradient_penalty({"img_real": xm, "img_fake": fake, "pred_real": D(xm)}, D, backward=True, eps=eps)
    e_gp = abs(gp - float(gp_ref)) / abs(float(gp_ref))
    # bias gradients of the penalty exist only through the second-order minibatch-stddev path:
    # tiny numbers, so their relative error is several times that of the weights
    worst = _check_param_grads(grads_by_name(D), dp_, 5 * GRAD_TOL[dtype], "GP", bias_factor=4.0)
    print(f"\n[{dtype}] scale {scale} width {width} bias_std {bias_std}: GP {gp:.5f} vs fp32 oracle {float(gp_ref):.5f} "
          f"rel {e_gp:.2e}; worst D-grad err {worst:.2e}")
    assert e_gp < _gp_tol(dtype, bias_std), e_gp
    assert D.decision_layer.module.bias.grad is None or float(D.decision_layer.module.bias.grad.abs().max()) == 0.0

    # ---- live loss: BCE + R1 (reference pggan/loss.py:16-27)
    dp_ = req(d_sd)
    xr = x.clone().requires_grad_()
    ld = O.loss_D_live(O.discriminator_forward(dp_, xr, alpha), O.discriminator_forward(dp_, fake, alpha), xr)
    ld["L_D"].backward()
    xm = x.clone().requires_grad_()
    D.zero_grad()
    L_D = lc.get_loss_D({"pred_real": D(xm), "pred_fake": D(fake), "img_real": xm, "img_fake": fake})
    L_D.backward()
    e_l = rel_l2(L_D, ld["L_D"])
    worst = _check_param_grads(grads_by_name(D), dp_, 5 * GRAD_TOL[dtype], "R1")
    print(f"    live L_D rel {e_l:.2e}; worst D-grad err {worst:.2e}")
    assert e_l < (1e-3 if dtype == torch.float32 else 1e-2)
    assert abs(lc.loss_dict["L_D"] - round(float(ld["L_D"]), 4)) < max(2e-3, 1e-2 * abs(float(ld["L_D"])))


@pytest.mark.parametrize("dtype", [torch.float32, BENCH])
def test_double_backward_masks_ride_in_gemm_epilogues(dtype, monkeypatch):
    """the penalty's backward: the LeakyReLU masks of D's conv layers are applied by the epilogue of the conv that
    produces the masked gradient (ops._MaskLink), not by passes of their own — same gradients as the unfused
    composition, and no mask kernel left for the conv blocks; the bias gradient of each block's second conv comes out
    of the pool-backward kernel in the plain backward"""
    import pggan_b200 as pg
    from pggan_b200 import ops
    pg.set_compute_dtype(dtype)
    scale, width, batch = 3, 64, 4
    G, D, g_sd, d_sd = build_pair(scale, [width] * (scale + 1), width, seed=3, bias_std=0.3)
    G.alpha = D.alpha = 0.5
    res = 4 * 2 ** scale
    torch.manual_seed(11)
    x = torch.rand(batch, 3, res, res, device="cuda") * 2 - 1
    fake = torch.rand(batch, 3, res, res, device="cuda") * 2 - 1
    eps = torch.rand(batch, 1)
    lc = pg.WGANGPLoss(pg.make_args())
    calls = {"mask": 0, "colsum": 0}
    real_mask, real_mc = ops.k_mask_mul, ops.k_mask_colsum

    def count_mask(*a, **k):
        calls["mask"] += 1
        return real_mask(*a, **k)

    def count_mc(*a, **k):
        calls["colsum"] += 1
        return real_mc(*a, **k)

    monkeypatch.setattr(ops, "k_mask_mul", count_mask)
    monkeypatch.setattr(ops, "k_mask_colsum", count_mc)

    def run(fuse):
        monkeypatch.setattr(ops, "FUSE_DOUBLE_BACKWARD_MASKS", fuse)
        calls["mask"] = calls["colsum"] = 0
        D.zero_grad()
        xm = x.clone().requires_grad_()
        gp = lc.get_gradient_penalty({"img_real": xm, "img_fake": fake, "pred_real": D(xm)}, D, backward=True, eps=eps)
        return gp, {k: v.clone() for k, v in grads_by_name(D).items() if v is not None}, dict(calls)

    gp1, g1, c1 = run(True)
    gp0, g0, c0 = run(False)
    print(f"\n[{dtype}] mask passes in the penalty's forward+backward: fused {c1['mask']}, unfused {c0['mask']}")
    assert abs(gp1 - gp0) <= 1e-5 * abs(gp0)                    # the first backward is the same program
    n_conv = 2 * scale                                           # conv layers of the blocks below the 4x4 head
    assert c0["mask"] - c1["mask"] >= n_conv, (c0, c1)
    assert sorted(g0) == sorted(g1)
    for k in g0:
        if float(g0[k].abs().max()) == 0.0:
            continue
        # (bias gradients of the penalty are the tiny second-order minibatch-stddev path: a looser bound)
        bound = (1e-5 if dtype == torch.float32 else 2e-3) * (5.0 if k.endswith("bias") else 1.0)
        assert rel_l2(g1[k], g0[k]) < bound, (k, rel_l2(g1[k], g0[k]))
    # plain backward of a critic loss: the second conv of every block gets its bias gradient from the pool backward
    calls["colsum"] = 0
    D.zero_grad()
    D(x).sum().backward()
    with_fused = calls["colsum"]
    gb1 = {k: v.clone() for k, v in grads_by_name(D).items() if v is not None}
    monkeypatch.setattr(ops, "upsample_colsum_fusable", lambda *a, **k: False)
    calls["colsum"] = 0
    D.zero_grad()
    D(x).sum().backward()
    assert calls["colsum"] - with_fused >= scale, (calls["colsum"], with_fused)
    for k, v in grads_by_name(D).items():
        if v is not None and k.endswith("bias") and float(v.abs().max()) > 0:
            assert rel_l2(gb1[k], v) < 1e-5, (k, rel_l2(gb1[k], v))


def _param_step_stats(mod, ref_params, sd0):
    """relative L2 of the parameter DELTAS of a network against a reference's, and of the deltas' magnitude.
    With beta1 = 0 Adam's first update is lr*g/(|g|+eps) = +-lr, so one sign disagreement on a near-zero gradient
    costs 2*lr in one element: a fraction f of flipped signs gives a relative L2 of sqrt(4f)."""
    num = den = 0.0
    for k, p in mod.named_parameters():
        d_my = (p.detach() - sd0[k]).double()
        d_ref = (ref_params[k].detach() - sd0[k]).double()
        num += float((d_my - d_ref).pow(2).sum())
        den += float(d_ref.pow(2).sum())
    return (num / max(den, 1e-300)) ** 0.5


def _moment_err(opt, ref_opt, ref_named, mod):
    """relative L2 of sqrt(exp_avg_sq) (the gradient magnitudes Adam has seen) against torch.optim.Adam's state"""
    idx = {id(p): i for i, p in enumerate(opt.grads.params)}
    num = den = 0.0
    for k, p in mod.named_parameters():
        if id(p) not in idx:
            continue
        st = ref_opt.state.get(ref_named[k])
        if not st:
            continue
        from pggan_b200.optim import FlatGrads
        off = opt.grads.offsets[idx[id(p)]]
        mine = FlatGrads.view_like(opt.exp_avg_sq, p, off).double().sqrt()      # logical shape (channels_last strides)
        ref = st["exp_avg_sq"].double().sqrt()
        num += float((mine - ref).pow(2).sum())
        den += float(ref.pow(2).sum())
    return (num / max(den, 1e-300)) ** 0.5


@pytest.mark.parametrize("variant", ["W", "L"])
@pytest.mark.parametrize("scale,width,batch", [(0, 64, 8), (2, 64, 8)])
def test_train_step_matches_oracle(scale, width, batch, variant):
    """one full D+G step (pggan/model.py:206-255) in the fp32 parity mode: losses, Adam moments and parameter deltas"""
    import pggan_b200 as pg
    pg.set_compute_dtype(torch.float32)
    alpha = 0.5 if scale else 0.0
    depths = [width] * (scale + 1)
    args = pg.make_args(latent_dim=width, depths=depths)
    tr = pg.ProgressiveGAN(args, 0, variant=variant)
    tr.set_scale(scale, alpha)
    g_sd = {k: v.detach().clone() for k, v in tr.G.state_dict().items()}
    d_sd = {k: v.detach().clone() for k, v in tr.D.state_dict().items()}
    otr = O.OracleTrainer(g_sd, d_sd, alpha=alpha, variant=variant, latent_dim=width)
    res = 4 * 2 ** scale
    torch.manual_seed(3)
    img = torch.rand(batch, 3, res, res, device="cuda") * 2 - 1
    z_d = torch.randn(batch, width, device="cuda")
    z_g = torch.randn(batch, width, device="cuda")
    eps = torch.rand(batch, 1)
    otr.train_step(img, z_d=z_d, z_g=z_g, eps=eps)
    tr.train_step(img, latent_D=z_d, latent_G=z_g, eps=eps.cuda())
    assert abs(tr.loss_collector.loss_dict["L_G"] - otr.loss_dict["L_G"]) < 2e-3 * max(1.0, abs(otr.loss_dict["L_G"]))
    if variant == "W":
        assert abs(tr.loss_collector.loss_dict["L_D_gp"] - otr.loss_dict["L_D_gp"]) < 2e-3 * abs(otr.loss_dict["L_D_gp"])
    for name, mod, ref, sd0, opt, ropt in (("G", tr.G, otr.g, g_sd, tr.opt_G, otr.opt_G),
                                           ("D", tr.D, otr.d, d_sd, tr.opt_D, otr.opt_D)):
        e_m = _moment_err(opt, ropt, ref, mod)
        e_d = _param_step_stats(mod, ref, sd0)
        print(f"\n{variant} scale {scale} {name}: Adam second-moment err {e_m:.2e}, parameter-delta rel L2 {e_d:.2e}")
        assert e_m < 2e-2, (name, e_m)          # gradient magnitudes: the fp32-mode gradient tolerance
        assert e_d < 0.25, (name, e_d)          # = at most 1.5 % of the update signs differ (near-zero gradients)


def _twin_trainers(scale, depths, latent, alpha, variant="W"):
    import pggan_b200 as pg
    args = pg.make_args(latent_dim=latent, depths=depths)
    a = pg.ProgressiveGAN(args, 0, variant=variant)
    a.set_scale(scale, alpha)
    b = pg.ProgressiveGAN(pg.make_args(latent_dim=latent, depths=depths), 0, variant=variant)
    b.set_scale(scale, alpha)
    b.G.load_state_dict(a.G.state_dict())
    b.D.load_state_dict(a.D.state_dict())
    return a, b


@pytest.mark.parametrize("deterministic", [True, False])
def test_graph_replay_matches_eager_and_oracle_over_three_steps(deterministic):
    """deterministic: pg.set_deterministic(True) keeps every weight gradient's summation order fixed (slice buffers summed
    in order instead of fp32 reductions from the GEMM epilogues): replay and eager then agree to rounding noise.
    The path bench.py times: benchmarked precision mode, C2's real widths (32x32, 512 channels), CUDA-graph replay,
    the batched real+fake D pass — against the same step launched eagerly and against the oracle, over THREE steps
    (Adam moments carried, packed-weight copies invalidated by every update), with alpha CHANGING between the steps:
    the graph captured at alpha 0.3 must follow G.alpha / D.alpha (device scalar), no re-capture."""
    import pggan_b200 as pg
    pg.set_deterministic(deterministic)
    pg.set_compute_dtype(BENCH)
    scale, width, batch = 3, 512, 8
    depths = [width] * (scale + 1)
    eager, graphed = _twin_trainers(scale, depths, width, 0.3)
    g_sd = {k: v.detach().clone() for k, v in eager.G.state_dict().items()}
    d_sd = {k: v.detach().clone() for k, v in eager.D.state_dict().items()}
    otr = O.OracleTrainer(g_sd, d_sd, alpha=0.3, variant="W", latent_dim=width)
    before = [p.detach().clone() for p in graphed.D.parameters()]
    replay = graphed.capture(batch)
    # capturing (3 warm-up steps on noise) must leave parameters, moments and step counters untouched
    for p, q in zip(graphed.D.parameters(), before):
        assert torch.equal(p, q)
    assert int(graphed.opt_D.step_count.item()) == 0 and float(graphed.opt_G.exp_avg_sq.abs().max()) == 0.0
    res = 4 * 2 ** scale
    gen = torch.Generator(device="cuda").manual_seed(11)
    for step, alpha in enumerate((0.3, 0.5, 0.7)):
        img = torch.rand(batch, 3, res, res, device="cuda", generator=gen) * 2 - 1
        z_d = torch.randn(batch, width, device="cuda", generator=gen)
        z_g = torch.randn(batch, width, device="cuda", generator=gen)
        eps = torch.rand(batch, 1, device="cuda", generator=gen)
        for t in (eager, graphed):
            t.G.alpha = t.D.alpha = alpha
        otr.alpha = alpha
        eager.train_step(img, latent_D=z_d, latent_G=z_g, eps=eps)
        replay(img, z_d, z_g, eps)
        otr.train_step(img, z_d=z_d, z_g=z_g, eps=eps.cpu())
        le, lg = eager.loss_collector.loss_dict, graphed.loss_collector.loss_dict
        for key in ("L_G", "L_D", "L_D_gp"):
            assert abs(le[key] - lg[key]) <= 2e-3 * max(1.0, abs(le[key])), (step, key, le[key], lg[key])   # read after EVERY replay
        assert abs(lg["L_D_gp"] - otr.loss_dict["L_D_gp"]) < 2e-3 * abs(otr.loss_dict["L_D_gp"]), (step, lg["L_D_gp"], otr.loss_dict["L_D_gp"])
        assert abs(lg["L_G"] - otr.loss_dict["L_G"]) < 1e-2 * max(1.0, abs(otr.loss_dict["L_G"]))
    torch.cuda.synchronize()
    assert int(graphed.opt_D.step_count.item()) == 3 and int(eager.opt_G.step_count.item()) == 3
    for name, sd0, get in (("G", g_sd, lambda t: t.G), ("D", d_sd, lambda t: t.D)):
        ref_eager = dict(get(eager).named_parameters())
        e_replay = _param_step_stats(get(graphed), ref_eager, sd0)
        e_oracle = _param_step_stats(get(graphed), otr.g if name == "G" else otr.d, sd0)
        opt, ropt, rnamed = (graphed.opt_G, otr.opt_G, otr.g) if name == "G" else (graphed.opt_D, otr.opt_D, otr.d)
        e_m = _moment_err(opt, ropt, rnamed, get(graphed))
        print(f"\n3 steps, {name}: replay vs eager delta rel L2 {e_replay:.2e}; vs oracle {e_oracle:.2e}; "
              f"Adam second-moment err vs oracle {e_m:.2e}")
        # replay and eager launch the same kernels; fp32 atomics (bias sums, RGB weight grads) reorder -> a few signs of
        # near-zero gradients may differ
        # (default mode: the weight-gradient reductions reorder too — with beta1 = 0 a 1e-7 difference in a near-zero
        # gradient flips an update of +-lr)
        # Measured: 1e-8 ... 2e-2 in BOTH modes depending on what ran before in the process — the fp32 atomics of the
        # bias / RGB-layer gradients land in an order that depends on block timing (tensor addresses differ between the
        # eager step and the graph's pool), whatever the order of the weight-gradient sums.
        assert e_replay < 5e-2, (name, e_replay, deterministic)
        assert e_m < GRAD_TOL[BENCH], (name, e_m)
        assert e_oracle < 0.45, (name, e_oracle)       # <= 5 % of the +-lr update signs (near-zero gradients)
    pg.set_deterministic(False)


def test_training_across_scale_jumps_matches_oracle():
    """check_jump -> change_scale / change_alpha (pggan/model.py:158-204) on the GPU: 4x4 -> 8x8 -> 16x16 with fading,
    new blocks, new Adam pair — losses of every step against the oracle trained on the same inputs
    (the oracle is re-seeded with the new randomly-initialised blocks at each jump, as a replica would be)."""
    import pggan_b200 as pg
    pg.set_compute_dtype(torch.float32)
    width, batch = 64, 8
    args = pg.make_args(latent_dim=width, depths=[width] * 3, max_step_at_scale=[2, 4, 4],
                        alpha_jump_start=[-1, 1, 1], alpha_jump_interval=[0, 1, 1], alpha_jump_Ntimes=[0, 2, 2])
    tr = pg.ProgressiveGAN(args, 0, variant="W")
    otr = None
    gen = torch.Generator(device="cuda").manual_seed(5)
    seen_scales, seen_alphas = [], []
    known_scale = -1
    for step in range(9):
        tr.check_jump(step)
        if tr.scale_index != known_scale:                       # (re)build the oracle after a scale jump
            known_scale = tr.scale_index
            g_sd = {k: v.detach().clone() for k, v in tr.G.state_dict().items()}
            d_sd = {k: v.detach().clone() for k, v in tr.D.state_dict().items()}
            otr = O.OracleTrainer(g_sd, d_sd, alpha=float(tr.G.alpha), variant="W", latent_dim=width)
            assert int(tr.opt_D.step_count.item()) == 0          # a fresh Adam pair
        otr.alpha = float(tr.G.alpha)
        assert tr.G.alpha == tr.D.alpha
        res = 4 * 2 ** tr.scale_index
        img = torch.rand(batch, 3, res, res, device="cuda", generator=gen) * 2 - 1
        z_d = torch.randn(batch, width, device="cuda", generator=gen)
        z_g = torch.randn(batch, width, device="cuda", generator=gen)
        eps = torch.rand(batch, 1, device="cuda", generator=gen)
        tr.train_step(img, latent_D=z_d, latent_G=z_g, eps=eps)
        otr.train_step(img, z_d=z_d, z_g=z_g, eps=eps.cpu())
        ld = tr.loss_collector.loss_dict
        seen_scales.append(tr.scale_index)
        seen_alphas.append(float(tr.G.alpha))
        for key in ("L_G", "L_D_gp", "L_D_eps"):
            ref = otr.loss_dict[key]
            assert abs(ld[key] - ref) < 3e-3 * max(1.0, abs(ref)) + 2e-4, (step, key, ld[key], ref)
    assert seen_scales == [0, 0, 1, 1, 1, 1, 2, 2, 2]
    assert seen_alphas == [0.0, 0.0, 0.0, 0.5, 1.0, 1.0, 0.0, 0.5, 1.0], seen_alphas
    assert len(tr.G.blocks) == 2 and len(tr.D.fromRGB_blocks) == 3
    # the first fromRGB / toRGB layers are stale now: not registered with Adam, no gradient
    assert tr.D.fromRGB_blocks[0].fromRGB.module.weight.grad is None


@pytest.mark.parametrize("dtype", [torch.float32, BENCH])
@pytest.mark.parametrize("name,scale,depths,batch", [
    ("C1 4x4", 0, [512], 16),
    ("C2 32x32", 3, [512, 512, 512, 512], 8),
    ("C3 128x128", 5, [512, 512, 512, 512, 256, 128], 4),
    ("C4 256x256", 6, [512, 512, 512, 512, 256, 128, 64], 4),
    ("C5 1024x1024 (paper widths)", 8, [512, 512, 512, 512, 256, 128, 64, 32, 16], 2),
])
def test_full_size_configs(name, scale, depths, batch, dtype):
    """BASELINE configs 1-5 at their real widths/resolutions (small batch): G images, D logits and the WGAN-GP value
    against the oracle at north_star's tolerances, in the fp32 parity mode AND in the benchmarked fp16 mode; plus a
    size-independent property: the first-order chain is linear in its seed, grad(2*sum D) == 2*grad(sum D)."""
    import pggan_b200 as pg
    pg.set_compute_dtype(dtype)
    alpha = 0.5 if scale else 0.0
    G, D, g_sd, d_sd = build_pair(scale, depths, 512, seed=40 + scale, bias_std=0.0)      # the reference initialisation
    G.alpha = D.alpha = alpha
    res = 4 * 2 ** scale
    torch.manual_seed(5)
    z = torch.randn(batch, 512, device="cuda")
    x = torch.rand(batch, 3, res, res, device="cuda") * 2 - 1
    with torch.no_grad():
        img_ref = O.generator_forward(g_sd, z, alpha)
        logit_ref = O.discriminator_forward(d_sd, x, alpha)
        img = G(z)
        logit = D(x)
    e_img, e_logit = rel_l2(img, img_ref), rel_l2(logit, logit_ref)
    print(f"\n[{dtype}] {name}: img {e_img:.2e} logit {e_logit:.2e}")
    assert img.shape == (batch, 3, res, res)
    assert e_img < OUT_TOL[dtype] and e_logit < OUT_TOL[dtype], (e_img, e_logit)
    # gradient penalty value (double backward through every resolution stage)
    eps = torch.rand(batch, 1)
    fake = img.detach()
    gp_ref = O.gradient_penalty(lambda t: O.discriminator_forward(d_sd, t, alpha), x.clone().requires_grad_(), fake,
                                10.0, eps=eps, backward=False)
    lc = pg.WGANGPLoss(pg.make_args())
    xm = x.clone().requires_grad_()
    gp = lc.get_gradient_penalty({"img_real": xm, "img_fake": fake, "pred_real": logit}, D, backward=True, eps=eps)
    e_gp = abs(gp - float(gp_ref)) / abs(float(gp_ref))
    print(f"    GP {gp:.4f} vs {float(gp_ref):.4f} rel {e_gp:.2e}")
    assert e_gp < GP_TOL, e_gp
    for pname, prm in D.named_parameters():
        if prm.grad is not None:
            assert torch.isfinite(prm.grad).all(), pname
    # linearity of the first-order chain in its seed.  At 16 bits the seed carries the trainer's static loss scale: at
    # 1024x1024 the unscaled per-element gradients are ~1e-7, below fp16's normal range (measured: 5 % error unscaled)
    S = 1.0 if dtype == torch.float32 else 1024.0
    xa = x.clone().requires_grad_()
    (g1,) = torch.autograd.grad(S * D(xa).sum(), xa)
    xb = x.clone().requires_grad_()
    (g2,) = torch.autograd.grad(2.0 * S * D(xb).sum(), xb)
    assert rel_l2(g2, 2.0 * g1) < (1e-5 if dtype == torch.float32 else 1e-2)
    # ... and the (unscaled) gradient itself against the oracle's
    xo = x.clone().requires_grad_()
    (g_ref,) = torch.autograd.grad(O.discriminator_forward(d_sd, xo, alpha).sum(), xo)
    e_g = rel_l2(g1 / S, g_ref)
    print(f"    dD/dx vs oracle {e_g:.2e}")
    # (16 bits: up to 17 LeakyReLU layers of mask flips between the logit and the image at 1024x1024; measured 2e-3 at
    # 4x4 ... 7.8e-2 at 1024x1024, twice the headroom asserted)
    assert e_g < (GRAD_TOL[dtype] if dtype == torch.float32 else 2 * GRAD_TOL[dtype]), e_g


@pytest.mark.parametrize("scale,width,batch", [(3, 512, 16), (5, 64, 4)])
def test_bf16_error_vs_stock_autocast(scale, width, batch):
    """(pure bf16 mode — no longer the default: the benchmarked mode is torch.float16, which meets the bars.)
    north_star asks <= 1e-2 for G images / D logits at 16 bits.  Beyond 16x16 pure bf16 measures 1.0-1.8e-2 — the
    question is whether that is the kernels or bf16 itself.  Yardstick: the SAME networks evaluated by stock
    PyTorch/cuDNN under torch.autocast(bfloat16) (the reference's own route to bf16), both against the fp32 oracle.
    Asserted: this build is no further from fp32 than 1.25x stock autocast (measured: it is closer, because its
    epilogues keep bias/He-scale/LeakyReLU in fp32 and round once per layer)."""
    import pggan_b200 as pg
    pg.set_compute_dtype(torch.bfloat16)
    depths = [width] * (scale + 1)
    G, D, g_sd, d_sd = build_pair(scale, depths, width, seed=7 + scale)
    G.alpha = D.alpha = 0.5
    res = 4 * 2 ** scale
    torch.manual_seed(5)
    z = torch.randn(batch, width, device="cuda")
    x = torch.rand(batch, 3, res, res, device="cuda") * 2 - 1
    with torch.no_grad():
        img_ref = O.generator_forward(g_sd, z, 0.5)
        logit_ref = O.discriminator_forward(d_sd, x, 0.5)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            img_ac = O.generator_forward(g_sd, z, 0.5).float()
            logit_ac = O.discriminator_forward(d_sd, x, 0.5).float()
        img, logit = G(z), D(x)
    e_img, e_img_ac = rel_l2(img, img_ref), rel_l2(img_ac, img_ref)
    e_log, e_log_ac = rel_l2(logit, logit_ref), rel_l2(logit_ac, logit_ref)
    print(f"\nbf16 vs fp32 oracle, scale {scale} width {width}: G images ours {e_img:.2e} / stock autocast {e_img_ac:.2e}; "
          f"D logits ours {e_log:.2e} / stock autocast {e_log_ac:.2e}")
    assert e_img < 1.25 * e_img_ac + 1e-3, (e_img, e_img_ac)
    assert e_log < 1.25 * e_log_ac + 1e-3, (e_log, e_log_ac)


@pytest.mark.parametrize("scale,width,batch", [(3, 512, 16), (5, 64, 4)])
def test_mixed_mode_networks(scale, width, batch):
    """the optional "mixed" mode (fp16 activations and weights, bf16 gradient tensors; every tensor-core GEMM runs in one
    format, the saved activation is re-rounded to bf16 for the weight-gradient GEMM): forward outputs meet the 16-bit
    bar like fp16 mode (same forward arithmetic); the gradient-penalty value carries the bf16 rounding of its
    first-order gradient chain: measured 3e-4 .. 2e-3 (largest where ||grad|| is near 1), asserted at 5e-3."""
    import pggan_b200 as pg
    pg.set_compute_dtype(MIXED)
    depths = [width] * (scale + 1)
    G, D, g_sd, d_sd = build_pair(scale, depths, width, seed=11 + scale)
    G.alpha = D.alpha = 0.5
    res = 4 * 2 ** scale
    torch.manual_seed(6)
    z = torch.randn(batch, width, device="cuda")
    x = torch.rand(batch, 3, res, res, device="cuda") * 2 - 1
    with torch.no_grad():
        img_ref = O.generator_forward(g_sd, z, 0.5)
        logit_ref = O.discriminator_forward(d_sd, x, 0.5)
        img, logit = G(z), D(x)
    e_img, e_log = rel_l2(img, img_ref), rel_l2(logit, logit_ref)
    print(f"\nmixed mode vs fp32 oracle, scale {scale} width {width}: G images {e_img:.2e}, D logits {e_log:.2e}")
    assert e_img < 1e-2 and e_log < 1e-2, (e_img, e_log)
    fake = torch.rand(batch, 3, res, res, device="cuda") * 2 - 1
    eps = torch.rand(batch, 1)
    lc = pg.WGANGPLoss(pg.make_args())
    dp_ = req(d_sd)
    gp_ref = O.gradient_penalty(lambda t: O.discriminator_forward(dp_, t, 0.5), x.clone().requires_grad_(), fake,
                                10.0, eps=eps)
    xm = x.clone().requires_grad_()
    D.zero_grad()
    gp = lc.get_gradient_penalty({"img_real": xm, "img_fake": fake, "pred_real": D(xm)}, D, backward=True, eps=eps)
    e_gp = abs(gp - float(gp_ref)) / abs(float(gp_ref))
    worst = _check_param_grads(grads_by_name(D), dp_, 5 * GRAD_TOL[MIXED], "GP", bias_factor=4.0)
    print(f"    GP {gp:.5f} vs fp32 oracle {float(gp_ref):.5f} rel {e_gp:.2e}; worst D-grad err {worst:.2e}")
    assert e_gp < 5e-3, (gp, float(gp_ref))
    for name, prm in D.named_parameters():
        if prm.grad is not None:
            assert torch.isfinite(prm.grad).all(), name


def test_static_loss_scale_is_transparent():
    """the trainer multiplies both losses (and the penalty's own backward) by loss_scale and the fused Adam kernel
    divides it out: one step with scale 1024 and one with scale 1 (bf16-free fp32 mode, where nothing under- or
    overflows either way) must move the parameters identically; in the fp16 mode the scale is on by default."""
    import pggan_b200 as pg
    pg.set_compute_dtype(BENCH)
    assert pg.ProgressiveGAN(pg.make_args(latent_dim=64, depths=[64]), 0).loss_scale == 1024.0
    pg.set_compute_dtype(torch.float32)
    a, b = _twin_trainers(1, [64, 64], 64, 0.5)
    assert a.loss_scale == 1.0
    b.loss_scale = 1024.0
    b.set_optimizers()                      # the optimizers fold 1/loss_scale
    torch.manual_seed(4)
    img = torch.rand(8, 3, 8, 8, device="cuda") * 2 - 1
    z_d, z_g, eps = torch.randn(8, 64, device="cuda"), torch.randn(8, 64, device="cuda"), torch.rand(8, 1, device="cuda")
    a.train_step(img, latent_D=z_d, latent_G=z_g, eps=eps)
    b.train_step(img, latent_D=z_d, latent_G=z_g, eps=eps)
    assert rel_l2(b.opt_D.grads.flat, 1024.0 * a.opt_D.grads.flat) < 1e-5        # .grad holds scale * gradient
    for (k, p), (_, q) in zip(a.D.named_parameters(), b.D.named_parameters()):
        assert rel_l2(q, p) < 1e-6, k
    for (k, p), (_, q) in zip(a.G.named_parameters(), b.G.named_parameters()):
        assert rel_l2(q, p) < 1e-6, k
    assert abs(a.loss_collector.loss_dict["L_D_gp"] - b.loss_collector.loss_dict["L_D_gp"]) < 1e-3


def test_demo_inference_entry(tmp_path):
    """reference demo.py:23-68 through pggan_b200.demo: a reference-format checkpoint ({"model": state_dict, ...}) is
    loaded with the reference's key-set check and the generated images equal the oracle's G on the same latents."""
    import pggan_b200 as pg
    from pggan_b200 import demo
    pg.set_compute_dtype(torch.float32)
    scale, width = 3, 64
    depths = [width] * (scale + 1)
    gen = torch.Generator().manual_seed(3)
    g_sd = O.init_generator_sd(width, depths, scale, generator=gen)
    ckpt = tmp_path / "G_latest.pt"
    torch.save({"model": g_sd, "optimizer": None, "args": {}, "global_step": 123, "alpha_G": 0.25}, ckpt)
    args = pg.make_args(latent_dim=width, depths=depths, ckpt_path=str(ckpt), n_samples=5, gpu=0,
                        save_path=str(tmp_path / "out"))
    G = demo.LoadModel(args)
    assert G.alpha == 0.25 and next(G.parameters()).is_cuda
    z = torch.randn(5, width)
    imgs = demo.test(args, G=G, latent_input=z)
    with torch.no_grad():
        ref = O.generator_forward({k: v.cuda() for k, v in g_sd.items()}, z.cuda(), 0.25)
    ref = (ref.permute(0, 2, 3, 1).cpu().numpy() * 0.5 + 0.5) * 255
    assert imgs.shape == (5, 32, 32, 3)
    assert rel_l2(torch.from_numpy(imgs), torch.from_numpy(ref)) < 1e-3
    assert len(list((tmp_path / "out").iterdir())) == 5
    # a checkpoint whose keys do not match the model is refused, as in the reference (demo.py:39-42)
    bad = dict(g_sd)
    bad.pop(next(iter(bad)))
    torch.save({"model": bad}, ckpt)
    with pytest.raises(AssertionError):
        demo.LoadModel(args)
