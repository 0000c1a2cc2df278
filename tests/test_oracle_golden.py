"""Pin oracle/pggan_oracle.py against fixtures produced by the reference's own modules
(oracle/make_golden.py).  CPU only."""
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import pggan_oracle as O

TOL = 2e-5   # fp32 CPU vs fp32 CPU; differences are summation-order only


def test_layers():
    g = load_golden("layers.pt")
    c = g["conv3"]
    assert rel_l2(O.eq_conv2d(c["x"], c["w"], c["b"], 1), c["y"]) < TOL
    c = g["conv1"]
    assert rel_l2(O.eq_conv2d(c["x"], c["w"], c["b"], 0), c["y"]) < TOL
    c = g["linear"]
    assert rel_l2(O.eq_linear(c["x"], c["w"], c["b"]), c["y"]) < TOL
    c = g["pixel_norm"]
    assert rel_l2(O.pixel_norm(c["x"]), c["y"]) < TOL
    c = g["upscale2d"]
    assert torch.equal(O.upscale2d(c["x"]), c["y"])            # exact indexing
    c = g["downscale2d"]
    assert rel_l2(O.downscale2d(c["x"]), c["y"]) < 1e-7
    for b in (1, 3, 4, 6, 8):
        c = g[f"mbstd_b{b}"]
        y = O.mbstd_concat(c["x"], 4)
        assert y.shape == c["y"].shape
        assert rel_l2(y, c["y"]) < TOL


def _check_grads(params, golden, tol):
    for k, gref in golden.items():
        g = params[k].grad
        if gref is None:
            assert g is None or float(g.abs().max()) == 0.0, k
            continue
        assert g is not None, k
        if float(gref.norm()) == 0.0:
            assert float(g.abs().max()) < 1e-6, k
        else:
            assert rel_l2(g, gref) < tol, (k, rel_l2(g, gref))


@pytest.mark.parametrize("name", ["nets_s0.pt", "nets_s1_w64.pt", "nets_s2.pt"])
def test_nets_and_losses(name):
    g = load_golden(name)
    a = g["alpha"]
    gp_ = {k: v.clone().requires_grad_() for k, v in g["g_sd"].items()}
    dp_ = {k: v.clone().requires_grad_() for k, v in g["d_sd"].items()}
    img = O.generator_forward(gp_, g["z"], a)
    assert rel_l2(img, g["img"]) < TOL
    logit, feat = O.discriminator_forward(dp_, g["x"], a, True)
    assert rel_l2(logit, g["logit"]) < TOL and rel_l2(feat, g["feat"]) < TOL

    (img ** 2).sum().backward()
    _check_grads(gp_, g["g_grads"], 1e-4)
    xr = g["x"].clone().requires_grad_()
    (O.discriminator_forward(dp_, xr, a) ** 2).sum().backward()
    _check_grads(dp_, g["d_grads"], 1e-4)
    assert rel_l2(xr.grad, g["d_grad_x"]) < 1e-4

    # live BCE + R1
    for p in dp_.values():
        p.grad = None
    xr = g["x"].clone().requires_grad_()
    ld = O.loss_D_live(O.discriminator_forward(dp_, xr, a), O.discriminator_forward(dp_, g["img"], a), xr)
    assert rel_l2(ld["L_D"], g["live_L_D"]) < 1e-4
    assert rel_l2(ld["L_reg"], g["live_r1"]) < 1e-4
    ld["L_D"].backward()
    _check_grads(dp_, g["live_d_grads"], 2e-4)

    # WGAN-GP
    for p in dp_.values():
        p.grad = None
    xr = g["x"].clone().requires_grad_()
    gp = O.gradient_penalty(lambda t: O.discriminator_forward(dp_, t, a), xr, g["img"], 10.0, eps=g["gp_eps"])
    assert rel_l2(gp, g["gp_value"]) < 1e-4
    _check_grads(dp_, g["gp_d_grads"], 2e-4)
    assert dp_["decision_layer.module.bias"].grad is None
    pred_real = O.discriminator_forward(dp_, g["x"], a)
    assert rel_l2(O.drift_loss(pred_real, 0.001), g["drift"]) < 1e-4
    assert rel_l2(O.loss_G(O.discriminator_forward(dp_, g["img"], a)), g["loss_G"]) < 1e-5


@pytest.mark.parametrize("name", ["step_W_s0.pt", "step_W_s1.pt", "step_L_s1.pt"])
def test_train_step(name):
    g = load_golden(name)
    tr = O.OracleTrainer(g["g_sd0"], g["d_sd0"], alpha=g["alpha"], variant=g["variant"], latent_dim=g["latent"])
    tr.train_step(g["img_real"], z_d=g["z_d"], z_g=g["z_g"], eps=g["eps"])
    assert abs(tr.loss_dict["L_G"] - float(g["loss_G"])) < 1e-4 * max(1.0, abs(float(g["loss_G"])))
    if g["variant"] == "W":
        assert abs(tr.loss_dict["L_D_gp"] - float(g["gp"])) < 2e-4 * abs(float(g["gp"]))
    # Adam's first step moves every live parameter by ~lr*sign(grad): compare the deltas
    for sd0, sd1, mine, lr in ((g["g_sd0"], g["g_sd1"], tr.g, 1e-4), (g["d_sd0"], g["d_sd1"], tr.d, 1e-5)):
        for k in sd0:
            ref_delta = sd1[k] - sd0[k]
            my_delta = mine[k].detach() - sd0[k]
            # sign flips happen only where the gradient is ~0; bound the mismatch fraction
            bad = ((ref_delta - my_delta).abs() > 0.5 * lr).float().mean().item()
            assert bad < 5e-3, (k, bad)
