"""Generate tests/golden/*.pt by EXECUTING THE REFERENCE's own modules on CPU.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

The fixtures pin oracle/pggan_oracle.py (tests/test_oracle_golden.py) and are replayed
through the CUDA path on the GPU (tests/test_gpu_golden.py).  Widths are kept small
(latent/depths 32, one case at 64 to hit the 64-channel tensor-core tile) so the
committed files stay small.

Reference entry points executed here (unmodified):
  pggan.nets.Generator / Discriminator      pggan/nets.py:10-276
  pggan.loss.WGANGPLoss                      pggan/loss.py:4-100
  lib.loss.Loss.get_BCE_loss / get_r1_reg    lib/loss.py:119-135
  lib.blocks.concatenate_stddev_channel      lib/blocks.py:204-233
  lib.utils.upscale2d / downscale2d          lib/utils.py:106-124
train_step is restated from pggan/model.py:206-255 around those modules because
ProgressiveGAN itself hard-requires CUDA, a dataset directory and wandb.
"""
import os
import sys
import types

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
REF = os.environ.get("PGGAN_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "submodel"))

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from pggan.nets import Generator, Discriminator  # noqa: E402
from pggan.loss import WGANGPLoss  # noqa: E402
from lib.loss import Loss  # noqa: E402
from lib.blocks import concatenate_stddev_channel  # noqa: E402
from lib.layers import PixelwiseVectorNorm, EqualizedConv2d, EqualizedLinear  # noqa: E402
from lib import utils as ref_utils  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
os.makedirs(OUT, exist_ok=True)

LATENT, DEPTHS = 32, [32, 32, 32]      # per-case override via set_width()
ARGS = types.SimpleNamespace(W_adv=1, W_gp=10, W_drift_D=0.001, max_step=1)

# pggan/loss.py:73 does eps.to(x.get_device()); get_device() is -1 on CPU and .to(-1)
# raises.  Make get_device() return the device object for this run only.
torch.Tensor.get_device = lambda self: self.device  # type: ignore[assignment]


def set_width(latent, width):
    global LATENT, DEPTHS
    LATENT, DEPTHS = latent, [width] * 3


def build(scale, seed):
    torch.manual_seed(seed)
    G = Generator(LATENT, DEPTHS[0], True, 0.2, True, None, 3, True)
    D = Discriminator(DEPTHS[0], True, 0.2, 1, True, 3, True)
    for s in range(1, scale + 1):
        G.add_block(DEPTHS[s])
        D.add_block(DEPTHS[s])
    # non-zero biases so the c*(Wx+b) bias path is exercised
    with torch.no_grad():
        for m in list(G.parameters()) + list(D.parameters()):
            if m.dim() == 1:
                m.normal_(0, 0.5)
    return G, D


def sd_of(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def grads_of(m):
    return {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in m.named_parameters()}


def case_layers():
    torch.manual_seed(10)
    out = {}
    conv = EqualizedConv2d(64, 128, 3, padding=1)
    conv.module.bias.data.normal_()
    x = torch.randn(2, 64, 8, 8)
    out["conv3"] = dict(w=conv.module.weight.data.clone(), b=conv.module.bias.data.clone(), x=x, y=conv(x).detach())
    rgb = EqualizedConv2d(64, 3, 1)
    rgb.module.bias.data.normal_()
    out["conv1"] = dict(w=rgb.module.weight.data.clone(), b=rgb.module.bias.data.clone(), x=x, y=rgb(x).detach())
    lin = EqualizedLinear(64, 96)
    lin.module.bias.data.normal_()
    xl = torch.randn(5, 64)
    out["linear"] = dict(w=lin.module.weight.data.clone(), b=lin.module.bias.data.clone(), x=xl, y=lin(xl).detach())
    out["pixel_norm"] = dict(x=x, y=PixelwiseVectorNorm()(x))
    out["upscale2d"] = dict(x=x, y=ref_utils.upscale2d(x))
    out["downscale2d"] = dict(x=x, y=ref_utils.downscale2d(x))
    for b in (1, 3, 4, 6, 8):
        xb = torch.randn(b, 16, 4, 4)
        out[f"mbstd_b{b}"] = dict(x=xb, y=concatenate_stddev_channel(xb, 4))
    torch.save(out, os.path.join(OUT, "layers.pt"))


def case_nets(scale, alpha, batch, seed, name):
    G, D = build(scale, seed)
    G.alpha = alpha
    D.alpha = alpha
    res = 4 * 2 ** scale
    torch.manual_seed(seed + 100)
    z = torch.randn(batch, LATENT)
    x = torch.rand(batch, 3, res, res) * 2 - 1
    img = G(z).detach()
    logit, feat = D(x, True)
    rec = dict(scale=scale, alpha=alpha, latent=LATENT, depths=list(DEPTHS), g_sd=sd_of(G), d_sd=sd_of(D), z=z, x=x,
               img=img, logit=logit.detach(), feat=feat.detach())

    # first-order grads of both nets (sum-of-squares heads so every output matters)
    G.zero_grad()
    (G(z) ** 2).sum().backward()
    rec["g_grads"] = grads_of(G)
    D.zero_grad()
    xr = x.clone().requires_grad_()
    (D(xr) ** 2).sum().backward()
    rec["d_grads"] = grads_of(D)
    rec["d_grad_x"] = xr.grad.detach().clone()

    # live D loss (BCE + R1), pggan/loss.py:16-27
    lc = WGANGPLoss(ARGS)
    D.zero_grad()
    xr = x.clone().requires_grad_()
    pred_real = D(xr)
    pred_fake = D(img)
    L_D = lc.get_loss_D({"pred_real": pred_real, "pred_fake": pred_fake, "img_real": xr, "img_fake": img})
    L_D.backward()
    rec["live_L_D"] = L_D.detach().clone()
    rec["live_d_grads"] = grads_of(D)
    rec["live_r1"] = Loss.get_r1_reg(Loss.get_BCE_loss(D(xr), True), xr).detach()

    # WGAN-GP penalty with its internal backward, pggan/loss.py:54-92
    D.zero_grad()
    xr = x.clone().requires_grad_()
    pred_real = D(xr)
    torch.manual_seed(seed + 200)
    eps = torch.rand(batch, 1)
    torch.manual_seed(seed + 200)          # get_gradient_penalty draws the same eps
    gp = lc.get_gradient_penalty({"img_real": xr, "img_fake": img, "pred_real": pred_real}, D, backward=True)
    rec["gp_eps"] = eps
    rec["gp_value"] = torch.tensor(gp)
    rec["gp_d_grads"] = grads_of(D)
    rec["drift"] = torch.tensor(lc.get_drift_loss({"pred_real": pred_real}))
    rec["loss_G"] = lc.get_loss_G({"pred_fake": D(img)}).detach()
    torch.save(rec, os.path.join(OUT, name))


def case_train_step(scale, alpha, batch, seed, variant, name):
    """pggan/model.py:206-255 restated around the reference modules + torch Adam."""
    G, D = build(scale, seed)
    G.alpha = alpha
    D.alpha = alpha
    res = 4 * 2 ** scale
    g0, d0 = sd_of(G), sd_of(D)
    opt_G = torch.optim.Adam(G.parameters(), lr=1e-4, betas=(0.0, 0.99))
    opt_D = torch.optim.Adam(D.parameters(), lr=1e-5, betas=(0.0, 0.99))
    lc = WGANGPLoss(ARGS)
    torch.manual_seed(seed + 300)
    img_real0 = torch.rand(batch, 3, res, res) * 2 - 1
    z_d = torch.randn(batch, LATENT)
    z_g = torch.randn(batch, LATENT)
    eps = torch.rand(batch, 1)

    img_real = img_real0
    if scale:
        low = F.avg_pool2d(img_real, (2, 2))
        low = F.interpolate(low, scale_factor=2, mode="nearest")
        img_real = (1 - D.alpha) * low + D.alpha * img_real
    img_real = img_real.detach().requires_grad_()
    pred_real = D(img_real)
    img_fake = G(z_d).detach()
    pred_fake = D(img_fake)
    D_dict = {"img_real": img_real, "img_fake": img_fake, "pred_real": pred_real, "pred_fake": pred_fake}
    if variant == "L":
        loss_D = lc.get_loss_D(D_dict)
        opt_D.zero_grad()
        loss_D.backward()
        opt_D.step()
    else:
        L_real = Loss.get_BCE_loss(pred_real, True)
        L_fake = Loss.get_BCE_loss(pred_fake, False)
        loss_D = L_real + L_fake
        opt_D.zero_grad()
        state = torch.get_rng_state()
        torch.manual_seed(0)
        # feed the pre-drawn eps through the reference code by replaying the generator
        _orig = torch.rand
        torch.rand = lambda *a, **k: eps.clone()
        try:
            gp = lc.get_gradient_penalty(D_dict, D, backward=True)
        finally:
            torch.rand = _orig
            torch.set_rng_state(state)
        drift = lc.get_drift_loss(D_dict)
        loss_D.backward()
        opt_D.step()
    img_fake = G(z_g)
    pred_fake, _ = D(img_fake, True)
    loss_G = lc.get_loss_G({"pred_fake": pred_fake})
    opt_G.zero_grad()
    loss_G.backward()
    opt_G.step()
    rec = dict(scale=scale, alpha=alpha, latent=LATENT, depths=list(DEPTHS), variant=variant, g_sd0=g0, d_sd0=d0, img_real=img_real0,
               z_d=z_d, z_g=z_g, eps=eps, g_sd1=sd_of(G), d_sd1=sd_of(D),
               loss_D=loss_D.detach(), loss_G=loss_G.detach(), img_fake=img_fake.detach())
    if variant == "W":
        rec["gp"] = torch.tensor(gp)
        rec["drift"] = torch.tensor(drift)
    torch.save(rec, os.path.join(OUT, name))


if __name__ == "__main__":
    case_layers()
    set_width(64, 64)                      # one case on the 64-channel tensor-core tile
    case_nets(1, 0.3, 8, 2, "nets_s1_w64.pt")
    set_width(32, 32)
    case_nets(0, 0.0, 4, 1, "nets_s0.pt")
    case_nets(2, 0.7, 8, 3, "nets_s2.pt")
    case_train_step(1, 0.5, 8, 4, "W", "step_W_s1.pt")
    case_train_step(1, 0.5, 8, 5, "L", "step_L_s1.pt")
    case_train_step(0, 0.0, 4, 6, "W", "step_W_s0.pt")
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
