"""TEST / BASELINE INFRASTRUCTURE — one PGGAN D+G training step on the REFERENCE'S OWN modules (oracle/_ref, staged
by oracle/stage_ref.py), on whatever device the tensors live on (the host CPU for bench.py's reference arm).

ProgressiveGAN itself (pggan/model.py) hard-requires CUDA, a dataset directory and wandb, so `train_step`
(pggan/model.py:206-255) is restated here line for line around the unmodified reference Generator / Discriminator /
WGANGPLoss / Loss — exactly as oracle/make_golden.py does for the golden vectors.  Step variant "W" wires the dormant
WGAN-GP code the way SURVEY.md §8(d) fixes it (zero_grad -> get_gradient_penalty (its own backward) ->
loss_D.backward -> Adam); "L" is the live BCE+R1 step.  Only bench.py (reference arm / cpu_baseline) and tests use this.
"""
from __future__ import annotations

import os
import sys
import types

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")
_mods = None


def load_reference():
    """import the staged reference modules (once); raises if oracle/_ref is absent"""
    global _mods
    if _mods is not None:
        return _mods
    if not os.path.exists(os.path.join(REFDIR, "MANIFEST.json")):
        raise FileNotFoundError("oracle/_ref is not staged (python oracle/stage_ref.py in the build container)")
    sys.dont_write_bytecode = True
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k in ("lib", "pggan", "submodel") or k.startswith(("lib.", "pggan.", "submodel."))}
    sys.path.insert(0, REFDIR)
    try:
        from pggan.nets import Generator, Discriminator
        from pggan.loss import WGANGPLoss
        from lib.loss import Loss
        _mods = types.SimpleNamespace(Generator=Generator, Discriminator=Discriminator, WGANGPLoss=WGANGPLoss, Loss=Loss)
    finally:
        sys.path.remove(REFDIR)
        for k in [k for k in sys.modules if k in ("lib", "pggan", "submodel") or k.startswith(("lib.", "pggan.", "submodel."))]:
            sys.modules.pop(k)
        sys.modules.update(saved)
    # pggan/loss.py:73 does eps.to(x.get_device()); get_device() is -1 on CPU and .to(-1) raises
    if not getattr(torch.Tensor.get_device, "_pggan_patched", False):
        orig = torch.Tensor.get_device

        def get_device(self):
            return self.device if self.device.type == "cpu" else orig(self)

        get_device._pggan_patched = True
        torch.Tensor.get_device = get_device
    return _mods


class ReferenceTrainer:
    def __init__(self, depths, scale, alpha=0.5, variant="W", latent_dim=512, device="cpu", seed=0):
        m = load_reference()
        torch.manual_seed(seed)
        self.G = m.Generator(latent_dim, depths[0], True, 0.2, True, None, 3, True)
        self.D = m.Discriminator(depths[0], True, 0.2, 1, True, 3, True)
        for s in range(1, scale + 1):
            self.G.add_block(depths[s])
            self.D.add_block(depths[s])
        self.G.to(device)
        self.D.to(device)
        self.G.alpha = self.D.alpha = alpha
        self.scale, self.variant, self.latent_dim, self.device = scale, variant, latent_dim, device
        # configs.yaml:24-27 (beta1 cast to float: torch >= 2 rejects the YAML's int 0)
        self.opt_G = torch.optim.Adam(self.G.parameters(), lr=1e-4, betas=(0.0, 0.99))
        self.opt_D = torch.optim.Adam(self.D.parameters(), lr=1e-5, betas=(0.0, 0.99))
        self.lc = m.WGANGPLoss(types.SimpleNamespace(W_adv=1, W_gp=10, W_drift_D=0.001, max_step=1))
        self.Loss = m.Loss

    def train_step(self, img_real):
        G, D, lc = self.G, self.D, self.lc
        n = len(img_real)
        if self.scale:                                                    # pggan/model.py:217-221
            low = F.avg_pool2d(img_real, (2, 2))
            low = F.interpolate(low, scale_factor=2, mode="nearest")
            img_real = (1 - D.alpha) * low + D.alpha * img_real
        img_real = img_real.detach().requires_grad_()                     # :223
        pred_real = D(img_real)                                           # :224
        img_fake = G(torch.randn(n, self.latent_dim).to(self.device)).detach()   # :226-227
        pred_fake = D(img_fake)                                           # :228
        D_dict = {"img_real": img_real, "img_fake": img_fake, "pred_real": pred_real, "pred_fake": pred_fake}
        if self.variant == "L":
            loss_D = lc.get_loss_D(D_dict)                                # :237 (live BCE + R1)
            self.opt_D.zero_grad()
            loss_D.backward()
            self.opt_D.step()
        else:
            loss_D = self.Loss.get_BCE_loss(pred_real, True) + self.Loss.get_BCE_loss(pred_fake, False)
            self.opt_D.zero_grad()
            lc.get_gradient_penalty(D_dict, D, backward=True)             # pggan/loss.py:54-92, backward inside
            lc.get_drift_loss(D_dict)                                     # pggan/loss.py:94-100
            loss_D.backward()
            self.opt_D.step()
        img_fake = G(torch.randn(n, self.latent_dim).to(self.device))     # :244-245
        pred_fake, _ = D(img_fake, True)                                  # :246
        loss_G = lc.get_loss_G({"pred_fake": pred_fake})                  # :252
        self.opt_G.zero_grad()
        loss_G.backward()
        self.opt_G.step()                                                 # :253
        return [img_real, img_fake]
