"""Losses of the PGGAN step — mirror of reference lib/loss.py (LossInterface, Loss.get_BCE_loss,
Loss.get_r1_reg) and pggan/loss.py (WGANGPLoss).

Loss-level math on [B,1] tensors stays in PyTorch autograd; the per-sample reductions
over image-sized gradients run in the fused reduction kernels (ops.RowsSumSq / ops.GradientPenalty).
`loss_dict` keeps the reference keys (L_G, L_D_real, L_D_fake, L_D, L_D_gp, L_D_eps); values are
fetched lazily (no host sync inside the step) and read back as rounded floats like the reference's.
"""
from __future__ import annotations

import abc
import time

import torch
import torch.nn.functional as F

from . import ops


class LazyLossDict(dict):
    """dict whose tensor values are read as round(x.item(), 4) (the reference stores the rounded float eagerly, which
    costs 4-6 host syncs per step — pggan/loss.py:12,23-25).  The TENSOR stays stored: under CUDA-graph replay the
    Python loss code does not run again, the graph refreshes the same tensors, and every read sees the current step."""

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        if isinstance(v, torch.Tensor):
            return round(float(v.item()), 4)
        return v

    def get(self, key, default=None):
        return self[key] if key in self else default

    def items(self):
        return [(k, self[k]) for k in dict.keys(self)]

    def values(self):
        return [self[k] for k in dict.keys(self)]

    def tensors(self):
        """the raw stored values (0-dim device tensors), no host sync"""
        return dict(dict.items(self))


class LossInterface(metaclass=abc.ABCMeta):
    """reference lib/loss.py:8-51"""

    def __init__(self, args):
        self.args = args
        self.start_time = time.time()
        self.loss_dict = LazyLossDict()

    def print_loss(self, global_step):
        seconds = int(time.time() - self.start_time)
        print("")
        print(f"[ {seconds//3600//24:02}d {(seconds//3600)%24:02}h {(seconds//60)%60:02}m {seconds%60:02}s ]")
        print(f'steps: {global_step:06} / {self.args.max_step}')
        print(f'lossD: {self.loss_dict["L_D"]} | lossG: {self.loss_dict["L_G"]}')

    @abc.abstractmethod
    def get_loss_G(self):
        pass

    @abc.abstractmethod
    def get_loss_D(self):
        pass


class Loss:
    """the two members of reference lib/loss.py:54-141 that the PGGAN path uses"""

    def get_BCE_loss(logits, target):
        """reference lib/loss.py:119-123"""
        assert target in [1, 0]
        targets = torch.full_like(logits, fill_value=target)
        return F.binary_cross_entropy_with_logits(logits, targets)

    def get_r1_reg(d_out, x_in):
        """zero-centred gradient penalty on real images — reference lib/loss.py:125-135.
        d_out may be the BCE loss scalar (as the reference calls it)."""
        with ops.input_grad_only():
            grad_dout = torch.autograd.grad(outputs=d_out.sum(), inputs=x_in, create_graph=True, retain_graph=True,
                                            only_inputs=True)[0]
        assert grad_dout.size() == x_in.size()
        if grad_dout.is_cuda:
            return 0.5 * ops.RowsSumSq.apply(grad_dout).mean(0)
        return 0.5 * grad_dout.pow(2).view(x_in.size(0), -1).sum(1).mean(0)

    def get_adversarial_loss(logits, target):
        """reference lib/loss.py:137-141"""
        return Loss.get_BCE_loss(logits, target)


class WGANGPLoss(LossInterface):
    """reference pggan/loss.py:4-100"""

    def get_loss_G(self, G_dict):
        """reference pggan/loss.py:5-14"""
        L_G = 0.0
        if self.args.W_adv:
            L_adv = Loss.get_BCE_loss(G_dict["pred_fake"], True)
            L_G += self.args.W_adv * L_adv
            self.loss_dict["L_G"] = L_G.detach()
        return L_G

    def get_loss_D(self, D_dict):
        """the LIVE discriminator loss: BCE(real,1) + BCE(fake,0) + R1 — reference pggan/loss.py:16-27"""
        L_D_real = Loss.get_BCE_loss(D_dict["pred_real"], True)
        L_D_fake = Loss.get_BCE_loss(D_dict["pred_fake"], False)
        L_reg = Loss.get_r1_reg(L_D_real, D_dict["img_real"])
        L_D = L_D_real + L_D_fake + L_reg
        self.loss_dict["L_D_real"] = L_D_real.detach().mean()
        self.loss_dict["L_D_fake"] = L_D_fake.detach().mean()
        self.loss_dict["L_D"] = L_D.detach()
        return L_D

    def get_loss_D_wgangp(self, D_dict, discriminator):
        """the dormant WGAN-GP discriminator loss of reference pggan/loss.py:29-51 (commented out
        there): BCE adversarial terms; GP back-propagates itself; drift is a reported value."""
        self.get_gradient_penalty(D_dict, discriminator, lazy=True)
        return self.get_loss_D_wgangp_terms(D_dict)

    def get_loss_D_wgangp_terms(self, D_dict):
        """the part of the above that is returned for `.backward()` (call get_gradient_penalty first)"""
        L_D_real = Loss.get_BCE_loss(D_dict["pred_real"], True)
        L_D_fake = Loss.get_BCE_loss(D_dict["pred_fake"], False)
        L_D = L_D_real + L_D_fake
        L_D_gp = getattr(self, "_last_gp", None)
        L_D_eps = getattr(self, "_fused_drift", None)
        self.loss_dict["L_D_real"] = L_D_real.detach().mean()
        self.loss_dict["L_D_fake"] = L_D_fake.detach().mean()
        total = L_D.detach()
        if L_D_gp is not None:
            self.loss_dict["L_D_gp"] = L_D_gp
            total = total + L_D_gp
        if L_D_eps is not None:
            self.loss_dict["L_D_eps"] = L_D_eps
            total = total + L_D_eps
        self.loss_dict["L_D"] = total
        return L_D

    def get_gradient_penalty(self, D_dict, discriminator, backward=True, eps=None, lazy=False, loss_scale=1.0):
        """reference pggan/loss.py:54-92.  W_gp * sum_b (||grad_xhat D(xhat)_b||_2 - 1)^2 on
        xhat = eps*real + (1-eps)*fake, eps ~ U(0,1) per sample drawn on the CPU generator exactly where
        the reference draws it (pass `eps` [B,1] to supply it).  Back-propagates itself when
        `backward` (of loss_scale * penalty: the trainer's static loss scale; the returned value is unscaled).
        Returns the float the reference returns (or a 0-dim tensor if `lazy`)."""
        gradient_penalty = None
        if self.args.W_gp:
            img_real, img_fake = D_dict["img_real"], D_dict["img_fake"]
            batchSize = img_real.size(0)
            if eps is None:
                eps = torch.rand(batchSize, 1)
            eps = eps.to(img_real.device, non_blocking=True).view(batchSize, 1, 1, 1)
            interpolates = eps * img_real + ((1 - eps) * img_fake)
            if not interpolates.requires_grad:          # the reference relies on img_real.requires_grad_()
                interpolates.requires_grad_(True)
            decisionInterpolate = discriminator(interpolates)
            decisionInterpolate = decisionInterpolate[:, 0].sum()
            # half-precision gradient tensors: seed the first-order chain with a power of two instead of 1 (at 1024x1024
            # the per-element gradients are ~1e-7, deep in fp16's subnormal range); the reduction divides it out
            seed = 1024.0 if (interpolates.is_cuda and ops._REAL[ops._fmt(ops.G)] == torch.float16) else 1.0
            with ops.input_grad_only():
                gradients = torch.autograd.grad(outputs=decisionInterpolate * seed if seed != 1.0 else decisionInterpolate,
                                                inputs=interpolates, create_graph=True, retain_graph=True)[0]
            pred = D_dict.get("pred_real") if self.args.W_drift_D else None
            res = ops.GradientPenalty.apply(gradients, None if pred is None else pred[:, 0], float(self.args.W_gp),
                                            float(self.args.W_drift_D or 0.0), seed)
            gradient_penalty = res[0]
            self._last_gp = res[0].detach()
            self._fused_drift = res[1].detach() if pred is not None else None
            if backward:
                (gradient_penalty * loss_scale if loss_scale != 1.0 else gradient_penalty).backward(retain_graph=True)
        if gradient_penalty is None:
            return None
        return gradient_penalty.detach() if lazy else gradient_penalty.item()

    def get_drift_loss(self, D_dict, lazy=False):
        """reference pggan/loss.py:94-100: W_drift_D * sum pred_real^2, value only"""
        if self.args.W_drift_D:
            drift = (D_dict["pred_real"].detach() ** 2).sum() * self.args.W_drift_D
            return drift if lazy else drift.item()
