"""CPU oracle for the PGGAN D+G training step (TEST INFRASTRUCTURE — not a product path).

A plain-PyTorch, functional restatement of the reference's hot path.  Nothing in
``pggan_b200/`` may import this module: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs do.

Parity pinning: the reference ships no golden vectors, so this
restatement is pinned against the reference's *own modules executed in the build
container* — ``oracle/make_golden.py`` imports ``/root/reference`` and writes the
fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every
function below against them.

Networks are expressed over a ``state_dict`` with the reference's parameter names
(SURVEY.md Appendix C), so the same dict drives the reference modules, this oracle
and the CUDA modules.

Reference citations (file:line, relative to the reference repo):
  he_constant            lib/layers.py:17-25
  eq_conv2d / eq_linear  lib/layers.py:58-63  (y = c * (module(x)); bias is scaled too)
  pixel_norm             lib/layers.py:13-14
  upscale2d/downscale2d  lib/utils.py:106-124
  mbstd_concat           lib/blocks.py:204-233
  generator_forward      pggan/nets.py:121-161, lib/blocks.py:145-170
  discriminator_forward  pggan/nets.py:248-276, lib/blocks.py:197-201,258-268,285-292
  bce_logits / r1_reg    lib/loss.py:119-135
  loss_G / loss_D_live   pggan/loss.py:5-27
  gradient_penalty       pggan/loss.py:54-92   (device line fixed for CPU)
  drift_loss             pggan/loss.py:94-100
  real_image_fade        pggan/model.py:217-221
  train_step             pggan/model.py:206-255 + lib/utils.py:72-75 + lib/model.py:95-97
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

# ---------------------------------------------------------------------------- emulation
# emulate(torch.bfloat16) makes the network functions below round tensors to bf16 at exactly the
# points where the CUDA path stores bf16 (weights of every tensor-core GEMM, every feature map written
# to HBM) while all arithmetic stays fp32 — i.e. "the reference's algorithm evaluated on bf16-rounded
# operands with fp32 accumulation".  With it, LeakyReLU masks agree with the CUDA path, so gradients
# can be compared tightly (against the plain fp32 oracle, gradient parity is limited by mask flips:
# rel err ~ 0.8*sqrt(fraction of flipped pre-activation signs)).
_EMU = None


class emulate:
    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        global _EMU
        self.prev = _EMU
        _EMU = self.dtype if self.dtype in (torch.bfloat16, torch.float16) else None

    def __exit__(self, *a):
        global _EMU
        _EMU = self.prev


def _q(t: Tensor) -> Tensor:
    return t if _EMU is None else t.to(_EMU).float()


# --------------------------------------------------------------------------- layers
def he_constant(weight: Tensor) -> float:
    fan_in = 1
    for s in weight.shape[1:]:
        fan_in *= int(s)
    return math.sqrt(2.0 / fan_in)


def eq_conv2d(x: Tensor, w: Tensor, b: Tensor, padding: int) -> Tensor:
    return F.conv2d(x, w, b, padding=padding) * he_constant(w)


def eq_linear(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    return F.linear(x, w, b) * he_constant(w)


def pixel_norm(x: Tensor, epsilon: float = 1e-8) -> Tensor:
    return x * torch.rsqrt((x * x).mean(dim=1, keepdim=True) + epsilon)


def upscale2d(x: Tensor, factor: int = 2) -> Tensor:
    if factor == 1:
        return x
    return x.repeat_interleave(factor, dim=2).repeat_interleave(factor, dim=3)


def downscale2d(x: Tensor, factor: int = 2) -> Tensor:
    if factor == 1:
        return x
    return F.avg_pool2d(x, (factor, factor))


def mbstd_concat(x: Tensor, subgroup_size: int = 4) -> Tensor:
    n, c, h, w = x.shape
    g = min(n, subgroup_size)
    if n % g != 0:
        g = n
    if g > 1:
        y = x.reshape(n // g, g, c, h, w)
        y = torch.sqrt(y.var(dim=1, unbiased=True) + 1e-8)        # [n/g, c, h, w]
        y = y.reshape(n // g, -1).mean(dim=1)                     # [n/g]
        y = y.repeat_interleave(g).reshape(n, 1, 1, 1).expand(n, 1, h, w)
    else:
        y = x.new_zeros(n, 1, h, w)
    return torch.cat([x, y], dim=1)


def lrelu(x: Tensor, slope: float = 0.2) -> Tensor:
    return F.leaky_relu(x, slope)


# ------------------------------------------------------------------------- networks
def _p(sd: StateDict, prefix: str) -> Tuple[Tensor, Tensor]:
    return sd[prefix + ".module.weight"], sd[prefix + ".module.bias"]


def count_blocks(sd: StateDict) -> int:
    n = 0
    while f"blocks.{n}.block.0.module.weight" in sd:
        n += 1
    return n


def generator_forward(sd: StateDict, z: Tensor, alpha: float, slope: float = 0.2,
                      apply_pixel_norm: bool = True) -> Tensor:
    nb = count_blocks(sd)
    pn = (lambda t: _q(pixel_norm(t))) if apply_pixel_norm else (lambda t: t)

    def conv_act(t, prefix, pad, sl):
        w, b = _p(sd, prefix)
        return _q(lrelu(eq_conv2d(t, _q(w), b, pad), sl))

    x = pn(_q(z))
    x = x.reshape(x.shape[0], -1)
    w, b = _p(sd, "latent_format_layer")
    x = _q(lrelu(eq_linear(x, _q(w), b), slope))
    x = x.reshape(x.shape[0], -1, 4, 4)
    x = pn(x)
    x = pn(conv_act(x, "first_block.block.0", 1, 0.2))
    x_up = None
    if nb == 1:
        x_up = upscale2d(eq_conv2d(x, *_p(sd, f"toRGB_blocks.{nb - 1}.toRGB"), 0))
    second = 3 if apply_pixel_norm else 2          # Sequential index of the 2nd conv
    for i in range(nb):
        x = upscale2d(x)
        x = pn(conv_act(x, f"blocks.{i}.block.0", 1, 0.2))
        x = pn(conv_act(x, f"blocks.{i}.block.{second}", 1, 0.2))
        if i == nb - 2:
            x_up = upscale2d(eq_conv2d(x, *_p(sd, f"toRGB_blocks.{nb - 1}.toRGB"), 0))
    x = eq_conv2d(x, *_p(sd, f"toRGB_blocks.{nb}.toRGB"), 0)
    if nb:
        x = (1.0 - alpha) * x_up + alpha * x
    return x


def discriminator_forward(sd: StateDict, x: Tensor, alpha: float, get_feature: bool = False,
                          apply_minibatch_norm: bool = True):
    nb = count_blocks(sd)

    def conv_act(t, prefix, pad):
        w, b = _p(sd, prefix)
        return _q(lrelu(eq_conv2d(t, _q(w), b, pad), 0.2))

    x_down = None
    if nb:      # the RGB 1x1 layers run in fp32 on fp32 weights; only their output is stored rounded
        x_down = _q(lrelu(eq_conv2d(downscale2d(x), *_p(sd, f"fromRGB_blocks.{nb - 1}.fromRGB"), 0), 0.2))
    x = _q(lrelu(eq_conv2d(x, *_p(sd, f"fromRGB_blocks.{nb}.fromRGB"), 0), 0.2))
    merge = nb > 0
    for i in reversed(range(nb)):
        x = conv_act(x, f"blocks.{i}.block.0", 1)
        x = conv_act(x, f"blocks.{i}.block.2", 1)
        x = F.avg_pool2d(x, (2, 2))
        if merge:
            merge = False
            x = (1 - alpha) * x_down + alpha * x
        x = _q(x)
    if apply_minibatch_norm:
        x = _q(mbstd_concat(x, 4))
    x = conv_act(x, "minibatch_normalization_block.conv", 1)
    x = x.reshape(x.shape[0], -1)
    w, b = _p(sd, "minibatch_normalization_block.linear")
    x = _q(lrelu(eq_linear(x, _q(w), b), 0.2))
    w, b = _p(sd, "decision_layer")
    out = eq_linear(x, _q(w), b)
    return (out, x) if get_feature else out


# --------------------------------------------------------------------------- losses
def bce_logits(logits: Tensor, target: int) -> Tensor:
    assert target in (0, 1)
    return F.binary_cross_entropy_with_logits(logits, torch.full_like(logits, float(target)))


def r1_reg(d_out: Tensor, x_in: Tensor) -> Tensor:
    (g,) = torch.autograd.grad(d_out.sum(), x_in, create_graph=True, retain_graph=True)
    return 0.5 * g.pow(2).reshape(x_in.shape[0], -1).sum(1).mean(0)


def loss_G(pred_fake: Tensor, w_adv: float = 1.0) -> Tensor:
    return w_adv * bce_logits(pred_fake, 1)


def loss_D_live(pred_real: Tensor, pred_fake: Tensor, img_real: Tensor) -> Dict[str, Tensor]:
    l_real = bce_logits(pred_real, 1)
    l_fake = bce_logits(pred_fake, 0)
    l_reg = r1_reg(l_real, img_real)
    return {"L_D_real": l_real, "L_D_fake": l_fake, "L_reg": l_reg, "L_D": l_real + l_fake + l_reg}


def gradient_penalty(d_fn, img_real: Tensor, img_fake: Tensor, w_gp: float = 10.0,
                     eps: Optional[Tensor] = None, backward: bool = True) -> Tensor:
    """Returns the penalty tensor (the reference returns ``.item()`` of it).  ``eps`` is
    the per-sample U(0,1) draw of shape [B,1]; if None it is drawn from the global CPU
    generator exactly where the reference draws it."""
    b = img_real.shape[0]
    if eps is None:
        eps = torch.rand(b, 1)
    eps = eps.to(img_real.device).reshape(b, 1, 1, 1).expand_as(img_real)
    x_hat = eps * img_real + (1 - eps) * img_fake
    d = d_fn(x_hat)[:, 0].sum()
    (g,) = torch.autograd.grad(d, x_hat, create_graph=True, retain_graph=True)
    norms = (g.reshape(b, -1) ** 2).sum(dim=1).sqrt()
    gp = ((norms - 1.0) ** 2).sum() * w_gp
    if backward:
        gp.backward(retain_graph=True)
    return gp.detach()


def drift_loss(pred_real: Tensor, w_drift: float = 0.001) -> Tensor:
    return (pred_real.detach() ** 2).sum() * w_drift


def real_image_fade(img: Tensor, alpha: float) -> Tensor:
    low = F.interpolate(F.avg_pool2d(img, (2, 2)), scale_factor=2, mode="nearest")
    return (1 - alpha) * low + alpha * img


# ------------------------------------------------------------------------ parameters
def init_generator_sd(latent_dim: int, depths: List[int], n_blocks: int, output_dim: int = 3,
                      generator: Optional[torch.Generator] = None) -> StateDict:
    """Reference initialisation (N(0,1) weights, zero biases — lib/layers.py:51-56)."""
    def w(*shape):
        return torch.randn(*shape, generator=generator)
    d0 = depths[0]
    sd: StateDict = {}

    def add(prefix, shape):
        sd[prefix + ".module.weight"] = w(*shape)
        sd[prefix + ".module.bias"] = torch.zeros(shape[0])

    add("latent_format_layer", (16 * d0, latent_dim))
    add("first_block.block.0", (d0, d0, 3, 3))
    add("toRGB_blocks.0.toRGB", (output_dim, d0, 1, 1))
    for i in range(n_blocks):
        cp, cn = depths[i], depths[i + 1]
        add(f"blocks.{i}.block.0", (cn, cp, 3, 3))
        add(f"blocks.{i}.block.3", (cn, cn, 3, 3))
        add(f"toRGB_blocks.{i + 1}.toRGB", (output_dim, cn, 1, 1))
    return sd


def init_discriminator_sd(depths: List[int], n_blocks: int, input_dim: int = 3,
                          apply_minibatch_norm: bool = True, decision_layer_size: int = 1,
                          generator: Optional[torch.Generator] = None) -> StateDict:
    def w(*shape):
        return torch.randn(*shape, generator=generator)
    d0 = depths[0]
    sd: StateDict = {}

    def add(prefix, shape):
        sd[prefix + ".module.weight"] = w(*shape)
        sd[prefix + ".module.bias"] = torch.zeros(shape[0])

    add("decision_layer", (decision_layer_size, d0))
    add("minibatch_normalization_block.conv", (d0, d0 + (1 if apply_minibatch_norm else 0), 3, 3))
    add("minibatch_normalization_block.linear", (d0, 16 * d0))
    add("fromRGB_blocks.0.fromRGB", (d0, input_dim, 1, 1))
    for i in range(n_blocks):
        cp, cn = depths[i], depths[i + 1]
        add(f"blocks.{i}.block.0", (cn, cn, 3, 3))
        add(f"blocks.{i}.block.2", (cp, cn, 3, 3))
        add(f"fromRGB_blocks.{i + 1}.fromRGB", (cn, input_dim, 1, 1))
    return sd


def live_param_names(sd: StateDict, net: str) -> List[str]:
    """Parameters that receive gradients (RGB layers older than [-2] are stale)."""
    nb = count_blocks(sd)
    rgb = "toRGB_blocks" if net == "G" else "fromRGB_blocks"
    out = []
    for k in sd:
        if k.startswith(rgb + "."):
            idx = int(k.split(".")[1])
            if idx < nb - 1:
                continue
        out.append(k)
    return out


# ------------------------------------------------------------------------- train step
class OracleTrainer:
    """Restatement of ProgressiveGAN.train_step (pggan/model.py:206-255).

    variant "L": the live step — BCE(real,1)+BCE(fake,0)+R1 (pggan/loss.py:16-27).
    variant "W": the WGAN-GP step of the commented-out caller (pggan/loss.py:29-51) wired the
                 way SURVEY.md §8(d) fixes it: zero_grad -> GP.backward -> loss_D.backward -> step,
                 drift reported as a value.
    """

    def __init__(self, g_sd: StateDict, d_sd: StateDict, alpha: float = 0.0, variant: str = "W",
                 lr_G: float = 1e-4, lr_D: float = 1e-5, beta1: float = 0.0, beta2: float = 0.99,
                 w_adv: float = 1.0, w_gp: float = 10.0, w_drift: float = 0.001, latent_dim: int = 512):
        self.g = {k: v.clone().requires_grad_(True) for k, v in g_sd.items()}
        self.d = {k: v.clone().requires_grad_(True) for k, v in d_sd.items()}
        self.alpha = alpha
        self.variant = variant
        self.w_adv, self.w_gp, self.w_drift = w_adv, w_gp, w_drift
        self.latent_dim = latent_dim
        self.opt_G = torch.optim.Adam(list(self.g.values()), lr=lr_G, betas=(float(beta1), float(beta2)))
        self.opt_D = torch.optim.Adam(list(self.d.values()), lr=lr_D, betas=(float(beta1), float(beta2)))
        self.loss_dict: Dict[str, float] = {}

    def G(self, z):
        return generator_forward(self.g, z, self.alpha)

    def D(self, x, get_feature=False):
        return discriminator_forward(self.d, x, self.alpha, get_feature)

    def train_step(self, img_real: Tensor, z_d: Optional[Tensor] = None, z_g: Optional[Tensor] = None,
                   eps: Optional[Tensor] = None):
        n = img_real.shape[0]
        dev = img_real.device
        if count_blocks(self.d):
            img_real = real_image_fade(img_real, self.alpha)
        img_real = img_real.detach().requires_grad_()
        pred_real = self.D(img_real)
        if z_d is None:
            z_d = torch.randn(n, self.latent_dim)
        img_fake = self.G(z_d.to(dev)).detach()
        pred_fake = self.D(img_fake)
        self.opt_D.zero_grad()
        if self.variant == "L":
            ld = loss_D_live(pred_real, pred_fake, img_real)
            loss_D = ld["L_D"]
        else:
            l_real = bce_logits(pred_real, 1)
            l_fake = bce_logits(pred_fake, 0)
            loss_D = l_real + l_fake
            gp = gradient_penalty(self.D, img_real, img_fake, self.w_gp, eps=eps, backward=True)
            drift = drift_loss(pred_real, self.w_drift)
            self.loss_dict["L_D_gp"] = float(gp)
            self.loss_dict["L_D_eps"] = float(drift)
        loss_D.backward()
        self.opt_D.step()
        self.loss_dict["L_D"] = float(loss_D.detach())

        if z_g is None:
            z_g = torch.randn(n, self.latent_dim)
        img_fake = self.G(z_g.to(dev))
        pred_fake, _ = self.D(img_fake, True)
        l_g = loss_G(pred_fake, self.w_adv)
        self.opt_G.zero_grad()
        l_g.backward()
        self.opt_G.step()
        self.loss_dict["L_G"] = float(l_g.detach())
        return [img_real.detach(), img_fake.detach()]
