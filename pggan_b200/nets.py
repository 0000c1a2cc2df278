"""Generator / Discriminator — mirror of reference pggan/nets.py (:10-276).

Same constructor signatures, attributes (`alpha`, `blocks`, `toRGB_blocks`, `fromRGB_blocks`,
`first_block`, `latent_format_layer`, `minibatch_normalization_block`, `decision_layer`,
`block_depths` / `depths`), `add_block` semantics and state_dict keys, so reference checkpoints load
unchanged.  `alpha` is a plain mutable float read at call time (a runtime kernel scalar).
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from .blocks import (ProgressiveGeneratorBlock, ProgressiveDiscriminatorBlock, LastProgressiveDiscriminatorBlock,
                     toRGBBlock, fromRGBBlock)
from .layers import EqualizedLinear, PixelwiseVectorNorm
from .utils import num_flat_features


class _AlphaMixin:
    """`alpha` stays the plain mutable float attribute of the reference (pggan/model.py:146-147,179-182 assign it from
    outside).  The kernels read it from a one-element device tensor that is refreshed whenever the float changed, so a
    captured CUDA graph follows change_alpha() without being captured again."""

    def alpha_ref(self) -> "ops.AlphaRef":
        dev = next(self.parameters()).device
        a = float(self.alpha)
        if dev.type != "cuda":
            return ops.AlphaRef(a, None)
        st = self.__dict__
        t = st.get("_alpha_dev")
        if t is None or t.device != dev:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("call alpha_ref() once before capturing a CUDA graph (the device scalar must exist)")
            st["_alpha_dev"] = t = torch.full((1,), a, dtype=torch.float32, device=dev)
            st["_alpha_host"] = a
        elif st["_alpha_host"] != a:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("alpha changed inside a CUDA graph capture")
            t.fill_(a)
            st["_alpha_host"] = a
        return ops.AlphaRef(a, t)


class Generator(_AlphaMixin, nn.Module):
    """reference pggan/nets.py:10-161"""

    def __init__(self, latent_dim, first_depth, init_bias_to_zero=True, LReLU_slope=0.2, apply_pixel_norm=True,
                 last_activation=None, output_dim=3, equalized_lr=True):
        super().__init__()
        self.equalized_lr = equalized_lr
        self.init_bias_to_zero = init_bias_to_zero
        self.leakyRelu = torch.nn.LeakyReLU(LReLU_slope)
        self.apply_pixel_norm = apply_pixel_norm
        self.pixel_norm = None
        if apply_pixel_norm:
            self.pixel_norm = PixelwiseVectorNorm()
        self.block_depths = [first_depth]
        self.first_depth = first_depth
        self.blocks = nn.ModuleList()
        self.toRGB_blocks = nn.ModuleList()
        self.init_format_layer(latent_dim)
        self.output_dim = output_dim
        self.first_block = ProgressiveGeneratorBlock(first_depth, first_depth, equalized_lr=self.equalized_lr,
                                                     init_bias_to_zero=self.init_bias_to_zero, norm=self.pixel_norm,
                                                     is_first=True)
        self.toRGB_blocks.append(toRGBBlock(first_depth, output_dim=self.output_dim, equalized_lr=self.equalized_lr,
                                            init_bias_to_zero=self.init_bias_to_zero))
        self.alpha = 0
        self.last_activation = last_activation

    def init_format_layer(self, latent_dim):
        """reference pggan/nets.py:83-93 (always equalized)"""
        self.latent_dim = latent_dim
        self.latent_format_layer = EqualizedLinear(self.latent_dim, 16 * self.first_depth,
                                                   init_bias_to_zero=self.init_bias_to_zero)

    def get_output_size(self):
        side = 4 * (2 ** (len(self.toRGB_blocks) - 1))
        return (side, side)

    def add_block(self, new_depth):
        """reference pggan/nets.py:102-119.  New modules are created on the device of the existing ones."""
        prev_depth = self.block_depths[-1]
        dev = self.latent_format_layer.module.weight.device
        self.block_depths.append(new_depth)
        self.blocks.append(ProgressiveGeneratorBlock(prev_depth, new_depth, equalized_lr=self.equalized_lr,
                                                     init_bias_to_zero=self.init_bias_to_zero,
                                                     norm=self.pixel_norm).to(dev))
        self.toRGB_blocks.append(toRGBBlock(new_depth, output_dim=self.output_dim, equalized_lr=self.equalized_lr,
                                            init_bias_to_zero=self.init_bias_to_zero).to(dev))

    def tiny(self) -> bool:
        """4x4 output: the network runs in the fp32 family whatever the precision mode (see ops.py)"""
        return len(self.blocks) == 0

    def live_parameters(self):
        """parameters that take part in forward at the current scale (RGB layers older than [-2] are
        stale — SURVEY.md D16), in reverse execution order (the order their gradients become ready)."""
        mods = [self.latent_format_layer, self.first_block] + list(self.blocks)
        rgb = list(self.toRGB_blocks[-2:]) if len(self.blocks) else [self.toRGB_blocks[-1]]
        out = []
        for m in reversed(mods + rgb):
            out.extend(reversed(list(m.parameters())))
        return out

    def forward(self, x):
        """reference pggan/nets.py:121-161"""
        pn = self.pixel_norm if self.apply_pixel_norm else None
        x = x.reshape(-1, num_flat_features(x))
        # the one place where the latent turns into the precision mode's 16-bit format; a 4x4 network (no blocks yet)
        # stays in fp32 (ops.py: precision follows the tensor)
        x = ops.enter(x, ops.A, f32=self.tiny())
        if pn is not None:
            x = ops.PixelNorm.apply(x, 1e-8)
        # format layer: Linear(latent -> 16*C) + LReLU, written directly in NHWC order by permuting the
        # output rows of the weight ((c,h,w) -> (h,w,c)); views only, so gradients map back for free
        lin = self.latent_format_layer.module
        C0 = self.first_depth
        slope = float(self.leakyRelu.negative_slope)
        layout = ("lin_out_perm", C0)
        if pn is not None:
            # pixel norm over the C0 channels of each of the 16 positions == over each contiguous C0 chunk
            # of the (h,w,c)-ordered output: run the fused conv+act+PN on a [B*16, C0] view
            x = ops.ConvF.apply(x, lin.weight, lin.bias, self.latent_format_layer.scale(), slope, False, layout)
            x = x.view(x.shape[0], 4, 4, C0).permute(0, 3, 1, 2)       # [B,C,4,4] logical, NHWC physical
            x = ops.PixelNorm.apply(x, 1e-8)
        else:
            x = ops.ConvF.apply(x, lin.weight, lin.bias, self.latent_format_layer.scale(), slope, False, layout)
            x = x.view(x.shape[0], 4, 4, C0).permute(0, 3, 1, 2)
        x = self.first_block(x)
        low = None
        nb = len(self.blocks)
        if nb == 1:
            low = self.toRGB_blocks[-2](x, apply_upscale=False)
        for i, block in enumerate(self.blocks):
            x = block(x)
            if i == nb - 2:
                low = self.toRGB_blocks[-2](x, apply_upscale=False)
        x = self.toRGB_blocks[-1](x)
        if nb:
            # (1-alpha) * upscale2d(low) + alpha * x in one pass (pggan/nets.py:155-156)
            x = ops.ImageFade.apply(x, low, self.alpha_ref())
        if self.last_activation is not None:
            x = self.last_activation(x)
        return x


class Discriminator(_AlphaMixin, nn.Module):
    """reference pggan/nets.py:164-276"""

    def __init__(self, last_depth, init_bias_to_zero=True, LReLU_slope=0.2, decision_layer_size=1,
                 apply_minibatch_norm=False, input_dim=3, equalized_lr=True):
        super().__init__()
        self.init_bias_to_zero = init_bias_to_zero
        self.equalized_lr = equalized_lr
        self.input_dim = input_dim
        self.depths = [last_depth]
        self.blocks = nn.ModuleList()
        self.fromRGB_blocks = nn.ModuleList()
        self.mergeLayers = nn.ModuleList()
        self.init_decision_layer(decision_layer_size)
        self.apply_minibatch_norm = apply_minibatch_norm
        self.minibatch_normalization_block = LastProgressiveDiscriminatorBlock(
            last_depth, equalized_lr=equalized_lr, init_bias_to_zero=init_bias_to_zero,
            apply_minibatch_norm=apply_minibatch_norm)
        self.fromRGB_blocks.append(fromRGBBlock(input_dim, last_depth, equalized_lr=equalized_lr,
                                                init_bias_to_zero=init_bias_to_zero))
        self.alpha = 0
        self.leakyRelu = torch.nn.LeakyReLU(LReLU_slope)

    def add_block(self, new_depth):
        """reference pggan/nets.py:227-239"""
        prev_depth = self.depths[-1]
        dev = self.decision_layer.module.weight.device
        self.depths.append(new_depth)
        self.blocks.append(ProgressiveDiscriminatorBlock(new_depth, prev_depth, equalized_lr=self.equalized_lr,
                                                         init_bias_to_zero=self.init_bias_to_zero).to(dev))
        self.fromRGB_blocks.append(fromRGBBlock(self.input_dim, new_depth, equalized_lr=self.equalized_lr,
                                                init_bias_to_zero=self.init_bias_to_zero).to(dev))

    def init_decision_layer(self, decision_layer_size):
        """reference pggan/nets.py:241-246"""
        self.decision_layer = EqualizedLinear(self.depths[0], decision_layer_size, equalized=self.equalized_lr,
                                              init_bias_to_zero=self.init_bias_to_zero)

    def tiny(self) -> bool:
        """4x4 input: the network runs in the fp32 family whatever the precision mode (see ops.py)"""
        return len(self.blocks) == 0

    def live_parameters(self):
        """live parameters in reverse execution order (decision layer first) — SURVEY.md D16 / §5"""
        rgb = list(self.fromRGB_blocks[-2:]) if len(self.blocks) else [self.fromRGB_blocks[-1]]
        mods = rgb + list(reversed(self.blocks)) + [self.minibatch_normalization_block, self.decision_layer]
        out = []
        for m in reversed(mods):
            out.extend(reversed(list(m.parameters())))
        return out

    def forward(self, x, get_feature=False):
        """reference pggan/nets.py:248-276"""
        img = x
        x_down = None
        if len(self.blocks):
            x_down = self.fromRGB_blocks[-2](img, apply_downscale=True)
        # with blocks present, fromRGB's LeakyReLU backward is applied by the first block's conv1 dgrad epilogue
        fuse_rgb = len(self.blocks) > 0 and self.blocks[-1].first_dgrad_fuses_mask(img.shape[0], img.shape[2], img.shape[3])
        x = self.fromRGB_blocks[-1](img, defer_mask=fuse_rgb, f32=self.tiny())
        apply_merge = len(self.blocks) > 0
        for block in reversed(self.blocks):
            if apply_merge:
                apply_merge = False
                x = block(x, blend_with=x_down, alpha=self.alpha_ref(), x_act_slope=0.2 if fuse_rgb else None)
            else:
                x = block(x)
        x = self.minibatch_normalization_block(x)
        out = self.decision_layer.fused(x, 1.0, raw_out=True)
        if not get_feature:
            return out
        return out, ops.ExportFloat.apply(x)
