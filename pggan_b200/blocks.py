"""PGGAN building blocks — mirror of the PGGAN half of reference lib/blocks.py (:113-292).

Module trees, constructor signatures (including the flipped argument order of the two Progressive
blocks) and state_dict keys are the reference's; `forward` runs the fused kernels.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .layers import EqualizedConv2d, EqualizedLinear, PixelwiseVectorNorm
from .utils import downscale2d, upscale2d, num_flat_features  # noqa: F401


def _conv_act_norm(layer, x, norm, upscale_input=False):
    """[upscale2d ->] conv3x3 -> c*(.+b) -> LeakyReLU(0.2) [-> pixel norm]   (reference lib/blocks.py:119-151)"""
    if isinstance(norm, PixelwiseVectorNorm):
        return layer.fused(x, 0.2, pixel_norm_eps=1e-8, upscale_input=upscale_input)
    x = layer.fused(x, 0.2, upscale_input=upscale_input)
    if norm is not None:
        x = norm(x)
    return x


class ProgressiveGeneratorBlock(nn.Module):
    """reference lib/blocks.py:113-151: [upscale2d] -> (conv3x3 -> LReLU -> PN) x2  (is_first: one conv)"""

    def __init__(self, prev_depth, new_depth, equalized_lr=True, init_bias_to_zero=True, norm=None, is_first=False):
        super().__init__()
        block = []
        if not is_first:
            block.append(EqualizedConv2d(prev_depth, new_depth, 3, padding=1, equalized=equalized_lr,
                                         init_bias_to_zero=init_bias_to_zero))
            block.append(nn.LeakyReLU(0.2))
            if norm:
                block.append(norm)
        block.append(EqualizedConv2d(new_depth, new_depth, 3, padding=1, equalized=equalized_lr,
                                     init_bias_to_zero=init_bias_to_zero))
        block.append(nn.LeakyReLU(0.2))
        if norm:
            block.append(norm)
        self.block = nn.Sequential(*block)
        self.is_first = is_first

    def _convs_and_norm(self):
        convs = [m for m in self.block if isinstance(m, EqualizedConv2d)]
        norm = next((m for m in self.block if isinstance(m, PixelwiseVectorNorm)), None)
        return convs, norm

    def forward(self, x):
        convs, norm = self._convs_and_norm()
        for i, conv in enumerate(convs):
            # upscale2d (lib/utils.py:106-118) is the PROLOGUE of the block's first convolution
            x = _conv_act_norm(conv, x, norm, upscale_input=(i == 0 and not self.is_first))
        return x


class toRGBBlock(nn.Module):
    """reference lib/blocks.py:153-170: 1x1 conv C -> output_dim (no activation) [+ upscale2d]"""

    def __init__(self, new_depth, output_dim=3, equalized_lr=True, init_bias_to_zero=True):
        super().__init__()
        self.toRGB = EqualizedConv2d(new_depth, output_dim, 1, equalized=equalized_lr,
                                     init_bias_to_zero=init_bias_to_zero)

    def forward(self, x, apply_upscale=False):
        m = self.toRGB.module
        W = m.weight
        Wc = W.view(W.shape[0], W.shape[1]).t()
        # the 1x1 conv commutes with nearest up-sampling, so the up-scaled image is written directly
        return ops.RgbReduce.apply(x, Wc, m.bias, bool(apply_upscale), self.toRGB.scale(), 1.0)


class ProgressiveDiscriminatorBlock(nn.Module):
    """reference lib/blocks.py:173-201: conv3x3(C->C) -> LReLU -> conv3x3(C->Cprev) -> LReLU -> AvgPool2"""

    def __init__(self, new_depth, prev_depth, equalized_lr=True, init_bias_to_zero=True):
        super().__init__()
        self.block = nn.Sequential(
            EqualizedConv2d(new_depth, new_depth, 3, padding=1, equalized=equalized_lr,
                            init_bias_to_zero=init_bias_to_zero),
            nn.LeakyReLU(0.2),
            EqualizedConv2d(new_depth, prev_depth, 3, padding=1, equalized=equalized_lr,
                            init_bias_to_zero=init_bias_to_zero),
            nn.LeakyReLU(0.2),
            nn.AvgPool2d((2, 2)),
        )

    def first_dgrad_fuses_mask(self, batch, h, w) -> bool:
        """may the producer of this block's input (a fromRGB block) defer its LeakyReLU mask to conv1's dgrad?"""
        c1 = self.block[0].module
        return ops.dgrad_mask_fusable(batch, h, w, c1.in_channels, c1.out_channels)

    def forward(self, x, blend_with=None, alpha=1.0, x_act_slope=None):
        """blend_with/alpha: optional feature-space fade fused with the pooling
        ((1-alpha)*blend_with + alpha*pool(x), reference pggan/nets.py:262-265).
        x_act_slope: x comes straight from a fromRGB block run with defer_mask=True.
        The three LeakyReLU backward passes of the block are fused away: conv1's mask is applied in conv2's dgrad
        epilogue, conv2's in the pooling backward, fromRGB's (if any) in conv1's dgrad epilogue."""
        # ...where that epilogue is free; a dgrad that runs on the swapped-orientation kernel keeps the separate pass
        b, _, h, w = x.shape
        c2 = self.block[2].module
        fuse2 = ops.dgrad_mask_fusable(b, h, w, c2.in_channels, c2.out_channels, f32=(x.dtype == torch.float32))
        x = self.block[0].fused(x, 0.2, x_act_slope=x_act_slope, defer_mask=fuse2)
        x = self.block[2].fused(x, 0.2, x_act_slope=0.2 if fuse2 else None, defer_mask=True)
        if blend_with is not None:
            return ops.PoolBlend.apply(x, blend_with, alpha if isinstance(alpha, ops.AlphaRef) else float(alpha), 0.2)
        return ops.Pool2x2.apply(x, 0.25, 0.2)


def _stddev_channel(x, subgroup_size=4):
    """the extra channel of reference lib/blocks.py:216-231 as a [N,1,H,W] tensor"""
    size = x.size()
    subgroup_size = min(size[0], subgroup_size)
    if size[0] % subgroup_size != 0:
        subgroup_size = size[0]
    if subgroup_size > 1:
        stat = ops.MbstdStat.apply(x, int(subgroup_size))                 # [N/G] fp32
        return ops.StatChannel.apply(stat, int(subgroup_size), size[0], size[2], size[3], x.dtype == torch.float32)
    return torch.zeros(size[0], 1, size[2], size[3], device=x.device, dtype=x.dtype)


def concatenate_stddev_channel(x, subgroup_size=4):
    """reference lib/blocks.py:204-233: unbiased variance over groups of `subgroup_size` consecutive
    samples, eps 1e-8 inside the sqrt, mean over (C,H,W); one group of N if N % size != 0; zeros if N == 1."""
    x = ops.feat(x)
    return ops.export_act(torch.cat([x, _stddev_channel(x, subgroup_size)], dim=1))


class LastProgressiveDiscriminatorBlock(nn.Module):
    """reference lib/blocks.py:236-268: mbstd -> conv3x3(C+1->C)@4x4 -> LReLU -> flatten -> Linear -> LReLU"""

    def __init__(self, depth, equalized_lr=True, init_bias_to_zero=True, apply_minibatch_norm=False):
        super().__init__()
        self.conv_list = []
        entry_dim = depth + 1 if apply_minibatch_norm else depth
        self.conv = EqualizedConv2d(entry_dim, depth, 3, padding=1, equalized=equalized_lr,
                                    init_bias_to_zero=init_bias_to_zero)
        self.linear = EqualizedLinear(depth * 16, depth, equalized=equalized_lr, init_bias_to_zero=init_bias_to_zero)
        self.activ = nn.LeakyReLU(0.2)
        self.apply_minibatch_norm = apply_minibatch_norm

    def forward(self, x, subgroup_size=4):
        x = ops.feat(x)
        W, b = self.conv.module.weight, self.conv.module.bias
        layout = None
        if self.apply_minibatch_norm:
            # the (C+1)-channel input is padded with zero channels up to the MMA K granule; the weight
            # packing pads to match (layout "pad_c"), so W and its gradient keep their [C, C+1, 3, 3] shape
            cin = x.shape[1] + 1
            gran = 64 if x.element_size() == 2 else 32
            cpad = (cin + gran - 1) // gran * gran
            parts = [x, _stddev_channel(x, subgroup_size)]
            if cpad > cin:
                parts.append(torch.zeros(x.shape[0], cpad - cin, x.shape[2], x.shape[3], device=x.device,
                                         dtype=x.dtype))
                layout = ("pad_c", cpad)
            x = torch.cat(parts, dim=1).contiguous(memory_format=torch.channels_last)
        x = ops.ConvF.apply(x, W, b, self.conv.scale(), 0.2, False, layout)
        n, c, h, w = x.shape
        # flatten: the reference flattens NCHW (c,h,w); our activations are NHWC, so the linear's K axis is
        # read in (h,w,c) order through the packing map ("lin_in_perm") instead of permuting the tensor
        x = x.permute(0, 2, 3, 1).reshape(n, h * w * c)
        lin = self.linear.module
        x = ops.ConvF.apply(x, lin.weight, lin.bias, self.linear.scale(), 0.2, False, ("lin_in_perm", c))
        return x


class fromRGBBlock(nn.Module):
    """reference lib/blocks.py:271-292: [avg_pool2] -> 1x1 conv 3 -> C -> LReLU(0.2)"""

    def __init__(self, input_dim=3, new_depth=None, equalized_lr=True, init_bias_to_zero=True):
        super().__init__()
        self.fromRGB = EqualizedConv2d(input_dim, new_depth, 1, equalized=equalized_lr,
                                       init_bias_to_zero=init_bias_to_zero)
        self.activ = nn.LeakyReLU(0.2)

    def forward(self, x, apply_downscale=False, defer_mask=False, f32=False):
        """f32: produce fp32 features (the fp32 family: a 4x4 discriminator) instead of the mode's 16-bit format"""
        m = self.fromRGB.module
        W = m.weight
        return ops.RgbExpand.apply(x, W.view(W.shape[0], W.shape[1]), m.bias, bool(apply_downscale),
                                   self.fromRGB.scale(), 0.2, bool(defer_mask), ops.A, bool(f32))
