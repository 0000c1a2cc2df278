"""Equalized-learning-rate layers and pixel norm — mirror of reference lib/layers.py
(and, through custom_layers.py, of its dead duplicate lib/custom_layers.py).

Same classes, constructor signatures, attribute names (`.module.weight/.bias`, float `.weight`) and
state_dict keys; the arithmetic runs in libpggan_b200.so.  Outputs are channels_last 16-bit (or fp32) tensors, like an
autocast region (see the precision modes in ops.py); the networks in nets.py convert their final outputs back to fp32.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
from numpy import prod

from . import ops


class PixelwiseVectorNorm(nn.Module):
    """reference lib/layers.py:8-14"""

    def __init__(self):
        super().__init__()

    def forward(self, x, epsilon=1e-8):
        return ops.export_act(ops.PixelNorm.apply(x, float(epsilon)))


def getLayerNormalizationFactor(x):
    """He's constant sqrt(2 / fan_in) — reference lib/layers.py:17-25"""
    size = x.weight.size()
    fan_in = prod(size[1:])
    return math.sqrt(2.0 / fan_in)


class ConstrainedLayer(nn.Module):
    """reference lib/layers.py:28-63: zero bias, N(0,1)/lrMul weights, He constant applied at run
    time to the module output INCLUDING the bias: y = c * (module(x))."""

    def __init__(self, module, equalized=True, lrMul=1.0, init_bias_to_zero=True):
        super().__init__()
        self.module = module
        self.equalized = equalized
        if init_bias_to_zero:
            self.module.bias.data.fill_(0)
        if self.equalized:
            self.module.weight.data.normal_(0, 1)
            self.module.weight.data /= lrMul
            self.weight = getLayerNormalizationFactor(self.module) * lrMul
        if isinstance(module, nn.Conv2d) and module.kernel_size == (3, 3):
            # Memory format only: the 3x3 weight is kept channels_last, i.e. physically [O][kh][kw][I] — the K-major
            # operand the tensor cores consume — while shape, values, state_dict keys and checkpoints stay OIHW.  The
            # packed fprop copy is then an elementwise cast (written by the fused Adam kernel) and weight gradients land
            # in the parameter's own layout without a transpose.
            self.module.weight.data = self.module.weight.data.contiguous(memory_format=torch.channels_last)

    # -- fused entry used by blocks / nets: lrelu(c * (W x + b), slope) in one kernel ------------
    def scale(self) -> float:
        return float(self.weight) if self.equalized else 1.0

    def fused(self, x, slope=1.0, raw_out=False, pixel_norm_eps=None, layout=None, x_act_slope=None,
              defer_mask=False, upscale_input=False):
        """lrelu(c*(W x + b), slope) [+ pixel norm if pixel_norm_eps is given] on the fused kernels.
        x_act_slope / defer_mask: LeakyReLU-backward fusion hints, see ops.ConvF.
        upscale_input: the layer reads upscale2d(x) (nearest 2x) — fused into the GEMM where ops.conv_up2_fusable."""
        m = self.module
        W, b = m.weight, m.bias
        c = self.scale()
        if isinstance(m, nn.Linear):
            x2 = x.reshape(x.shape[0], -1) if x.dim() != 2 else x
            return ops.ConvF.apply(x2, W, b, c, float(slope), raw_out, layout)
        if not isinstance(m, nn.Conv2d):
            raise TypeError("ConstrainedLayer supports nn.Conv2d and nn.Linear modules")
        k = m.kernel_size[0]
        if m.kernel_size != (k, k) or m.stride != (1, 1) or m.dilation != (1, 1) or m.groups != 1 \
                or m.padding != (k // 2, k // 2) or k not in (1, 3):
            raise NotImplementedError("only 3x3/pad 1 and 1x1/pad 0 stride-1 convolutions are on the PGGAN path")
        if k == 1 and W.shape[1] <= 4:                       # fromRGB-shaped: image in, features out
            return ops.RgbExpand.apply(x, W.view(W.shape[0], W.shape[1]), b, False, c, float(slope), bool(defer_mask))
        if k == 1 and W.shape[0] <= 4:                       # toRGB-shaped: features in, image out
            if slope != 1.0:
                raise NotImplementedError("toRGB-shaped 1x1 convolutions have no fused activation")
            return ops.RgbReduce.apply(x, W.view(W.shape[0], W.shape[1]).t(), b, False, c, 1.0)
        if pixel_norm_eps is not None:
            return ops.ConvActPN.apply(x, W, b, c, float(slope), float(pixel_norm_eps), layout, bool(upscale_input))
        if upscale_input:
            x = ops.Upsample2x.apply(x, 1.0)
        return ops.ConvF.apply(x, W, b, c, float(slope), raw_out, layout, x_act_slope, bool(defer_mask))

    def forward(self, x):
        """the reference's public call: c * module(x).  (In mixed-precision mode the result is exported as an honest
        fp16 tensor; the blocks / networks call `fused`, whose results stay in the kernels' internal container.)"""
        return ops.export_act(self.fused(x, 1.0))


class EqualizedConv2d(ConstrainedLayer):
    """reference lib/layers.py:66-89"""

    def __init__(self, nChannelsPrevious, nChannels, kernelSize, padding=0, bias=True, **kwargs):
        ConstrainedLayer.__init__(self, nn.Conv2d(nChannelsPrevious, nChannels, kernelSize, padding=padding,
                                                  bias=bias), **kwargs)


class EqualizedLinear(ConstrainedLayer):
    """reference lib/layers.py:92-109"""

    def __init__(self, nChannelsPrevious, nChannels, bias=True, **kwargs):
        ConstrainedLayer.__init__(self, nn.Linear(nChannelsPrevious, nChannels, bias=bias), **kwargs)
