"""Mirror of reference lib/custom_layers.py — the duplicate of lib/layers.py that differs only in
constructor argument names (in_channels/out_channels/kernel_size).  One implementation (layers.py)."""
from __future__ import annotations

import torch.nn as nn

from .layers import ConstrainedLayer, getLayerNormalizationFactor  # noqa: F401


class EqualizedConv2d(ConstrainedLayer):
    """reference lib/custom_layers.py:56-74"""

    def __init__(self, in_channels, out_channels, kernel_size, padding=0, bias=True, **kwargs):
        ConstrainedLayer.__init__(self, nn.Conv2d(in_channels, out_channels, kernel_size, padding=padding,
                                                  bias=bias), **kwargs)


class EqualizedLinear(ConstrainedLayer):
    """reference lib/custom_layers.py:77-91"""

    def __init__(self, in_channels, out_channels, bias=True, **kwargs):
        ConstrainedLayer.__init__(self, nn.Linear(in_channels, out_channels, bias=bias), **kwargs)
