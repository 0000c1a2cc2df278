"""pggan_b200 — B200-native PGGAN D+G training step behind the reference's PyTorch API.

Python host code mirrors yukyeongleee/pggan's hot-path modules (lib/layers.py, lib/custom_layers.py,
lib/blocks.py, lib/utils.py, lib/loss.py, pggan/nets.py, pggan/loss.py, pggan/model.py:train_step);
all arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of include/pggan_b200.h.
"""
from . import _lib
from .ops import set_compute_dtype, get_compute_dtype, input_grad_only, set_deterministic  # noqa: F401
from .layers import (PixelwiseVectorNorm, ConstrainedLayer, EqualizedConv2d, EqualizedLinear,  # noqa: F401
                     getLayerNormalizationFactor)
from .blocks import (ProgressiveGeneratorBlock, ProgressiveDiscriminatorBlock, LastProgressiveDiscriminatorBlock,  # noqa: F401
                     toRGBBlock, fromRGBBlock, concatenate_stddev_channel)
from .utils import upscale2d, downscale2d, num_flat_features, update_net  # noqa: F401
from .nets import Generator, Discriminator  # noqa: F401
from .loss import Loss, LossInterface, WGANGPLoss  # noqa: F401
from .optim import FusedAdam, FlatGrads  # noqa: F401
from .dist import GradSync, broadcast_parameters  # noqa: F401
from .trainer import ProgressiveGAN, make_args, DEPTHS_1024  # noqa: F401
from . import demo  # noqa: F401
from .data import GPUInputPipeline  # noqa: F401
from .ops import get_precision_mode  # noqa: F401

__version__ = "0.1.0"
