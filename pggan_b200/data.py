"""GPU input pipeline — the step BEFORE the hot path: the transform chain of the reference's
UnsupervisedDataset (lib/dataset.py:86-127: Resize((s,s)) with s = 2**(scale_index+2), RandomHorizontalFlip(0.5),
ColorJitter(0.2, 0.2, 0.2, 0.01), ToTensor, Normalize(0.5, 0.5)) fused with the real-image fade that train_step applies
to the batch (pggan/model.py:217-221), on the GPU.

The reference runs this chain in 8 PIL worker processes per rank (lib/model.py:51), which starves eight B200s at the low
resolutions (a 32x32 step takes 19 ms for 64 images).  Here a batch of DECODED images (uint8 [B,3,Hs,Ws] on the device —
file reading / JPEG decoding stay outside) goes through three bandwidth kernels (csrc/input.cu).  Random parameters
are drawn on the CPU generator in exactly the order torchvision draws them (flip; randperm(4); the four factors), so a
run is reproducible against `transforms.Compose` applied per image under the same seed.  The arithmetic is that of
torchvision's float-tensor backend; the PIL backend additionally quantises to uint8 after every operation.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib as L
from . import ops

# transforms.ColorJitter(0.2, 0.2, 0.2, 0.01) -> the sampling ranges torchvision derives from those arguments
JITTER_RANGES = ((0.8, 1.2), (0.8, 1.2), (0.8, 1.2), (-0.01, 0.01))
FLIP_P = 0.5


def sample_params(batch: int, generator: Optional[torch.Generator] = None, jitter=JITTER_RANGES, flip_p=FLIP_P):
    """per-sample random parameters, drawn like the reference's transform objects draw them for ONE image, image after
    image: RandomHorizontalFlip.forward (torch.rand(1) < p), then ColorJitter.get_params (torch.randperm(4), then
    brightness, contrast, saturation, hue factors by uniform_).  Returns (int32 [B,5], float32 [B,4]) CPU tensors."""
    ip = torch.empty(batch, 5, dtype=torch.int32)
    fp = torch.empty(batch, 4, dtype=torch.float32)
    for b in range(batch):
        ip[b, 0] = int(torch.rand(1, generator=generator) < flip_p)
        ip[b, 1:] = torch.randperm(4, generator=generator).to(torch.int32)
        for k, (lo, hi) in enumerate(jitter):
            fp[b, k] = float(torch.empty(1).uniform_(lo, hi, generator=generator))
    return ip, fp


def resize_antialias(images: torch.Tensor, size: int) -> torch.Tensor:
    """transforms.Resize((size, size)) on a uint8 or float [B,3,Hs,Ws] CUDA batch -> fp32 [B,3,size,size] in [0,1]
    (uint8 input is scaled by 1/255 like ToTensor does)"""
    if not images.is_cuda:
        raise L.PgganLibraryError("pggan_b200.data runs on CUDA tensors only (no CPU fallback)")
    is_u8 = images.dtype == torch.uint8
    src = images.contiguous() if is_u8 else images.float().contiguous()
    B, C, Hs, Ws = src.shape
    tmp = torch.empty((B * C, Hs, size), dtype=torch.float32, device=src.device)
    dst = torch.empty((B, C, size, size), dtype=torch.float32, device=src.device)
    L.check(L.load().pggan_resize_aa(ctypes.c_void_p(src.data_ptr()), int(is_u8), ctypes.c_void_p(tmp.data_ptr()),
                                     ctypes.c_void_p(dst.data_ptr()), B * C, Hs, Ws, size, size,
                                     1.0 / 255.0 if is_u8 else 1.0, ops._s()), "pggan_resize_aa")
    return dst


def jitter_normalize_fade(img01: torch.Tensor, iparams: torch.Tensor, fparams: torch.Tensor, fade_alpha=None):
    """[B,3,S,S] fp32 in [0,1] -> the network's input: colour jitter (per-sample order and factors), horizontal flip,
    Normalize(0.5, 0.5) and — if fade_alpha is given (a float or an ops.AlphaRef) — the real-image fade
    (1-alpha)*up2(avgpool2(x)) + alpha*x of pggan/model.py:217-221, in one pass."""
    img01 = img01.float().contiguous()
    B, C, S, S2 = img01.shape
    if C != 3 or S != S2:
        raise ValueError("expected a [B,3,S,S] batch")
    dev = img01.device
    ip = iparams.to(dev, torch.int32).contiguous()
    fp = fparams.to(dev, torch.float32).contiguous()
    mean = torch.empty((B,), dtype=torch.float32, device=dev)
    out = torch.empty_like(img01)
    L.check(L.load().pggan_jitter_normalize_fade(ctypes.c_void_p(img01.data_ptr()), ctypes.c_void_p(ip.data_ptr()),
                                                 ctypes.c_void_p(fp.data_ptr()), ctypes.c_void_p(mean.data_ptr()),
                                                 ctypes.c_void_p(out.data_ptr()), B, S, int(fade_alpha is not None),
                                                 ops._coef(0.0, 1.0, fade_alpha if fade_alpha is not None else 1.0),
                                                 ops._s()), "pggan_jitter_normalize_fade")
    return out


class GPUInputPipeline:
    """UnsupervisedDataset's transforms (lib/dataset.py:100-111) + the real-image fade, for one training stage.

        pipe = GPUInputPipeline(scale_index)
        img_real = pipe(decoded_uint8_batch, fade_alpha=model.D.alpha_ref() if scale_index else None)

    `img_real` is what train_step would have after its own fade (pass it with the trainer's fade disabled, or pass
    fade_alpha=None here and let train_step fade)."""

    def __init__(self, scale_index: int = 0, generator: Optional[torch.Generator] = None):
        self.set_scale(scale_index)
        self.generator = generator

    def set_scale(self, scale_index: int):
        self.scale_index = scale_index
        self.size = 2 ** (scale_index + 2)                    # lib/dataset.py:100

    def __call__(self, images: torch.Tensor, fade_alpha=None, params=None):
        ip, fp = params if params is not None else sample_params(images.shape[0], self.generator)
        x = resize_antialias(images, self.size)
        return jitter_normalize_fade(x, ip, fp, fade_alpha)
