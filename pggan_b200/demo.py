"""G-only inference entry — mirror of the reference's demo.py:23-68.

`LoadModel(args)` builds the generator from `args` exactly as the reference does (same constructor arguments, the
same `add_block` replay over `args.depths[1:]`), loads a reference-format checkpoint (`{"model": state_dict, ...}`,
lib/checkpoint.py:22-34) after the reference's key-set check, and moves it to the GPU; `test(args)` draws
`args.n_samples` latents, runs G under no_grad and returns / writes the images converted as the reference converts
them ((x*0.5+0.5)*255, HWC).  The arithmetic runs on the same CUDA kernels as training (no CPU path)."""
from __future__ import annotations

import os

import numpy as np
import torch

from .nets import Generator


def LoadModel(args):
    """reference demo.py:23-46"""
    G = Generator(args.latent_dim, args.depths[0], args.init_bias_to_zero, args.LReLU_slope, args.apply_pixel_norm,
                  args.generator_last_activation, args.output_dim, args.equalized_lr)
    for depth in args.depths[1:]:
        G.add_block(depth)
    ckpt_dict = torch.load(args.ckpt_path, map_location="cpu", weights_only=False)
    A = set(ckpt_dict["model"].keys())
    B = set(G.state_dict().keys())
    assert A == B, f"checkpoint does not match the model: {sorted(A ^ B)[:8]}"
    G.load_state_dict(ckpt_dict["model"])
    if "alpha_G" in ckpt_dict:                      # pggan/model.py:54-64 stores the fade state next to the weights
        G.alpha = float(ckpt_dict["alpha_G"])
    gpu = getattr(args, "gpu", 0)
    return G.to(gpu if isinstance(gpu, (str, torch.device)) else f"cuda:{int(gpu)}")


def test(args, G=None, latent_input=None, write=True):
    """reference demo.py:48-68; returns the float image array [n, H, W, 3] in 0..255 (the reference only writes files)"""
    G = (G if G is not None else LoadModel(args)).eval()
    dev = next(G.parameters()).device
    with torch.no_grad():
        if latent_input is None:
            latent_input = torch.randn(args.n_samples, args.latent_dim)
        generated_imgs = G(latent_input.to(dev)).float().cpu().numpy()
    generated_imgs = np.transpose(generated_imgs, axes=[0, 2, 3, 1])
    generated_imgs = (generated_imgs * 0.5 + 0.5) * 255
    if write and getattr(args, "save_path", None):
        os.makedirs(args.save_path, exist_ok=True)
        try:
            from PIL import Image
        except Exception:                            # no image library: keep the arrays
            Image = None
        for i in range(generated_imgs.shape[0]):
            arr = np.clip(generated_imgs[i], 0, 255).astype(np.uint8)
            if Image is not None:
                Image.fromarray(arr).save(f"{args.save_path}/result_{i:02d}.png")
            else:
                np.save(f"{args.save_path}/result_{i:02d}.npy", arr)
    return generated_imgs
