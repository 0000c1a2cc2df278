"""helpers shared by the GPU parity tests"""
import torch

from oracle import pggan_oracle as O


def build_pair(scale, depths, latent, seed=0, bias_std=0.3, device="cuda"):
    """reference-initialised state dicts (N(0,1) weights) + the pggan_b200 modules loaded with them.
    bias_std > 0 randomises the biases (the reference initialises them to 0, lib/layers.py:51-53) so the bias path is
    live; bias_std = 0 is the reference initialisation itself."""
    import pggan_b200 as pg
    gen = torch.Generator().manual_seed(seed)
    g_sd = O.init_generator_sd(latent, depths, scale, generator=gen)
    d_sd = O.init_discriminator_sd(depths, scale, generator=gen)
    for sd in (g_sd, d_sd):
        for k in sd:
            if k.endswith(".bias") and bias_std > 0:
                sd[k] = torch.randn(sd[k].shape, generator=gen) * bias_std
    G = pg.Generator(latent, depths[0])
    D = pg.Discriminator(depths[0], apply_minibatch_norm=True)
    for s in range(1, scale + 1):
        G.add_block(depths[s])
        D.add_block(depths[s])
    G.load_state_dict(g_sd)
    D.load_state_dict(d_sd)
    G.to(device)
    D.to(device)
    g_sd = {k: v.to(device) for k, v in g_sd.items()}
    d_sd = {k: v.to(device) for k, v in d_sd.items()}
    return G, D, g_sd, d_sd


def req(sd):
    return {k: v.clone().requires_grad_() for k, v in sd.items()}


def grads_by_name(module):
    return {k: p.grad for k, p in module.named_parameters()}
