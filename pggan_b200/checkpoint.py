"""Checkpoint wire format of the reference — mirror of lib/checkpoint.py:5-34 (same function signatures) and of the
glue in pggan/model.py:50-101.  A checkpoint is `torch.save` of
{model, optimizer, args, global_step, alpha_G/alpha_D, alpha_index, alpha_jump_value, next_alpha_jump_step,
scale_index, next_scale_jump_step}; parameter names/shapes are the reference's and `optimizer` is a
torch.optim.Adam-format state dict (optim.FusedAdam writes and reads that format), so files written by either side load
in the other."""
from __future__ import annotations

import os

import torch


def load_checkpoint(args, name):
    """reference lib/checkpoint.py:5-19: returns the checkpoint dict of `{save_root}/{ckpt_id}/ckpt/{name}_{step}.pt`
    (step = args.ckpt_step or 'latest').  The reference swallows every error and returns 0, which crashes later at
    `G_ckpt_dict["args"]`; here a missing or unreadable file raises with the path in the message."""
    ckpt_step = getattr(args, "ckpt_step", None)
    if ckpt_step is None:
        ckpt_step = "latest"
    ckpt_path = f"{args.save_root}/{args.ckpt_id}/ckpt/{name}_{ckpt_step}.pt"
    if not os.path.exists(ckpt_path):
        raise FileNotFoundError(f"Failed to load checkpoint of {name}: {ckpt_path} does not exist")
    device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
    return torch.load(ckpt_path, map_location=device, weights_only=False)


def save_checkpoint(model, optimizer, name, ckpt_dict={}):
    """reference lib/checkpoint.py:22-34 (the shared mutable default included: callers pass their own dict)"""
    ckpt_dict["model"] = model.state_dict()
    ckpt_dict["optimizer"] = optimizer.state_dict()
    dir_path = f'{ckpt_dict["args"]["save_root"]}/{ckpt_dict["args"]["run_id"]}/ckpt'
    os.makedirs(dir_path, exist_ok=True)
    ckpt_path = dir_path + f'/{name}_{ckpt_dict["global_step"]}.pt'
    torch.save(ckpt_dict, ckpt_path)
    latest_ckpt_path = dir_path + f"/{name}_latest.pt"
    torch.save(ckpt_dict, latest_ckpt_path)


def grow_to_checkpoint(G, D, depths, scale_index):
    """replay add_block `scale_index` times before loading weights (reference pggan/model.py:86-92; the reference
    indexes depths[index] there, which only works because its shipped depths repeat — the block added at scale s has
    depths[s], as in change_scale, pggan/model.py:163-164)"""
    while len(G.blocks) < scale_index:
        d = depths[len(G.blocks) + 1]
        G.add_block(d)
        D.add_block(d)
