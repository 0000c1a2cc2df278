"""Checkpoint wire format of the reference — mirror of lib/checkpoint.py:5-34 and the glue in
pggan/model.py:50-101.  A checkpoint is `torch.save` of
{model, optimizer, args, global_step, alpha_G/alpha_D, alpha_index, alpha_jump_value,
next_alpha_jump_step, scale_index, next_scale_jump_step}; parameter names/shapes are the reference's,
so files written by either side load in the other."""
from __future__ import annotations

import os

import torch


def save_checkpoint(args, model, optimizer, name, global_step, extra=None):
    """reference lib/checkpoint.py:22-34 (+ the schedule variables of pggan/model.py:54-64 via `extra`)"""
    ckpt_dict = {"model": model.state_dict(),
                 "optimizer": optimizer.state_dict() if optimizer is not None else None,
                 "args": vars(args) if hasattr(args, "__dict__") else dict(args),
                 "global_step": global_step}
    if extra:
        ckpt_dict.update(extra)
    dir_path = f"{args.save_root}/{args.run_id}/ckpt"
    os.makedirs(dir_path, exist_ok=True)
    torch.save(ckpt_dict, f"{dir_path}/{name}_{global_step}.pt")
    torch.save(ckpt_dict, f"{dir_path}/{name}_latest.pt")
    return ckpt_dict


def load_checkpoint(args, model, optimizer, name, global_step=-1, strict=False):
    """reference lib/checkpoint.py:5-19; returns the checkpoint dict (the reference returns 0 on failure
    and crashes later — here a missing file raises)"""
    ckpt_step = "latest" if global_step == -1 else global_step
    path = f"{args.save_root}/{args.ckpt_id}/ckpt/{name}_{ckpt_step}.pt"
    ckpt_dict = torch.load(path, map_location="cpu", weights_only=False)
    model.load_state_dict(ckpt_dict["model"], strict=strict)
    if optimizer is not None and ckpt_dict.get("optimizer") is not None and hasattr(optimizer, "load_state_dict"):
        try:
            optimizer.load_state_dict(ckpt_dict["optimizer"])
        except Exception:
            pass          # optimizer kinds differ (torch Adam vs FusedAdam): moments restart, like a scale jump
    return ckpt_dict


def grow_to_checkpoint(G, D, depths, scale_index):
    """replay add_block `scale_index` times before loading weights (reference pggan/model.py:86-92)"""
    while len(G.blocks) < scale_index:
        d = depths[len(G.blocks) + 1]
        G.add_block(d)
        D.add_block(d)
