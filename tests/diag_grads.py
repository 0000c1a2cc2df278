"""diagnostic (not a test): per-parameter gradient errors of G and D against the oracle"""
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import rel_l2
from gpu_util import build_pair, req
from oracle import pggan_oracle as O
import pggan_b200 as pg

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
scale, width, batch, alpha = int(sys.argv[1]), int(sys.argv[2]), 8, 0.5
for dtype in (torch.float32, torch.bfloat16):
    pg.set_compute_dtype(dtype)
    depths = [width] * (scale + 1)
    G, D, g_sd, d_sd = build_pair(scale, depths, width, seed=scale)
    G.alpha = D.alpha = alpha
    res = 4 * 2 ** scale
    torch.manual_seed(100 + scale)
    z = torch.randn(batch, width, device="cuda")
    x = torch.rand(batch, 3, res, res, device="cuda") * 2 - 1
    gp_, dp_ = req(g_sd), req(d_sd)
    img_ref = O.generator_forward(gp_, z, alpha)
    img = G(z)
    wimg = torch.randn_like(img_ref)
    (img_ref * wimg).sum().backward()
    (img * wimg).sum().backward()
    print(f"==== {dtype} scale {scale} width {width}: img err {rel_l2(img, img_ref):.2e}")
    for k, p in G.named_parameters():
        if gp_[k].grad is not None and p.grad is not None:
            print(f"  G {k:45s} {rel_l2(p.grad, gp_[k].grad):.2e}")
    xr = x.clone().requires_grad_()
    logit_ref = O.discriminator_forward(dp_, xr, alpha)
    xm = x.clone().requires_grad_()
    logit = D(xm)
    wl = torch.randn_like(logit_ref)
    (logit_ref * wl).sum().backward()
    (logit * wl).sum().backward()
    print(f"  logit err {rel_l2(logit, logit_ref):.2e}  dD/dx err {rel_l2(xm.grad, xr.grad):.2e}")
    for k, p in D.named_parameters():
        if dp_[k].grad is not None and p.grad is not None:
            print(f"  D {k:45s} {rel_l2(p.grad, dp_[k].grad):.2e}")
