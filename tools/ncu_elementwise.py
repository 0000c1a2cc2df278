"""One launch each of the bandwidth kernels at the 1024^2 (C = 16) and 256^2 (C = 64) stage sizes, bracketed by
cudaProfilerStart/Stop for `ncu --profile-from-start off`:
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_ew python tools/ncu_elementwise.py
"""
import sys
import torch
sys.path.insert(0, ".")
import pggan_b200 as pg
from pggan_b200 import ops

pg.set_compute_dtype(torch.bfloat16)
dev = "cuda"
work = []
for (B, R, C) in [(4, 1024, 16), (16, 256, 64), (64, 32, 512)]:
    f = torch.randn(B, C, R, R, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    g = torch.randn_like(f)
    img = torch.rand(B, 3, R, R, device=dev)
    Wc = torch.randn(C, 3, device=dev)
    bias = torch.zeros(C, device=dev)
    lo = torch.randn(B, C, R // 2, R // 2, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    work += [
        lambda img=img, Wc=Wc, bias=bias, C=C: ops._k_rgb_expand(img, Wc, bias, C, False, 0.1, 0.2, torch.bfloat16),
        lambda f=f, Wc=Wc: ops._k_rgb_reduce(f, Wc.t().contiguous(), None, 3, False, 0.1, 1.0),
        lambda f=f, img=img: ops._k_rgb_wgrad(f, img, False, 0.1, want_bias=True),
        lambda lo=lo: ops.k_upsample(lo, 1.0),
        lambda lo=lo, f=f: ops.k_upsample(lo, 0.25, mask_src=f, slope=0.2),
        lambda f=f: ops.k_pool(f, 0.25),
        lambda g=g, f=f: ops.k_mask_colsum(g, None, 0.2, True, 0.1),
        lambda g=g, f=f: ops.k_mask_mul(g, f, 0.2),
    ]
for fn in work:
    fn()
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flush.zero_()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for fn in work:
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
