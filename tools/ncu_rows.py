"""One launch each of the smem-resident-rows kernels (conv_rows / wgrad_rows) on the narrow-layer shapes and of the
main conv_gemm / wgrad_gemm kernels on the 32^2 512->512 layer (the bench line's dominant kernel), bracketed
by cudaProfilerStart/Stop so that `ncu --profile-from-start off` captures exactly these:
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_rows python tools/ncu_rows.py
"""
import sys
import torch
sys.path.insert(0, ".")
import pggan_b200 as pg
from pggan_b200 import ops

pg.set_compute_dtype(torch.bfloat16)
dev = "cuda"
SHAPES = [(16, 256, 256, 64, 64), (4, 1024, 1024, 16, 16), (4, 512, 512, 32, 32), (64, 32, 32, 512, 512)]
if len(sys.argv) > 1 and sys.argv[1] == "c4":       # the mid-width layers of the 128^2 / 256^2 stages (256 x N tiles)
    SHAPES = [(16, 256, 256, 128, 64), (16, 128, 128, 128, 128), (16, 128, 128, 256, 128)]
work = []
for (B, H, W, C, N) in SHAPES:
    x = torch.randn(B, C, H, W, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    g = torch.randn(B, N, H, W, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    Wt = torch.nn.Parameter(torch.randn(N, C, 3, 3, device=dev))
    wp = ops._pack(Wt, "f", torch.bfloat16)
    bias = torch.zeros(N, device=dev)
    work.append((lambda x=x, wp=wp, bias=bias, N=N: ops.k_conv(x, wp, bias, N, 9, 0.02, 0.2)))
    work.append((lambda g=g, x=x, N=N, C=C: ops.k_wgrad(g, x, 0.02, (N, C, 3, 3))))
for fn in work:          # warm-up (module load, attribute setting)
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
