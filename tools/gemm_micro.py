"""Micro-benchmark / profiling driver for the two tensor-core kernels on the C2 layer shapes.
    python tools/gemm_micro.py [reps]
Times each shape with CUDA events (after warm-up) and prints TFLOP/s; run it under ncu for
per-kernel counters (profiles/)."""
import sys
import torch
sys.path.insert(0, ".")
import pggan_b200 as pg
from pggan_b200 import ops

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
pg.set_compute_dtype(torch.bfloat16)
dev = "cuda"
SHAPES = [(64, 32, 32, 512, 512), (128, 32, 32, 512, 512), (64, 16, 16, 512, 512), (64, 8, 8, 512, 512),
          (128, 8, 8, 512, 512), (64, 4, 4, 512, 512)]
if len(sys.argv) > 2 and sys.argv[2] == "c4":       # the 128^2 / 256^2 stage layer shapes (B, H, W, Cin, Cout)
    SHAPES = [(16, 256, 256, 64, 64), (16, 256, 256, 128, 64), (16, 256, 256, 64, 128), (16, 128, 128, 128, 128),
              (16, 128, 128, 256, 128), (16, 128, 128, 128, 256), (16, 64, 64, 256, 256)]
if len(sys.argv) > 2 and sys.argv[2] == "narrow":   # the 256^2 / 512^2 / 1024^2 narrow layers (conv_rows.cu)
    SHAPES = [(16, 256, 256, 64, 64), (16, 256, 256, 64, 128), (4, 512, 512, 32, 32), (4, 512, 512, 32, 64),
              (4, 512, 512, 64, 32), (4, 1024, 1024, 16, 16), (4, 1024, 1024, 16, 32), (4, 1024, 1024, 32, 16)]
if len(sys.argv) > 2 and sys.argv[2] == "small":    # the 4^2 .. 32^2 trunk layers at the C4 / C3 / C5 batch sizes (few M tiles, long K)
    SHAPES = [(16, 4, 4, 512, 512), (16, 8, 8, 512, 512), (16, 16, 16, 512, 512), (32, 4, 4, 512, 512),
              (32, 8, 8, 512, 512), (4, 8, 8, 512, 512), (4, 16, 16, 512, 512), (4, 32, 32, 512, 512),
              (4, 64, 64, 256, 256), (4, 64, 64, 512, 256)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (B, H, W, C, N) in SHAPES:
    x = torch.randn(B, C, H, W, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    g = torch.randn(B, N, H, W, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    Wt = torch.nn.Parameter(torch.randn(N, C, 3, 3, device=dev))
    wp = ops._pack(Wt, "f", torch.bfloat16)
    bias = torch.zeros(N, device=dev)
    fl = 2.0 * B * H * W * C * N * 9
    for name, fn in (("conv_fprop", lambda: ops.k_conv(x, wp, bias, N, 9, 0.02, 0.2)),
                     ("wgrad", lambda: ops.k_wgrad(g, x, 0.02, (N, C, 3, 3)))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()                      # evict L2 between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        med = ts[len(ts) // 2]
        byt = 2.0 * B * H * W * (C + N)
        print(f"{name:11s} B{B} {H}x{W} C{C}->N{N}: median {med*1e3:8.1f} us  {fl/med/1e9:8.1f} TFLOP/s  "
              f"{byt/med/1e6:7.0f} GB/s (min {ts[0]*1e3:.1f} us)", flush=True)
