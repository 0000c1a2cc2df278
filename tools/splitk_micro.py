"""GPU-side time of the few-tile trunk layers under split-K choices: each variant is captured into a CUDA graph of 20
launches and the replay is timed (tools/gemm_micro.py times eager calls, which are CPU-bound below ~25 us).
    python tools/splitk_micro.py            # env PGGAN_FORCE_SPLITK is set per variant by this script
"""
import os
import sys
import torch
sys.path.insert(0, ".")
import pggan_b200 as pg
from pggan_b200 import ops

pg.set_compute_dtype(torch.float16)
dev = "cuda"
SHAPES = [(16, 4, 4, 512, 512), (16, 8, 8, 512, 512), (32, 4, 4, 512, 512), (32, 8, 8, 512, 512), (16, 16, 16, 512, 512),
          (4, 4, 4, 512, 512), (4, 8, 8, 512, 512), (4, 16, 16, 512, 512), (4, 32, 32, 512, 512)]
VARIANTS = [None, "0,1", "64,2", "64,3", "64,4", "64,6", "64,8", "64,9", "128,2", "128,3", "128,4", "128,6", "128,8", "128,12",
            "256,4", "256,8"]
REP = 20
for (B, H, W, C, N) in SHAPES:
    x = torch.randn(B, C, H, W, device=dev).to(torch.float16).contiguous(memory_format=torch.channels_last)
    Wt = torch.nn.Parameter(torch.randn(N, C, 3, 3, device=dev))
    wp = ops._pack(Wt, "f")
    bias = torch.zeros(N, device=dev)
    fl = 2.0 * B * H * W * C * N * 9
    line = []
    for v in VARIANTS:
        if v is None:
            os.environ.pop("PGGAN_FORCE_SPLITK", None)
        else:
            os.environ["PGGAN_FORCE_SPLITK"] = v
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3):
                ops.k_conv(x, wp, bias, N, 9, 0.02, 0.2)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for _ in range(REP):
                    y = ops.k_conv(x, wp, bias, N, 9, 0.02, 0.2)
            g.replay()
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s)
                g.replay()
                e1.record(s)
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1000 / REP)
        line.append(f"{v or 'auto'}:{min(ts):.1f}")
    print(f"B{B} {H}x{W} C{C}->N{N} ({fl/1e9:.1f} GF): " + "  ".join(line), flush=True)
