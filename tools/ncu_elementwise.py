"""One launch each of the bandwidth kernels at the C4 (256^2 x 64, 128^2 x 128) and C5 (1024^2 x 16) stage sizes in the
benchmarked fp16 mode, bracketed by cudaProfilerStart/Stop for `ncu --profile-from-start off`:
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_ew python tools/ncu_elementwise.py
Without ncu it prints CUDA-event times per kernel (warm, L2 flushed between launches) — the numbers the tuning used.
"""
import sys
import torch
sys.path.insert(0, ".")
import pggan_b200 as pg
from pggan_b200 import ops

pg.set_compute_dtype(torch.float16)
dev = "cuda"
H = torch.float16
work = []
for (B, R, C) in [(16, 256, 64), (16, 128, 128), (4, 1024, 16)]:
    f = torch.randn(B, C, R, R, device=dev).to(H).contiguous(memory_format=torch.channels_last)
    g = torch.randn_like(f)
    img = torch.rand(B, 3, R, R, device=dev)
    img2 = torch.rand(B, 3, 2 * R, 2 * R, device=dev) if R < 1024 else None
    Wc = torch.randn(C, 3, device=dev)
    bias = torch.zeros(C, device=dev)
    lo = torch.randn(B, C, R // 2, R // 2, device=dev).to(H).contiguous(memory_format=torch.channels_last)
    tag = f"{R}x{R}x{C}"
    work += [
        (f"rgb_expand {tag}", lambda img=img, Wc=Wc, bias=bias, C=C: ops._k_rgb_expand(img, Wc, bias, C, False, 0.1, 0.2, ops.A)),
        (f"rgb_reduce {tag}", lambda f=f, Wc=Wc: ops._k_rgb_reduce(f, Wc.t().contiguous(), None, 3, False, 0.1, 1.0, ops.A)),
        (f"rgb_wgrad {tag}", lambda f=f, img=img: ops._k_rgb_wgrad(f, img, False, 0.1, ops.G, want_bias=True)),
        (f"upsample {tag}", lambda lo=lo: ops.k_upsample(lo, 1.0)),
        (f"upsample_mask {tag}", lambda lo=lo, f=f: ops.k_upsample(lo, 0.25, mask_src=f, slope=0.2, cls=ops.G)),
        (f"pool {tag}", lambda f=f: ops.k_pool(f, 0.25)),
        (f"colsum_only {tag}", lambda g=g: ops.k_mask_colsum(g, None, 0.2, True, 0.1)),
        (f"mask_colsum {tag}", lambda g=g, f=f: ops.k_mask_colsum(g, f, 0.2, True, 0.1)),
        (f"mask_mul {tag}", lambda g=g, f=f: ops.k_mask_mul(g, f, 0.2)),
    ]
    if img2 is not None:
        work += [
            (f"rgb_expand_pool {tag}", lambda img2=img2, Wc=Wc, bias=bias, C=C: ops._k_rgb_expand(img2, Wc, bias, C, True, 0.1, 0.2, ops.A)),
            (f"rgb_wgrad_pool {tag}", lambda f=f, img2=img2: ops._k_rgb_wgrad(f, img2, True, 0.1, ops.G, want_bias=True)),
        ]
# D's decision head: linear 8192 -> 512 weight gradient un-pack (the generic map path) and its operand pack
Wl = torch.nn.Parameter(torch.randn(512, 8192, device=dev))
gl = torch.randn(16, 512, device=dev).to(H)
xl = torch.randn(16, 8192, device=dev).to(H)
lay = ("lin_in_perm", 512)
work += [
    ("linear wgrad+unpack 8192->512", lambda: ops.k_wgrad(gl, xl, 0.1, tuple(Wl.shape), lay)),
    ("linear pack f", lambda: (ops.clear_pack_cache(), ops._pack(Wl, "f", None, lay))[1]),
    ("linear pack t", lambda: (ops.clear_pack_cache(), ops._pack(Wl, "t", None, lay))[1]),
]
for _, fn in work:
    fn()
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
if "--time" in sys.argv:
    # GPU time per call: 10 calls captured into a CUDA graph, the replay timed (eager calls of the short kernels are
    # bound by the host-side wrapper)
    REP = 10
    st = torch.cuda.Stream()
    for name, fn in work:
        if name.startswith("linear pack"):
            continue                      # (drops the pack cache: a host-side effect a graph would not replay)
        with torch.cuda.stream(st):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                for _ in range(REP):
                    fn()
            g.replay()
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(st)
                g.replay()
                b.record(st)
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1000 / REP)
        print(f"{name:40s} {min(ts):8.1f} us per call (graph of {REP}, min of 5 replays; L2 warm where the tensor fits)")
    sys.exit(0)
flush.zero_()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _, fn in work:
    fn()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
