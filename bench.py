#!/usr/bin/env python
"""bench.py — PGGAN D+G WGAN-GP training step throughput (BASELINE.json metric: train imgs/sec).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU

Workload at N=1 = BASELINE.json configs[1]: the 32x32 stage with alpha fade-in (alpha 0.5), batch 64,
latent 512, depths 512x4, one full D+G step of the WGAN-GP variant (SURVEY.md §8d "W": BCE adversarial
terms + gradient penalty with its own backward + drift, Adam on D and G), bf16 operands / fp32
accumulate, synthetic uniform images, reference initialisation.  N>1: one rank per GPU (torchrun),
same per-GPU batch (weak scaling), bucketed gradient all-reduce overlapped with backward.

One JSON line on stdout (rank 0).  `value` = images/s with inputs resident in HBM; `e2e` = the same
step driven from pinned HOST buffers (H2D of images/latents/eps and D2H of the losses inside the timed
region); `roofline` = the dominant kernel (3x3 conv implicit GEMM, tcgen05) against the measured bf16
tensor peak; `cpu_baseline` = the oracle restatement of the reference step on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (scale_index, per-GPU batch, depths, W-step GFLOP per image ("W minimal":
    #        D weight-gradient skipped in the G step, which this build does))
    "c1_4x4_b16": (0, 16, [512], 1.5),
    "c2_32x32_b64_fade": (3, 64, [512, 512, 512, 512], 229.9),
    "c3_128x128_b32": (5, 32, [512, 512, 512, 512, 256, 128], 752.0),
    "c4_256x256_b16": (6, 16, [512, 512, 512, 512, 256, 128, 64], 1013.2),
    # paper CelebA-HQ widths
    "c5_1024x1024_b4": (8, 4, [512, 512, 512, 512, 256, 128, 64, 32, 16], 1537.0),
}
METRIC = "train imgs/sec (D+G step incl. WGAN-GP)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region: the sampler is started before the warm-up (so
    nvidia-smi is already streaming when the timed steps begin), every sample carries nvidia-smi's own timestamp, and
    only the samples between mark_begin() and mark_end() are reported."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.idx = gpu_index
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        import datetime
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in out.strip().splitlines():
            f = [t.strip() for t in line.split(",")]
            if len(f) < 7:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), f[3:7]))
            except ValueError:
                continue
        inside = [r for r in rows if self.t0 is not None and self.t0 <= r[0] <= (self.t1 or r[0])]
        note = None
        if not inside and rows and self.t0 is not None:
            # timed region shorter than the sampling period: take the samples closest to it (warm-up steps of the
            # same workload run right before it, so these are under the same load)
            rows.sort(key=lambda r: abs(r[0] - 0.5 * (self.t0 + (self.t1 or self.t0))))
            inside = rows[:3]
            note = "timed region shorter than the 20 ms sampling period: nearest samples (warm-up of the same step)"
        clocks = [r[1] for r in inside]
        reasons = set()
        for r in inside:
            for n, v in zip(names, r[3]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        res = {"sm_mhz": statistics.median(clocks) if clocks else None, "sm_max_mhz": inside[0][2] if inside else None,
               "reasons": sorted(reasons), "samples": len(clocks)}
        if note:
            res["note"] = note
        return res


def dist_setup(n_gpus: int):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    return rank, world, local


# ------------------------------------------------------------------------------------- our arm
def run_ours(args):
    rank, world, local = dist_setup(args.gpus)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import pggan_b200 as pg
    from pggan_b200 import _lib, ops

    scale, batch, depths, gflop_img = WORKLOADS[args.workload]
    if args.batch:
        batch = args.batch
    dtype = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[args.dtype]
    pg.set_compute_dtype(dtype)
    torch.manual_seed(1234 + rank)
    targs = pg.make_args(latent_dim=512, depths=depths)
    tr = pg.ProgressiveGAN(targs, local, variant=args.variant)
    alpha = 0.5 if scale else 0.0
    tr.set_scale(scale, alpha)
    res = 4 * 2 ** scale

    g = torch.Generator(device="cpu").manual_seed(99 + rank)
    host = {
        "img": (torch.rand(batch, 3, res, res, generator=g) * 2 - 1).pin_memory(),
        "z_d": torch.randn(batch, 512, generator=g).pin_memory(),
        "z_g": torch.randn(batch, 512, generator=g).pin_memory(),
        "eps": torch.rand(batch, 1, generator=g).pin_memory(),
    }
    devb = {k: v.to(dev) for k, v in host.items()}

    def step_eager():
        tr.train_step(devb["img"], latent_D=devb["z_d"], latent_G=devb["z_g"], eps=devb["eps"])

    # launches per step, counted on an eager step with the library's own counters
    for _ in range(2):
        step_eager()
    torch.cuda.synchronize()
    c0 = _lib.launch_counts()
    step_eager()
    torch.cuda.synchronize()
    c1 = _lib.launch_counts()
    launches_per_step = {k: c1[k] - c0[k] for k in c1}

    replay, graph_err = None, None
    if args.graph and (world == 1 or not args.no_graph_dp):
        try:
            replay = tr.capture(batch)
        except Exception as e:          # fall back to eager launches, and say so in the JSON line
            graph_err = f"{type(e).__name__}: {e}"[:200]
            replay = None
            torch.cuda.synchronize()

    def step_resident():
        if replay is not None:
            replay()                    # inputs already sit in the graph's static device buffers
        else:
            step_eager()

    loss_host = torch.empty(3, dtype=torch.float32).pin_memory()

    def step_e2e():
        if replay is not None:
            replay(host["img"], host["z_d"], host["z_g"], host["eps"])     # pinned host -> static device buffers
        else:
            d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
            tr.train_step(d["img"], latent_D=d["z_d"], latent_G=d["z_g"], eps=d["eps"])
        ld = dict.items(tr.loss_collector.loss_dict)
        vals = {k: v for k, v in ld}
        t = torch.stack([vals["L_D"].float().reshape(()), vals["L_G"].float().reshape(()),
                         vals.get("L_D_gp", vals["L_D"]).float().reshape(())])
        loss_host.copy_(t, non_blocking=True)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    warm = max(args.warmup, 3)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(warm):
        step_resident()
    torch.cuda.synchronize()

    sampler.mark_begin()
    ms = timed(step_resident, args.steps)
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None

    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    torch.cuda.synchronize()
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = loss_host.numel() * loss_host.element_size()

    # ---- per-kernel timing of the GEMM-class launches over real steps (CUDA events on the launch stream)
    roof = None
    # every rank runs these steps (they contain collectives); only rank 0 records the events
    nprof = max(1, min(3, args.steps))
    if rank == 0:
        ops.PROFILE = {}
    for _ in range(nprof):
        step_eager()
    torch.cuda.synchronize()
    if rank == 0:
        prof, ops.PROFILE = ops.PROFILE, None
        agg = {}
        for key, evs in prof.items():
            tot = sum(a.elapsed_time(b) for a, b in evs)
            agg[key] = (tot, len(evs))
        step_total = sum(t for t, _ in agg.values())
        top_key, (top_ms, top_n) = max(agg.items(), key=lambda kv: kv[1][0])
        kind, (b_, h_, w_, c_, n_, taps_) = top_key
        flops = 2.0 * b_ * h_ * w_ * c_ * n_ * taps_
        breakdown = []
        for (kd, shp), (t_ms, n_l) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
            fl = 2.0 * shp[0] * shp[1] * shp[2] * shp[3] * shp[4] * shp[5]
            breakdown.append({"kernel": kd, "B,H,W,C,N,taps": list(shp), "launches_per_step": n_l // nprof,
                              "avg_ms": round(t_ms / n_l, 4), "ms_per_step": round(t_ms / nprof, 3),
                              "tflops": round(fl / (t_ms / n_l * 1e-3) / 1e12, 1)})
        peak_burst, peak_sus, hbm, src = measured_peaks()
        achieved = flops / (top_ms / top_n * 1e-3) / 1e12
        # algorithmic bytes of the launch: activations in + out once (16-bit), packed weights once; wgrad reads g and x
        es = 2 if dtype != torch.float32 else 4
        pix = b_ * h_ * w_
        abytes = es * pix * (c_ + n_) + (es if kind == "conv_gemm" else 4) * taps_ * c_ * n_
        ridge = peak_sus * 1e12 / (hbm * 1e9)               # FLOP per byte where the two roofs meet
        hbm_bound = flops / abytes < ridge
        achieved_gbs = abytes / (top_ms / top_n * 1e-3) / 1e9
        roof = {"bound": "hbm" if hbm_bound else "tensor",
                "kernel": f"{kind} B{b_} {h_}x{w_} C{c_}->N{n_} taps{taps_}",
                "achieved": round(achieved_gbs if hbm_bound else achieved, 1), "peak": hbm if hbm_bound else peak_sus,
                "unit": "GB/s" if hbm_bound else "TFLOP/s",
                "frac": round((achieved_gbs / hbm) if hbm_bound else (achieved / peak_sus), 4),
                "peak_source": (f"hbm_gbs ({src})" if hbm_bound else f"bf16_tflops_sustained ({src})"),
                "arithmetic_intensity_flop_per_byte": round(flops / abytes, 1), "algorithmic_bytes": abytes,
                "achieved_tflops": round(achieved, 1),
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel/shape from the committed
                # `ncu --set full` capture profiles/r1_ncu_gemm_full.md (71.9 MB read + 22.3 MB written; the
                # algorithmic bytes are 67.1 MB in + 4.7 MB weights + 67.1 MB out, most of the output is still
                # in L2 when the kernel ends)
                # (and profiles/r1_ncu_rows.md for the 1024^2 16->16 conv_rows launch: 139.8 MB read + 82.1 MB written)
                "traffic": {("conv_gemm", 64, 32, 32, 512, 512, 9): 94.2e6,
                            ("conv_gemm", 4, 1024, 1024, 16, 16, 9): 221.9e6}.get((kind, b_, h_, w_, c_, n_, taps_)),
                "launches_per_step": top_n // nprof,
                "avg_launch_ms": round(top_ms / top_n, 4),
                "share_of_gemm_time": round(top_ms / step_total, 3),
                "gemm_ms_per_step": round(step_total / nprof, 3),
                "gemm_breakdown": breakdown,
                "whole_step_frac_of_peak": round(gflop_img * batch / (ms / args.steps * 1e-3) / 1e3 / peak_sus, 4)}

    # ---- CPU baseline: the oracle restatement of the reference step on the host cores (bounded sample)
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.workload, args.variant, sample_batch=args.cpu_sample, steps=2)

    if rank == 0:
        n_img = batch * world * args.steps
        out = {
            "metric": METRIC, "value": round(n_img / (ms * 1e-3), 2), "unit": "img/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": round(ms / args.steps, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {torch.bfloat16: "bf16", torch.float16: "fp16", torch.float32: "fp32(3xTF32)"}[dtype], "data": "synthetic",
            "config": {"workload": args.workload, "resolution": res, "batch_per_gpu": batch, "alpha": alpha,
                       "step_variant": args.variant, "latent_dim": 512, "depths": depths,
                       "parallelism": f"dp{world}", "cuda_graph": replay is not None,
                       **({"cuda_graph_error": graph_err} if graph_err else {}),
                       "l2_note": "per-step working set (activations of 3 D + 2 G passes and their gradients) is "
                                  "several GB, far above the 126 MB L2; no explicit flush",
                       "flop_count": f"{gflop_img} GFLOP/img (BASELINE.md W-minimal: D wgrad skipped in the G step)",
                       "precision_note": "bf16 is the default (BASELINE config); --dtype fp16 runs the same kernels at "
                                         "the same speed with an 11-bit significand (forward parity 1e-3 instead of "
                                         "1e-2, DESIGN.md section 5)"},
            "clocks": clocks,
            "e2e": {"value": round(n_img / (ms_e2e * 1e-3), 2), "unit": "img/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": round(ms_e2e / args.steps, 3)},
            "gpu_launches": int(sum(launches_per_step.values()) * args.steps),
            "launches_per_step": launches_per_step,
            "roofline": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        import gc
        import torch.distributed as dist
        sys.stdout.flush()
        # drop the captured graph (it holds NCCL work) before leaving; destroy_process_group() after a graph
        # that captured collectives was seen to hang, so the ranks leave through a barrier + hard exit
        replay = None
        tr._graph = None
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


# ------------------------------------------------------------------------- reference / CPU arm
def cpu_baseline(workload, variant, sample_batch, steps, warmup=0):
    """the reference's algorithm (oracle port) timed on the host cores: one W-step on `sample_batch`
    images of the same workload (the full batch takes minutes on a CPU)"""
    from oracle import pggan_oracle as O
    scale, batch, depths, _ = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    gen = torch.Generator().manual_seed(0)
    g_sd = O.init_generator_sd(512, depths, scale, generator=gen)
    d_sd = O.init_discriminator_sd(depths, scale, generator=gen)
    alpha = 0.5 if scale else 0.0
    tr = O.OracleTrainer(g_sd, d_sd, alpha=alpha, variant=variant, latent_dim=512)
    res = 4 * 2 ** scale
    b = min(sample_batch, batch)
    img = torch.rand(b, 3, res, res, generator=gen) * 2 - 1
    for _ in range(warmup):
        tr.train_step(img)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.train_step(img)
    dt = time.perf_counter() - t0
    return {"value": round(b * steps / dt, 4), "unit": "img/s", "cores": cores, "kind": "port",
            "sample": f"{steps} W-step(s) of {workload} on {b} of the {batch} images, fp32, torch CPU, "
                      f"{cores} threads, {dt:.1f} s", "seconds": round(dt, 2)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (the oracle port — the
    Python reference cannot travel to the GPU box), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warm = 1 if args.warmup else 0
    t0 = time.perf_counter()
    cb = cpu_baseline(args.workload, args.variant, sample_batch=args.cpu_sample, steps=steps, warmup=warm)
    scale, batch, depths, _ = WORKLOADS[args.workload]
    out = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "img/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": warm,
        "ms_per_step": round(cb["seconds"] / steps * 1e3, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": args.workload, "resolution": 4 * 2 ** scale, "batch_per_gpu": batch,
                   "alpha": 0.5 if scale else 0.0, "step_variant": args.variant, "latent_dim": 512, "depths": depths},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": round(time.perf_counter() - t0, 1),
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2_32x32_b64_fade", choices=sorted(WORKLOADS))
    ap.add_argument("--variant", default="W", choices=["W", "L"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="launch eagerly instead of a CUDA graph")
    ap.add_argument("--no-graph-dp", action="store_true", help="under data parallelism launch eagerly (default: the "
                    "NCCL all-reduces are captured inside the CUDA graph too)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
