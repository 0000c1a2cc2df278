#!/usr/bin/env python
"""bench.py — PGGAN D+G WGAN-GP training step throughput (BASELINE.json metric: train imgs/sec).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own modules on the host CPU

Headline workload = BASELINE.json configs[3], the configuration north_star sets its 60 %-of-peak target on: the 256x256
stage, batch 16 per GPU, latent 512, depths [512,512,512,512,256,128,64], alpha 0.5, one full D+G step of the WGAN-GP
variant (SURVEY.md §8d "W": BCE adversarial terms + gradient penalty with its own backward + drift, Adam on D and G),
in fp16 mode (IEEE-half activations, weights and gradient tensors, fp32 accumulate, static loss scale folded into
Adam — the 16-bit mode that meets north_star's parity tolerances), synthetic uniform images, reference initialisation.  The other BASELINE configs
(C1 4x4, C2 32x32, C3 128x128, C5 1024x1024) are measured in the same run and reported in `per_config`.
N>1: one rank per GPU (torchrun), same per-GPU batch (weak scaling), bucketed gradient all-reduce overlapped with backward.

One JSON line on stdout (rank 0).  `value` = images/s with inputs resident in HBM; `e2e` = the same step driven from
pinned HOST buffers (H2D of images/latents/eps and D2H of the losses inside the timed region); `roofline` = the
kernel class with the largest share of the step's GEMM time against the measured bf16 tensor peak (or HBM peak when
its arithmetic intensity is below the ridge); `cpu_baseline` = the reference step on the host cores (bounded sample).
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
HEADLINE = "c4_256x256_b16"
METRIC = "train imgs/sec (D+G step incl. WGAN-GP)"
DTYPES = {"mixed": "mixed", "bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}
DTYPE_NAMES = {"mixed": "fp16 activations/weights x bf16 gradients, fp32 accumulate (mixed)", "bf16": "bf16",
               "fp16": "fp16 (IEEE-half activations, weights and gradient tensors; fp32 accumulate; static loss scale 1024)",
               "fp32": "fp32(3xTF32)"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


def committed_traffic():
    """dram bytes per launch of the profiled kernels, from the committed `ncu --set full` captures
    (profiles/r2_ncu_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep of this round)"""
    p = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh)
    return {}


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


def dist_setup():
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
def kernel_class(kind, shape, dtype_enum_fn):
    """the kernel variant the library runs a GEMM-class launch on (conv_gemm / conv_gemm_swapped / conv_rows /
    conv_direct, wgrad_gemm / wgrad_rows / wgrad_simt) — asked from the library itself"""
    import ctypes
    from pggan_b200 import _lib as L
    b, h, w, c, n, taps = shape
    if kind == "conv_gemm":
        p = L.ConvParams()
        p.B, p.H, p.W, p.C, p.N, p.taps = b, h, w, c, n, taps
        p.dtype = p.out_dtype = dtype_enum_fn("a")
        k = L.load().pggan_conv_kernel_kind(ctypes.byref(p))
        return {L.PG_CONV_KIND_PLAIN: "conv_gemm", L.PG_CONV_KIND_SWAPPED: "conv_gemm_swapped",
                L.PG_CONV_KIND_ROWS: "conv_rows"}.get(k, "conv_direct(SIMT)")
    rows_ok = taps == 9 and w % 128 == 0 and h >= 2 and c in (16, 32, 64) and n in (16, 32, 64) and dtype_enum_fn("g") != L.PG_F32
    if rows_ok:
        return "wgrad_rows"
    return "wgrad_gemm" if (c % 64 == 0 and n % 8 == 0) else "wgrad_simt"


def bench_config(args, workload, steps, warm, full, rank, world, local):
    """one BASELINE configuration: returns the dict of its measurements (rank 0; other ranks run the same steps)"""
    import pggan_b200 as pg
    from pggan_b200 import _lib, ops
    dev = torch.device("cuda", local)
    # nvidia-smi needs up to a second before its first sample: start it now (model build, eager steps and the graph
    # capture follow), so that it is streaming long before the timed region
    sampler = ClockSampler(local)
    if rank == 0 and full:
        sampler.start()
    scale, batch, depths, gflop_img = WORKLOADS[workload]
    if args.batch and full:
        batch = args.batch
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
    if os.environ.get("PGGAN_PROFILE_ONE_STEP"):
        # profiler hook (tools/gpu_ncu_step.sh): exactly ONE warm eager step between cudaProfilerStart/Stop, then exit —
        # `ncu --profile-from-start off` then lists the launches of one step and nothing else.  Not a bench value.
        torch.cuda.profiler.start()
        step_eager()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        if rank == 0:
            print(json.dumps({"profiled_one_step": workload, "launches_per_step": launches_per_step}))
        sys.exit(0)

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
        vals = tr.loss_collector.loss_dict.tensors()
        t = torch.stack([vals["L_D"].float().reshape(()), vals["L_G"].float().reshape(()),
                         vals.get("L_D_gp", vals["L_D"]).float().reshape(())])
        loss_host.copy_(t, non_blocking=True)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
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

    for _ in range(warm):
        step_resident()
    torch.cuda.synchronize()
    sampler.mark_begin()
    ms = timed(step_resident, steps)
    sampler.mark_end()
    clocks = sampler.stop() if (rank == 0 and full) else None

    e2e = None
    if full:
        for _ in range(2):
            step_e2e()
        ms_e2e = timed(step_e2e, steps)
        torch.cuda.synchronize()
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        d2h = loss_host.numel() * loss_host.element_size()
        e2e = {"value": round(batch * world * steps / (ms_e2e * 1e-3), 2), "unit": "img/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": round(ms_e2e / steps, 3)}

    # ---- per-kernel timing of the GEMM-class launches over real (eagerly launched) steps: CUDA events on the launch
    #      stream around every launch.  Every rank runs these steps (they contain collectives); rank 0 records.
    nprof = 2 if full else 1
    if rank == 0:
        ops.PROFILE = {}
    for _ in range(nprof):
        step_eager()
    torch.cuda.synchronize()
    roof = None
    if rank == 0:
        prof, ops.PROFILE = ops.PROFILE, None
        peak_burst, peak_sus, hbm, src = measured_peaks()
        es = 2 if args.dtype != "fp32" else 4
        groups, shapes = {}, []
        for (kind, shp), evs in prof.items():
            t_ms = sum(a.elapsed_time(b) for a, b in evs)
            b_, h_, w_, c_, n_, taps_ = shp
            fl = 2.0 * b_ * h_ * w_ * c_ * n_ * taps_
            pix = b_ * h_ * w_
            # algorithmic bytes of one launch: activations in + out once (16-bit), packed weights once; wgrad reads g
            # and x and writes the fp32 weight gradient
            by = es * pix * (c_ + n_) + (es if kind == "conv_gemm" else 4) * taps_ * c_ * n_
            cls = kernel_class(kind, shp, ops._fmt)
            grp = groups.setdefault(cls, {"ms": 0.0, "launches": 0, "flops": 0.0, "bytes": 0.0})
            grp["ms"] += t_ms
            grp["launches"] += len(evs)
            grp["flops"] += fl * len(evs)
            grp["bytes"] += by * len(evs)
            shapes.append({"kernel": cls, "B,H,W,C,N,taps": list(shp), "launches_per_step": len(evs) // nprof,
                           "avg_ms": round(t_ms / len(evs), 4), "ms_per_step": round(t_ms / nprof, 3),
                           "tflops": round(fl / (t_ms / len(evs) * 1e-3) / 1e12, 1),
                           "gbs": round(by / (t_ms / len(evs) * 1e-3) / 1e9, 1)})
        gemm_total = sum(v["ms"] for v in groups.values())
        top = max(groups, key=lambda k: groups[k]["ms"])            # the kernel class with the largest TIME share
        gt = groups[top]
        avg_ms = gt["ms"] / gt["launches"]
        ach_tf = gt["flops"] / (gt["ms"] * 1e-3) / 1e12
        ach_gb = gt["bytes"] / (gt["ms"] * 1e-3) / 1e9
        ridge = peak_sus * 1e12 / (hbm * 1e9)
        hbm_bound = gt["flops"] / gt["bytes"] < ridge
        traffic = committed_traffic().get(workload, {}).get(top)
        roof = {"bound": "hbm" if hbm_bound else "tensor", "kernel": top,
                "achieved": round(ach_gb if hbm_bound else ach_tf, 1), "peak": hbm if hbm_bound else peak_sus,
                "unit": "GB/s" if hbm_bound else "TFLOP/s",
                "frac": round((ach_gb / hbm) if hbm_bound else (ach_tf / peak_sus), 4),
                "peak_source": (f"hbm_gbs ({src})" if hbm_bound else
                                f"bf16_tflops_sustained ({src}): a kernel timed inside a long step"),
                "frac_of_burst_peak": None if hbm_bound else round(ach_tf / peak_burst, 4),
                "arithmetic_intensity_flop_per_byte": round(gt["flops"] / gt["bytes"], 1),
                "algorithmic_bytes_per_launch": round(gt["bytes"] / gt["launches"]),
                "algorithmic_flops_per_launch": round(gt["flops"] / gt["launches"]),
                "traffic": traffic,
                "launches_per_step": gt["launches"] // nprof, "avg_launch_ms": round(avg_ms, 4),
                "share_of_gemm_time": round(gt["ms"] / gemm_total, 3),
                "share_of_step_time": round(gt["ms"] / nprof / (ms / steps), 3),
                "gemm_ms_per_step_eager_events": round(gemm_total / nprof, 3),
                "kernel_classes": {k: {"ms_per_step": round(v["ms"] / nprof, 3), "launches_per_step": v["launches"] // nprof,
                                       "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1),
                                       "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1)}
                                   for k, v in sorted(groups.items(), key=lambda kv: -kv[1]["ms"])},
                "top_shapes": sorted(shapes, key=lambda d: -d["ms_per_step"])[:12 if full else 5],
                "whole_step_frac_of_peak": round(gflop_img * batch / (ms / steps * 1e-3) / 1e3 / peak_sus, 4)}

    # ---- CUPTI timeline of two replayed steps: true warm kernel durations by kernel name (explains the step)
    breakdown = None
    if full and rank == 0 and args.timeline:
        try:
            breakdown = cupti_breakdown(step_resident, 2)
        except Exception as e:
            breakdown = {"error": f"{type(e).__name__}: {e}"[:160]}
    elif full and world > 1 and args.timeline:
        for _ in range(2):              # keep the ranks in step with rank 0's profiled replays
            step_resident()
        torch.cuda.synchronize()

    out = {"workload": workload, "resolution": res, "batch_per_gpu": batch, "alpha": alpha, "depths": depths,
           "value": round(batch * world * steps / (ms * 1e-3), 2), "unit": "img/s", "ms_per_step": round(ms / steps, 3),
           "steps": steps, "cuda_graph": replay is not None,
           "flop_count": f"{gflop_img} GFLOP/img (BASELINE.md W-minimal: D wgrad skipped in the G step)",
           "launches_per_step": launches_per_step, "roofline": roof}
    if graph_err:
        out["cuda_graph_error"] = graph_err
    if roof:
        out["whole_step_frac_of_peak"] = roof["whole_step_frac_of_peak"]
    if full:
        out.update({"clocks": clocks, "e2e": e2e, "step_breakdown": breakdown})
    # free this configuration before the next one (graphs hold their private pools)
    replay = None
    tr._graph = None
    del tr
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return out


def cupti_breakdown(step, n):
    import collections
    import re
    from torch.profiler import profile, ProfilerActivity
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    kern = [(e.time_range.start, e.time_range.end, e.name) for e in evs if "Memcpy" not in e.name and "Memset" not in e.name]
    if not kern:
        return None
    kern.sort(key=lambda t: t[0])
    t0, t1 = kern[0][0], max(k[1] for k in kern)
    busy = sum(k[1] - k[0] for k in kern)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for s, e, name in kern:
        name = re.sub(r"pg::|void |\(.*", "", name)
        name = re.sub(r"at::native::.*", "at::native (eager ATen elementwise/reduce)", name)[:64]
        agg[name][0] += 1
        agg[name][1] += e - s
    rows = [{"kernel": k, "ms_per_step": round(t / n / 1e3, 3), "launches_per_step": round(c / n, 1),
             "share": round(t / busy, 4)} for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]]
    pdl = os.environ.get("PGGAN_PDL", "0") not in ("", "0")
    out = {"span_ms_per_step": round((t1 - t0) / n / 1e3, 3), "kernel_busy_ms_per_step": round(busy / n / 1e3, 3),
           "launches_per_step": round(len(kern) / n, 1), "programmatic_dependent_launch": pdl, "kernels": rows}
    if pdl:
        out["note"] = ("kernels are pre-launched under programmatic dependent launch: a kernel's duration here includes its "
                       "wait for the predecessor, so durations overlap (kernel-busy > span); run with PGGAN_PDL=0 for "
                       "exclusive per-kernel times (profiles/r2_step_breakdown.md)")
    return out


def run_ours(args):
    # the CPU baseline runs FIRST, alone, on rank 0 of a single-rank launch only (under torchrun the other ranks would
    # spin in an NCCL barrier on the same host cores and contaminate it)
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    cpu = None
    if world_env == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.workload, args.variant, sample_batch=args.cpu_sample, steps=1)
    rank, world, local = dist_setup()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    import pggan_b200 as pg
    pg.set_compute_dtype(DTYPES[args.dtype])
    warm = max(args.warmup, 3)
    main = bench_config(args, args.workload, args.steps, warm, True, rank, world, local)
    per_config = []
    if not args.no_per_config:
        sub_steps = max(2, min(args.steps, 5))
        for wl in WORKLOADS:
            if wl == args.workload:
                continue
            try:
                r = bench_config(args, wl, sub_steps, 3, False, rank, world, local)
                if r["roofline"]:
                    rf = r["roofline"]
                    r["roofline"] = {k: rf[k] for k in ("bound", "kernel", "achieved", "peak", "unit", "frac", "share_of_gemm_time",
                                                        "share_of_step_time", "avg_launch_ms", "kernel_classes")}
                per_config.append(r)
            except Exception as e:
                per_config.append({"workload": wl, "error": f"{type(e).__name__}: {e}"[:200]})
                torch.cuda.synchronize()
    if rank == 0:
        out = {
            "metric": METRIC, "value": main["value"], "unit": "img/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": main["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE_NAMES[args.dtype], "data": "synthetic",
            "config": {"workload": args.workload, "resolution": main["resolution"], "batch_per_gpu": main["batch_per_gpu"],
                       "alpha": main["alpha"], "step_variant": args.variant, "latent_dim": 512, "depths": main["depths"],
                       "parallelism": f"dp{world}", "cuda_graph": main["cuda_graph"],
                       **({"cuda_graph_error": main["cuda_graph_error"]} if "cuda_graph_error" in main else {}),
                       "l2_note": "per-step working set (activations of 3 D + 2 G passes and their gradients) is "
                                  "several GB, far above the 126 MB L2; no explicit flush",
                       "flop_count": main["flop_count"],
                       "precision_note": "fp16 = IEEE-half activations, packed weights and gradient tensors, fp32 accumulate "
                                         "and epilogues, saturating conversions, static loss scale 1024 divided out by "
                                         "the Adam kernel; meets north_star's 1e-2 (images/logits) and 1e-3 (gradient "
                                         "penalty) bars at C1-C5 real widths (tests/test_gpu_nets.py)"},
            "clocks": main["clocks"], "e2e": main["e2e"],
            "gpu_launches": int(sum(main["launches_per_step"].values()) * args.steps),
            "launches_per_step": main["launches_per_step"],
            "whole_step_frac_of_peak": main.get("whole_step_frac_of_peak"),
            "roofline": main["roofline"], "cpu_baseline": cpu, "step_breakdown": main.get("step_breakdown"),
            "per_config": per_config,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        import gc
        import torch.distributed as dist
        sys.stdout.flush()
        # captured graphs held NCCL work; destroy_process_group() after such a graph was seen to hang, so the ranks
        # leave through a barrier + hard exit
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


# ------------------------------------------------------------------------- reference / CPU arm
def cpu_baseline(workload, variant, sample_batch, steps, warmup=0):
    """the reference step timed on the host cores: `steps` W-step(s) on `sample_batch` images of the workload (the full
    batch takes minutes on a CPU).  Runs the reference's OWN modules staged in oracle/_ref (kind "reference"); if the
    stage is absent, the oracle port of the same algorithm (kind "port")."""
    from oracle import stage_ref
    scale, batch, depths, _ = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    alpha = 0.5 if scale else 0.0
    res = 4 * 2 ** scale
    b = min(sample_batch, batch)
    gen = torch.Generator().manual_seed(0)
    img = torch.rand(b, 3, res, res, generator=gen) * 2 - 1
    if stage_ref.available():
        from oracle.ref_step import ReferenceTrainer
        tr = ReferenceTrainer(depths, scale, alpha=alpha, variant=variant, latent_dim=512, device="cpu")
        kind = "reference"
    else:
        from oracle import pggan_oracle as O
        g_sd = O.init_generator_sd(512, depths, scale, generator=gen)
        d_sd = O.init_discriminator_sd(depths, scale, generator=gen)
        tr = O.OracleTrainer(g_sd, d_sd, alpha=alpha, variant=variant, latent_dim=512)
        kind = "port"
    for _ in range(warmup):
        tr.train_step(img)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.train_step(img)
    dt = time.perf_counter() - t0
    return {"value": round(b * steps / dt, 4), "unit": "img/s", "cores": cores, "kind": kind,
            "sample": f"{steps} W-step(s) of {workload} on {b} of the {batch} images per step, fp32, torch CPU, "
                      f"{cores} threads, {dt:.1f} s" + (" (the reference's own Generator/Discriminator/WGANGPLoss from "
                                                        "oracle/_ref)" if kind == "reference" else " (oracle port)"),
            "seconds": round(dt, 2)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (its modules staged in oracle/_ref, else the
    oracle port), all host threads, bounded sample per step.  Rank 0 alone runs; other ranks exit 0."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 2))
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
    ap.add_argument("--workload", default=HEADLINE, choices=sorted(WORKLOADS))
    ap.add_argument("--variant", default="W", choices=["W", "L"])
    ap.add_argument("--dtype", default="fp16", choices=sorted(DTYPES))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=4, help="images per step of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-per-config", action="store_true", help="measure only --workload")
    ap.add_argument("--no-timeline", dest="timeline", action="store_false", help="skip the CUPTI kernel-name breakdown")
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
