"""Context measurement (not a pytest test, not a bench line): the reference algorithm — the oracle port of
`pggan/model.py:206-255` — on the SAME GPU with stock PyTorch/cuDNN kernels, fp32 with TF32 allowed and under
bf16 autocast.  This is BASELINE.md §4's "reference GPU path (the number to beat)".

    python tests/torch_gpu_baseline.py [workload] [variant]

Lives under tests/ because only tests/, smoke() and bench.py's CPU legs may execute oracle/ code.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402
from oracle import pggan_oracle as O  # noqa: E402


def run(workload="c2_32x32_b64_fade", variant="W", steps=10, warmup=5):
    scale, batch, depths, _ = WORKLOADS[workload]
    out = {"workload": workload, "variant": variant, "batch": batch}
    for mode in ("tf32", "bf16_autocast"):
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.benchmark = True
        gen = torch.Generator().manual_seed(0)
        g_sd = {k: v.cuda() for k, v in O.init_generator_sd(512, depths, scale, generator=gen).items()}
        d_sd = {k: v.cuda() for k, v in O.init_discriminator_sd(depths, scale, generator=gen).items()}
        tr = O.OracleTrainer(g_sd, d_sd, alpha=0.5 if scale else 0.0, variant=variant, latent_dim=512)
        res = 4 * 2 ** scale
        img = torch.rand(batch, 3, res, res, device="cuda") * 2 - 1
        zd, zg = torch.randn(batch, 512, device="cuda"), torch.randn(batch, 512, device="cuda")
        eps = torch.rand(batch, 1, device="cuda")

        def step():
            if mode == "bf16_autocast":
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    tr.train_step(img, z_d=zd, z_g=zg, eps=eps)
            else:
                tr.train_step(img, z_d=zd, z_g=zg, eps=eps)
        try:
            for _ in range(warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[mode] = {"ms_per_step": round(ms, 2), "img_per_s": round(batch / ms * 1e3, 1)}
        except Exception as e:          # e.g. an op without a bf16 double-backward
            out[mode] = {"error": f"{type(e).__name__}: {e}"[:160]}
        del tr, g_sd, d_sd
        torch.cuda.empty_cache()
    out["note"] = "stock PyTorch/cuDNN on the same GPU, includes the reference's float() host syncs of the loss values"
    return out


if __name__ == "__main__":
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2_32x32_b64_fade"
    var = sys.argv[2] if len(sys.argv) > 2 else "W"
    print(json.dumps(run(wl, var)))
