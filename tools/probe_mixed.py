"""GPU probe: which of the round-2 kernel features executes (each case in its own process; a fault kills the context)"""
import subprocess
import sys

CASES = {
    "fp16_fprop_satfinite": """
pg.set_compute_dtype(torch.float16)
x = ops.enter(torch.randn(2, 64, 8, 8, device='cuda'), 'a'); w = torch.randn(64, 64, 3, 3, device='cuda')
y = ops.k_conv(x, ops._pack(w, 'f'), None, 64, 9, 0.05, 0.2); torch.cuda.synchronize(); print(float(y.float().abs().mean()))
""",
    "mixed_fprop_f16xf16": """
pg.set_compute_dtype('mixed')
x = ops.enter(torch.randn(2, 64, 8, 8, device='cuda'), 'a'); w = torch.randn(64, 64, 3, 3, device='cuda')
y = ops.k_conv(x, ops._pack(w, 'f'), None, 64, 9, 0.05, 0.2); torch.cuda.synchronize(); print(float(ops.to_float(y).abs().mean()))
""",
    "mixed_dgrad_bf16xf16": """
pg.set_compute_dtype('mixed')
g = ops.enter(torch.randn(2, 64, 8, 8, device='cuda'), 'g'); w = torch.randn(64, 64, 3, 3, device='cuda')
y = ops.k_conv(g, ops._pack(w, 't', fmt=ops._fmt('g')), None, 64, 9, 0.05, 1.0, xc='g'); torch.cuda.synchronize()
ref = torch.nn.functional.conv_transpose2d(g.float(), w.bfloat16().float(), padding=1) * 0.05
print('rel err', float((y.float() - ref).norm() / ref.norm()))
""",
    "mixed_wgrad_bf16xf16": """
pg.set_compute_dtype('mixed')
g = ops.enter(torch.randn(2, 64, 8, 8, device='cuda'), 'g'); x = ops.enter(torch.randn(2, 64, 8, 8, device='cuda'), 'a')
gw = ops.k_wgrad(g, x, 1.0, (64, 64, 3, 3)); torch.cuda.synchronize()
xr = ops.to_float(x).requires_grad_(False); wr = torch.zeros(64, 64, 3, 3, device='cuda', requires_grad=True)
torch.nn.functional.conv2d(xr, wr, padding=1).backward(g.float())
print('rel err', float((gw - wr.grad).norm() / wr.grad.norm()))
""",
    "mixed_wgrad_rows": """
pg.set_compute_dtype('mixed')
g = ops.enter(torch.randn(1, 64, 8, 128, device='cuda'), 'g'); x = ops.enter(torch.randn(1, 64, 8, 128, device='cuda'), 'a')
gw = ops.k_wgrad(g, x, 1.0, (64, 64, 3, 3)); torch.cuda.synchronize(); print(float(gw.abs().mean()))
""",
    "mixed_pixelnorm_bwd": """
pg.set_compute_dtype('mixed')
x = torch.randn(2, 64, 8, 8, device='cuda', requires_grad=True)
y = ops.PixelNorm.apply(x, 1e-8); y.backward(ops.enter(torch.randn(2, 64, 8, 8, device='cuda'), 'g')); torch.cuda.synchronize(); print(float(x.grad.abs().mean()))
""",
    "mixed_mbstd": """
pg.set_compute_dtype('mixed')
x = torch.randn(8, 64, 4, 4, device='cuda', requires_grad=True)
s = ops.MbstdStat.apply(ops.enter(x, 'a'), 4); (g,) = torch.autograd.grad(s.sum(), x, create_graph=True); (g.float() ** 2).sum().backward()
torch.cuda.synchronize(); print(float(x.grad.abs().mean()))
""",
    "mixed_elementwise": """
pg.set_compute_dtype('mixed')
x = ops.enter(torch.randn(2, 64, 8, 8, device='cuda'), 'a')
a = ops.k_pool(x, 0.25); b = ops.k_upsample(x, 1.0); c = ops.k_axpby(x, x, 0.5, 0.5); torch.cuda.synchronize(); print('ok')
img = torch.rand(2, 3, 8, 8, device='cuda'); o = ops.ImageFade.apply(img, None, 0.3); torch.cuda.synchronize(); print('fade ok')
""",
}
HEAD = "import torch, sys\nsys.path.insert(0, '.')\nimport pggan_b200 as pg\nfrom pggan_b200 import ops\n"
for name, body in CASES.items():
    r = subprocess.run([sys.executable, "-c", HEAD + body], capture_output=True, text=True, env={**__import__('os').environ, "CUDA_LAUNCH_BLOCKING": "1"})
    tail = (r.stdout.strip().splitlines() or [""])[-1]
    err = [ln for ln in r.stderr.strip().splitlines() if "Error" in ln or "error" in ln][-1:] if r.returncode else []
    print(f"{name:28s} rc={r.returncode} {tail} {err}", flush=True)
