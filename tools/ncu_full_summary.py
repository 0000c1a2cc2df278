"""Summarise an `ncu --set full` capture (ncu -i X.ncu-rep --page raw --csv) of the tensor-core kernels:
duration, tensor-pipe activity, DRAM bytes and throughput, grid, per launch.
    python tools/ncu_full_summary.py gpurun_out/ncu_gemm_full_r2.csv > profiles/r2_ncu_gemm_full.md"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}


def find(pat):
    for h in hdr:
        if re.search(pat, h):
            return col[h]
    return None


c_name = col["Kernel Name"]
c_dur = find(r"^gpu__time_duration.sum$")
c_tensor = find(r"^sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active$")
c_rd, c_wr = find(r"^dram__bytes_read.sum$"), find(r"^dram__bytes_write.sum$")
c_dramp = find(r"^gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed$") or find(r"dram__throughput.avg.pct_of_peak_sustained_elapsed")
c_grid = find(r"^launch__grid_size$")
c_clk = find(r"^sm__cycles_elapsed.max.per_second$")
c_regs = find(r"^launch__registers_per_thread$")
c_smem = find(r"^launch__shared_mem_per_block_dynamic$")
c_issue = find(r"sm__inst_executed.avg.per_cycle_elapsed$") or find(r"smsp__issue_active.avg.pct")
print("# r2 — `ncu --set full --clock-control none --import-source on` of the tensor-core kernels inside the C4 step\n")
print("Command: see tools/gpu_evidence2.sh (20 launches of conv_gemm / wgrad_gemm / conv_rows / wgrad_rows after 250 skipped: D's forward over the real + fake batch (B = 32) and the interpolates (B = 16) of an eager C4 step, fp16 mode; the split-K launches of the 4^2 / 8^2 layers are the 64-wide tiles with 64-128 CTAs).")
print("Every kernel is replayed ~40 times by ncu with cold caches: durations are NOT bench values; the tensor-pipe activity and the\n"
      "DRAM byte counts are what this capture is for.  (The kernel template argument `__nv_bfloat16` is the shared 16-bit instantiation;\n"
      "the operand format — IEEE half here — is a run-time field of the instruction descriptor.)\n")
print("| kernel | grid | duration us | tensor pipe active % | DRAM read MB | DRAM write MB | DRAM throughput % of peak | SM clock GHz | regs | dyn smem KB |")
print("|---|---|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    name = re.sub(r"^void ", "", r[c_name])
    name = re.sub(r"\(.*", "", name).replace("pg::", "")

    def g(c, fmt="{:.1f}", scale=1.0):
        if c is None or r[c] in ("", "n/a"):
            return "-"
        try:
            return fmt.format(float(r[c].replace(",", "")) * scale)
        except ValueError:
            return r[c]
    print(f"| `{name}` | {g(c_grid, '{:.0f}')} | {g(c_dur)} | {g(c_tensor)} | {g(c_rd, '{:.2f}')} | {g(c_wr, '{:.2f}')} | {g(c_dramp)} | "
          f"{g(c_clk, '{:.2f}')} | {g(c_regs, '{:.0f}')} | {g(c_smem, '{:.0f}', 1/1024 if False else 1.0)} |")
