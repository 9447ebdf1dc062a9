"""Summarise an ncu report (run here, no GPU): one row per profiled launch with duration, pipe utilisation, DRAM traffic and
achieved GB/s / TFLOP/s against the measured peaks.      python tools/ncu_summary.py gpurun_out/x.ncu-rep [out.txt]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
want = [("gpu__time_duration.sum", "us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu%"),
        ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid")]


def val(r, k):
    if k not in ix:
        return float("nan")
    try:
        v = float(r[ix[k]].replace(",", ""))
    except ValueError:
        return float("nan")
    u = units[ix[k]]
    if k.startswith("dram__bytes"):
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    if k == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    return v


lines = [f"# {os.path.basename(rep)}: per-launch ncu metrics (cold-cache, serialised under the profiler); HBM peak {peaks['hbm_gbs']} GB/s, "
         f"bf16 burst peak {peaks['bf16_tflops']} TFLOP/s (MEASURED_PEAKS.json)",
         "kernel | " + " | ".join(n for _, n in want) + " | dram GB/s | % of HBM peak"]
for r in body:
    name = r[ix["Kernel Name"]][:46]
    v = [val(r, k) for k, _ in want]
    gbs = (v[6] + v[7]) / v[0] * 1e-3 * 1e3 if v[0] else float("nan")     # MB / us = TB/s -> GB/s * 1e3
    gbs = (v[6] + v[7]) / v[0] * 1e3
    lines.append(f"{name} | " + " | ".join(f"{x:.1f}" for x in v) + f" | {gbs:.0f} | {100 * gbs / peaks['hbm_gbs']:.1f}")
txt = "\n".join(lines)
print(txt)
if out:
    open(out, "w").write(txt + "\n")
