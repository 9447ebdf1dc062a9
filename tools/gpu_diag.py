"""Run every GPU test node in its own process with a timeout (a trapping / hanging kernel must not take the other
checks down) and write a summary to gpurun_out/diag.txt.   python tools/gpu_diag.py [pytest -k expr] """
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
files = [a for a in sys.argv[1:] if a.endswith(".py")] or ["tests/test_kernels_gpu.py"]
kexpr = [a for a in sys.argv[1:] if not a.endswith(".py")]
col = subprocess.run([sys.executable, "-m", "pytest", "--collect-only", "-q", "-m", "gpu", *files] +
                     (["-k", kexpr[0]] if kexpr else []), cwd=ROOT, capture_output=True, text=True)
nodes = [l.strip() for l in col.stdout.splitlines() if "::" in l]
print(f"{len(nodes)} nodes", flush=True)
lines = []
for n in nodes:
    t0 = time.time()
    try:
        r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-s", "-m", "gpu", n, "--no-header", "-p", "no:cacheprovider"],
                           cwd=ROOT, capture_output=True, text=True, timeout=240)
        ok = r.returncode == 0
        for l in r.stdout.splitlines():
            if '{"' in l and len(l) < 600:
                print(l, flush=True)
                lines.append(l)
        tail = "" if ok else "\n".join((r.stdout + r.stderr).splitlines()[-25:])
    except subprocess.TimeoutExpired:
        ok, tail = False, "TIMEOUT"
    msg = f"{'PASS' if ok else 'FAIL'} {time.time() - t0:6.1f}s {n}"
    print(msg, flush=True)
    lines.append(msg)
    if not ok:
        print(tail, flush=True)
        lines.append(tail)
with open(os.path.join(ROOT, "gpurun_out", "diag.txt"), "w") as f:
    f.write("\n".join(lines) + "\n")
