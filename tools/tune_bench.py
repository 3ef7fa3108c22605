"""Runs a workload with every tuning build under enhancedbayesiannetworks.jl_b200/tune/ (one
subprocess per library: EBN_LIB selects it) and prints samples/s. GPU box only.
    python tools/tune_bench.py [w1|w2] [n]"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = sorted(glob.glob(os.path.join(ROOT, "enhancedbayesiannetworks.jl_b200", "tune", "*.so")))
wl = sys.argv[1] if len(sys.argv) > 1 else "w1"
n = sys.argv[2] if len(sys.argv) > 2 else ("1e8" if wl == "w1" else "1e9")
code = f"""
import sys; sys.path.insert(0, {ROOT!r})
import ebn_b200
from ebn_b200 import workloads as W
ebn_b200.init()
job = W.w1_vehicle(n=int({n})) if {wl!r} == 'w1' else W.w2_straub(n=int({n}))
best = 1e30
for _ in range(5):
    cnt, pf, _ = ebn_b200.pf_batch(job)
    best = min(best, ebn_b200.last_stats()['kernel_ms'])
print(f"{{job.S * job.n / best / 1e-3:.4e}} samples/s  {{best:.3f}} ms  pf0={{pf[0]:.6f}}")
"""
for lib in libs:
    env = dict(os.environ, EBN_LIB=lib)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(f"{os.path.basename(lib):36s} {out.stdout.strip() or out.stderr[-300:]}", flush=True)
