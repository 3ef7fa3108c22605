"""Runs the W1 step with every tuning build under enhancedbayesiannetworks.jl_b200/tune/ (one
subprocess per library: EBN_LIB selects it) and prints samples/s. GPU box only."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = sorted(glob.glob(os.path.join(ROOT, "enhancedbayesiannetworks.jl_b200", "tune", "*.so")))
n = sys.argv[1] if len(sys.argv) > 1 else "1e8"
for lib in libs:
    env = dict(os.environ, EBN_LIB=lib)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3",
                          "--no-cpu-baseline", "--n-per-scenario", n], env=env, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print(f"{os.path.basename(lib):32s} {d['value']:.4e} samples/s  {d['ms_per_step']:.3f} ms  pf0={d['pf_first_scenarios'][0]:.6f}", flush=True)
    except Exception as e:
        print(os.path.basename(lib), "FAILED", out.stderr[-300:], flush=True)
