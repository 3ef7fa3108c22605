"""Runs a workload with every tuning build under enhancedbayesiannetworks.jl_b200/tune/ (one
subprocess per library: EBN_LIB selects it) and prints samples/s. GPU box only.
    python tools/tune_bench.py [w1|w2|w3|poly|hist|w1t2] [n]
(poly / hist: A + B through the interpreted polynomial functor and the runtime-dispatched plan, count and
histogram mode; w1t2: W1 with C truncated and K log-normal, runtime-dispatched plan; w3: copula, looping plan)"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = sorted(glob.glob(os.path.join(ROOT, "enhancedbayesiannetworks.jl_b200", "tune", "*.so")))
wl = sys.argv[1] if len(sys.argv) > 1 else "w1"
n = sys.argv[2] if len(sys.argv) > 2 else {"w2": "1e9", "w3": "2e7"}.get(wl, "1e8")
code = f"""
import sys; sys.path.insert(0, {ROOT!r})
import ebn_b200
from ebn_b200 import workloads as W
ebn_b200.init()
import numpy as np
wl, n = {wl!r}, int({n})
if wl == 'w1':
    job = W.w1_vehicle(n=n)
elif wl == 'w2':
    job = W.w2_straub(n=n)
elif wl == 'w3':
    job = W.w3_straub_copula(n=n)
    job.jit_plan = False
elif wl == 'w1t2':
    job = W.w1_vehicle(n=n)
    job.inputs[:, 3]["lo"] = 400.0
    job.inputs[:, 5]["kind"], job.inputs[:, 5]["p"] = ebn_b200._lib.IN_LOGNORMAL, (4.0, 0.18, 0, 0)
    job.jit_plan = False
else:
    prog = np.array([1, 2, 1.0, 1, 0, 1.0, 1, 1])
    rows = [[(0, a, 0, 0, 0, -np.inf, np.inf), (1, 0.0, 1.0, 0, 0, -np.inf, np.inf)] for a in np.linspace(0, 3, 20)]
    job = ebn_b200.Job(ebn_b200._lib.LS_POLYNOMIAL, prog, ebn_b200.make_inputs(rows), n=n, seed=3, jit_plan=False)
best = 1e30
for _ in range(5):
    if wl == 'hist':
        pf = ebn_b200.hist_batch(job, np.linspace(-8, 11, 1025))[0][:, 0] / job.n
    else:
        cnt, pf, _ = ebn_b200.pf_batch(job)
    best = min(best, ebn_b200.last_stats()['kernel_ms'])
print(f"{{job.S * job.n / best / 1e-3:.4e}} samples/s  {{best:.3f}} ms  pf0={{pf[0]:.6f}}")
"""
for lib in libs:
    env = dict(os.environ, EBN_LIB=lib)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(f"{os.path.basename(lib):36s} {out.stdout.strip() or out.stderr[-300:]}", flush=True)
