"""Throughput of the secondary BASELINE configurations (W2 Straub, W3 copula, W5 grid, the histogram
path and the strict-arithmetic variant) on one GPU: device time from the library's CUDA events.
GPU box only.  python tools/bench_workloads.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ebn_b200  # noqa: E402
from ebn_b200 import workloads as W  # noqa: E402

ebn_b200.init()


def run(name, job, reps=3, hist_edges=None):
    best = 1e30
    for _ in range(reps + 1):
        if hist_edges is None:
            cnt, pf, _ = ebn_b200.pf_batch(job)
        else:
            ebn_b200.hist_batch(job, hist_edges)
        st = ebn_b200.last_stats()
        best = min(best, st["kernel_ms"])
    total = job.S * job.n
    extra = "" if hist_edges is not None else f" pf[0]={pf[0]:.6f}"
    print(f"{name:46s} S={job.S:5d} n={job.n:.1e}  {best:9.3f} ms  {total / best / 1e-3:.4e} samples/s  "
          f"static_plan={st['specialised']}{extra}", flush=True)


run("W1 vehicle (fast)", W.w1_vehicle(n=10**8))
run("W1 vehicle (strict: Julia-order g, 6 div + sqrt)", W.w1_vehicle(n=10**8, strict=True))
run("W2 Straub frame S=1", W.w2_straub(n=10**9))
run("W3 Straub-style copula, 64 scenarios", W.w3_straub_copula(n=2 * 10**7))
run("W5 vehicle grid, 1024 scenarios", W.w5_vehicle_grid(n=4 * 10**6))
run("W4 vehicle, V fixed: 4 states x 25 outer-loop values", W.w4_vehicle_fixed_v(np.linspace(7, 12, 25), n=2 * 10**7))
j = W.w1_vehicle(n=10**8)
j.inputs[:, 3]["lo"] = 400.0   # truncated normals: inverse-CDF path through the generic plan
run("W1 with C truncated at 400 (generic plan)", j)
prog = np.array([1, 2, 1.0, 1, 0, 1.0, 1, 1])
rows = [[(0, a, 0, 0, 0, -np.inf, np.inf), (1, 0.0, 1.0, 0, 0, -np.inf, np.inf)] for a in np.linspace(0, 3, 20)]
hj = ebn_b200.Job(ebn_b200._lib.LS_POLYNOMIAL, prog, ebn_b200.make_inputs(rows), n=10**8, seed=3)
run("histogram path A+B, 2048 edges, 20 scenarios", hj, hist_edges=np.linspace(-8, 11, 1025))
run("count path A+B (polynomial functor, generic plan)", hj)
