"""Runs one secondary workload a few times (for ncu): python tools/run_workload.py w1|w2|w3|w1t2|poly|hist [total samples]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ebn_b200  # noqa: E402
from ebn_b200 import workloads as W  # noqa: E402

ebn_b200.init()
wl = sys.argv[1]
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10**8
if wl == "w1":
    job = W.w1_vehicle(n=n // 20)
elif wl == "w2":
    job = W.w2_straub(n=n)
elif wl == "w3":
    job = W.w3_straub_copula(n=n // 64)
    job.jit_plan = False
elif wl == "w1t2":   # W1 with C truncated and K log-normal: runtime-dispatched plan, Phi^-1 cells in dynamic shared memory
    job = W.w1_vehicle(n=n // 20)
    job.inputs[:, 3]["lo"] = 400.0
    job.inputs[:, 5]["kind"], job.inputs[:, 5]["p"] = ebn_b200._lib.IN_LOGNORMAL, (4.0, 0.18, 0, 0)
    job.jit_plan = False
else:
    prog = np.array([1, 2, 1.0, 1, 0, 1.0, 1, 1])
    rows = [[(0, a, 0, 0, 0, -np.inf, np.inf), (1, 0.0, 1.0, 0, 0, -np.inf, np.inf)] for a in np.linspace(0, 3, 20)]
    job = ebn_b200.Job(ebn_b200._lib.LS_POLYNOMIAL, prog, ebn_b200.make_inputs(rows), n=n // 20, seed=3)
for _ in range(4):
    if wl == "hist":
        ebn_b200.hist_batch(job, np.linspace(-8, 11, 1025))
    else:
        cnt, pf, _ = ebn_b200.pf_batch(job)
    print(ebn_b200.last_stats()["kernel_ms"])
