"""The two HBM-bound kernels of the path against the measured copy bandwidth (MEASURED_PEAKS.json:
hbm_gbs): matrix_kernel (ebn_eval_matrix: limit state over a resident column-major matrix, d*8 B read per
row, +8 B written with g_out) and replay_kernel (ebn_sample_matrix: d*8 B written per row). Device
time from the library's CUDA events; the host<->device copies around them are not in it. GPU box only."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ebn_b200  # noqa: E402
from ebn_b200 import workloads as W  # noqa: E402

ebn_b200.init()
peak = None
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
n = 40_000_000
rng = np.random.default_rng(0)
X = np.empty((n, 6), order="F")
X[:, 0], X[:, 1] = 0.8, 0.27
X[:, 2] = rng.uniform(7, 12, n)
for j, mu in zip((3, 4, 5), (431.7221, 1475.5503, 55.0406)):
    X[:, j] = rng.normal(mu, 10, n)
k = W.vehicle_consts()
for strict in (False, True):
    for want_g in (False, True):
        best = 1e30
        for _ in range(3):
            ebn_b200.eval_matrix(ebn_b200._lib.LS_VEHICLE_SUSPENSION, k, X, strict=strict, want_g=want_g)
            best = min(best, ebn_b200.last_stats()["kernel_ms"])
        b = n * (6 * 8 + (8 if want_g else 0))
        gbs = b / best / 1e6
        print(f"matrix_kernel vehicle strict={int(strict)} g_out={int(want_g)}: {best:7.3f} ms  {gbs:7.0f} GB/s"
              + (f"  = {gbs / peak:.2f} of the measured copy bandwidth ({peak:.0f} GB/s)" if peak else ""), flush=True)
