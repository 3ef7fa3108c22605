"""How far the fp32 Box-Muller draws (EBN_JOB_FP32_SAMPLING) are from the fp64 draws of the same
words: percentiles of |dz| over 4e6 normals. GPU box only."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ebn_b200  # noqa: E402
from ebn_b200 import workloads as W  # noqa: E402
from oracle import replay  # noqa: E402

ebn_b200.init()
n = 2_000_000
a = ebn_b200.sample_matrix(W.w1_vehicle(n=n), 3, 0, n)[:, 3:]
b = ebn_b200.sample_matrix(W.w1_vehicle(n=n, fp32_sampling=True), 3, 0, n)[:, 3:]
dz = np.abs(a - b).ravel() / 10.0
print("device fp32 vs device fp64, |dz| percentiles 50/99/99.9/99.99/max:",
      [float(f"{v:.3g}") for v in np.percentile(dz, [50, 99, 99.9, 99.99, 100])])
j = W.w1_vehicle(n=n, fp32_sampling=True)
t = replay.sample_matrix(j.inputs[3], 3, j.seed, 0, 200_000, fp32=True)[:, 3:]
dt = np.abs(t - b[:200_000]).ravel() / 10.0
print("device fp32 vs NumPy float32 twin, |dz| percentiles 50/99/99.9/max:",
      [float(f"{v:.3g}") for v in np.percentile(dt, [50, 99, 99.9, 100])])
z = (b - np.array([431.7221, 1475.5503, 55.0406])) / 10.0
print("fp32 draws: mean", z.mean(0).round(5), "std", z.std(0).round(5), "max |z|", np.abs(z).max().round(3))
