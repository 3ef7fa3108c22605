"""Small end-to-end run for compute-sanitizer (one tool per gpurun call): every kernel family once."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ebn_b200  # noqa: E402
from ebn_b200 import workloads as W  # noqa: E402

ebn_b200.init()
print(ebn_b200.pf_batch(W.w1_vehicle(n=20_001))[1][:3])
print(ebn_b200.pf_batch(W.w1_vehicle(n=20_001, strict=True))[1][:3])
print(ebn_b200.pf_batch(W.w2_straub(n=30_000))[1])
print(ebn_b200.pf_batch(W.w3_straub_copula(n=5_000, bits=2))[1])
j = W.w1_vehicle(n=10_000)
j.inputs[:, 3]["lo"] = 400.0
print(ebn_b200.pf_batch(j)[1][:3])
X = ebn_b200.sample_matrix(W.w2_straub(n=1000), 0, 3, 501, want_g=True)
print(X[0].shape, ebn_b200.eval_matrix(2, W.straub_consts(), X[0]))
prog = np.array([1, 2, 1.0, 1, 0, 1.0, 1, 1])
rows = [[(0, a, 0, 0, 0, -np.inf, np.inf), (1, 0.0, 1.0, 0, 0, -np.inf, np.inf)] for a in (0.0, 1.0)]
hj = ebn_b200.Job(ebn_b200._lib.LS_POLYNOMIAL, prog, ebn_b200.make_inputs(rows), n=50_001, seed=3)
print(ebn_b200.hist_batch(hj, np.linspace(-4, 5, 100))[0].sum(axis=1))
print(ebn_b200.sample_matrix(hj, 1, 0, 10, ncols=3)[:2])
ebn_b200.shutdown()
print("done")
