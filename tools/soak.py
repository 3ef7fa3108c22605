"""Lifecycle soak: repeated ebn_init / ebn_shutdown with compiled-in and run-time compiled kernels in
between; device memory in use must not grow. GPU box only."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402  (only for mem_get_info)
import ebn_b200  # noqa: E402
from ebn_b200 import expr, workloads as W  # noqa: E402

prog = expr.compile_program(["A", "B"], [("C", "df.A .+ df.B")], "2 .- df.C")
rows = [[(0, a, 0, 0, 0, -np.inf, np.inf), (1, 0.0, 1.0, 0, 0, -np.inf, np.inf)] for a in (1.0, 2.0)]
ej = ebn_b200.Job(ebn_b200._lib.LS_EXPRESSION, prog.consts, ebn_b200.make_inputs(rows), n=100_000, seed=1)
vj = W.w1_vehicle(n=100_000)
used = []
ref = None
for it in range(150):
    ebn_b200.init()
    a = ebn_b200.pf_batch(vj)[0]
    b = ebn_b200.pf_batch(ej)[0]
    h = ebn_b200.hist_batch(ej, np.linspace(-3, 5, 33))[0]
    x = ebn_b200.sample_matrix(ej, 1, 0, 1000)
    if ref is None:
        ref = (a.copy(), b.copy(), h.copy(), x.copy())
    assert np.array_equal(a, ref[0]) and np.array_equal(b, ref[1]) and np.array_equal(h, ref[2]) and np.array_equal(x, ref[3])
    ebn_b200.shutdown()
    free, total = torch.cuda.mem_get_info()
    used.append((total - free) / 2**20)
print(f"150 init/shutdown cycles: device memory in use after shutdown {used[5]:.0f} MiB (cycle 5) -> {used[-1]:.0f} MiB (cycle 150)")
assert used[-1] - used[5] < 64, used[::10]
print("ok")
