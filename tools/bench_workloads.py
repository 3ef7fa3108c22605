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
run("W1 vehicle, normals drawn in fp32 (EBN_JOB_FP32_SAMPLING)", W.w1_vehicle(n=10**8, fp32_sampling=True))
run("W1 vehicle (strict: Julia-order g, 6 div + sqrt)", W.w1_vehicle(n=10**8, strict=True))
run("W2 Straub frame S=1", W.w2_straub(n=10**9))
w2f = W.w2_straub(n=10**9)
w2f.fp32_sampling = True
run("W2 Straub frame, normals drawn in fp32", w2f)
w3 = W.w3_straub_copula(n=2 * 10**7)
w3.jit_plan = False
run("W3 Straub-style copula, 64 scenarios (looping copula plan)", w3)
w3.jit_plan = True
run("W3 copula compiled at run time (unrolled plan, coefficient table rewritten as a formula)", w3)
from ebn_b200 import expr  # noqa: E402
import dataclasses  # noqa: E402
run("W1 vehicle, quasi-Monte Carlo (EBN_JOB_QMC_SOBOL: Owen-scrambled Sobol, inverse-CDF marginals)", dataclasses.replace(W.w1_vehicle(n=10**8), qmc=True))
run("W5 vehicle grid, 1024 scenarios", W.w5_vehicle_grid(n=4 * 10**6))
run("W4 vehicle, V fixed: 4 states x 25 outer-loop values", W.w4_vehicle_fixed_v(np.linspace(7, 12, 25), n=2 * 10**7))
j = W.w1_vehicle(n=10**8)
j.inputs[:, 3]["lo"] = 400.0   # truncated normal: inverse CDF on the truncation window
run("W1 with C truncated at 400 (compiled-in plan P,P,U,I,N,N)", j)
j.inputs[:, 4]["hi"] = 1500.0  # two truncated normals: compiled in too (all eight combinations of C, Ck, K are)
run("W1 with C and Ck truncated (compiled-in plan P,P,U,I,I,N)", j)
j.inputs[:, 4]["hi"] = np.inf
j.inputs[:, 5]["kind"], j.inputs[:, 5]["p"] = ebn_b200._lib.IN_LOGNORMAL, (4.0, 0.18, 0, 0)  # K log-normal: no compiled-in plan
j.jit_plan = False
run("W1 with C truncated and K log-normal (runtime-dispatched plan)", j)
j.jit_plan = True
run("W1 with C truncated and K log-normal (plan compiled at run time)", j)
j = W.w1_vehicle(n=10**8)
for col, (lo, hi) in ((3, (400.0, 460.0)), (4, (1440.0, 1500.0)), (5, (30.0, 80.0))):
    j.inputs[:, col]["lo"], j.inputs[:, col]["hi"] = lo, hi
run("W1 with C, Ck and K truncated (compiled-in plan P,P,U,I,I,I)", j)
k = W.vehicle_consts()
VEHICLE_TEXT = [
    ("g1", "1 .- (pi .* m .* df.V .* df.A) ./ (df.b0 .* df.K .* g .^ 2) .* ((df.Ck ./ (m .+ M) .- (df.C ./ M)) .^ 2"
           " .+ df.C .^ 2 ./ (m .* M) .+ df.Ck .* df.K .^ 2 ./ (m .* M .^ 2))"),
    ("g2", "4000 .* df.C .* Mg15 .- 8.6394"),
    ("g3", "2 .* sqrt(M .* g .* (df.K .^ 2 .* df.Ck ./ (df.C .* (m .+ M)) .+ df.C)) .- 1"),
    ("g4", "df.Ck .- gMm"),
    ("y", "minimum([g1[1], g2, g3, g4[1]])"),
]
vp = expr.compile_program(["A", "b0", "V", "C", "Ck", "K"], VEHICLE_TEXT, "y",
                          params={"M": k[0], "m": k[1], "g": k[2], "Mg15": k[3], "gMm": k[4]})
w1 = W.w1_vehicle(n=10**8)
run("W1 vehicle, closures as text (run-time compiled, fast)",
    ebn_b200.Job(ebn_b200._lib.LS_EXPRESSION, vp.consts, w1.inputs, n=10**8, seed=w1.seed))
run("W1 vehicle, closures as text (run-time compiled, strict)",
    ebn_b200.Job(ebn_b200._lib.LS_EXPRESSION, vp.consts, w1.inputs, n=10**8, seed=w1.seed, strict=True))
lam, zeta = W.lognormal_params(150.0, 30.0)
cap = "exp(rand(Normal(lam + rho * zeta * df.Ur, sqrt((1 - rho^2) * zeta^2))))"
sp = expr.compile_program(["Ur", "V", "H"], [(f"r{i}", cap) for i in range(1, 6)] + [("G", "minimum([r1 + r2 + r4 + r5 - 5 * df.H, r2 + 2 * r3 + r4 - 5 * df.V, r1 + 2 * r3 + 2 * r4 + r5 - 5 * df.H - 5 * df.V])")], "df.G", params={"lam": lam, "zeta": zeta, "rho": W.STRAUB_RHO})
w2 = W.w2_straub(n=10**9)
run("W2 Straub, closures as text (run-time compiled, fast)",
    ebn_b200.Job(ebn_b200._lib.LS_EXPRESSION, sp.consts, w2.inputs, n=10**9, seed=w2.seed))
prog = np.array([1, 2, 1.0, 1, 0, 1.0, 1, 1])
rows = [[(0, a, 0, 0, 0, -np.inf, np.inf), (1, 0.0, 1.0, 0, 0, -np.inf, np.inf)] for a in np.linspace(0, 3, 20)]
hj = ebn_b200.Job(ebn_b200._lib.LS_POLYNOMIAL, prog, ebn_b200.make_inputs(rows), n=10**8, seed=3, jit_plan=False)
run("histogram path A+B, 2048 edges, 20 scenarios", hj, hist_edges=np.linspace(-8, 11, 1025))
run("count path A+B (polynomial functor, generic plan)", hj)
hj.jit_plan = True   # what the library does on its own above 5e10 samples: the program rewritten as a formula
run("histogram path A+B, polynomial program compiled at run time", hj, hist_edges=np.linspace(-8, 11, 1025))
run("count path A+B, polynomial program compiled at run time", hj)
