"""The second half of the BASELINE metric: pf error of the GPU path against the CPU oracle in standard
errors, z = (pf_gpu - pf_ref) / sqrt(var_gpu + var_ref) per scenario, for W1 (20
scenarios), W2 (Straub) and W5 (1024 scenarios). The oracle draws with its own generator
(xoshiro256++), so this is an independent-RNG comparison. GPU box only (the oracle runs on its host
cores).    python tools/pf_error_report.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ebn_b200  # noqa: E402
from ebn_b200 import workloads as W  # noqa: E402
from oracle import cpu as ocpu  # noqa: E402

ebn_b200.init()
threads = os.cpu_count() or 8


def report(name, job, ls, n_ref):
    t = time.time()
    cnt, pf, var = ebn_b200.pf_batch(job)
    t_gpu = time.time() - t
    t = time.time()
    ref = ocpu.pf_batch(ls, job.consts, job.inputs.view(ocpu.INPUT_DTYPE), n_ref, seed=2026, threads=threads, block=1 << 20)
    t_cpu = time.time() - t
    pf_ref = ref / n_ref
    var_ref = (pf_ref - pf_ref**2) / n_ref
    z = (pf - pf_ref) / np.sqrt(var + var_ref + 1e-300)
    S = job.S
    print(f"{name}: S={S} n_gpu={job.n:.0e} ({t_gpu:.2f} s wall) n_ref={n_ref:.0e} ({t_cpu:.1f} s on {threads} host threads)")
    print(f"  pf range gpu [{pf.min():.3e}, {pf.max():.3e}]  ref [{pf_ref.min():.3e}, {pf_ref.max():.3e}]")
    print(f"  |z| max {np.abs(z).max():.2f}  mean {np.abs(z).mean():.2f}  rms {np.sqrt((z**2).mean()):.2f}  "
          f"#(|z|>2) {int((np.abs(z) > 2).sum())} (expected {0.0455 * S:.1f})  #(|z|>3) {int((np.abs(z) > 3).sum())} "
          f"(expected {0.0027 * S:.2f})", flush=True)
    return z


report("W1 vehicle (BASELINE config 2)", W.w1_vehicle(n=10**8), ocpu.LS_VEHICLE, 10**7)
report("W2 Straub (BASELINE config 1)", W.w2_straub(n=10**8), ocpu.LS_STRAUB, 2 * 10**7)
report("W5 vehicle grid (BASELINE config 5 shape)", W.w5_vehicle_grid(n=10**7), ocpu.LS_VEHICLE, 2 * 10**6)
