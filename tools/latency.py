"""Wall-clock latency of one ebn_pf_batch call through the ctypes binding for small jobs (the
regime of the imprecise outer loop: many small batches). GPU box only."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ebn_b200  # noqa: E402
from ebn_b200 import workloads as W  # noqa: E402

ebn_b200.init()
for n in (1, 1000, 100_000, 1_000_000):
    job = W.w1_vehicle(n=n)
    for _ in range(20):
        ebn_b200.pf_batch(job)
    t = time.perf_counter()
    reps = 200
    for _ in range(reps):
        ebn_b200.pf_batch(job)
    dt = (time.perf_counter() - t) / reps
    print(f"W1 n={n:>8}: {dt * 1e6:8.1f} us per call (kernel {ebn_b200.last_stats()['kernel_ms'] * 1e3:7.1f} us)", flush=True)
