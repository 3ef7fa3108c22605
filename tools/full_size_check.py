"""Size-independent parity properties at BASELINE config 5's FULL size (1024 scenarios x n samples,
default n = 1e10: 1.024e13 limit-state evaluations) on one GPU:
  (1) the eight shards of an 8-rank job, run one after the other with EBN_JOB_PARTIAL, add up to exactly
      the counts of the unsharded job (what the NCCL all-reduce computes on 8 GPUs);
  (2) pf of every scenario lies within 5 standard errors of a short independent run (other seed);
a SHA-256 of the count vector is printed so runs on other machines / GPU counts can be compared.
GPU box only.    python tools/full_size_check.py [n per scenario]"""
import hashlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ebn_b200  # noqa: E402
from ebn_b200 import workloads as W  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**10
ebn_b200.init()
job = W.w5_vehicle_grid(n=n)
t = time.time()
whole, pf, var = ebn_b200.pf_batch(job)
print(f"unsharded: S={job.S} n={n:.1e} -> {job.S * n:.3e} samples in {time.time() - t:.1f} s "
      f"(kernel {ebn_b200.last_stats()['kernel_ms'] / 1e3:.1f} s); sha256(counts) = {hashlib.sha256(whole.tobytes()).hexdigest()[:16]}",
      flush=True)
acc = np.zeros_like(whole)
t = time.time()
for r in range(8):
    shard = W.w5_vehicle_grid(n=n)
    shard.shard_rank, shard.shard_world, shard.partial = r, 8, True
    acc += ebn_b200.pf_batch(shard)[0]
print(f"8 shards one after the other: {time.time() - t:.1f} s; sum == unsharded: {bool(np.array_equal(acc, whole))}", flush=True)
short = W.w5_vehicle_grid(n=10**7, seed=12345)
_, pf_s, var_s = ebn_b200.pf_batch(short)
z = (pf - pf_s) / np.sqrt(var + var_s + 1e-300)
print(f"against an independent 1e7-sample run: max |z| = {np.abs(z).max():.2f}, rms = {np.sqrt((z * z).mean()):.2f}, "
      f"pf range [{pf.min():.3e}, {pf.max():.3e}]")
assert np.array_equal(acc, whole) and np.abs(z).max() < 5.0
print("ok")
