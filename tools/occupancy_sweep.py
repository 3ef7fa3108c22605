import os, subprocess, sys
ROOT="/root/repo"
code = """
import sys; sys.path.insert(0, %r)
import ebn_b200
from ebn_b200 import workloads as W
ebn_b200.init()
job = W.w1_vehicle(n=10**8)
best = 1e30
for _ in range(5):
    cnt, pf, _ = ebn_b200.pf_batch(job)
    best = min(best, ebn_b200.last_stats()['kernel_ms'])
print(f"{job.S * job.n / best / 1e-3:.4e} samples/s  {best:.3f} ms")
""" % ROOT
for b in ["1","2","3","4"]:
    env=dict(os.environ, EBN_BLOCKS_PER_SM=b)
    out=subprocess.run([sys.executable,"-c",code],env=env,capture_output=True,text=True)
    print("blocks/SM",b,out.stdout.strip() or out.stderr[-300:],flush=True)
