#!/usr/bin/env python
"""bench.py — limit-state Monte Carlo samples/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): W1 = BASELINE.json configs[1], the vehicle-suspension functional
node (demo/vehicle_suspension.jl) — 20 discrete-parent scenarios x 10^8 samples each on one
GPU. At N > 1 (torchrun, one process per GPU) samples per scenario grow with N (weak scaling:
2*10^9 samples per GPU per step); the (scenario, chunk) work items are dealt round-robin to
the ranks and the int64 count vector is all-reduced with NCCL inside the C library.

A step = one `ebn_pf_batch` call through the C ABI with HOST buffers in and out: descriptor
upload, the fused sample→limit-state→count kernel, (all-reduce), count download.
  value  = samples / device time of the kernels (CUDA events on the launching stream)
  e2e    = samples / wall time of the same calls (includes the copies and the all-reduce)
The CPU arm (`--impl reference`, and `cpu_baseline` in the default run) times the oracle's
structural restatement of the reference path on the host cores: the reference itself is
Julia + UncertaintyQuantification.jl and cannot run in this image.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# FP64-pipe instructions per sample of the W1 kernel: (DFMA 68.1 + DMUL 38.0 + DADD 13.1) warp
# instructions per sample PAIR executed in the hot loop, from the ncu source page of
# profiles/r01_s_* (= sm__inst_executed_pipe_fp64), i.e. 59.6 per sample. See DESIGN.md
# "Roofline". The SURVEY's ceiling for this figure is 339; the first version executed 151.5.
# The figure is re-frozen whenever the kernel changes: removing fp64 work raises samples/s and
# LOWERS the fraction, which is why DESIGN.md tabulates both for every version.
F_ALG_W1 = 59.6
# dram__bytes_read.sum + dram__bytes_write.sum of one mc_kernel launch (ncu --set full,
# profiles/r01_s_ncu_full_w1_n1e7_summary.txt): 52.0 KB read, 0 B written — the plan and tables.
DRAM_TRAFFIC_BYTES_PER_LAUNCH = 51968
METRIC = "limit_state_mc_samples_per_sec"
UNIT = "samples/s"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # pragma: no cover - NVML missing
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_reference_run(n_per_scenario: int, threads: int):
    """The oracle's structural restatement of the reference CPU path on W1 (bounded sample)."""
    from oracle import cpu as ocpu
    import ebn_b200
    from ebn_b200 import workloads as W
    job = W.w1_vehicle(n=n_per_scenario)
    inp = job.inputs.view(ocpu.INPUT_DTYPE)
    t0 = time.perf_counter()
    # columns are materialised like UQ.jl's `sample` does, in blocks of at most 5e6 rows so that all
    # host threads together stay within a few GB
    cnt = ocpu.pf_batch(ocpu.LS_VEHICLE, job.consts, inp, n_per_scenario, seed=12345, threads=threads,
                        block=min(n_per_scenario, 5_000_000))
    dt = time.perf_counter() - t0
    return job.S * n_per_scenario / dt, dt, cnt


def run_reference(args):
    """`--impl reference`: the reference's CPU path (oracle port) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = args.cpu_n
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_reference_run(max(n // 10, 1000), cores)
    times = []
    for _ in range(args.steps):
        _, dt, _ = cpu_reference_run(n, cores)
        times.append(dt)
    total = 20 * n * args.steps
    value = total / sum(times)
    sample = f"W1 vehicle suspension, 20 scenarios x {n} samples per step (full workload: 20 x 1e8)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": "W1 vehicle_suspension (BASELINE configs[1]): S=20 scenarios, bounded CPU "
                                                    f"sample of {n} samples per scenario per step",
                                        "n_per_scenario": n, "scenarios": 20},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is Julia+UQ.jl (not runnable here); this is the oracle's structural C port",
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-per-scenario", type=float, default=1e8, help="per GPU (weak scaling)")
    ap.add_argument("--cpu-n", type=int, default=50_000_000,
                    help="CPU arm: samples per scenario per step (20 scenarios; ~10 s on 16 cores)")
    ap.add_argument("--strict", action="store_true", help="Julia-order limit state (no FMA contraction)")
    ap.add_argument("--workload", default="w1", choices=["w1", "w5"],
                    help="w1: BASELINE configs[1] (default, weak scaling); w5: configs[4], 1024 scenarios x "
                         "--n-per-scenario samples in total, sharded over the ranks (strong scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import ebn_b200
    from ebn_b200 import workloads as W

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libebncuda has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ebn_b200.init([local_rank])
    if world > 1:
        # torch.distributed is plumbing only: it carries the 128-byte NCCL id; the count
        # all-reduce itself runs inside libebncuda on its own communicator.
        uid = [ebn_b200._lib.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ebn_b200._lib.comm_init_rank(uid[0], rank, world)

    if args.workload == "w5":
        n_per_scenario = int(args.n_per_scenario)      # strong scaling: total work fixed
        job = W.w5_vehicle_grid(n=n_per_scenario, strict=args.strict)
        wl_name = f"W5 vehicle_suspension grid (BASELINE configs[4]): S=1024 scenarios x {n_per_scenario:.3g} samples"
    else:
        n_per_scenario = int(args.n_per_scenario) * world  # weak scaling: per-GPU work fixed
        job = W.w1_vehicle(n=n_per_scenario, strict=args.strict)
        wl_name = ("W1 vehicle_suspension (BASELINE configs[1]): S=20 scenarios x "
                   f"{n_per_scenario:.3g} samples, seed 0x5EED0001")
    job.shard_rank, job.shard_world = rank, world
    total_samples = job.S * n_per_scenario

    peak_dfma = ebn_b200.peak_fp64(0.5) if rank == 0 else 0.0

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step():
        cnt, pf, var = ebn_b200.pf_batch(job)
        st = ebn_b200.last_stats()
        return cnt, pf, st

    for _ in range(max(args.warmup, 3)):
        step()
        flush.fill_(1)
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    kernel_ms = 0.0
    wall = 0.0
    launches = 0
    h2d = d2h = 0.0
    cnt = pf = None
    for _ in range(args.steps):
        t0 = time.perf_counter()
        cnt, pf, st = step()
        wall += time.perf_counter() - t0
        kernel_ms += st["kernel_ms"]
        launches += st["launches"]
        h2d, d2h = st["h2d_bytes"], st["d2h_bytes"]
        flush.fill_(1)  # L2 flush between timed iterations (not timed)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    clocks = sampler.stop()

    if dist is not None:
        t = torch.tensor([kernel_ms, wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kernel_ms, wall = t.tolist()
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt.item())

    if rank == 0:
        value = total_samples * args.steps / (kernel_ms * 1e-3)
        e2e = total_samples * args.steps / wall
        per_gpu = value / world
        achieved_tf = per_gpu * F_ALG_W1 * 2.0 / 1e12
        peak_tf = peak_dfma * 2.0 / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": kernel_ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if args.workload == "w5" else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": wl_name,
                "n_per_scenario": n_per_scenario, "scenarios": job.S, "strict_fp": bool(args.strict),
                "parallelism": f"(scenario,chunk) items round-robin over {world} rank(s); NCCL int64 all-reduce",
                "l2": "256 MiB flush written between timed iterations; inputs are O(KB) descriptors",
            },
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * wall / args.steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {
                "bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": DRAM_TRAFFIC_BYTES_PER_LAUNCH,
                "traffic_unit": "bytes per launch (ncu, n=1e7 launch; algorithmic traffic is O(S) descriptors)",
                "fp64_slots_per_sample": F_ALG_W1,
                "ncu": {"fp64_pipe_active_pct": 48.8, "issue_slots_active_pct": 64.2, "tensor_pipe_active_pct": 0.0,
                        "instructions_per_sample_pair": 314, "fp64_instructions_per_sample_pair": 119,
                        "source": "profiles/r01_s_ncu_full_w1_n1e7_summary.txt (static: taken from the committed capture)"},
                "peak_source": "measured in this run by ebn_peak_fp64 (independent DFMA chains, burst); "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "note": "compute-bound path: algorithmic DRAM traffic is ~0 B/sample, no tensor cores",
            },
            "pf_first_scenarios": [float(x) for x in pf[:5]],
        }
        if world == 1 and args.workload == "w1" and not args.strict:
            # reported separately, outside the timed region (north star: "FP32 where tolerated"): the same
            # workload with the untruncated normals drawn through the fp32 Box-Muller (same random words)
            job.fp32_sampling = True
            step()
            ms32 = 0.0
            for _ in range(3):
                flush.fill_(1)
                _, pf32, st32 = step()
                ms32 += st32["kernel_ms"]
            job.fp32_sampling = False
            line["fp32_sampling_variant"] = {
                "value": total_samples * 3 / (ms32 * 1e-3), "unit": UNIT,
                "max_abs_pf_diff_vs_fp64": float(np.max(np.abs(pf32 - pf))),
                "note": "EBN_JOB_FP32_SAMPLING: sampling transform in fp32, limit state in fp64; not the headline"}
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            v, dt, _ = cpu_reference_run(args.cpu_n, cores)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"W1, 20 scenarios x {args.cpu_n} samples ({dt:.1f} s); oracle structural C port "
                          "of the reference path (Julia/UQ.jl cannot run here)"}
        print(json.dumps(line), flush=True)

    if dist is not None:
        ebn_b200._lib.comm_destroy()
        dist.destroy_process_group()
    ebn_b200.shutdown()


if __name__ == "__main__":
    main()
