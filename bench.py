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

PROFILE = os.path.join(ROOT, "profiles", "w1_kernel_profile.json")


def load_profile(build_id: str) -> dict:
    """The frozen, machine-readable profile of the headline kernel (tools/freeze_profile.py): the
    algorithmic fp64 work per sample that `roofline.achieved` is computed from, and the ncu figures the
    line quotes. It names the build (`ebn_build_id()`: hash of the kernel sources and tuning flags) it
    was taken from; quoting instruction counts of a different kernel would silently mis-state
    `roofline.frac`, so a mismatch stops the bench."""
    with open(PROFILE) as f:
        prof = json.load(f)
    if prof["build_id"] != build_id:
        raise SystemExit(f"{PROFILE} was frozen for build {prof['build_id']} but libebncuda.so is build {build_id}: "
                         "run `python tools/freeze_profile.py` (and re-capture ncu) after changing the kernels")
    return prof


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
    emit(line)


_RESULT_FD = None


def keep_stdout_for_the_result():
    """The contract is ONE JSON line on stdout. Libraries write there too (with NCCL_DEBUG set in the box's
    environment NCCL prints its version banner on fd 1): keep a private copy of stdout for the result and point
    fd 1 at stderr for everything else."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _RESULT_FD if _RESULT_FD is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    keep_stdout_for_the_result()
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
    ap.add_argument("--single-process", action="store_true",
                    help="one process drives all --gpus devices (ebn_init(devices) + ncclCommInitAll inside the "
                         "library): the mode a Julia caller uses. Not launched under torchrun.")
    ap.add_argument("--strong-n", type=float, default=1e9,
                    help="samples per scenario of the strong-scaling W5 call (1024 scenarios; total fixed for any N)")
    ap.add_argument("--sustained-s", type=float, default=5.0, help="target length of the sustained call (0: skip)")
    ap.add_argument("--no-extras", action="store_true", help="headline loop only (no strong / sustained / variants)")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return

    import hashlib
    import torch
    import ebn_b200
    from ebn_b200 import workloads as W

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    single = args.single_process and args.gpus > 1
    if single and world > 1:
        raise SystemExit("--single-process is not a torchrun mode")
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libebncuda has no CPU fallback")
    prof = load_profile(ebn_b200._lib.build_id())
    f_alg = float(prof["sass"]["fp64_instructions_per_sample"])
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if single:
        if torch.cuda.device_count() < args.gpus:
            raise SystemExit(f"--single-process --gpus {args.gpus}: only {torch.cuda.device_count()} device(s) visible")
        ebn_b200.init(list(range(args.gpus)))
    else:
        ebn_b200.init([local_rank])
    n_gpus = args.gpus if single else world
    if world > 1:
        # torch.distributed is plumbing only: it carries the 128-byte NCCL id; the count
        # all-reduce itself runs inside libebncuda on its own communicator.
        uid = [ebn_b200._lib.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ebn_b200._lib.comm_init_rank(uid[0], rank, world)

    def shard(job):
        job.shard_rank, job.shard_world = rank, world
        return job

    if args.workload == "w5":
        n_per_scenario = int(args.n_per_scenario)      # strong scaling: total work fixed
        job = W.w5_vehicle_grid(n=n_per_scenario, strict=args.strict)
        wl_name = f"W5 vehicle_suspension grid (BASELINE configs[4]): S=1024 scenarios x {n_per_scenario:.3g} samples"
    else:
        n_per_scenario = int(args.n_per_scenario) * n_gpus  # weak scaling: per-GPU work fixed
        job = W.w1_vehicle(n=n_per_scenario, strict=args.strict)
        wl_name = ("W1 vehicle_suspension (BASELINE configs[1]): S=20 scenarios x "
                   f"{n_per_scenario:.3g} samples, seed 0x5EED0001")
    shard(job)
    total_samples = job.S * n_per_scenario

    peak_dfma = ebn_b200.peak_fp64(0.5) if rank == 0 else 0.0

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step(j=None):
        cnt, pf, var = ebn_b200.pf_batch(j or job)
        st = ebn_b200.last_stats()
        return cnt, pf, st

    def max_over_ranks(*vals):
        if dist is None:
            return list(vals)
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def timed(j, reps):
        """`reps` calls of job j outside the headline loop: (device ms, wall s), max over ranks, L2 flushed."""
        ms = wall = 0.0
        out = None
        for _ in range(reps):
            flush.fill_(1)
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            cnt, pf, st = step(j)
            wall += time.perf_counter() - t0
            ms += st["kernel_ms"]
            out = (cnt, pf)
        ms, wall = max_over_ranks(ms, wall)
        return ms, wall, out

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

    kernel_ms, wall = max_over_ranks(kernel_ms, wall)
    if dist is not None:
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt.item())

    # ---- beside the headline (every rank takes part: the calls all-reduce inside the library) ----------
    extras = {}
    if not args.no_extras and args.workload == "w1" and not args.strict:
        # (1) strict variant: the Julia-order limit state (one IEEE operation per Julia operation, six
        # divisions and a square root per sample) — the mode the bit-exact parity bar is defined on
        sjob = shard(W.w1_vehicle(n=n_per_scenario, strict=True))
        step(sjob)
        ms, _, (scnt, _) = timed(sjob, 3)
        extras["strict_variant"] = {
            "value": total_samples * 3 / (ms * 1e-3), "unit": UNIT, "ms": ms / 3,
            "max_abs_count_diff_vs_fast": int(np.max(np.abs(scnt - cnt))),
            "note": "strict_fp: g evaluated operation by operation in the reference's order (bit-exact vs the oracle "
                    "on replayed matrices); the headline runs the sign-only fast predicate, which may differ on ties"}
        # (2) BASELINE configs[4] as STRONG scaling: 1024 scenarios x strong_n samples in total, whatever N is
        w5 = shard(W.w5_vehicle_grid(n=int(args.strong_n)))
        step(shard(W.w5_vehicle_grid(n=1_000_000)))
        ms, wl, _ = timed(w5, 1)
        extras["strong"] = {
            "workload": f"W5 vehicle_suspension grid (BASELINE configs[4] at {args.strong_n / 1e10:.3g} of its size): 1024 scenarios x "
                        f"{int(args.strong_n):.3g} samples = {1024 * int(args.strong_n):.4g} limit-state evaluations in "
                        "total, sharded over the ranks",
            "scaling": "strong", "n_gpus": n_gpus, "value": 1024 * int(args.strong_n) / (ms * 1e-3), "unit": UNIT,
            "ms": ms, "e2e_value": 1024 * int(args.strong_n) / wl, "steps": 1}
        # (3) cross-N identity: the int64 count vector of a fixed job must not depend on the number of GPUs
        fixed = shard(W.w5_vehicle_grid(n=4_000_000, seed=0x5EED0055))
        fcnt, _, _ = step(fixed)
        extras["strong"]["counts_sha256"] = hashlib.sha256(np.ascontiguousarray(fcnt, dtype="<i8").tobytes()).hexdigest()
        extras["strong"]["counts_job"] = "W5 grid, 1024 scenarios x 4e6 samples, seed 0x5EED0055: identical digest at every N"
        # (4) one sustained call with the NVML clock record
        if args.sustained_s > 0:
            per_gpu_rate = total_samples * args.steps / (kernel_ms * 1e-3) / n_gpus
            n_long = int(per_gpu_rate * args.sustained_s * 1.05 / 20) * n_gpus
            ljob = shard(W.w1_vehicle(n=n_long))
            s2 = ClockSampler(local_rank)
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            s2.start()
            t0 = time.perf_counter()
            _, _, st = step(ljob)
            wl = time.perf_counter() - t0
            c2 = s2.stop()
            ms, wl = max_over_ranks(st["kernel_ms"], wl)
            extras["sustained"] = {"workload": f"W1, 20 scenarios x {n_long:.4g} samples in ONE call",
                                   "seconds": ms * 1e-3, "value": 20 * n_long / (ms * 1e-3), "unit": UNIT,
                                   "e2e_value": 20 * n_long / wl, "clocks": c2}

    if rank == 0:
        value = total_samples * args.steps / (kernel_ms * 1e-3)
        e2e = total_samples * args.steps / wall
        per_gpu = value / n_gpus
        achieved_tf = per_gpu * f_alg * 2.0 / 1e12
        peak_tf = peak_dfma * 2.0 / 1e12
        ncu = prof.get("ncu") or {}
        ncu_line = {k: ncu.get(k) for k in ("fp64_pipe_active_pct", "alu_pipe_active_pct", "fmaheavy_pipe_active_pct",
                                            "issue_slots_active_pct", "tensor_pipe_active_pct", "shared_wavefronts_pct_of_peak",
                                            "registers_per_thread", "capture")}
        if ncu.get("fp64_warp_instructions") and ncu.get("samples_in_capture"):
            # SURVEY 8d cross-check: executed fp64 thread instructions per sample in the capture vs the SASS figure
            ncu_line["fp64_thread_instructions_per_sample_executed"] = round(
                ncu["fp64_warp_instructions"] * 32.0 / ncu["samples_in_capture"], 3)
        ncu_line["stale"] = ncu.get("build_id") != prof["build_id"]
        ncu_line["source"] = "profiles/w1_kernel_profile.json (tools/freeze_profile.py --ncu; build id checked above)"
        mode = (f"one process, {n_gpus} devices (ncclCommInitAll inside the library)" if single
                else f"(scenario,chunk) items round-robin over {world} rank(s); NCCL int64 all-reduce")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": kernel_ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if args.workload == "w5" else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": wl_name,
                "n_per_scenario": n_per_scenario, "scenarios": job.S, "strict_fp": bool(args.strict),
                "parallelism": mode,
                "l2": "256 MiB flush written between timed iterations; inputs are O(KB) descriptors",
                "rng": "Philox4x32-10, 32-bit uniforms, Box-Muller from 64 bits per normal pair (42-bit radius, 22-bit full-circle angle); "
                       "coarser than Julia's 52-bit rand, see DESIGN.md section 4",
            },
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * wall / args.steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {
                "bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf if peak_tf else None,
                "traffic": (ncu.get("dram_bytes_read", 0) or 0) + (ncu.get("dram_bytes_written", 0) or 0) if ncu else None,
                "traffic_unit": "bytes per launch (ncu, n=1e7 launch; algorithmic traffic is O(S) descriptors)",
                "fp64_slots_per_sample": f_alg,
                "instructions_per_sample_pair": prof["sass"]["instructions_per_sample_pair"],
                "fp64_instructions_per_sample_pair": prof["sass"]["fp64_instructions_per_sample_pair"],
                "build_id": prof["build_id"],
                "ncu": ncu_line,
                "peak_source": "measured in this run by ebn_peak_fp64 (independent DFMA chains, burst); "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "note": "compute-bound path: algorithmic DRAM traffic is ~0 B/sample, no tensor cores; the bound on this "
                        "kernel is the sub-partition's issue / register-file traffic (DESIGN.md section 5)",
            },
            "pf_first_scenarios": [float(x) for x in pf[:5]],
        }
        line.update(extras)
        if n_gpus == 1 and args.workload == "w1" and not args.strict and not args.no_extras:
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
        if not args.no_cpu_baseline and n_gpus == 1:
            cores = os.cpu_count() or 1
            v, dt, _ = cpu_reference_run(args.cpu_n, cores)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"W1, 20 scenarios x {args.cpu_n} samples ({dt:.1f} s) — a bounded sample of the GPU arm's "
                          f"20 x {n_per_scenario:.3g} (throughput metric: samples/s does not depend on n); oracle structural "
                          "C port of the reference path (Julia/UQ.jl cannot run here)"}
        emit(line)

    if dist is not None:
        ebn_b200._lib.comm_destroy()
        dist.destroy_process_group()
    ebn_b200.shutdown()


if __name__ == "__main__":
    main()
