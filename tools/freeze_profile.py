"""Freezes the machine-readable profile of the headline kernel: profiles/w1_kernel_profile.json.

bench.py takes the algorithmic fp64 work per sample (roofline.achieved = samples/s x that figure) and
the ncu figures it quotes from this file, and REFUSES to run when the file belongs to another build
of the library (`ebn_build_id()`, a hash of the kernel sources and tuning flags). So a kernel change
must be followed by

    python tools/freeze_profile.py                      # CPU box: SASS audit of the built objects
    python tools/freeze_profile.py --ncu X.ncu-rep      # additionally ingest an `ncu --set full` capture
                                                        # of `tools/run_workload.py w1 1e7` on this build

The SASS half needs only cuobjdump; the ncu half is kept from the previous file when no capture is
given, with the build id it was measured on (bench.py reports it as stale when the ids differ)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sass_audit as A  # noqa: E402

OUT = os.path.join(ROOT, "profiles", "w1_kernel_profile.json")
OBJ = os.path.join(ROOT, "enhancedbayesiannetworks.jl_b200", "csrc", "ebn_ls_vehicle.o")
KERNEL = r"mc_kernel.*VehicleSuspension.*FastOps.*StaticPlanIJLi0ELi0ELi1ELi2ELi2ELi2EEEELi0"
PAIRS_PER_TRIP = 1.0


def sass_half():
    import collections
    import re
    ks = {k: v for k, v in A.kernels(OBJ).items() if re.search(KERNEL, k)}
    assert len(ks) == 1, list(ks)
    name, ins = next(iter(ks.items()))
    body = A.hot_loop(ins)
    mix, pipe, dsrc = collections.Counter(), collections.Counter(), collections.Counter()
    for _, text in body:
        op, t = A.opcode(text)
        mix[op] += 1
        pipe[A.PIPE.get(op, "other")] += 1
        if A.PIPE.get(op) == "fp64":
            regs = set()
            for o in [x.strip() for x in t.split(None, 1)[1].split(",")][1:]:
                m = re.match(r"[-|~]*R(\d+)", o)
                if m and ".reuse" not in o:
                    regs.add(int(m.group(1)))
            dsrc[str(len(regs))] += 1
    n = len(body) / PAIRS_PER_TRIP
    fp64 = pipe["fp64"] / PAIRS_PER_TRIP
    with open(os.path.join(ROOT, "profiles", "w1_hot_loop.sass"), "w") as f:
        f.write(f"// {name}\n// interior count loop: {len(body)} instructions per trip, {PAIRS_PER_TRIP:g} sample pair(s) per trip\n")
        for a, t in body:
            f.write(f"/*{a:04x}*/ {t} ;\n")
    return {"kernel": name, "instructions_per_sample_pair": n, "fp64_instructions_per_sample_pair": fp64,
            "fp64_instructions_per_sample": fp64 / 2.0, "per_pipe": dict(pipe), "opcodes": dict(mix),
            "fp64_by_distinct_register_pair_sources": dict(dsrc), "listing": "profiles/w1_hot_loop.sass"}


SAMPLES_IN_CAPTURE = 2e8     # tools/run_workload.py w1 2e8: 20 scenarios x 1e7 samples in the captured launch


def fp64_executed(rep):
    """Executed warp instructions on the fp64 pipe, summed over the SASS lines of the source page
    (`--set full` carries the pipe's utilisation but not its instruction count)."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    src, ex = rows[head].index("Source"), rows[head].index("Instructions Executed")
    total = 0
    for r in rows[head + 1:]:
        if len(r) > ex and A.PIPE.get(A.opcode(r[src].strip())[0]) == "fp64":
            total += int(r[ex])
    return float(total)


def ncu_half(rep, build_id):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    best = max(rows[2:], key=lambda r: float(r[rows[0].index("gpu__time_duration.sum")] or 0))
    m = dict(zip(rows[0], best))
    g = lambda k: float(m[k])  # noqa: E731
    stalls = {k.split("issue_stalled_")[1].split("_per_issue")[0]: round(float(v), 4) for k, v in m.items()
              if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")}
    return {
        "build_id": build_id, "capture": os.path.basename(rep), "kernel": m.get("Kernel Name"),
        "fp64_pipe_active_pct": g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "alu_pipe_active_pct": g("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
        "fmaheavy_pipe_active_pct": g("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "issue_slots_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "tensor_pipe_active_pct": g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "shared_wavefronts_pct_of_peak": g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        "dram_bytes_read": g("dram__bytes_read.sum") * {"Kbyte": 1e3, "Mbyte": 1e6, "byte": 1.0}.get(
            rows[1][rows[0].index("dram__bytes_read.sum")], 1.0),
        "dram_bytes_written": g("dram__bytes_write.sum") * {"Kbyte": 1e3, "Mbyte": 1e6, "byte": 1.0}.get(
            rows[1][rows[0].index("dram__bytes_write.sum")], 1.0),
        "registers_per_thread": g("launch__registers_per_thread"),
        "warp_instructions": g("smsp__inst_executed.sum"),
        "fp64_warp_instructions": fp64_executed(rep),
        "samples_in_capture": SAMPLES_IN_CAPTURE,
        "warps_issue_stalled_per_issue": stalls,
    }


def main():
    import ebn_b200
    bid = ebn_b200._lib.build_id()
    prev = json.load(open(OUT)) if os.path.exists(OUT) else {}
    prof = {"build_id": bid, "workload": "W1 vehicle suspension, fast mode, compile-time plan (P,P,U,N,N,N)",
            "sass": sass_half(), "ncu": prev.get("ncu")}
    if "--ncu" in sys.argv:
        prof["ncu"] = ncu_half(sys.argv[sys.argv.index("--ncu") + 1], bid)
    with open(OUT, "w") as f:
        json.dump(prof, f, indent=1)
        f.write("\n")
    print(f"{OUT}: build {bid}, {prof['sass']['instructions_per_sample_pair']:.0f} instructions / "
          f"{prof['sass']['fp64_instructions_per_sample_pair']:.0f} fp64 per sample pair; ncu half: "
          f"{'none' if not prof['ncu'] else prof['ncu']['build_id'] + (' (this build)' if prof['ncu']['build_id'] == bid else ' (STALE)')}")


if __name__ == "__main__":
    main()
