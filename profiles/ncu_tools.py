"""Helpers to summarise ncu reports (read on the CPU box): opcode mix weighted by executed
count and stall reasons from `ncu -i X.ncu-rep --page source --csv`, key metrics from `--page raw`."""
import collections
import csv
import re
import subprocess
import sys


def source_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    return hdr, [r for r in rows[2:] if len(r) == len(hdr) and r[0].startswith("0x")]


def opcode_mix(rep, pairs_warps=None):
    hdr, rows = source_rows(rep)
    i_src, i_exec = hdr.index("Source"), hdr.index("Instructions Executed")
    i_samp = hdr.index("# Samples")
    mix, samp = collections.Counter(), collections.Counter()
    for r in rows:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[i_src])
        op = m.group(2) if m else "?"
        mix[op] += int(r[i_exec])
        samp[op] += int(r[i_samp])
    tot = sum(mix.values())
    print(f"total warp instructions: {tot}")
    for op, c in mix.most_common(30):
        extra = f"  per pair-warp {c / pairs_warps:8.1f}" if pairs_warps else ""
        print(f"{op:10s} {c:14d} {100 * c / tot:6.2f}%  stall samples {samp[op]:8d}{extra}")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    st = collections.Counter()
    for r in rows:
        for i in stall_cols:
            st[hdr[i]] += int(r[i] or 0)
    tots = sum(st.values())
    print("stall reasons (all samples):", ", ".join(f"{k[6:]} {100 * v / tots:.1f}%" for k, v in st.most_common(8)))


def raw_metrics(rep, patterns):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    for h, u, v in zip(rows[0], rows[1], rows[2]):
        if any(re.search(p, h) for p in patterns):
            print(f"{h} [{u}] = {v}")


if __name__ == "__main__":
    rep = sys.argv[1]
    pw = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw_metrics(rep, [r"^gpu__time_duration.sum$", r"^dram__bytes_(read|write).sum$", r"^smsp__inst_executed.sum$",
                      r"^sm__inst_executed_pipe_(fp64|fma|alu|xu|uniform|fmaheavy|fmalite)\.sum$",
                      r"^smsp__inst_executed_pipe_fp64", r"^sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
                      r"^sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", r"^smsp__issue_active.avg.pct",
                      r"^sm__warps_active.avg.pct_of_peak_sustained_active", r"^launch__registers_per_thread$",
                      r"^launch__grid_size", r"^smsp__cycles_active.avg$", r"sm__cycles_elapsed.avg$"])
    opcode_mix(rep, pw)
