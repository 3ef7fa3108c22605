"""Issue-cost model of the W1 count loop, checked against every kernel variant measured in round 2.

Hypothesis: on B200 an instruction holds the sub-partition's dispatch for as many
cycles as it writes 32-bit destination words (at least one): DFMA/DMUL/DADD/IMAD.WIDE/I2F.F64 = 2,
LDS.128 = 4, everything else = 1. The cost of one trip of the loop is the sum over its instructions, and
no number of resident warps can go below it (occupancy sweep: one warp per sub-partition already reaches
88 % of the best throughput). Predicted cycles per sample-pair-warp per sub-partition = that sum;
measured = 148 SMs x 4 sub-partitions x 1.965 GHz x 64 samples / (samples/s).

    python tools/issue_cost_model.py VARIANT_DIR=MEASURED_SAMPLES_PER_S[:PAIRS_PER_TRIP] ...
"""
import re
import sys

sys.path.insert(0, __file__.rsplit("/", 1)[0])
import sass_audit as A  # noqa: E402

KERNEL = r"mc_kernel.*VehicleSuspension.*FastOps.*StaticPlanIJLi0ELi0ELi1ELi2ELi2ELi2EEEELi0"


def dest_words(text):
    op, t = A.opcode(text)
    full = t.split()[0]
    if op in ("DFMA", "DMUL", "DADD"):
        return 2
    if op == "IMAD" and ".WIDE" in full:
        return 2
    if op == "I2F" and "F64" in full:
        return 2
    if op in ("LDS", "LDG", "LDC"):
        m = re.search(r"\.(64|128)", full)
        return {"64": 2, "128": 4}.get(m.group(1), 1) if m else 1
    return 1


def main():
    print(f"{'variant':14s} {'instr/pair':>10s} {'fp64':>6s} {'sum dest words':>15s} {'measured cycles':>16s} {'ratio':>6s}")
    for arg in sys.argv[1:]:
        path, rest = arg.split("=")
        val, _, ppt = rest.partition(":")
        ppt = float(ppt or 1)
        ks = {k: v for k, v in A.kernels(path + "/pkg/csrc/ebn_ls_vehicle.o").items() if re.search(KERNEL, k)}
        body = A.hot_loop(next(iter(ks.values())))
        cost = sum(dest_words(t) for _, t in body) / ppt
        fp64 = sum(1 for _, t in body if A.PIPE.get(A.opcode(t)[0]) == "fp64") / ppt
        cyc = 148 * 4 * 1.965e9 * 64 / float(val)
        print(f"{path.rsplit('_', 1)[-1]:14s} {len(body) / ppt:10.1f} {fp64:6.1f} {cost:15.1f} {cyc:16.1f} {cyc / cost:6.3f}")


if __name__ == "__main__":
    main()
