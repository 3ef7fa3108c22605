"""SASS audit of a kernel's hot loop (runs on the CPU box: cuobjdump only).

Finds the innermost backward branch with the largest body in the named kernel, prints the opcode
mix per pipe and classifies every fp64 instruction by how many DISTINCT register-pair sources it
reads (operands served by the constant bank `c[..]`, a uniform register `UR`, an immediate or a
`.reuse` slot do not go through the register file read ports).

    python tools/sass_audit.py enhancedbayesiannetworks.jl_b200/csrc/ebn_ls_vehicle.o \
        'mc_kernel.*VehicleSuspension.*FastOps.*StaticPlanIJLi0ELi0ELi1ELi2ELi2ELi2EEEELi0' [--dump loop.sass]
"""
import collections
import re
import subprocess
import sys

PIPE = {
    "DFMA": "fp64", "DMUL": "fp64", "DADD": "fp64", "DSETP": "fp64", "DMNMX": "fp64",
    "LOP3": "alu", "SHF": "alu", "IADD3": "alu", "LEA": "alu", "ISETP": "alu", "SEL": "alu", "VIADD": "alu",
    "PRMT": "alu", "MOV": "alu", "IABS": "alu", "FLO": "xu", "POPC": "xu", "FSETP": "alu",
    "IMAD": "fma", "FFMA": "fma", "FMUL": "fma", "FADD": "fma",
    "MUFU": "xu", "I2F": "xu", "F2I": "xu", "F2F": "xu", "I2FP": "xu",
    "LDS": "lsu", "STS": "lsu", "LDG": "lsu", "STG": "lsu", "ATOMS": "lsu", "RED": "lsu", "LDC": "adu",
    "BRA": "cbu", "BAR": "cbu", "EXIT": "cbu", "BSSY": "cbu", "BSYNC": "cbu", "CALL": "cbu", "RET": "cbu",
    "UMOV": "uniform", "UIADD3": "uniform", "ULOP3": "uniform", "USHF": "uniform", "UIMAD": "uniform",
    "LDCU": "uniform", "REDUX": "alu", "S2R": "xu", "S2UR": "uniform", "NOP": "none",
}


def kernels(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, body = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            body[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return body


def hot_loop(ins, which=0):
    """Innermost loops = backward branches whose body holds no other backward branch. The count kernels
    hold two copies of the sample loop (boundary items mask samples, interior items do not): `which`
    picks among the loops of more than 100 instructions, smallest first (0 = the unmasked interior loop)."""
    loops = []
    for a, text in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", text)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= a:
            continue
        body = [(x, t) for x, t in ins if tgt <= x <= a]
        inner = sum(1 for x, t in body[:-1] if re.search(r"BRA", t) and
                    int(re.search(r"(0x[0-9a-f]+)", t).group(1), 16) < x)
        if inner == 0 and len(body) > 100:
            loops.append(body)
    loops.sort(key=len)
    return loops[min(which, len(loops) - 1)]


def opcode(text):
    t = re.sub(r"^@!?U?P\d+\s+", "", text)
    return t.split()[0].split(".")[0], t


def audit(body, pairs_per_trip=1.0):
    mix, pipe = collections.Counter(), collections.Counter()
    dsrc = collections.Counter()
    for _, text in body:
        op, t = opcode(text)
        mix[op] += 1
        pipe[PIPE.get(op, "other")] += 1
        if PIPE.get(op) == "fp64":
            ops = [o.strip() for o in t.split(None, 1)[1].split(",")][1:]     # sources
            regs = set()
            for o in ops:
                if ".reuse" in o:
                    continue
                m = re.match(r"[-|~]*R(\d+)", o)
                if m:
                    regs.add(int(m.group(1)))
            dsrc[len(regs)] += 1
    n = len(body)
    print(f"hot loop: {n} instructions per trip = {n / pairs_per_trip:.1f} per sample pair")
    print("per pipe :", ", ".join(f"{k} {v / pairs_per_trip:.1f}" for k, v in pipe.most_common()))
    print("opcodes  :", ", ".join(f"{k} {v / pairs_per_trip:.1f}" for k, v in mix.most_common(16)))
    tot = sum(dsrc.values())
    print(f"fp64 instructions by distinct register-pair sources (of {tot / pairs_per_trip:.1f}): " +
          ", ".join(f"{k} pairs: {v / pairs_per_trip:.1f} ({100 * v / tot:.0f}%)" for k, v in sorted(dsrc.items())))
    return n


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    dump = sys.argv[sys.argv.index("--dump") + 1] if "--dump" in sys.argv else None
    ppt = float(sys.argv[sys.argv.index("--pairs") + 1]) if "--pairs" in sys.argv else 1.0
    ks = {k: v for k, v in kernels(obj).items() if re.search(pat, k)}
    if not ks:
        sys.exit(f"no kernel matches {pat}")
    for name, ins in ks.items():
        print(name)
        body = hot_loop(ins)
        audit(body, ppt)
        if dump:
            with open(dump, "w") as f:
                f.write(f"// {name}\n// hot loop, {len(body)} instructions per trip ({ppt:g} sample pairs per trip)\n")
                for a, t in body:
                    f.write(f"/*{a:04x}*/ {t} ;\n")


if __name__ == "__main__":
    main()
