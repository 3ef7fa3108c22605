"""Julia cannot run here, so the `ccall` wrapper is checked statically against include/ebncuda.h:
struct fields in the same order with matching widths, the limit-state ids, marginal kinds, opcode range
and job flags it hard-codes, and every C symbol it calls being declared (and exported)."""
import os
import re

import ebn_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = open(os.path.join(ROOT, "include", "ebncuda.h")).read()
JL = open(os.path.join(ROOT, "enhancedbayesiannetworks.jl_b200", "julia", "src", "cuda", "EBNCuda.jl")).read()

CTYPE_TO_JL = {"int32_t": "Int32", "uint32_t": "UInt32", "int64_t": "Int64", "uint64_t": "UInt64", "double": "Float64"}


def _c_struct(name):
    body = re.search(r"typedef struct %s \{(.*?)\} %s_t;" % (name, name), HDR, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        m = re.match(r"(const )?(\w+)(\*?) ?(\*?)(\w+)(\[\d+\])?$", decl)
        assert m, decl
        ctype, ptr, fname, arr = m.group(2), m.group(3) or m.group(4), m.group(5), m.group(6)
        fields.append((fname, "Ptr" if ptr else CTYPE_TO_JL[ctype] + (arr or "")))
    return fields


def _jl_struct(name):
    body = re.search(r"struct %s\n(.*?)\nend" % name, JL, re.S).group(1)
    out = []
    for line in body.splitlines():
        line = line.split("#")[0].strip()
        if line:
            fname, ftype = [t.strip() for t in line.split("::")]
            if ftype.startswith("Ptr"):
                ftype = "Ptr"
            m = re.match(r"NTuple\{(\d+),(\w+)\}", ftype)
            if m:
                ftype = f"{m.group(2)}[{m.group(1)}]"
            out.append((fname, ftype))
    return out


def test_structs_match_the_header_field_by_field():
    assert _jl_struct("EbnInput") == _c_struct("ebn_input")
    cj = _c_struct("ebn_job")
    jj = _jl_struct("EbnJob")
    assert [t for _, t in jj] == [t for _, t in cj]
    assert [n for n, _ in jj] == [n for n, _ in cj]
    assert "@assert sizeof(EbnInput) == 56" in JL and "@assert sizeof(EbnJob) == 104" in JL


def test_constants_match_the_header():
    ls = dict(re.findall(r"EBN_LS_(\w+) = (\d+)", HDR))
    jl_ls = dict(re.findall(r":(\w+) => (\d+)", re.search(r"const EBN_LS = Dict\((.*?)\)\n", JL, re.S).group(1)))
    assert {k.lower(): v for k, v in ls.items()} == jl_ls
    kinds = re.findall(r"EBN_IN_(\w+) = (\d+)", HDR)
    assert [int(v) for _, v in kinds] == list(range(8)) and "IN_TABULATED = 0:7" in JL.replace("\n", " ")
    nops = int(re.search(r"EBN_OP_NOPS = (\d+)", HDR).group(1))
    jl_ops = re.search(r"const (OP_INPUT.*?) = 0:(\d+)", JL, re.S)
    names = [n.strip() for n in jl_ops.group(1).replace("\n", " ").split(",")]
    hdr_ops = re.findall(r"EBN_OP_(\w+) = (\d+)", HDR)[:-1]
    assert int(jl_ops.group(2)) == nops - 1 and len(names) == nops
    assert [n.replace("OP_", "") for n in names] == [n for n, _ in hdr_ops]
    assert [int(v) for _, v in hdr_ops] == list(range(nops))
    flags = dict(re.findall(r"#define EBN_JOB_(\w+) (\d+)u", HDR))
    jl_flags = re.search(r"const (JOB_\w+(?:, JOB_\w+)*) = UInt32\.\(\(([\d, ]+)\)\)", JL)
    jl_names = [n.strip().replace("JOB_", "") for n in jl_flags.group(1).split(",")]
    jl_vals = [v.strip() for v in jl_flags.group(2).split(",")]
    assert dict(zip(jl_names, jl_vals)) == flags


def test_every_symbol_the_wrapper_calls_is_declared_and_exported():
    called = set(re.findall(r"ccall\(\(:(\w+),", JL))
    assert called >= {"ebn_pf_batch", "ebn_hist_batch", "ebn_init", "ebn_shutdown", "ebn_last_error"}
    lib = ebn_b200._lib.load()
    for sym in called:
        assert re.search(r"\b%s\(" % sym, HDR), sym
        getattr(lib, sym)


def test_every_job_field_and_every_marginal_kind_is_populated_by_some_wrapper_path():
    """The wrapper builds ebn_job_t in ONE place (with_job): every pointer field must be fed from Julia data
    (not hard-wired to C_NULL), and every EBN_IN_* kind must be produced by some `abi` method."""
    ctor = re.search(r"job = Ref\(EbnJob\((.*?)\)\)\n", JL, re.S).group(1)
    assert JL.count("Ref(EbnJob(") == 1
    for needle in ("pointer(consts)", "pointer(inputs)", "pointer(chol)", "pointer(stream)", "pointer(tables)",
                   "length(tables)", "sim.n", "sim.seed", "sim.strict", "job_flags(sim)", "jd.S", "jd.d"):
        assert needle in ctor, needle
    # 17 fields in the header, 17 arguments in the constructor call (commas at depth 0)
    depth, args = 0, 1
    for ch in ctor:
        depth += ch in "([" 
        depth -= ch in ")]"
        args += ch == "," and depth == 0
    assert args == len(_c_struct("ebn_job")) == 17
    for kind in ("IN_PARAM", "IN_NORMAL", "IN_LOGNORMAL", "IN_UNIFORM", "IN_GUMBEL", "IN_GAMMA", "IN_EXP_AFFINE", "IN_TABULATED"):
        assert re.search(r"EbnInput\(%s," % kind, JL), kind
    # the imprecise simulations and the copula reach the job: stream ids, Cholesky factors, table pool
    for needle in ("struct CudaDoubleLoop", "struct CudaRandomSlicing", "stream=stream", "cholesky(Symmetric(R))",
                   "tabulated(d, lo, hi, pool)", "ImpreciseDiscreteProbability", "JointDistribution"):
        assert needle in JL, needle


def test_wrapper_is_syntactically_balanced():
    """No Julia here: at least every block opener has its `end` and brackets balance (comments and strings stripped)."""
    src = re.sub(r'"""(.*?)"""', '""', JL, flags=re.S)
    src = re.sub(r'"(?:\\.|[^"\\])*"', '""', src)
    src = re.sub(r"#.*", "", src)
    for o, c in ("()", "[]", "{}"):
        assert src.count(o) == src.count(c), (o, src.count(o), src.count(c))
    lines = [ln.strip() for ln in src.splitlines()]
    openers = sum(bool(re.match(r"(?:function|struct|mutable struct|if|for|while|let|module|try|quote)\b", ln)) for ln in lines)
    openers += sum(bool(re.search(r"\bbegin$", ln)) for ln in lines)              # `x = begin`, `GC.@preserve a b begin`
    openers += sum(bool(re.search(r"\bdo(\s+\w+)?$", ln)) for ln in lines)         # `f(args) do job`
    ends = sum(ln == "end" or ln.startswith("end ") for ln in lines)
    assert openers == ends, (openers, ends)
