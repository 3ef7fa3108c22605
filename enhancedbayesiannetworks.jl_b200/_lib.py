"""ctypes binding of libebncuda.so (the C ABI in include/ebncuda.h).

There is no fallback of any kind: if the shared library is missing, or no CUDA
device is visible, the calls raise. The CPU oracle under ``oracle/`` is test
infrastructure and is never imported from here.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EBN_LIB", os.path.join(_HERE, "libebncuda.so"))  # EBN_LIB: tuning builds only

# ---- constants mirrored from include/ebncuda.h ------------------------------
EBN_ABI_VERSION = 2
EBN_MAX_INPUTS = 16
EBN_MAX_EDGES = 1025

LS_VEHICLE_SUSPENSION = 1
LS_STRAUB_FRAME = 2
LS_HYDROGEN_RADIUS = 3
LS_POLYNOMIAL = 4
LS_MIN_AFFINE = 5
LS_STRAUB_FRAME_R45 = 6
LS_STRAUB_CAPACITY = 7
LS_EXPRESSION = 8

# expression program opcodes (EBN_OP_*)
(OP_INPUT, OP_PARAM, OP_CONST, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_ABS, OP_SQRT, OP_EXP, OP_LOG,
 OP_SIN, OP_COS, OP_MIN, OP_MAX, OP_POW, OP_POWI, OP_LT, OP_SELECT, OP_STORE) = range(21)

IN_PARAM = 0
IN_NORMAL = 1
IN_LOGNORMAL = 2
IN_UNIFORM = 3
IN_GUMBEL = 4
IN_GAMMA = 5
IN_EXPONENTIAL_AFFINE = 6
IN_TABULATED = 7

JOB_COPULA_SHARED = 1
JOB_PARTIAL = 2
JOB_JIT_PLAN = 4
JOB_NO_JIT_PLAN = 8
JOB_FP32_SAMPLING = 16
JOB_QMC_SOBOL = 32

ERROR_NAMES = {0: "EBN_OK", -1: "EBN_EINVAL", -2: "EBN_ELIMITSTATE", -3: "EBN_EMARGINAL",
               -4: "EBN_ECUDA", -5: "EBN_ENCCL", -6: "EBN_ENODEV", -7: "EBN_ENOMEM", -8: "EBN_ESTATE",
               -9: "EBN_EJIT"}

INPUT_DTYPE = np.dtype(
    [("kind", "<i4"), ("flags", "<i4"), ("p", "<f8", (4,)), ("lo", "<f8"), ("hi", "<f8")]
)
assert INPUT_DTYPE.itemsize == 56


class EbnJob(C.Structure):
    _fields_ = [
        ("limit_state", C.c_int32), ("n_consts", C.c_int32), ("consts", C.c_void_p),
        ("n_scenarios", C.c_int32), ("n_inputs", C.c_int32), ("inputs", C.c_void_p),
        ("n_per_scenario", C.c_int64), ("seed", C.c_uint64), ("strict", C.c_int32),
        ("flags", C.c_uint32), ("corr_chol", C.c_void_p), ("stream_id", C.c_void_p),
        ("tables", C.c_void_p), ("n_tables", C.c_int64), ("sample_offset", C.c_int64),
        ("shard_rank", C.c_int32), ("shard_world", C.c_int32),
    ]


class EbnStats(C.Structure):
    _fields_ = [
        ("kernel_ms", C.c_double), ("h2d_bytes", C.c_double), ("d2h_bytes", C.c_double),
        ("samples", C.c_int64), ("launches", C.c_int32), ("devices", C.c_int32),
        ("specialised", C.c_int32), ("jit_compiles", C.c_int32),
    ]


assert C.sizeof(EbnJob) == 104 and C.sizeof(EbnStats) == 48

EXPORTS = [
    "ebn_abi_version", "ebn_build_id", "ebn_device_count", "ebn_init", "ebn_shutdown", "ebn_set_stream",
    "ebn_nccl_unique_id", "ebn_comm_init_rank", "ebn_comm_destroy", "ebn_pf_batch",
    "ebn_hist_batch", "ebn_eval_matrix", "ebn_sample_matrix", "ebn_sample_matrix_ex", "ebn_peak_fp64", "ebn_last_stats",
    "ebn_partition", "ebn_last_error", "ebn_limit_state_id", "ebn_limit_state_name", "ebn_limit_state_arity",
    "ebn_jit_available", "ebn_expression_check", "ebn_expression_source", "ebn_jit_compile",
]


class EbnError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {message}")
        self.code = code
        self.message = message


_lib = None


def load():
    """Loads libebncuda.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C enhancedbayesiannetworks.jl_b200/csrc`). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i64p, dp = C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_double)
    L.ebn_abi_version.restype = C.c_int
    L.ebn_build_id.restype = C.c_char_p
    L.ebn_device_count.restype = C.c_int
    L.ebn_init.argtypes = [C.POINTER(C.c_int), C.c_int]
    L.ebn_set_stream.argtypes = [vp]
    L.ebn_nccl_unique_id.argtypes = [vp]
    L.ebn_comm_init_rank.argtypes = [vp, C.c_int, C.c_int]
    L.ebn_pf_batch.argtypes = [C.POINTER(EbnJob), i64p, dp, dp]
    L.ebn_hist_batch.argtypes = [C.POINTER(EbnJob), dp, C.c_int, i64p, dp, dp]
    L.ebn_eval_matrix.argtypes = [C.c_int, dp, C.c_int, dp, C.c_int64, C.c_int, C.c_int64, C.c_int, i64p, dp]
    L.ebn_sample_matrix.argtypes = [C.POINTER(EbnJob), C.c_int, C.c_int64, C.c_int64, dp, dp]
    L.ebn_sample_matrix_ex.argtypes = [C.POINTER(EbnJob), C.c_int, C.c_int64, C.c_int64, C.c_int, dp, dp]
    L.ebn_peak_fp64.argtypes = [C.c_double, dp]
    L.ebn_partition.argtypes = [C.POINTER(EbnJob), C.c_int, i64p]
    L.ebn_last_stats.argtypes = [C.POINTER(EbnStats)]
    L.ebn_last_error.restype = C.c_char_p
    L.ebn_limit_state_id.argtypes = [C.c_char_p]
    L.ebn_limit_state_name.restype = C.c_char_p
    L.ebn_limit_state_name.argtypes = [C.c_int]
    L.ebn_limit_state_arity.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.ebn_jit_available.restype = C.c_int
    L.ebn_expression_check.argtypes = [dp, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.ebn_expression_source.restype = C.c_int64
    L.ebn_expression_source.argtypes = [dp, C.c_int, C.c_int, C.c_char_p, C.c_int64]
    L.ebn_jit_compile.argtypes = [C.POINTER(EbnJob), C.c_int, i64p]
    if L.ebn_abi_version() != EBN_ABI_VERSION:
        raise RuntimeError("libebncuda.so ABI version mismatch")
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise EbnError(rc, load().ebn_last_error().decode())


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _i64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


_initialised = False


def build_id() -> str:
    """ebn_build_id: hash of the kernel sources and tuning flags the library was built from."""
    return load().ebn_build_id().decode()


def device_count() -> int:
    return load().ebn_device_count()


def init(devices: Optional[Sequence[int]] = None):
    """ebn_init. devices=None → device 0."""
    global _initialised
    L = load()
    if devices is None:
        _check(L.ebn_init(None, 0))
    else:
        arr = (C.c_int * len(devices))(*devices)
        _check(L.ebn_init(arr, len(devices)))
    _initialised = True


def ensure_init():
    if not _initialised:
        init()


def shutdown():
    global _initialised
    load().ebn_shutdown()
    _initialised = False


def set_stream(ptr: Optional[int]):
    _check(load().ebn_set_stream(C.c_void_p(ptr) if ptr else None))


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(load().ebn_nccl_unique_id(buf))
    return buf.raw


def comm_init_rank(uid: bytes, rank: int, world: int):
    assert len(uid) == 128
    buf = C.create_string_buffer(uid, 128)
    _check(load().ebn_comm_init_rank(buf, rank, world))


def comm_destroy():
    load().ebn_comm_destroy()


def make_inputs(rows) -> np.ndarray:
    """rows[s][j] = (kind, p0, p1, p2, p3, lo, hi) → structured array [S, d]."""
    S, d = len(rows), len(rows[0])
    a = np.zeros((S, d), dtype=INPUT_DTYPE)
    for s, row in enumerate(rows):
        if len(row) != d:
            raise ValueError("ragged scenario rows")
        for j, t in enumerate(row):
            a[s, j]["kind"] = t[0]
            a[s, j]["p"] = list(t[1:5])
            a[s, j]["lo"] = t[5]
            a[s, j]["hi"] = t[6]
    return a


@dataclass
class Job:
    """Python-side owner of the buffers an ebn_job_t points at."""
    limit_state: int
    consts: np.ndarray
    inputs: np.ndarray            # [S, d] INPUT_DTYPE
    n: int
    seed: int = 0
    strict: bool = False
    corr_chol: Optional[np.ndarray] = None   # [S, d, d] or [d, d]
    stream_id: Optional[np.ndarray] = None   # [S] uint32
    tables: Optional[np.ndarray] = None
    sample_offset: int = 0
    shard_rank: int = 0
    shard_world: int = 1
    partial: bool = False          # EBN_JOB_PARTIAL: this shard's counts only
    jit_plan: Optional[bool] = None  # True: EBN_JOB_JIT_PLAN, False: EBN_JOB_NO_JIT_PLAN, None: library decides
    fp32_sampling: bool = False    # EBN_JOB_FP32_SAMPLING: untruncated normals through the fp32 Box-Muller
    qmc: bool = False              # EBN_JOB_QMC_SOBOL: Owen-scrambled Sobol points instead of the Philox streams

    def __post_init__(self):
        self.consts = np.ascontiguousarray(self.consts, dtype=np.float64).ravel()
        self.inputs = np.ascontiguousarray(self.inputs)
        if self.inputs.dtype != INPUT_DTYPE or self.inputs.ndim != 2:
            raise ValueError("inputs must be a [S, d] array of INPUT_DTYPE")
        if self.corr_chol is not None:
            self.corr_chol = np.ascontiguousarray(self.corr_chol, dtype=np.float64)
        if self.stream_id is not None:
            self.stream_id = np.ascontiguousarray(self.stream_id, dtype=np.uint32)
            if self.stream_id.shape != (self.inputs.shape[0],):
                raise ValueError("stream_id must have one entry per scenario")
        if self.tables is not None:
            self.tables = np.ascontiguousarray(self.tables, dtype=np.float64).ravel()

    @property
    def S(self) -> int:
        return self.inputs.shape[0]

    @property
    def d(self) -> int:
        return self.inputs.shape[1]

    def c_struct(self) -> EbnJob:
        j = EbnJob()
        j.limit_state = self.limit_state
        j.n_consts = self.consts.size
        j.consts = self.consts.ctypes.data if self.consts.size else None
        j.n_scenarios, j.n_inputs = self.S, self.d
        j.inputs = self.inputs.ctypes.data
        j.n_per_scenario = self.n
        j.seed = self.seed & 0xFFFFFFFFFFFFFFFF
        j.strict = 1 if self.strict else 0
        j.flags = JOB_PARTIAL if self.partial else 0
        if self.jit_plan is not None:
            j.flags |= JOB_JIT_PLAN if self.jit_plan else JOB_NO_JIT_PLAN
        if self.fp32_sampling:
            j.flags |= JOB_FP32_SAMPLING
        if self.qmc:
            j.flags |= JOB_QMC_SOBOL
        if self.corr_chol is not None:
            j.corr_chol = self.corr_chol.ctypes.data
            if self.corr_chol.ndim == 2:
                j.flags |= JOB_COPULA_SHARED
        if self.stream_id is not None:
            j.stream_id = self.stream_id.ctypes.data
        if self.tables is not None:
            j.tables = self.tables.ctypes.data
            j.n_tables = self.tables.size
        j.sample_offset = self.sample_offset
        j.shard_rank, j.shard_world = self.shard_rank, self.shard_world
        return j


def pf_batch(job: Job):
    """ebn_pf_batch → (fail_count int64[S], pf f64[S], var f64[S])."""
    ensure_init()
    S = job.S
    cnt = np.zeros(S, dtype=np.int64)
    pf = np.zeros(S)
    var = np.zeros(S)
    cj = job.c_struct()
    _check(load().ebn_pf_batch(C.byref(cj), _i64p(cnt), _dp(pf), _dp(var)))
    return cnt, pf, var


def pf_batch_resumable(job: Job, chunk: int, checkpoint: str, max_chunks: Optional[int] = None):
    """A long ebn_pf_batch run in pieces of `chunk` samples per scenario with a checkpoint after every
    piece: (seed, next sample offset, int64 counts) — a few KB even for 1e13 samples. A run that is
    interrupted continues from the file; because every sample is a pure function of (seed, stream,
    index), the final counts are bit-identical to one uninterrupted call (SURVEY.md §5
    checkpoint / resume). Returns (fail_count, pf, var, done): done is False when `max_chunks`
    stopped the run early."""
    import dataclasses
    import json
    S, n, first = job.S, job.n, job.sample_offset
    state = {"seed": int(job.seed), "limit_state": int(job.limit_state), "S": S, "n": n, "first": first,
             "next": first, "counts": [0] * S}
    if os.path.exists(checkpoint):
        with open(checkpoint) as f:
            saved = json.load(f)
        same = all(saved.get(k) == state[k] for k in ("seed", "limit_state", "S", "n", "first"))
        if not same:
            raise ValueError(f"{checkpoint} belongs to a different job")
        state = saved
    counts = np.asarray(state["counts"], dtype=np.int64)
    done_chunks = 0
    while state["next"] < first + n and (max_chunks is None or done_chunks < max_chunks):
        m = min(chunk, first + n - state["next"])
        piece = dataclasses.replace(job, n=m, sample_offset=state["next"])
        counts = counts + pf_batch(piece)[0]
        state["next"] += m
        state["counts"] = [int(c) for c in counts]
        tmp = checkpoint + ".tmp"
        with open(tmp, "w") as f:
            json.dump(state, f)
        os.replace(tmp, checkpoint)
        done_chunks += 1
    finished = state["next"] >= first + n
    seen = state["next"] - first
    pf = counts / max(seen, 1)
    return counts, pf, (pf - pf * pf) / max(seen, 1), finished


def hist_batch(job: Job, edges):
    """ebn_hist_batch → (counts int64[S, nedges+1], sum f64[S], sumsq f64[S])."""
    ensure_init()
    edges = np.ascontiguousarray(edges, dtype=np.float64)
    S = job.S
    counts = np.zeros((S, edges.size + 1), dtype=np.int64)
    s1 = np.zeros(S)
    s2 = np.zeros(S)
    cj = job.c_struct()
    _check(load().ebn_hist_batch(C.byref(cj), _dp(edges), edges.size, _i64p(counts), _dp(s1), _dp(s2)))
    return counts, s1, s2


def eval_matrix(limit_state: int, consts, X, strict: bool = True, want_g: bool = False):
    """ebn_eval_matrix on an (n, d) matrix → fail count (and g)."""
    ensure_init()
    X = np.asfortranarray(X, dtype=np.float64)
    n, d = X.shape
    consts = np.ascontiguousarray(consts, dtype=np.float64).ravel()
    cnt = C.c_int64(0)
    g = np.empty(n) if want_g else None
    _check(load().ebn_eval_matrix(limit_state, _dp(consts) if consts.size else None, consts.size,
                                  _dp(X) if n else None, n, d, max(n, 0), 1 if strict else 0,
                                  C.byref(cnt), _dp(g)))
    return (cnt.value, g) if want_g else cnt.value


def sample_matrix(job: Job, scenario: int, first: int, n: int, want_g: bool = False, ncols: int = 0):
    """ebn_sample_matrix(_ex) → X (n, ncols or d) exactly as ebn_pf_batch draws it (and g)."""
    ensure_init()
    ncols = ncols or job.d
    X = np.empty((n, ncols), order="F")
    g = np.empty(n) if want_g else None
    cj = job.c_struct()
    _check(load().ebn_sample_matrix_ex(C.byref(cj), scenario, first, n, ncols, _dp(X), _dp(g)))
    return (X, g) if want_g else X


def partition(job: Job, histogram: bool = False) -> dict:
    """ebn_partition: the static (scenario, chunk) work grid of a job. Needs no GPU."""
    out = np.zeros(4, dtype=np.int64)
    cj = job.c_struct()
    _check(load().ebn_partition(C.byref(cj), 1 if histogram else 0, _i64p(out)))
    return {"pairs_per_item": int(out[0]), "items_per_scenario": int(out[1]),
            "local_items": int(out[2]), "pair_base": int(out[3])}


def peak_fp64(seconds: float = 0.25) -> float:
    ensure_init()
    out = C.c_double(0.0)
    _check(load().ebn_peak_fp64(seconds, C.byref(out)))
    return out.value


def last_stats() -> dict:
    st = EbnStats()
    _check(load().ebn_last_stats(C.byref(st)))
    return {k: getattr(st, k) for k, _ in EbnStats._fields_}


def jit_available() -> bool:
    """ebn_jit_available: can kernels be compiled at run time (libnvrtc loadable)? Needs no GPU."""
    return bool(load().ebn_jit_available())


def expression_check(consts, d: int) -> int:
    """ebn_expression_check → number of columns (d + stored model outputs). Raises on a bad program."""
    consts = np.ascontiguousarray(consts, dtype=np.float64).ravel()
    ncol = C.c_int(0)
    _check(load().ebn_expression_check(_dp(consts) if consts.size else None, consts.size, d, C.byref(ncol)))
    return ncol.value


def expression_source(consts, d: int) -> str:
    """ebn_expression_source: the CUDA functor generated for an expression program."""
    consts = np.ascontiguousarray(consts, dtype=np.float64).ravel()
    L = load()
    n = L.ebn_expression_source(_dp(consts) if consts.size else None, consts.size, d, None, 0)
    if n < 0:
        raise EbnError(int(n), L.ebn_last_error().decode())
    buf = C.create_string_buffer(int(n) + 1)
    L.ebn_expression_source(_dp(consts), consts.size, d, buf, int(n) + 1)
    return buf.value.decode()


def jit_compile(job: Job, role: int = 0) -> int:
    """ebn_jit_compile → bytes of machine code (0: the job runs a compiled-in kernel). Needs no GPU."""
    out = C.c_int64(0)
    cj = job.c_struct()
    _check(load().ebn_jit_compile(C.byref(cj), role, C.byref(out)))
    return out.value


def limit_state_id(name: str) -> int:
    rc = load().ebn_limit_state_id(name.encode())
    if rc < 0:
        raise EbnError(rc, load().ebn_last_error().decode())
    return rc


def limit_state_name(ls: int) -> Optional[str]:
    r = load().ebn_limit_state_name(ls)
    return r.decode() if r else None


def limit_state_arity(ls: int):
    a, b = C.c_int(0), C.c_int(0)
    _check(load().ebn_limit_state_arity(ls, C.byref(a), C.byref(b)))
    return a.value, b.value
