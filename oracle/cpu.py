"""ctypes front end of oracle/ebn_oracle.c (test infrastructure, see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libebn_oracle.so")

# numpy mirror of ebn_input_t (include/ebncuda.h), 56 bytes
INPUT_DTYPE = np.dtype(
    [("kind", "<i4"), ("flags", "<i4"), ("p", "<f8", (4,)), ("lo", "<f8"), ("hi", "<f8")]
)
assert INPUT_DTYPE.itemsize == 56

LS_VEHICLE, LS_STRAUB, LS_HYDROGEN, LS_POLYNOMIAL, LS_MIN_AFFINE, LS_STRAUB_R45, LS_STRAUB_CAPACITY = 1, 2, 3, 4, 5, 6, 7
LS_EXPRESSION = 8
IN_PARAM, IN_NORMAL, IN_LOGNORMAL, IN_UNIFORM, IN_GUMBEL, IN_GAMMA, IN_EXP_AFFINE, IN_TABULATED = range(8)


def build(force: bool = False) -> str:
    """Compiles the C oracle (gcc, a second) if the shared object is missing or stale."""
    src = os.path.join(_HERE, "ebn_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        L.oracle_eval_matrix.restype = C.c_int64
        L.oracle_eval_matrix.argtypes = [C.c_int, dp, dp, C.c_int64, C.c_int, C.c_int64, dp]
        L.oracle_pf_scenario.restype = C.c_int64
        L.oracle_pf_scenario.argtypes = [C.c_int, dp, C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_int64]
        L.oracle_pf_batch.restype = C.c_int
        L.oracle_pf_batch.argtypes = [C.c_int, dp, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_uint64,
                                      C.c_int, C.c_int64, C.POINTER(C.c_int64)]
        L.oracle_phi_inv.restype = C.c_double
        L.oracle_phi_inv.argtypes = [C.c_double]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def eval_matrix(ls: int, consts, X: np.ndarray, want_g: bool = False):
    """X is (n, d); returns fail count (and g). `sum(performance(df) .< 0)`."""
    X = np.asfortranarray(X, dtype=np.float64)
    n, d = X.shape
    consts = np.ascontiguousarray(consts, dtype=np.float64)
    g = np.empty(n, dtype=np.float64) if want_g else None
    cnt = lib().oracle_eval_matrix(ls, _dp(consts), _dp(X), n, d, max(n, 1), _dp(g) if want_g else None)
    return (int(cnt), g) if want_g else int(cnt)


def make_inputs(rows) -> np.ndarray:
    """rows: list (scenarios) of lists of (kind, p0, p1, p2, p3, lo, hi) tuples → [S, d] array."""
    S = len(rows)
    d = len(rows[0])
    a = np.zeros((S, d), dtype=INPUT_DTYPE)
    for s, row in enumerate(rows):
        assert len(row) == d
        for j, t in enumerate(row):
            kind, p, lo, hi = t[0], list(t[1:5]), t[5], t[6]
            a[s, j]["kind"] = kind
            a[s, j]["p"] = p
            a[s, j]["lo"] = lo
            a[s, j]["hi"] = hi
    return a


def pf_batch(ls: int, consts, inputs: np.ndarray, n: int, seed: int = 0, threads: int = 1,
             block: int = 1 << 16) -> np.ndarray:
    """Structural MC over all scenarios; returns int64 fail counts [S]."""
    inputs = np.ascontiguousarray(inputs)
    assert inputs.dtype == INPUT_DTYPE and inputs.ndim == 2
    S, d = inputs.shape
    consts = np.ascontiguousarray(consts, dtype=np.float64)
    out = np.zeros(S, dtype=np.int64)
    rc = lib().oracle_pf_batch(ls, _dp(consts), inputs.ctypes.data, S, d, n, seed, threads, block,
                               out.ctypes.data_as(C.POINTER(C.c_int64)))
    if rc != 0:
        raise ValueError("oracle: unsupported marginal")
    return out


def phi_inv(p: float) -> float:
    return float(lib().oracle_phi_inv(p))
