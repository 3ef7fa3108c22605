"""Every hand-written limit-state functor of csrc/ebn_limit_states.cuh compiled for the HOST from its own source
text (g++ behind a shim for the handful of CUDA intrinsics it uses) and compared with the C oracle — no GPU:

  * strict eval() of every functor equals the oracle bit for bit (the functor text IS Julia's evaluation order;
    exp is the same libm on both sides here, on the device it is CUDA's and the GPU suite allows its last ulp),
  * fast eval() stays within rounding,
  * eval_pair() of the two program-driven functors (polynomial program, min of affine forms) returns exactly what
    two eval() calls return, in both arithmetic policies.
The vehicle functor's sign-only fails() has its own file (test_vehicle_predicate_cpu.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from ebn_b200 import workloads as W
from oracle import cpu as ocpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = r"""
#include <math.h>
#include <string.h>
#define __device__
#define __forceinline__ inline
#define EBN_MAX_INPUTS 16
#define EBN_BLOCKS_FOR_WARPS(w) 1
#define EBN_LS_VEHICLE_SUSPENSION 1
#define EBN_LS_STRAUB_FRAME 2
#define EBN_LS_HYDROGEN_RADIUS 3
#define EBN_LS_POLYNOMIAL 4
#define EBN_LS_MIN_AFFINE 5
#define EBN_LS_STRAUB_FRAME_R45 6
#define EBN_LS_STRAUB_CAPACITY 7
namespace ebn {
struct FastMathSmem {};
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dsqrt_rn(double a) { return sqrt(a); }
inline double __ldg(const double* p) { return *p; }
inline int __double2hiint(double d) { long long v; memcpy(&v, &d, 8); return (int)(v >> 32); }
inline double exp_fast(double x, const FastMathSmem&) { return exp(x); }
template <bool S> struct OpsT {
  static constexpr bool strict = S;
  static double add(double a, double b) { return a + b; }
  static double sub(double a, double b) { return a - b; }
  static double mul(double a, double b) { return a * b; }
  static double div(double a, double b) { return a / b; }
  static double sqrt(double a) { return ::sqrt(a); }
  static double min(double a, double b) { return (b < a || b != b) ? b : a; }
  static double max(double a, double b) { return (b > a || b != b) ? b : a; }
};
typedef OpsT<true> StrictOps;
typedef OpsT<false> FastOps;
"""
DRIVER = r"""
template <class F>
static void run_t(const double* k, int nk, const double* X, long n, int d, double* g, double* gp) {
  FastMathSmem sm;
  typename F::Ctx c;
  F::setup(c, k, nk, d, &sm);
  for (long i = 0; i < n; ++i) {
    double x[EBN_MAX_INPUTS + 1] = {0};
    for (int j = 0; j < d; ++j) x[j] = X[i * d + j];
    g[i] = F::eval(c, x);
  }
  (void)gp;
}
template <class F>
static void pair_t(const double* k, int nk, const double* X, long n, int d, double* g, double* gp) {
  FastMathSmem sm;
  typename F::Ctx c;
  F::setup(c, k, nk, d, &sm);
  for (long i = 0; i + 1 < n; i += 2) {
    double x0[EBN_MAX_INPUTS + 1] = {0}, x1[EBN_MAX_INPUTS + 1] = {0};
    for (int j = 0; j < d; ++j) { x0[j] = X[i * d + j]; x1[j] = X[(i + 1) * d + j]; }
    F::eval_pair(c, x0, x1, gp[i], gp[i + 1]);
  }
  run_t<F>(k, nk, X, n, d, g, gp);
}
}  // namespace ebn
extern "C" void run(int ls, int strict, const double* k, int nk, const double* X, long n, int d, double* g, double* gp) {
  using namespace ebn;
#define CASE(ID, T, FN) case ID: if (strict) FN<T<StrictOps> >(k, nk, X, n, d, g, gp); else FN<T<FastOps> >(k, nk, X, n, d, g, gp); break;
  switch (ls) {
    CASE(1, VehicleSuspension, run_t)
    CASE(2, StraubFrame, run_t)
    CASE(3, HydrogenRadius, run_t)
    CASE(4, PolynomialProgram, pair_t)
    CASE(5, MinAffine, pair_t)
    CASE(6, StraubFrameR45, run_t)
    CASE(7, StraubCapacity, run_t)
  }
}
"""


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("functors")
    text = open(os.path.join(ROOT, "enhancedbayesiannetworks.jl_b200", "csrc", "ebn_limit_states.cuh")).read()
    a = text.index("// ---------------------------------------------------------------------------\n// Vehicle suspension")
    b = text.rindex("}  // namespace ebn")
    src = tmp / "functors.cpp"
    src.write_text(SHIM + text[a:b] + DRIVER)
    so = tmp / "functors.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas", "-shared", "-fPIC", str(src), "-o", str(so)],
                   check=True)
    lib = C.CDLL(str(so))
    dp = C.POINTER(C.c_double)
    lib.run.argtypes = [C.c_int, C.c_int, dp, C.c_int, dp, C.c_long, C.c_int, dp, dp]

    def run(ls, consts, X, strict):
        X = np.ascontiguousarray(X, dtype=np.float64)
        k = np.ascontiguousarray(consts, dtype=np.float64)
        if k.size == 0:
            k = np.zeros(1)
        g, gp = np.full(len(X), np.nan), np.full(len(X), np.nan)
        lib.run(ls, int(strict), k.ctypes.data_as(dp), int(np.asarray(consts).size), X.ctypes.data_as(dp), len(X), X.shape[1],
                g.ctypes.data_as(dp), gp.ctypes.data_as(dp))
        return g, gp
    return run


def _poly(stages):
    out = [float(len(stages))]
    for terms in stages:
        out.append(float(len(terms)))
        for coef, cols in terms:
            out += [float(coef), float(len(cols))] + [float(c) for c in cols]
    return np.array(out)


def test_every_functor_on_the_host_against_the_oracle(host):
    rng = np.random.default_rng(11)
    n = 200_000
    lam, zeta = W.lognormal_params(150.0, 30.0)
    straub_k = np.array([lam, zeta, W.STRAUB_RHO])
    cases = [
        (ocpu.LS_VEHICLE, W.vehicle_consts(), np.column_stack([rng.choice([0.15915, 0.8], n), rng.choice([0.27, 0.5], n), rng.uniform(7, 12, n),
                                                              rng.normal(431.7, 60, n), rng.normal(1475.5, 60, n), rng.normal(55.0, 30, n)])),
        (ocpu.LS_STRAUB, straub_k, np.column_stack([rng.normal(0, 1, n), rng.gamma(25, 2.4, n), rng.gumbel(41.0, 15.6, n)] +
                                                   [rng.normal(0, 1, n) for _ in range(5)])),
        (ocpu.LS_STRAUB_R45, straub_k, np.column_stack([rng.normal(0, 1, n), rng.gamma(25, 2.4, n), rng.gumbel(41.0, 15.6, n),
                                                       rng.lognormal(lam, zeta, n), rng.lognormal(lam, zeta, n)] +
                                                      [rng.normal(0, 1, n) for _ in range(3)])),
        (ocpu.LS_STRAUB_CAPACITY, straub_k, rng.normal(0, 1, (n, 2))),
        (ocpu.LS_HYDROGEN, np.zeros(0), np.column_stack([rng.integers(0, 5, n).astype(float), rng.normal(10, 3, n), rng.normal(9, 3, n),
                                                        rng.normal(11, 3, n)])),
        (ocpu.LS_POLYNOMIAL, _poly([[(1.0, [0, 0]), (-0.7, []), (1.0, [1])], [(0.5, [2]), (1.0, [0, 1, 1]), (-0.25, [])],
                                    [(2.0, []), (-1.0, [3]), (0.125, [2, 3])]]), rng.normal(0.2, 1.3, (n, 2))),
        (ocpu.LS_MIN_AFFINE, np.array([3.0, 1.0, 1.0, -2.0, 0.5, -0.3, 0.25, 0.25, 0.25, 2.0, -1.0, 0.0, 1.5]), rng.normal(0.0, 2.0, (n, 3))),
    ]
    for ls, k, X in cases:
        want_cnt, want_g = ocpu.eval_matrix(ls, k, X, want_g=True)
        gs, gps = host(ls, k, X, strict=True)
        assert np.array_equal(gs, want_g, equal_nan=True), ls
        assert int(np.sum(gs < 0)) == want_cnt, (ls, want_cnt)
        assert 0 < want_cnt < n or ls == ocpu.LS_STRAUB_CAPACITY, (ls, want_cnt)      # (a capacity is positive)
        gf, gpf = host(ls, k, X, strict=False)
        fin = np.isfinite(want_g)
        assert np.array_equal(np.isnan(gf), np.isnan(want_g)), ls
        assert np.all(np.abs(gf[fin] - want_g[fin]) <= 1e-9 * np.maximum(1.0, np.abs(want_g[fin]))), ls
        if ls in (ocpu.LS_POLYNOMIAL, ocpu.LS_MIN_AFFINE):      # both samples of a pair in one pass == two passes
            assert np.array_equal(gps, gs, equal_nan=True) and np.array_equal(gpf, gf, equal_nan=True), ls
