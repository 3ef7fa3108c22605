"""The hand-written vehicle functor (csrc/ebn_limit_states.cuh: VehicleSuspension) — the code the headline bench
times — compiled for the HOST from its own source text behind a few lines of shim, against the C oracle
(oracle/ebn_oracle.c, the restatement of demo/vehicle_suspension.jl:17-25) on adversarial rows. No GPU:

  * eval<StrictOps> equals the oracle bit for bit (Julia's evaluation order, no contraction),
  * eval<FastOps> stays within rounding of it,
  * fails() — the division-free, square-root-free sign test of fast mode — equals `g < 0` with NaN as safe on every
    row that is not a tie, for both signs of K and C, for negative square-root arguments, and on near-degenerate rows.

The GPU suite checks the same predicate inside the fused kernel (tests/test_gpu_adversarial.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

from ebn_b200 import workloads as W
from oracle import cpu as ocpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = r"""
#include <math.h>
#include <string.h>
#define __device__
#define __forceinline__ inline
#define EBN_LS_VEHICLE_SUSPENSION 1
#define EBN_BLOCKS_FOR_WARPS(w) 1
namespace ebn {
struct FastMathSmem {};
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline int __double2hiint(double d) { long long v; memcpy(&v, &d, 8); return (int)(v >> 32); }
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
}
extern "C" void run(const double* k, const double* X, long n, double* g_strict, double* g_fast, unsigned char* f) {
  ebn::FastMathSmem sm;
  ebn::VehicleSuspension<ebn::StrictOps>::Ctx cs;
  ebn::VehicleSuspension<ebn::FastOps>::Ctx cf;
  ebn::VehicleSuspension<ebn::StrictOps>::setup(cs, k, 5, 6, &sm);
  ebn::VehicleSuspension<ebn::FastOps>::setup(cf, k, 5, 6, &sm);
  for (long i = 0; i < n; ++i) {
    g_strict[i] = ebn::VehicleSuspension<ebn::StrictOps>::eval(cs, X + 6 * i);
    g_fast[i] = ebn::VehicleSuspension<ebn::FastOps>::eval(cf, X + 6 * i);
    f[i] = ebn::VehicleSuspension<ebn::FastOps>::fails(cf, X + 6 * i) ? 1 : 0;
  }
}
"""


def _host_vehicle(tmp_path):
    text = open(os.path.join(ROOT, "enhancedbayesiannetworks.jl_b200", "csrc", "ebn_limit_states.cuh")).read()
    a = text.index("template <class Ops>\nstruct VehicleSuspension {")
    b = text.index("// ---------------------------------------------------------------------------\n// Straub frame")
    src = tmp_path / "vehicle.cpp"
    src.write_text(SHIM + text[a:b] + DRIVER)
    so = tmp_path / "vehicle.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", str(src), "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    dp = C.POINTER(C.c_double)
    lib.run.argtypes = [dp, dp, C.c_long, dp, dp, C.POINTER(C.c_ubyte)]

    def run(X):
        X = np.ascontiguousarray(X, dtype=np.float64)
        k = np.ascontiguousarray(W.vehicle_consts(), dtype=np.float64)
        gs, gf, f = np.empty(len(X)), np.empty(len(X)), np.empty(len(X), dtype=np.uint8)
        lib.run(k.ctypes.data_as(dp), X.ctypes.data_as(dp), len(X), gs.ctypes.data_as(dp), gf.ctypes.data_as(dp),
                f.ctypes.data_as(C.POINTER(C.c_ubyte)))
        return gs, gf, f.astype(bool)
    return run


def _rows(rng, n, c, ck, k):
    return np.column_stack([rng.choice([0.15915, 0.8], n), rng.choice([0.27, 0.5], n), rng.uniform(7, 12, n),
                            rng.normal(*c, n), rng.normal(*ck, n), rng.normal(*k, n)])


def test_vehicle_functor_on_the_host_against_the_oracle(tmp_path):
    run = _host_vehicle(tmp_path)
    rng = np.random.default_rng(20261020)
    k = W.vehicle_consts()
    blocks = [
        _rows(rng, 400_000, (431.7221, 10.0), (1475.5503, 10.0), (55.0406, 10.0)),        # the demo's marginals
        _rows(rng, 600_000, (0.0, 300.0), (1442.64, 300.0), (0.0, 40.0)),                 # both signs of C and K, g4 threshold
        _rows(rng, 100_000, (391.0, 1e-3), (k[4], 1e-9), (55.0, 1e-6)),                    # sitting on the g2 / g4 thresholds
        _rows(rng, 100_000, (-5.0, 5.0), (1500.0, 1.0), (1e-6, 1e-6)),                     # tiny K of both signs
        _rows(rng, 100_000, (1e4, 1e4), (-1442.0, 3000.0), (-300.0, 500.0)),               # huge spreads
    ]
    for i, X in enumerate(blocks):
        want_cnt, want_g = ocpu.eval_matrix(ocpu.LS_VEHICLE, k, X, want_g=True)
        gs, gf, f = run(X)
        assert np.array_equal(gs, want_g, equal_nan=True), i                              # strict: bit for bit
        fin = np.isfinite(want_g)
        assert np.array_equal(np.isnan(gf), np.isnan(want_g)), i
        assert np.all(np.abs(gf[fin] - want_g[fin]) <= 1e-9 * np.maximum(1.0, np.abs(want_g[fin]))), i
        differ = f != (want_g < 0)
        ties = np.abs(want_g) < 1e-9
        assert not differ[~ties].any(), (i, X[differ & ~ties][:3], want_g[differ & ~ties][:3])
        if i == 1:
            assert np.isnan(want_g).mean() > 0.1 and 0.2 < f.mean() < 0.8
            # failures with K of either sign (D = b0 K changes sign under the division-free g1 test); a negative
            # C puts a negative number under g3's square root: NaN, safe, whatever the other mechanisms say
            assert (X[f, 5] < 0).any() and (X[f, 5] > 0).any() and not f[X[:, 3] < 0].any()
