"""The functor the library generates for an expression program (ebn_expression_source), compiled for the HOST with
g++ behind a few lines of shim, and run on random rows — no GPU, no NVRTC:

  * eval() in strict mode equals the oracle's stack machine bit for bit (the generated code IS the program),
  * fails(), the sign-only predicate derived for fast mode (ebn_expr.cu: FailsGen — min -> or, square roots against
    constants -> comparisons of the radicand, quotients -> signs of numerator and denominator), equals `eval() < 0`
    on every row that is not a tie, NaN rows (negative radicands, logarithms of negatives) included,
  * and the predicate's own statements hold no division and no square root.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from ebn_b200 import _lib as L
from ebn_b200 import expr, workloads as W
from oracle import cpu as ocpu

import formula_fuzz

SHIM = r"""
#include <math.h>
#include <string.h>
#define __device__
#define __forceinline__ inline
#define EBN_LS_EXPRESSION 8
#define EBN_BLOCKS_FOR_WARPS(w) 1
namespace ebn {
struct FastMathSmem {};
inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
inline double exp_guarded(double x, const FastMathSmem&) { return exp(x); }
inline double log_guarded(double x, const FastMathSmem&) { return log(x); }
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
}
"""
DRIVER = r"""
template <class F>
static void run_t(const double* consts, int nc, const double* X, long n, int d, double* g, unsigned char* f) {
  ebn::FastMathSmem sm;
  typename F::Ctx c;
  F::setup(c, consts, nc, d, &sm);
  for (long i = 0; i < n; ++i) {
    double x[F::XCAP + 1], y[F::XCAP + 1];
    for (int j = 0; j < d; ++j) x[j] = y[j] = X[i * d + j];
    g[i] = F::eval(c, x);
    f[i] = F::fails(c, y) ? 1 : 0;
  }
}
extern "C" void run(int strict, const double* consts, int nc, const double* X, long n, int d, double* g, unsigned char* f) {
  if (strict) run_t<ebn::JitExpr<ebn::StrictOps> >(consts, nc, X, n, d, g, f);
  else run_t<ebn::JitExpr<ebn::FastOps> >(consts, nc, X, n, d, g, f);
}
"""


def _host_functor(tmp_path, prog, d, tag):
    src = L.expression_source(prog.consts, d)
    cu = tmp_path / f"f{tag}.cpp"
    cu.write_text(SHIM + src + DRIVER)
    so = tmp_path / f"f{tag}.so"
    # no contraction: strict mode must be one IEEE operation per instruction, as __dmul_rn & co. are on the device
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", str(cu), "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    dp = C.POINTER(C.c_double)
    lib.run.argtypes = [C.c_int, dp, C.c_int, dp, C.c_long, C.c_int, dp, C.POINTER(C.c_ubyte)]

    def run(X, strict):
        X = np.ascontiguousarray(X, dtype=np.float64)
        g = np.empty(len(X))
        f = np.empty(len(X), dtype=np.uint8)
        k = np.ascontiguousarray(prog.consts, dtype=np.float64)
        lib.run(int(strict), k.ctypes.data_as(dp), k.size, X.ctypes.data_as(dp), len(X), d, g.ctypes.data_as(dp),
                f.ctypes.data_as(C.POINTER(C.c_ubyte)))
        return g, f.astype(bool)
    return src, run


def _predicate_statements(src):
    body = src[src.index("static bool fails"):]
    return [ln for ln in body.splitlines() if ln.strip().startswith(("int p", "const double r"))]


def test_vehicle_text_predicate_has_no_division_and_agrees_with_g(tmp_path):
    k = W.vehicle_consts()
    models = [("g1", "1 .- (pi .* m .* df.V .* df.A) ./ (df.b0 .* df.K .* g .^ 2) .* ((df.Ck ./ (m .+ M) .- (df.C ./ M)) .^ 2"
                     " .+ df.C .^ 2 ./ (m .* M) .+ df.Ck .* df.K .^ 2 ./ (m .* M .^ 2))"),
              ("g2", "4000 .* df.C .* Mg15 .- 8.6394"),
              ("g3", "2 .* sqrt(M .* g .* (df.K .^ 2 .* df.Ck ./ (df.C .* (m .+ M)) .+ df.C)) .- 1"),
              ("g4", "df.Ck .- gMm")]
    prog = expr.compile_program(["A", "b0", "V", "C", "Ck", "K"], models, "minimum([g1[1], g2, g3, g4[1]])",
                                params={"M": k[0], "m": k[1], "g": k[2], "Mg15": k[3], "gMm": k[4]})
    src, run = _host_functor(tmp_path, prog, 6, "veh")
    assert "HAS_FAILS = !Ops::strict" in src
    stm = _predicate_statements(src)
    assert len(stm) > 10 and not any("/" in s or "sqrt" in s or "Ops::div" in s for s in stm), stm
    rng = np.random.default_rng(5)
    n = 200_000
    X = np.column_stack([rng.choice([0.15915, 0.8], n), rng.choice([0.27, 0.5], n), rng.uniform(7, 12, n),
                         rng.normal(0, 300, n), rng.normal(1442.64, 300, n), rng.normal(0, 40, n)])   # both signs of C, K
    want_cnt, want_g = ocpu.eval_matrix(ocpu.LS_EXPRESSION, prog.consts, X, want_g=True)
    g, f = run(X, strict=True)
    assert np.array_equal(g, want_g, equal_nan=True)                   # strict eval() == the oracle's stack machine
    gf, ff = run(X, strict=False)
    assert np.isnan(want_g).mean() > 0.1 and 0.1 < (want_g < 0).mean() < 0.9
    differ = ff != (want_g < 0)
    assert not differ[~(np.abs(want_g) < 1e-9)].any(), X[differ][:3]


def test_generated_predicates_against_the_value_on_random_chains(tmp_path):
    rng = np.random.default_rng(4242)
    names = ["x0", "x1", "x2", "x3"]
    n = 20_000
    X = np.column_stack([rng.normal(0.3, 1.5, n), rng.normal(-0.2, 1.0, n), rng.uniform(-1.0, 2.0, n), rng.normal(1.0, 0.7, n)])
    nan_rows = fails = rewritten = 0
    for trial in range(30):
        models, perf, consts = formula_fuzz.chain(rng, names)
        prog = expr.compile_program(names, models, perf, consts)
        src, run = _host_functor(tmp_path, prog, 4, str(trial))
        _, want_g = ocpu.eval_matrix(ocpu.LS_EXPRESSION, prog.consts, X, want_g=True)
        g, _ = run(X, strict=True)
        assert np.array_equal(g, want_g, equal_nan=True), (trial, models, perf)
        _, ff = run(X, strict=False)
        differ = ff != (want_g < 0)
        if differ.any():   # only rows whose sign is a matter of rounding may differ (ties, cancellation)
            oracle_g = lambda Y: ocpu.eval_matrix(ocpu.LS_EXPRESSION, prog.consts, Y, want_g=True)[1]  # noqa: E731
            ok = ~formula_fuzz.unstable_rows(oracle_g, X, want_g)
            assert not differ[ok].any(), (trial, models, perf, X[differ & ok][:3], want_g[differ & ok][:3])
        nan_rows += int(np.isnan(want_g).sum())
        fails += int((want_g < 0).sum())
        rewritten += any(s.strip().startswith("const double r") for s in _predicate_statements(src))   # fractions carried
    assert nan_rows > 0.05 * 30 * n and 0.05 < fails / (30 * n) < 0.95 and rewritten >= 10


def test_zero_and_infinite_divisors_take_the_value(tmp_path):
    """A Parameter state of exactly 0 under a division (a switch variable, an absent area), a formula that cancels
    itself, a literal -0.0: IEEE arithmetic answers with infinities, signed zeros and NaNs, which the sign rules of
    the predicate cannot see. Samples with such a divisor run eval()'s statements (the cold block of fails()): the
    predicate must equal `eval() < 0` EXACTLY there."""
    names = ["a", "b", "c"]
    cases = ["1 - a / (b * c)", "a / b - 1", "2 * sqrt(a / (b / -0.0)) - 0.3", "minimum([a / b, c / (b - b), 1.0]) - 0.5",
             "1 - (a / b) * (c / (a - a))", "0.5 - 1.5 * sqrt((a / b) + c)", "max(a / b, c / b) - 2", "(a / b)^2 - 4",
             "-(a / (b * 0.0))", "1 - (c / (a / b)) * a"]
    vals = [0.0, -0.0, 1.0, -2.0, 0.5, 3.0]
    X = np.array([[x, y, z] for x in vals for y in vals for z in vals])
    seen_inf = seen_nan = 0
    for i, text in enumerate(cases):
        prog = expr.compile_program(names, [], text, {})
        src, run = _host_functor(tmp_path, prog, 3, f"z{i}")
        _, want_g = ocpu.eval_matrix(ocpu.LS_EXPRESSION, prog.consts, X, want_g=True)
        gs, _ = run(X, strict=True)
        assert np.array_equal(gs, want_g, equal_nan=True), text
        gf, ff = run(X, strict=False)
        assert np.array_equal(ff, gf < 0), (text, X[ff != (gf < 0)][:4])          # the predicate IS the fast value's sign here
        assert np.array_equal(ff, want_g < 0), (text, X[ff != (want_g < 0)][:4])  # and the strict one's (exact inputs)
        seen_inf += int(np.isinf(want_g).sum())
        seen_nan += int(np.isnan(want_g).sum())
    assert seen_inf > 100 and seen_nan > 100
