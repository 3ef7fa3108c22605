// Limit states compiled into the kernels as device functors.
//
// Every functor has two arithmetic policies:
//   StrictOps — one IEEE-754 operation per Julia operation, in Julia's
//               evaluation order, no FMA contraction (__dmul_rn & co. are never
//               fused by nvcc). This is what ebn_eval_matrix(strict=1) uses for
//               bit-exact fail counts on matrices exported from the reference.
//   FastOps   — same mathematics, re-associated: reciprocals of constants
//               hoisted, FMA contraction, shared subexpressions. |g| differs
//               from strict by a few ulp; count differences can only come from
//               samples with |g| ~ 1e-13 (reported as ties by the tests).
#pragma once
#include "ebn_rtc_compat.h"
#include "../../include/ebncuda.h"
#include "ebn_fastmath.cuh"

namespace ebn {

struct StrictOps {
  static constexpr bool strict = true;
  __device__ __forceinline__ static double add(double a, double b) { return __dadd_rn(a, b); }
  __device__ __forceinline__ static double sub(double a, double b) { return __dsub_rn(a, b); }
  __device__ __forceinline__ static double mul(double a, double b) { return __dmul_rn(a, b); }
  __device__ __forceinline__ static double div(double a, double b) { return __ddiv_rn(a, b); }
  __device__ __forceinline__ static double sqrt(double a) { return __dsqrt_rn(a); }
  // Julia's min/minimum propagate NaN.
  __device__ __forceinline__ static double min(double a, double b) {
    return (b < a || b != b) ? b : a;
  }
  __device__ __forceinline__ static double max(double a, double b) {
    return (b > a || b != b) ? b : a;
  }
};

struct FastOps {
  static constexpr bool strict = false;
  __device__ __forceinline__ static double add(double a, double b) { return a + b; }
  __device__ __forceinline__ static double sub(double a, double b) { return a - b; }
  __device__ __forceinline__ static double mul(double a, double b) { return a * b; }
  __device__ __forceinline__ static double div(double a, double b) { return a / b; }
  __device__ __forceinline__ static double sqrt(double a) { return ::sqrt(a); }
  __device__ __forceinline__ static double min(double a, double b) {
    return (b < a || b != b) ? b : a;
  }
  __device__ __forceinline__ static double max(double a, double b) {
    return (b > a || b != b) ? b : a;
  }
};

// ---------------------------------------------------------------------------
// Vehicle suspension — reference demo/vehicle_suspension.jl:17-25
// (same body at demo/vehicle_suspension_imprecise.jl:43-51).
// x = (A, b0, V, C, Ck, K); consts = (M, m, g, (M*g)^-1.5, (g*(M+m))^0.877).
// ---------------------------------------------------------------------------
template <class Ops>
struct VehicleSuspension {
  static constexpr int ID = EBN_LS_VEHICLE_SUSPENSION;
  static constexpr int D = 6;    // fixed arity
  static constexpr int XCAP = 6; // columns held per sample
  static constexpr int NC = 5;
  static constexpr bool HAS_FAILS = !Ops::strict;  // sign-only test in fast mode
#ifndef EBN_VEHICLE_BLOCKS
#define EBN_VEHICLE_BLOCKS EBN_BLOCKS_FOR_WARPS(8)
#endif
  // A launch bound of 2 blocks per SM caps the kernel at 255 registers per thread; ptxas then uses ~124 and
  // four blocks are resident anyway (the host sizes the grid from the occupancy query). The cap matters for
  // the schedule ptxas picks, not for occupancy: bound 4 (128 registers) 1.565e11, bound 3 1.674e11,
  // bound 2 1.692e11 samples/s; one resident warp per sub-partition already gives 88 % of that
  // (profiles/r02_tune_w1.txt, r02_occupancy_sweep.txt).
  // strict mode (IEEE divisions and a square root per sample) is latency-bound instead: 16 warps per SM
  static constexpr int MIN_BLOCKS = Ops::strict ? EBN_BLOCKS_FOR_WARPS(16) : EBN_VEHICLE_BLOCKS;
  struct Ctx {
    double M, m, gg, pow_a, pow_b;     // strict
    double pim, mM, iM, imM, immM, mpM, Mg;
    double k1, k2, k3, k4;             // fast
    double thr2, c3;                   // fast, sign-only
  };
  __device__ static void setup(Ctx& c, const double* k, int, int, const FastMathSmem*) {
    const double M = k[0], m = k[1], g = k[2];
    c.M = M;
    c.m = m;
    c.gg = __dmul_rn(g, g);              // g .^ 2 (g is an Int in the demo: exact)
    c.pow_a = k[3];
    c.pow_b = k[4];
    c.pim = __dmul_rn(3.141592653589793, m);  // Float64(pi) * m
    c.mpM = __dadd_rn(m, M);
    c.mM = __dmul_rn(m, M);
    c.immM = __dmul_rn(m, __dmul_rn(M, M));   // m * M^2
    c.Mg = __dmul_rn(M, g);
    if (!Ops::strict) {
      c.iM = 1.0 / M;
      c.imM = 1.0 / c.mM;
      c.k1 = 1.0 / c.mpM;
      c.k2 = 1.0 / c.immM;
      c.k3 = 4000.0 * k[3];
      c.k4 = c.pim / c.gg;
      c.thr2 = 8.6394 / c.k3;            // g2 < 0  <=>  C < thr2
      c.c3 = 0.25 * c.mpM / c.Mg;        // g3 < 0  <=>  K^2 Ck + C^2 (m+M) < c3 * C   (C > 0)
    }
  }
  __device__ __forceinline__ static double eval(const Ctx& c, const double* x) {
    const double A = x[0], b0 = x[1], V = x[2], C = x[3], Ck = x[4], K = x[5];
    if (Ops::strict) {
      // g1 = 1 - (pi*m*V*A)/(b0*K*g^2) * [(Ck/(m+M) - C/M)^2 + C^2/(m*M) + Ck*K^2/(m*M^2)]
      const double num = Ops::mul(Ops::mul(c.pim, V), A);
      const double den = Ops::mul(Ops::mul(b0, K), c.gg);
      const double e = Ops::sub(Ops::div(Ck, c.mpM), Ops::div(C, c.M));
      const double t1 = Ops::mul(e, e);
      const double t2 = Ops::div(Ops::mul(C, C), c.mM);
      const double t3 = Ops::div(Ops::mul(Ck, Ops::mul(K, K)), c.immM);
      const double br = Ops::add(Ops::add(t1, t2), t3);
      const double g1 = Ops::sub(1.0, Ops::mul(Ops::div(num, den), br));
      // g2 = 4000*C*(M*g)^-1.5 - 8.6394
      const double g2 = Ops::sub(Ops::mul(Ops::mul(4000.0, C), c.pow_a), 8.6394);
      // g3 = 2*sqrt(M*g*(K^2*Ck/(C*(m+M)) + C)) - 1
      const double in3 =
          Ops::add(Ops::div(Ops::mul(Ops::mul(K, K), Ck), Ops::mul(C, c.mpM)), C);
      const double g3 = Ops::sub(Ops::mul(2.0, Ops::sqrt(Ops::mul(c.Mg, in3))), 1.0);
      // g4 = Ck - (g*(M+m))^0.877
      const double g4 = Ops::sub(Ck, c.pow_b);
      return Ops::min(Ops::min(Ops::min(g1, g2), g3), g4);
    } else {
      // one reciprocal serves both divisions by a random variable: 1/(K*C)
      const double rKC = 1.0 / (K * C);
      const double iK = C * rKC, iC = K * rKC;
      const double K2Ck = K * K * Ck;
      const double e = Ck * c.k1 - C * c.iM;
      const double br = e * e + C * C * c.imM + K2Ck * c.k2;
      const double g1 = 1.0 - (c.k4 * A) * (V * iK) * br / b0;
      const double g2 = C * c.k3 - 8.6394;
      const double g3 = 2.0 * ::sqrt(c.Mg * (K2Ck * iC * c.k1 + C)) - 1.0;
      const double g4 = Ck - c.pow_b;
      return Ops::min(Ops::min(Ops::min(g1, g2), g3), g4);
    }
  }
  // Sign of g without its value (count mode, fast arithmetic): no division, no sqrt.
  //   g1 < 0 <=> q/D > 1 with q = (pi m/g^2) A V [..], D = b0 K      (sign of D respected)
  //   g2 < 0 <=> C < 8.6394/(4000 (Mg)^-1.5)
  //   g3 < 0 <=> M g (K^2 Ck/(C (m+M)) + C) < 1/4                   (sign of C respected)
  //   g4 < 0 <=> Ck < (g (M+m))^0.877
  // A negative sqrt argument makes g NaN in the reference arithmetic, and NaN is safe.
  __device__ __forceinline__ static bool fails(const Ctx& c, const double* x) {
    const double A = x[0], b0 = x[1], V = x[2], C = x[3], Ck = x[4], K = x[5];
    const double e = Ck * c.k1 - C * c.iM;
    const double K2Ck = K * K * Ck;
    const double CC = C * C;
    const double br = fma(e, e, fma(CC, c.imM, K2Ck * c.k2));
    const double q = (c.k4 * A) * V * br;
    const double D = b0 * K;
    const double lhs3 = fma(CC, c.mpM, K2Ck);
    // The comparisons are taken from the sign bits of differences with integer logic:
    // DSETP occupies the fp64 pipe about twice as long as a DADD (tools/microbench).
    //   f1: q/D > 1      <=> sign(D - q) != sign(D)
    //   f2: C < thr2     <=> sign(C - thr2)
    //   f3: lhs3*C < c3*C*C ... <=> sign(lhs3 - c3*C) != sign(C)
    //   f4: Ck < pow_b   <=> sign(Ck - pow_b)
    //   nan3 (negative sqrt argument) <=> sign(lhs3) != sign(C)
    const int sC = __double2hiint(C);
    const int t1 = __double2hiint(D - q) ^ __double2hiint(D);
    const int t2 = __double2hiint(C - c.thr2);
    const int t3 = __double2hiint(fma(-c.c3, C, lhs3)) ^ sC;
    const int t4 = __double2hiint(Ck - c.pow_b);
    const int n3 = __double2hiint(lhs3) ^ sC;
    return (((t1 | t2 | t3 | t4) & ~n3) < 0);
  }
};

// ---------------------------------------------------------------------------
// Straub frame — reference demo/straub_example.jl:16-39 (== test/inference/
// variableselimination.jl:148-171) after transmission.jl:9-29 fused R1..R5
// into the frame node. x = (Ur, V, H, e1..e5); consts = (lambda, zeta, rho).
// ---------------------------------------------------------------------------
template <class Ops>
struct StraubFrame {
  static constexpr int ID = EBN_LS_STRAUB_FRAME;
  static constexpr bool HAS_FAILS = false;
// Resident warps per SM the variadic limit states (polynomial program, min of affine forms) are compiled
// for: 65536 / (32 warps) registers per thread.
#ifndef EBN_GENERIC_WARPS
#define EBN_GENERIC_WARPS 24
#endif
#ifndef EBN_STRAUB_BLOCKS
#define EBN_STRAUB_BLOCKS EBN_BLOCKS_FOR_WARPS(12)  // register cap 168; measured (profiles/r02_tune_w2.txt): bound 2 4.30e10, 3 4.35e10, 4 4.16e10 samples/s
#endif
  static constexpr int MIN_BLOCKS = EBN_STRAUB_BLOCKS;
  static constexpr int D = 8;
  static constexpr int XCAP = 8;
  static constexpr int NC = 3;
  struct Ctx {
    double lambda, rz, sd;
    const FastMathSmem* fm;
  };
  __device__ static void setup(Ctx& c, const double* k, int, int, const FastMathSmem* fm) {
    c.fm = fm;
    const double lambda = k[0], zeta = k[1], rho = k[2];
    c.lambda = lambda;
    c.rz = __dmul_rn(rho, zeta);  // rho * zeta (then * u_r)
    // sqrt((1 - rho^2) * zeta^2)
    c.sd = __dsqrt_rn(__dmul_rn(__dsub_rn(1.0, __dmul_rn(rho, rho)), __dmul_rn(zeta, zeta)));
  }
  __device__ __forceinline__ static double eval(const Ctx& c, const double* x) {
    const double ur = x[0], v = x[1], h = x[2];
    const double mu = Ops::add(c.lambda, Ops::mul(c.rz, ur));
    double r[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const double a = Ops::add(mu, Ops::mul(c.sd, x[3 + i]));
      r[i] = Ops::strict ? exp(a) : exp_fast(a, *c.fm);   // strict: libdevice exp (< 1 ulp, like Julia's)
    }
    const double h5 = Ops::mul(5.0, h), v5 = Ops::mul(5.0, v);
    const double g1 = Ops::sub(Ops::add(Ops::add(Ops::add(r[0], r[1]), r[3]), r[4]), h5);
    const double g2 = Ops::sub(Ops::add(Ops::add(r[1], Ops::mul(2.0, r[2])), r[3]), v5);
    const double g3 = Ops::sub(
        Ops::sub(Ops::add(Ops::add(Ops::add(r[0], Ops::mul(2.0, r[2])), Ops::mul(2.0, r[3])), r[4]),
                 h5),
        v5);
    return Ops::min(Ops::min(g1, g2), g3);
  }
};

// ---------------------------------------------------------------------------
// Straub frame with R4, R5 as observed columns — test/inference/variableselimination.jl:207-217
// (`frame_model.(df.r1, df.r2, df.r3, df.R4, df.R5, df.V, df.H)`).
// x = (Ur, V, H, R4, R5, e1, e2, e3); consts = (lambda, zeta, rho).
// ---------------------------------------------------------------------------
template <class Ops>
struct StraubFrameR45 {
  static constexpr int ID = EBN_LS_STRAUB_FRAME_R45;
  static constexpr bool HAS_FAILS = false;
  static constexpr int MIN_BLOCKS = EBN_STRAUB_BLOCKS;
  static constexpr int D = 8;
  static constexpr int XCAP = 8;
  static constexpr int NC = 3;
  using Ctx = typename StraubFrame<Ops>::Ctx;
  __device__ static void setup(Ctx& c, const double* k, int a, int b, const FastMathSmem* fm) {
    StraubFrame<Ops>::setup(c, k, a, b, fm);
  }
  __device__ __forceinline__ static double eval(const Ctx& c, const double* x) {
    const double ur = x[0], v = x[1], h = x[2];
    const double mu = Ops::add(c.lambda, Ops::mul(c.rz, ur));
    double r[5];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double a = Ops::add(mu, Ops::mul(c.sd, x[5 + i]));
      r[i] = Ops::strict ? exp(a) : exp_fast(a, *c.fm);
    }
    r[3] = x[3];
    r[4] = x[4];
    const double h5 = Ops::mul(5.0, h), v5 = Ops::mul(5.0, v);
    const double g1 = Ops::sub(Ops::add(Ops::add(Ops::add(r[0], r[1]), r[3]), r[4]), h5);
    const double g2 = Ops::sub(Ops::add(Ops::add(r[1], Ops::mul(2.0, r[2])), r[3]), v5);
    const double g3 = Ops::sub(
        Ops::sub(Ops::add(Ops::add(Ops::add(r[0], Ops::mul(2.0, r[2])), Ops::mul(2.0, r[3])), r[4]),
                 h5),
        v5);
    return Ops::min(Ops::min(g1, g2), g3);
  }
};

// ---------------------------------------------------------------------------
// One plastic moment capacity — demo/straub_example.jl:16-26 (`plastic_moment_capacities`).
// x = (Ur, e); consts = (lambda, zeta, rho); value r.
// ---------------------------------------------------------------------------
template <class Ops>
struct StraubCapacity {
  static constexpr int ID = EBN_LS_STRAUB_CAPACITY;
  static constexpr bool HAS_FAILS = false;
  static constexpr int MIN_BLOCKS = EBN_BLOCKS_FOR_WARPS(24);
  static constexpr int D = 2;
  static constexpr int XCAP = 2;
  static constexpr int NC = 3;
  using Ctx = typename StraubFrame<Ops>::Ctx;
  __device__ static void setup(Ctx& c, const double* k, int a, int b, const FastMathSmem* fm) {
    StraubFrame<Ops>::setup(c, k, a, b, fm);
  }
  __device__ __forceinline__ static double eval(const Ctx& c, const double* x) {
    const double a = Ops::add(Ops::add(c.lambda, Ops::mul(c.rz, x[0])), Ops::mul(c.sd, x[1]));
    return Ops::strict ? exp(a) : exp_fast(a, *c.fm);
  }
};

// ---------------------------------------------------------------------------
// Hydrogen — reference demo/Hydrogen_new/ebn.jl:98-111 (`hyram_performance`).
// x = (scenario, r_operators, burn_r_C_3th, explosion_r_C). The two radii come
// from HyRAM, which is not in the reference repository; here they are inputs.
// ---------------------------------------------------------------------------
template <class Ops>
struct HydrogenRadius {
  static constexpr int ID = EBN_LS_HYDROGEN_RADIUS;
  static constexpr bool HAS_FAILS = false;
  static constexpr int MIN_BLOCKS = EBN_BLOCKS_FOR_WARPS(24);
  static constexpr int D = 4;
  static constexpr int XCAP = 4;
  static constexpr int NC = 0;
  struct Ctx {};
  __device__ static void setup(Ctx&, const double*, int, int, const FastMathSmem*) {}
  __device__ __forceinline__ static double eval(const Ctx&, const double* x) {
    const double sc = x[0];
    if (sc == 1.0) return Ops::sub(x[1], x[2]);
    if (sc == 2.0) return Ops::sub(x[1], x[3]);
    return 1.0;
  }
};

// ---------------------------------------------------------------------------
// Polynomial program — the toy models of test/networks/ebn/evaluate/
// evaluate_node.jl:96-98, evaluate_net.jl:18-33, discretize.jl:275-277.
// Encoding documented in include/ebncuda.h. The program is warp-uniform, the
// column file x[] is per thread.
// ---------------------------------------------------------------------------
template <class Ops>
struct PolynomialProgram {
  static constexpr int ID = EBN_LS_POLYNOMIAL;
  static constexpr bool HAS_FAILS = false;
  static constexpr int MIN_BLOCKS = EBN_BLOCKS_FOR_WARPS(EBN_GENERIC_WARPS);
  static constexpr int D = -1;  // variadic
  static constexpr int XCAP = EBN_MAX_INPUTS;
  static constexpr int NC = -1;
  struct Ctx {
    const double* prog;
    int d;
  };
  __device__ static void setup(Ctx& c, const double* k, int, int d, const FastMathSmem*) {
    c.prog = k;
    c.d = d;
  }
  // y_out (optional) receives the value of the last MODEL stage (the column
  // a ContinuousFunctionalNode keeps); the return value is the last stage.
  __device__ __forceinline__ static double eval(const Ctx& c, double* x) {
    const double* p = c.prog;
    const int T = (int)p[0];
    int pos = 1, col = c.d;
    double acc = 0.0;
    for (int t = 0; t < T; ++t) {
      const int nterms = (int)p[pos++];
      acc = 0.0;
      for (int k = 0; k < nterms; ++k) {
        double term = p[pos++];
        const int nf = (int)p[pos++];
        for (int f = 0; f < nf; ++f) term = Ops::mul(term, x[(int)p[pos++]]);
        acc = (k == 0) ? term : Ops::add(acc, term);
      }
      if (col < XCAP) x[col] = acc;
      ++col;
    }
    return acc;
  }
  // Both samples of a pair in ONE pass over the (warp-uniform) program: the decoding — loads, integer
  // conversions, loop control — is paid once, every sample still sees the operations of eval() in order.
  static constexpr bool HAS_EVAL_PAIR = true;
  __device__ __forceinline__ static void eval_pair(const Ctx& c, double* x0, double* x1, double& g0, double& g1) {
    const double* p = c.prog;
    const int T = (int)__ldg(p);
    int pos = 1, col = c.d;
    double a0 = 0.0, a1 = 0.0;
    for (int t = 0; t < T; ++t) {
      const int nterms = (int)__ldg(p + pos++);
      a0 = 0.0;
      a1 = 0.0;
      for (int k = 0; k < nterms; ++k) {
        const double coef = __ldg(p + pos);
        const int nf = (int)__ldg(p + pos + 1);
        pos += 2;
        double t0 = coef, t1 = coef;
        for (int f = 0; f < nf; ++f) {
          const int j = (int)__ldg(p + pos++);
          t0 = Ops::mul(t0, x0[j]);
          t1 = Ops::mul(t1, x1[j]);
        }
        a0 = (k == 0) ? t0 : Ops::add(a0, t0);
        a1 = (k == 0) ? t1 : Ops::add(a1, t1);
      }
      if (col < XCAP) {
        x0[col] = a0;
        x1[col] = a1;
      }
      ++col;
    }
    g0 = a0;
    g1 = a1;
  }
};

// ---------------------------------------------------------------------------
// min over affine forms: g = min_k (c_k0 + sum_j c_kj x_j).
// consts = (K, K rows of (1+d)).
// ---------------------------------------------------------------------------
template <class Ops>
struct MinAffine {
  static constexpr int ID = EBN_LS_MIN_AFFINE;
  static constexpr bool HAS_FAILS = false;
  static constexpr int MIN_BLOCKS = EBN_BLOCKS_FOR_WARPS(EBN_GENERIC_WARPS);
  static constexpr int D = -1;
  static constexpr int XCAP = EBN_MAX_INPUTS;
  static constexpr int NC = -1;
  struct Ctx {
    const double* k;
    int d;
  };
  __device__ static void setup(Ctx& c, const double* k, int, int d, const FastMathSmem*) {
    c.k = k;
    c.d = d;
  }
  __device__ __forceinline__ static double eval(const Ctx& c, double* x) {
    const int K = (int)c.k[0];
    double g = 0.0;
    for (int r = 0; r < K; ++r) {
      const double* row = c.k + 1 + r * (1 + c.d);
      double acc = row[0];
      for (int j = 0; j < c.d; ++j) acc = Ops::add(acc, Ops::mul(row[1 + j], x[j]));
      g = (r == 0) ? acc : Ops::min(g, acc);
    }
    return g;
  }
  static constexpr bool HAS_EVAL_PAIR = true;   // see PolynomialProgram::eval_pair
  __device__ __forceinline__ static void eval_pair(const Ctx& c, double* x0, double* x1, double& g0, double& g1) {
    const int K = (int)__ldg(c.k);
    double m0 = 0.0, m1 = 0.0;
    for (int r = 0; r < K; ++r) {
      const double* row = c.k + 1 + r * (1 + c.d);
      const double b = __ldg(row);
      double a0 = b, a1 = b;
      for (int j = 0; j < c.d; ++j) {
        const double w = __ldg(row + 1 + j);
        a0 = Ops::add(a0, Ops::mul(w, x0[j]));
        a1 = Ops::add(a1, Ops::mul(w, x1[j]));
      }
      m0 = (r == 0) ? a0 : Ops::min(m0, a0);
      m1 = (r == 0) ? a1 : Ops::min(m1, a1);
    }
    g0 = m0;
    g1 = m1;
  }
};

}  // namespace ebn
