// In-register marginal transforms over a packed Philox word stream.
// Replaces `rand(dist, n)` per input inside UQ.jl's `sample(inputs, sim)`
// (reference call sites src/networks/ebn/evaluation/evaluate_node.jl:24,83;
// the marginals that reach it are listed in SURVEY.md F7 / rows 2 and 5).
//
// Random streams, layout v4: for scenario stream s and sample pair
// q = sample index >> 1 the word stream is
//     W[4c .. 4c+3] = Philox4x32-10(counter = (q lo, q hi, s, c), key = seed)
// and the inputs consume words in input order:
//     Parameter                         0 words
//     Uniform                           2 words  (one 32-bit uniform per sample)
//     untruncated Normal / LogNormal    2 words  (Box-Muller across the pair:
//                                       42-bit radius, 22-bit angle on the full circle; box_muller_front)
// Bit placement: U52(lo, hi) = ((hi & 0xfffff) * 2^32 + lo + 0.5) * 2^-52,
//                U32(w) = (rotl(w, 12) + 0.5) * 2^-32.
//     inverse-CDF marginals             4 words  (one 52-bit uniform per sample)
//     Gamma (Marsaglia-Tsang)           0 words  (own counters, see gamma_mt_pair)
// Under a Gaussian copula every random input takes 2 words (its Box-Muller pair).
#pragma once
#include "../../include/ebncuda.h"
#include "ebn_device.cuh"
#include "ebn_fastmath.cuh"
#include "ebn_rng.cuh"
#include "ebn_qmc.cuh"

namespace ebn {

struct PairCtr {
  uint32_t qlo, qhi;  // pair index
  uint32_t stream;    // scenario stream id
  uint32_t k0, k1;    // key
  int ncalls;         // run-time dispatched plans: Philox calls one pair of the scenario consumes (pair_calls)
};

// 52-bit uniform in (0,1) from two words: mantissa = (hi & 0xfffff) : lo.
__device__ __forceinline__ double u52w(uint32_t lo, uint32_t hi) {
  return __hiloint2double((int)((hi & 0x000FFFFFu) | 0x3FF00000u), (int)lo) - 0x1.fffffffffffffp-1;
}
// 32-bit uniform in (0,1): (rot(w) + 0.5) * 2^-32 with rot(w) = w rotated left by 12 (a
// bijection of the 32-bit word, so still exactly uniform); masks only, no conversion.
__device__ __forceinline__ double u32w(uint32_t w) {
  return __hiloint2double((int)((w & 0x000FFFFFu) | 0x3FF00000u), (int)(w & 0xFFF00000u)) - (1.0 - 0x1p-33);
}

// the [1,2) double behind U32(w), before the offset is removed
__device__ __forceinline__ double u32w_raw(uint32_t w) {
  return __hiloint2double((int)hi_of_one_two(w), (int)(w & 0xFFF00000u));
}

// Word sources -----------------------------------------------------------------
// Static: all Philox calls of the pair are made up front, words are picked with
// compile-time indices (registers only).
template <int OFF>
struct StaticWords {
  const Philox4* ph;
  template <int K>
  __device__ __forceinline__ uint32_t w() const {
    constexpr int W = OFF + K;
    const Philox4& p = ph[W >> 2];
    return (W & 3) == 0 ? p.w0 : (W & 3) == 1 ? p.w1 : (W & 3) == 2 ? p.w2 : p.w3;
  }
};
// Dynamic: the offsets are known at run time only. All Philox calls of the pair are made first, by ONE
// loop, into a word buffer in local memory; an input then reads words at its offset (a load each).
// (A call generated on demand behind every word access put a copy of the generator and a branch at each
// of them: 400 of the 1150 instructions a W1 pair with two truncated normals took.)
constexpr int kMaxPairWords = 4 * EBN_MAX_INPUTS;
struct BufWords {
  const uint32_t* words;
  __device__ __forceinline__ uint32_t at(int k) const { return words[k]; }
  template <int K>
  __device__ __forceinline__ uint32_t w() const { return words[K]; }
};
// Philox calls per pair of a scenario: inputs consume words in order (woff), so the last input ends the
// stream. Under a copula every random input owns the two words of its Box-Muller pair.
__device__ __forceinline__ int pair_calls(const DevInput* sp, int d, bool copula) {
  int tw = 0;
  for (int j = 0; j < d; ++j) {
    const int k = sp[j].kind;
    const int n = copula ? (k == DK_CONST ? 0 : 2) : words_of_kind(k);
    tw = max(tw, sp[j].woff + n);
  }
  return (tw + 3) >> 2;
}
// wbuf: 16-byte aligned, kMaxPairWords words
__device__ __forceinline__ void pair_words(const PairCtr& pc, uint32_t* wbuf) {
#pragma unroll 1
  for (int c = 0; c < pc.ncalls; ++c) {
    const Philox4 ph = philox4x32_10(pc.qlo, pc.qhi, pc.stream, (uint32_t)c, pc.k0, pc.k1);
    *reinterpret_cast<uint4*>(wbuf + 4 * c) = make_uint4(ph.w0, ph.w1, ph.w2, ph.w3);
  }
}

// Marsaglia & Tsang (2000) for shape >= 1 (shape < 1 boosted by u^(1/shape)),
// the algorithm class Distributions.jl's Gamma sampler uses. Attempt t of sample
// e of input j reads its own Philox call, counter word 3 =
// 0x80000000 | j<<16 | e<<15 | t: normal from (w0,w1,w2) (first Box-Muller
// output), acceptance uniform from w3. The boost uniform is call t = 0x7fff.
// in.c = d = shape' - 1/3 and in.d = c = 1/sqrt(9 d) are precomputed on the host.
// The acceptance test of one Marsaglia-Tsang attempt: true and v = (1 + c z)^3 when accepted.
__device__ __forceinline__ bool gamma_accept(double dd, double cc, double z, double u, const FastMathSmem& sm,
                                             double& v) {
  const double t1 = fma(cc, z, 1.0);
  v = t1 * t1 * t1;
  const double z2 = z * z;
  // squeeze, then the exact test; both evaluated branch-free (in a warp some lane needs
  // the logarithms almost always), t1 <= 0 gives NaN/false below and is rejected explicitly
  const bool squeeze = u < fma(-0.0331 * z2, z2, 1.0);
  const bool exact = -neg_log01(u, sm) < fma(dd, 1.0 - v - neg_log_pos(fmax(v, 0x1p-1000), sm), 0.5 * z2);
  return t1 > 0.0 && (squeeze || exact);
}

// A retry (attempt t >= 1 of one sample): its own Philox call, normal from (w0, w1) (first Box-Muller
// output), acceptance uniform from w3.
__device__ __forceinline__ bool gamma_attempt(double dd, double cc, uint32_t qlo, uint32_t qhi,
                                              uint32_t stream, uint32_t slot, uint32_t k0,
                                              uint32_t k1, const FastMathSmem& sm, double& v) {
  const Philox4 w = philox4x32_10(qlo, qhi, stream, slot, k0, k1);
  double z, z_unused;
  box_muller_words(w.w0, w.w1, sm, z, z_unused);
  return gamma_accept(dd, cc, z, u32w(w.w3), sm, v);
}

// Slow path (about 0.3 % of the samples at shape 25): attempts 1..63, all arguments by value
// so that nothing lives in local memory.
static __device__ __noinline__ double gamma_retry(double dd, double cc, uint32_t qlo, uint32_t qhi,
                                                  uint32_t stream, uint32_t base, uint32_t k0,
                                                  uint32_t k1, const FastMathSmem* sm) {
  double v = 1.0;
#pragma unroll 1
  for (uint32_t t = 1; t < 64u; ++t)
    if (gamma_attempt(dd, cc, qlo, qhi, stream, base | t, k0, k1, *sm, v)) break;
  return v;
}

static __device__ __noinline__ double gamma_boost(double shape, uint32_t qlo, uint32_t qhi,
                                                  uint32_t stream, uint32_t base, uint32_t k0,
                                                  uint32_t k1) {
  const Philox4 w = philox4x32_10(qlo, qhi, stream, base | 0x7FFFu, k0, k1);
  return pow(u52w(w.w0, w.w1), 1.0 / shape);
}

// Both samples of a pair. in.c = d = shape' - 1/3 and in.d = c = 1/sqrt(9 d) are precomputed on the host.
// Layout v4: the FIRST attempt of the two samples shares one Philox call (counter word 3 =
// 0x80000000 | j<<16): one Box-Muller gives z for sample 0 and z for sample 1, the acceptance uniforms are
// w2 and w3 — half the generator calls and half the transforms of v3 on the common path. A rejected sample
// retries on its own counters (0x80000000 | j<<16 | e<<15 | t, t = 1..63), the shape < 1 boost reads
// call t = 0x7fff of its sample.
__device__ __forceinline__ void gamma_mt_pair(const DevInput& in, const PairCtr& pc, uint32_t j,
                                              const FastMathSmem& sm, double& x0, double& x1) {
  const uint32_t base = 0x80000000u | (j << 16);
  const Philox4 w = philox4x32_10(pc.qlo, pc.qhi, pc.stream, base, pc.k0, pc.k1);
  double z0, z1, v0, v1;
  box_muller_words(w.w0, w.w1, sm, z0, z1);
  const bool ok0 = gamma_accept(in.c, in.d, z0, u32w(w.w2), sm, v0);
  const bool ok1 = gamma_accept(in.c, in.d, z1, u32w(w.w3), sm, v1);
  if (!ok0) v0 = gamma_retry(in.c, in.d, pc.qlo, pc.qhi, pc.stream, base, pc.k0, pc.k1, &sm);
  if (!ok1) v1 = gamma_retry(in.c, in.d, pc.qlo, pc.qhi, pc.stream, base | (1u << 15), pc.k0, pc.k1, &sm);
  x0 = in.c * v0;
  x1 = in.c * v1;
  if (in.a < 1.0) {
    x0 *= gamma_boost(in.a, pc.qlo, pc.qhi, pc.stream, base, pc.k0, pc.k1);
    x1 *= gamma_boost(in.a, pc.qlo, pc.qhi, pc.stream, base | (1u << 15), pc.k0, pc.k1);
  }
  x0 *= in.b;
  x1 *= in.b;
}

__device__ __forceinline__ double table_quantile(const DevInput& in, const double* tables,
                                                 double u) {
  const double* q = tables + in.toff;
  const double pos = u * (double)(in.npts - 1);
  int i = (int)pos;
  i = i < 0 ? 0 : (i > in.npts - 2 ? in.npts - 2 : i);
  const double f = pos - (double)i;
  const double q0 = q[i], q1 = q[i + 1];
  return q0 + f * (q1 - q0);
}

// x = Q(u') for the inverse-CDF kinds, u' = c + u*d already mapped into the
// truncation window.
// CELLS: the kernel holds the Phi^-1 cell table in shared memory (compile-time sampling plans that can
// draw a truncated normal); otherwise the same cells are read from global memory (same bits, slower).
template <int KIND, bool CELLS = false>
__device__ __forceinline__ double icdf(const DevInput& in, const double* tables,
                                       const FastMathSmem& sm, double u) {
  const double up = fma(u, in.d, in.c);
  if (KIND == DK_NORMAL_ICDF || KIND == DK_LOGNORMAL_ICDF) {
    const double z = normcdfinv_fast<CELLS>(up);
    const double y = fma(z, in.b, in.a);
    return KIND == DK_LOGNORMAL_ICDF ? exp_fast(y, sm) : y;
  }
  if (KIND == DK_GUMBEL) return fma(in.b, neg_log_pos(neg_log01(up, sm), sm), in.a);  // a - b ln(-ln u)
  if (KIND == DK_EXP_AFFINE) {   // a - b ln(1 - u'): the lean logarithm unless 1 - u' rounds to zero
    const double t = 1.0 - up;
    return t > 0x1p-1000 ? fma(in.b, neg_log_pos(t, sm), in.a) : in.a - in.b * log1p(-up);
  }
  return table_quantile(in, tables, up);
}

// Per-input state between the front end (word consumption, one pair ahead in the pipelined
// loop) and the back end (fp64 transform) of a draw. Unused members are dead code per kind.
struct InState {
  uint32_t w0, w1, w2, w3;
  BmFront bm;
};

template <int KIND, class WS>
__device__ __forceinline__ InState draw_front(const WS& ws, const FastMathSmem& sm) {
  InState st;
  if (KIND == DK_UNIFORM) {
    st.w0 = ws.template w<0>();
    st.w1 = ws.template w<1>();
  } else if (KIND == DK_NORMAL_BM || KIND == DK_LOGNORMAL_BM) {
    st.bm = box_muller_front(ws.template w<0>(), ws.template w<1>(), sm);
  } else if (KIND == DK_NORMAL_BM32 || KIND == DK_LOGNORMAL_BM32) {
    st.w0 = ws.template w<0>();
    st.w1 = ws.template w<1>();
  } else if (KIND != DK_CONST && KIND != DK_GAMMA_MT) {
    st.w0 = ws.template w<0>();
    st.w1 = ws.template w<1>();
    st.w2 = ws.template w<2>();
    st.w3 = ws.template w<3>();
  }
  return st;
}

template <int KIND, bool CELLS = false>
__device__ __forceinline__ void draw_back(const DevInput& in, const InState& st, const PairCtr& pc,
                                          uint32_t j, const double* tables,
                                          const FastMathSmem& sm, double& x0, double& x1) {
  if (KIND == DK_CONST) {
    x0 = in.a;
    x1 = in.a;
  } else if (KIND == DK_UNIFORM) {
    // x = a + b*U32(w) with U32(w) = d - (1 - 2^-33), d in [1,2): folded into one FMA, the
    // host stores c = a - b*(1 - 2^-33)
    x0 = fma(u32w_raw(st.w0), in.b, in.c);
    x1 = fma(u32w_raw(st.w1), in.b, in.c);
  } else if (KIND == DK_NORMAL_BM || KIND == DK_LOGNORMAL_BM) {
    box_muller_back_affine(st.bm, sm, in.a, in.b, x0, x1);
    if (KIND == DK_LOGNORMAL_BM) {
      x0 = exp_fast(x0, sm);
      x1 = exp_fast(x1, sm);
    }
  } else if (KIND == DK_NORMAL_BM32 || KIND == DK_LOGNORMAL_BM32) {
    double z0, z1;
    box_muller_words_f32(st.w0, st.w1, z0, z1);
    x0 = fma(z0, in.b, in.a);
    x1 = fma(z1, in.b, in.a);
    if (KIND == DK_LOGNORMAL_BM32) {
      x0 = exp_fast(x0, sm);
      x1 = exp_fast(x1, sm);
    }
  } else if (KIND == DK_GAMMA_MT) {
    gamma_mt_pair(in, pc, j, sm, x0, x1);
  } else if (KIND == DK_NORMAL_ICDF || KIND == DK_LOGNORMAL_ICDF) {
    double z0, z1;
    normcdfinv_fast2<CELLS>(fma(u52w(st.w0, st.w1), in.d, in.c), fma(u52w(st.w2, st.w3), in.d, in.c), z0, z1);
    x0 = fma(z0, in.b, in.a);
    x1 = fma(z1, in.b, in.a);
    if (KIND == DK_LOGNORMAL_ICDF) {
      x0 = exp_fast(x0, sm);
      x1 = exp_fast(x1, sm);
    }
  } else {
    x0 = icdf<KIND>(in, tables, sm, u52w(st.w0, st.w1));
    x1 = icdf<KIND>(in, tables, sm, u52w(st.w2, st.w3));
  }
}

// Draws input j for both samples of a pair. KIND is a DevKind, WS a word source
// positioned at the input's first word.
template <int KIND, class WS, bool CELLS = false>
__device__ __forceinline__ void draw_pair(const DevInput& in, const WS& ws, const PairCtr& pc,
                                          uint32_t j, const double* tables,
                                          const FastMathSmem& sm, double& x0, double& x1) {
  draw_back<KIND, CELLS>(in, draw_front<KIND>(ws, sm), pc, j, tables, sm, x0, x1);
}

// Runtime (warp-uniform) dispatch on the kind; the generic, slower path.
__device__ __forceinline__ void draw_pair_dyn_body(const DevInput& in, const PairCtr& pc,
                                                   const uint32_t* wbuf, uint32_t j,
                                                   const double* tables, const FastMathSmem& sm,
                                                   double& x0, double& x1) {
  BufWords ws{wbuf + in.woff};
  switch (in.kind) {
    case DK_CONST: draw_pair<DK_CONST>(in, ws, pc, j, tables, sm, x0, x1); break;
    case DK_UNIFORM: draw_pair<DK_UNIFORM>(in, ws, pc, j, tables, sm, x0, x1); break;
    case DK_NORMAL_BM: draw_pair<DK_NORMAL_BM>(in, ws, pc, j, tables, sm, x0, x1); break;
    case DK_LOGNORMAL_BM: draw_pair<DK_LOGNORMAL_BM>(in, ws, pc, j, tables, sm, x0, x1); break;
    case DK_NORMAL_ICDF: draw_pair<DK_NORMAL_ICDF>(in, ws, pc, j, tables, sm, x0, x1); break;
    case DK_LOGNORMAL_ICDF: draw_pair<DK_LOGNORMAL_ICDF>(in, ws, pc, j, tables, sm, x0, x1); break;
    case DK_GUMBEL: draw_pair<DK_GUMBEL>(in, ws, pc, j, tables, sm, x0, x1); break;
    case DK_EXP_AFFINE: draw_pair<DK_EXP_AFFINE>(in, ws, pc, j, tables, sm, x0, x1); break;
    case DK_GAMMA_MT: draw_pair<DK_GAMMA_MT>(in, ws, pc, j, tables, sm, x0, x1); break;
    case DK_NORMAL_BM32: draw_pair<DK_NORMAL_BM32>(in, ws, pc, j, tables, sm, x0, x1); break;
    case DK_LOGNORMAL_BM32: draw_pair<DK_LOGNORMAL_BM32>(in, ws, pc, j, tables, sm, x0, x1); break;
    default: draw_pair<DK_TABLE>(in, ws, pc, j, tables, sm, x0, x1); break;
  }
}
// out of line: the unrolled per-input loops of the fixed-arity limit states would hold a copy per input
static __device__ __noinline__ void draw_pair_dyn(const DevInput& in, const PairCtr& pc,
                                                  const uint32_t* wbuf, uint32_t j,
                                                  const double* tables, const FastMathSmem& sm,
                                                  double& x0, double& x1) {
  draw_pair_dyn_body(in, pc, wbuf, j, tables, sm, x0, x1);
}

// One input of a sample pair under quasi-Monte Carlo (ebn_qmc.cuh): Sobol coordinate v of the even sample
// (the odd one differs by the first direction number of the dimension) -> Owen-scrambled uniform -> the
// marginal's inverse CDF on its truncation window. Untruncated normals take Phi^-1 like the truncated
// ones: a Box-Muller pair would tie two samples to one radius and spend two dimensions per normal.
static __device__ __noinline__ void draw_pair_qmc(const DevInput& in, uint32_t v, uint32_t odd_xor, uint32_t key,
                                                  const double* tables, const FastMathSmem& sm,
                                                  double& x0, double& x1) {
  const double u0 = qmc_unit(owen_scramble(v, key));
  const double u1 = qmc_unit(owen_scramble(v ^ odd_xor, key));
  switch (in.kind) {
    case DK_UNIFORM:
      x0 = fma(u0, in.b, in.a);
      x1 = fma(u1, in.b, in.a);
      break;
    case DK_NORMAL_BM: case DK_LOGNORMAL_BM: case DK_NORMAL_BM32: case DK_LOGNORMAL_BM32: {
      double z0, z1;                         // no window: c = 0, d = 1
      normcdfinv_fast2<true>(u0, u1, z0, z1);
      x0 = fma(z0, in.b, in.a);
      x1 = fma(z1, in.b, in.a);
      if (in.kind == DK_LOGNORMAL_BM || in.kind == DK_LOGNORMAL_BM32) {
        x0 = exp_fast(x0, sm);
        x1 = exp_fast(x1, sm);
      }
      break;
    }
    case DK_NORMAL_ICDF: x0 = icdf<DK_NORMAL_ICDF, true>(in, tables, sm, u0); x1 = icdf<DK_NORMAL_ICDF, true>(in, tables, sm, u1); break;
    case DK_LOGNORMAL_ICDF: x0 = icdf<DK_LOGNORMAL_ICDF, true>(in, tables, sm, u0); x1 = icdf<DK_LOGNORMAL_ICDF, true>(in, tables, sm, u1); break;
    case DK_GUMBEL: x0 = icdf<DK_GUMBEL, true>(in, tables, sm, u0); x1 = icdf<DK_GUMBEL, true>(in, tables, sm, u1); break;
    case DK_EXP_AFFINE: x0 = icdf<DK_EXP_AFFINE, true>(in, tables, sm, u0); x1 = icdf<DK_EXP_AFFINE, true>(in, tables, sm, u1); break;
    default: x0 = icdf<DK_TABLE, true>(in, tables, sm, u0); x1 = icdf<DK_TABLE, true>(in, tables, sm, u1); break;   // DK_TABLE
  }
}

// Gaussian copula (UQ.jl JointDistribution + GaussianCopula): eps ~ N(0,I) by
// Box-Muller (2 words per random input, in.woff), z = L*eps, then per marginal
// x = a + b*z for the normal family and x = Q(Phi(z)) otherwise. L is row-major
// lower-triangular [d*d].
static __device__ __noinline__ void draw_pair_copula(const DevInput* plan, const double* L, int d,
                                                     const PairCtr& pc, const double* tables,
                                                     const FastMathSmem& sm, double* x0,
                                                     double* x1) {
  double2 ee[EBN_MAX_INPUTS];   // the pair's two independent scores of input k side by side: one 16-byte load per use
  __align__(16) uint32_t wbuf[kMaxPairWords];
  pair_words(pc, wbuf);
  for (int j = 0; j < d; ++j) {
    double a = 0.0, b = 0.0;
    if (plan[j].kind != DK_CONST) {
      const uint32_t* w = wbuf + plan[j].woff;
      box_muller_words(w[0], w[1], sm, a, b);
    }
    ee[j] = make_double2(a, b);
  }
  for (int j = 0; j < d; ++j) {
    const DevInput in = plan[j];
    if (in.kind == DK_CONST) {
      x0[j] = in.a;
      x1[j] = in.a;
      continue;
    }
    double z0 = 0.0, z1 = 0.0;
#pragma unroll 4
    for (int k = 0; k <= j; ++k) {
      const double l = L[j * d + k];
      const double2 e = ee[k];
      z0 = fma(l, e.x, z0);
      z1 = fma(l, e.y, z1);
    }
    double a0, a1;
    switch (in.kind) {
      case DK_NORMAL_BM:
        a0 = fma(z0, in.b, in.a);
        a1 = fma(z1, in.b, in.a);
        break;
      case DK_LOGNORMAL_BM:
        a0 = exp_fast(fma(z0, in.b, in.a), sm);
        a1 = exp_fast(fma(z1, in.b, in.a), sm);
        break;
      case DK_UNIFORM:
        a0 = fma(phi_clamped(z0, sm), in.b, in.a);
        a1 = fma(phi_clamped(z1, sm), in.b, in.a);
        break;
      case DK_NORMAL_ICDF:
        a0 = icdf<DK_NORMAL_ICDF>(in, tables, sm, phi_clamped(z0, sm));
        a1 = icdf<DK_NORMAL_ICDF>(in, tables, sm, phi_clamped(z1, sm));
        break;
      case DK_LOGNORMAL_ICDF:
        a0 = icdf<DK_LOGNORMAL_ICDF>(in, tables, sm, phi_clamped(z0, sm));
        a1 = icdf<DK_LOGNORMAL_ICDF>(in, tables, sm, phi_clamped(z1, sm));
        break;
      case DK_GUMBEL:
        a0 = icdf<DK_GUMBEL>(in, tables, sm, phi_clamped(z0, sm));
        a1 = icdf<DK_GUMBEL>(in, tables, sm, phi_clamped(z1, sm));
        break;
      case DK_EXP_AFFINE:
        a0 = icdf<DK_EXP_AFFINE>(in, tables, sm, phi_clamped(z0, sm));
        a1 = icdf<DK_EXP_AFFINE>(in, tables, sm, phi_clamped(z1, sm));
        break;
      default:  // DK_TABLE (Gamma under a copula is tabulated by the caller)
        a0 = icdf<DK_TABLE>(in, tables, sm, phi_clamped(z0, sm));
        a1 = icdf<DK_TABLE>(in, tables, sm, phi_clamped(z1, sm));
        break;
    }
    x0[j] = a0;
    x1[j] = a1;
  }
}

}  // namespace ebn
