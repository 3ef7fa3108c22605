// In-register marginal transforms: one Philox call per (input, sample pair).
// Replaces `rand(dist, n)` per input inside UQ.jl's `sample(inputs, sim)`
// (reference call sites src/networks/ebn/evaluation/evaluate_node.jl:24,83;
// the marginals that reach it are listed in SURVEY.md F7 / rows 2 and 5).
#pragma once
#include "../../include/ebncuda.h"
#include "ebn_device.cuh"
#include "ebn_rng.cuh"

namespace ebn {

struct PairCtr {
  uint32_t qlo, qhi;  // pair index
  uint32_t stream;    // scenario stream id
  uint32_t k0, k1;    // key
};

// Marsaglia & Tsang (2000) for shape >= 1 (shape < 1 boosted by u^(1/shape)),
// the same algorithm class Distributions.jl's Gamma sampler uses. Attempt t of
// sample e in the pair reads counter slot j + 16*(1 + 2*t + e): radius from
// w1:w0 (52 bits), angle from w2, acceptance uniform from w3.
static __device__ __noinline__ double gamma_mt(const DevInput& in, const PairCtr& pc, uint32_t j,
                                        int e, double uboost) {
  const double shape = in.a;
  const double a = shape < 1.0 ? shape + 1.0 : shape;
  const double dd = a - 1.0 / 3.0;
  const double cc = 1.0 / sqrt(9.0 * dd);
  double v = 1.0;
  for (uint32_t t = 0; t < 64u; ++t) {
    const Philox4 w =
        philox4x32_10(pc.qlo, pc.qhi, pc.stream, j + 16u * (1u + 2u * t + (uint32_t)e), pc.k0, pc.k1);
    const double r = sqrt(-2.0 * log(u52(w.w0, w.w1)));
    const double z = r * cospi(2.0 * u32(w.w2));
    const double t1 = 1.0 + cc * z;
    if (t1 <= 0.0) continue;
    v = t1 * t1 * t1;
    const double u = u32(w.w3);
    const double z2 = z * z;
    if (u < 1.0 - 0.0331 * z2 * z2) break;
    if (log(u) < 0.5 * z2 + dd * (1.0 - v + log(v))) break;
  }
  double x = dd * v;
  if (shape < 1.0) x *= pow(uboost, 1.0 / shape);
  return x * in.b;
}

__device__ __forceinline__ double table_quantile(const DevInput& in, const double* tables,
                                                 double u) {
  const double* q = tables + in.off;
  const double pos = u * (double)(in.npts - 1);
  int i = (int)pos;
  i = i < 0 ? 0 : (i > in.npts - 2 ? in.npts - 2 : i);
  const double f = pos - (double)i;
  const double q0 = q[i], q1 = q[i + 1];
  return q0 + f * (q1 - q0);
}

// Draws input j for both samples of a pair. KIND is a DevKind.
template <int KIND>
__device__ __forceinline__ void draw_pair(const DevInput& in, const PairCtr& pc, uint32_t j,
                                          const double* tables, double& x0, double& x1) {
  if (KIND == DK_CONST) {
    x0 = in.a;
    x1 = in.a;
    return;
  }
  const Philox4 w = philox4x32_10(pc.qlo, pc.qhi, pc.stream, j, pc.k0, pc.k1);
  const double ua = u52(w.w0, w.w1);
  const double ub = u52(w.w2, w.w3);
  if (KIND == DK_UNIFORM) {
    x0 = fma(ua, in.b, in.a);
    x1 = fma(ub, in.b, in.a);
  } else if (KIND == DK_NORMAL_BM || KIND == DK_LOGNORMAL_BM) {
    double z0, z1;
    box_muller(ua, ub, z0, z1);
    x0 = fma(z0, in.b, in.a);
    x1 = fma(z1, in.b, in.a);
    if (KIND == DK_LOGNORMAL_BM) {
      x0 = exp(x0);
      x1 = exp(x1);
    }
  } else if (KIND == DK_NORMAL_ICDF || KIND == DK_LOGNORMAL_ICDF) {
    x0 = fma(normcdfinv(fma(ua, in.d, in.c)), in.b, in.a);
    x1 = fma(normcdfinv(fma(ub, in.d, in.c)), in.b, in.a);
    if (KIND == DK_LOGNORMAL_ICDF) {
      x0 = exp(x0);
      x1 = exp(x1);
    }
  } else if (KIND == DK_GUMBEL) {
    x0 = in.a - in.b * log(-log(fma(ua, in.d, in.c)));
    x1 = in.a - in.b * log(-log(fma(ub, in.d, in.c)));
  } else if (KIND == DK_EXP_AFFINE) {
    x0 = in.a - in.b * log1p(-fma(ua, in.d, in.c));
    x1 = in.a - in.b * log1p(-fma(ub, in.d, in.c));
  } else if (KIND == DK_GAMMA_MT) {
    x0 = gamma_mt(in, pc, j, 0, ua);
    x1 = gamma_mt(in, pc, j, 1, ub);
  } else if (KIND == DK_TABLE) {
    x0 = table_quantile(in, tables, fma(ua, in.d, in.c));
    x1 = table_quantile(in, tables, fma(ub, in.d, in.c));
  }
}

// Runtime (warp-uniform) dispatch on the kind.
static __device__ __noinline__ void draw_pair_dyn(const DevInput& in, const PairCtr& pc, uint32_t j,
                                              const double* tables, double& x0, double& x1) {
  switch (in.kind) {
    case DK_CONST: draw_pair<DK_CONST>(in, pc, j, tables, x0, x1); break;
    case DK_UNIFORM: draw_pair<DK_UNIFORM>(in, pc, j, tables, x0, x1); break;
    case DK_NORMAL_BM: draw_pair<DK_NORMAL_BM>(in, pc, j, tables, x0, x1); break;
    case DK_LOGNORMAL_BM: draw_pair<DK_LOGNORMAL_BM>(in, pc, j, tables, x0, x1); break;
    case DK_NORMAL_ICDF: draw_pair<DK_NORMAL_ICDF>(in, pc, j, tables, x0, x1); break;
    case DK_LOGNORMAL_ICDF: draw_pair<DK_LOGNORMAL_ICDF>(in, pc, j, tables, x0, x1); break;
    case DK_GUMBEL: draw_pair<DK_GUMBEL>(in, pc, j, tables, x0, x1); break;
    case DK_EXP_AFFINE: draw_pair<DK_EXP_AFFINE>(in, pc, j, tables, x0, x1); break;
    case DK_GAMMA_MT: draw_pair<DK_GAMMA_MT>(in, pc, j, tables, x0, x1); break;
    default: draw_pair<DK_TABLE>(in, pc, j, tables, x0, x1); break;
  }
}

// Gaussian copula (UQ.jl JointDistribution + GaussianCopula): eps ~ N(0,I) by
// Box-Muller, z = L*eps, then per marginal x = a + b*z for the normal family
// and x = Q(Phi(z)) otherwise. L is row-major lower-triangular [d*d].
static __device__ __noinline__ void draw_pair_copula(const DevInput* plan, const double* L, int d,
                                                 const PairCtr& pc, const double* tables,
                                                 double* x0, double* x1) {
  double e0[EBN_MAX_INPUTS], e1[EBN_MAX_INPUTS];
  for (int j = 0; j < d; ++j) {
    e0[j] = 0.0;
    e1[j] = 0.0;
    if (plan[j].kind != DK_CONST) {
      const Philox4 w = philox4x32_10(pc.qlo, pc.qhi, pc.stream, (uint32_t)j, pc.k0, pc.k1);
      box_muller(u52(w.w0, w.w1), u52(w.w2, w.w3), e0[j], e1[j]);
    }
  }
  for (int j = 0; j < d; ++j) {
    const DevInput in = plan[j];
    if (in.kind == DK_CONST) {
      x0[j] = in.a;
      x1[j] = in.a;
      continue;
    }
    double z0 = 0.0, z1 = 0.0;
    for (int k = 0; k <= j; ++k) {
      const double l = L[j * d + k];
      z0 = fma(l, e0[k], z0);
      z1 = fma(l, e1[k], z1);
    }
    double a0, a1;
    switch (in.kind) {
      case DK_NORMAL_BM:
      case DK_NORMAL_ICDF:
        // (a truncated normal under a copula goes through Phi like the others)
        if (in.kind == DK_NORMAL_BM) {
          a0 = fma(z0, in.b, in.a);
          a1 = fma(z1, in.b, in.a);
        } else {
          a0 = fma(normcdfinv(fma(normcdf(z0), in.d, in.c)), in.b, in.a);
          a1 = fma(normcdfinv(fma(normcdf(z1), in.d, in.c)), in.b, in.a);
        }
        break;
      case DK_LOGNORMAL_BM:
        a0 = exp(fma(z0, in.b, in.a));
        a1 = exp(fma(z1, in.b, in.a));
        break;
      case DK_LOGNORMAL_ICDF:
        a0 = exp(fma(normcdfinv(fma(normcdf(z0), in.d, in.c)), in.b, in.a));
        a1 = exp(fma(normcdfinv(fma(normcdf(z1), in.d, in.c)), in.b, in.a));
        break;
      case DK_UNIFORM:
        a0 = fma(normcdf(z0), in.b, in.a);
        a1 = fma(normcdf(z1), in.b, in.a);
        break;
      case DK_GUMBEL:
        a0 = in.a - in.b * log(-log(fma(normcdf(z0), in.d, in.c)));
        a1 = in.a - in.b * log(-log(fma(normcdf(z1), in.d, in.c)));
        break;
      case DK_EXP_AFFINE:
        a0 = in.a - in.b * log1p(-fma(normcdf(z0), in.d, in.c));
        a1 = in.a - in.b * log1p(-fma(normcdf(z1), in.d, in.c));
        break;
      default:  // DK_TABLE (Gamma under a copula is tabulated by the host)
        a0 = table_quantile(in, tables, fma(normcdf(z0), in.d, in.c));
        a1 = table_quantile(in, tables, fma(normcdf(z1), in.d, in.c));
        break;
    }
    x0[j] = a0;
    x1[j] = a1;
  }
}

}  // namespace ebn
