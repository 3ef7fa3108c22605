// Device-side plan structures shared by the kernels and the host API.
#pragma once
#include "ebn_rtc_compat.h"

// Threads per block of every kernel. Measured on B200 with the vehicle kernel at 16 warps per SM
// (profiles/r01_m_tune_block_shape.txt): 64x8 1.466e11, 128x4 1.471e11, 256x2 1.424e11 samples/s —
// smaller blocks lose less at the per-item reduction barrier.
#ifndef EBN_THREADS
#define EBN_THREADS 128
#endif
// blocks per SM that give `warps` resident warps per SM
#define EBN_BLOCKS_FOR_WARPS(warps) ((warps) * 32 / EBN_THREADS)

namespace ebn {

// Canonical per-input sampling recipe, produced on the host from ebn_input_t
// (truncation windows are mapped to probability space there, in fp64).
enum DevKind : int32_t {
  DK_CONST = 0,           // x = a                          (UQ Parameter)
  DK_UNIFORM = 1,         // x = a + u*b
  DK_NORMAL_BM = 2,       // x = a + b*z,  z by Box-Muller across the sample pair
  DK_LOGNORMAL_BM = 3,    // x = exp(a + b*z)
  DK_NORMAL_ICDF = 4,     // x = a + b*Phi^-1(c + u*d)      (truncated)
  DK_LOGNORMAL_ICDF = 5,  // x = exp(a + b*Phi^-1(c + u*d))
  DK_GUMBEL = 6,          // x = a - b*log(-log(c + u*d))
  DK_EXP_AFFINE = 7,      // x = a - b*log1p(-(c + u*d))    (b carries sign*theta)
  DK_GAMMA_MT = 8,        // Marsaglia-Tsang, shape a, scale b (untruncated)
  DK_TABLE = 9,           // x = Q(c + u*d), Q linear on a uniform grid of npts points
  DK_NORMAL_BM32 = 10,    // as DK_NORMAL_BM, the Box-Muller transform evaluated in fp32
  DK_LOGNORMAL_BM32 = 11, // (EBN_JOB_FP32_SAMPLING; same words, z agrees to ~1e-6)
  DK_NKINDS = 12
};

struct DevInput {
  int32_t kind;
  int32_t npts;
  double a, b, c, d;
  int32_t toff;  // DK_TABLE: offset of Q[0] in the tables pool
  int32_t woff;  // first word of this input in the pair's Philox word stream
};

// Philox words one input consumes per sample pair (stream layout v3).
#ifdef __CUDACC__
__host__ __device__
#endif
constexpr int words_of_kind(int kind) {
  return kind == DK_CONST ? 0
       : kind == DK_UNIFORM ? 2
       : (kind == DK_NORMAL_BM || kind == DK_LOGNORMAL_BM || kind == DK_NORMAL_BM32 || kind == DK_LOGNORMAL_BM32) ? 2
       : kind == DK_GAMMA_MT ? 0
       : 4;
}
static_assert(sizeof(DevInput) == 48, "DevInput layout");

struct McParams {
  const DevInput* plan;         // [S*d]
  const double* consts;         // [n_consts]
  const double* tables;         // pool
  const double* chol;           // nullable: copula factors
  const uint32_t* stream_id;    // nullable [S]
  unsigned long long* counts;   // PF: [S]; HIST: [S*(nedges+1)]
  const double* edges;          // HIST: [nedges]
  double* partial;              // HIST: [items*2] per-item (sum, sumsq)
  int64_t first;                // first global sample index
  int64_t last;                 // one past the last global sample index
  int64_t pairs_per_item;       // pairs in one work item
  int64_t items_per_scenario;
  int64_t local_items;          // items this rank processes
  int64_t pair_base;            // first >> 1
  int32_t shard_rank, shard_world;
  int32_t S, d, n_consts, nedges;
  uint32_t k0, k1;
  int32_t chol_shared;
  int32_t ncols;  // replay: columns written per sample (d inputs + model outputs)
  int32_t hist_uniform;  // HIST: edges are equally spaced -> direct bin index + exact fix-up
  double hist_e0, hist_inv_h;
  int32_t icdf_dyn;  // run-time dispatched plans: the launch carries kIcdfDynBytes of dynamic shared memory for the Phi^-1 cells
};

// Dynamic shared memory of mc_kernel (ebn_mc_impl.cuh): the histogram's edges and bins, rounded up to 16
// bytes, then — for run-time dispatched plans that draw truncated normals — the Phi^-1 cell table.
constexpr int kIcdfDynBytes = 257 * 5 * 16;   // EBN_ICDF_ROWS rows, kIcdfStride double2 apart (ebn_fastmath.cuh)
__host__ __device__ constexpr size_t hist_dyn_bytes(int nedges) {
  return nedges > 0 ? (((size_t)nedges * sizeof(double) + (size_t)(nedges + 1) * sizeof(unsigned)) + 15) / 16 * 16 : 0;
}

}  // namespace ebn
