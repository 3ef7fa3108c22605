// Randomised quasi-Monte Carlo: an Owen-scrambled Sobol sequence in place of the Philox word stream.
//
// UQ.jl's quasi-Monte Carlo simulations (SobolSampling & co., subtypes of AbstractMonteCarlo, which
// `ContinuousFunctionalNode.simulation::AbstractMonteCarlo` and `DiscreteFunctionalNode.simulation`
// accept: reference src/nodes/functionalnodes.jl:4,31) draw a low-discrepancy point set in the unit
// cube and map every coordinate through the marginal's quantile. Here:
//   sample i of scenario stream s, random input number r (Parameters take no dimension):
//     v = Sobol_r(i)                       32-bit, gray-code order, Joe & Kuo direction numbers
//     w = owen(v, key(seed, s, r))         hash-based nested uniform scrambling (Laine & Karras 2011,
//                                          Burley 2020): a random permutation of every dyadic
//                                          sub-interval, so the scrambled set is again a (t,m,d)-net
//     u = (w + 1/2) 2^-32,  x = Q_j(c + u d)      every marginal through its inverse CDF
// The point of sample i is a pure function of (seed, stream, i), like the Monte Carlo streams: counts do
// not depend on the grid, the shard or a resume offset. Threads walk their samples in strides, so the
// state is advanced by XOR-ing the direction numbers of the gray-code bits that changed (2-3 per step).
#pragma once
#include "../../include/ebncuda.h"
#include "ebn_device.cuh"
#include "ebn_fastmath.cuh"
#include "ebn_rng.cuh"
#include "ebn_sobol_tables.inc"

namespace ebn {

struct QmcState {
  uint32_t v[EBN_MAX_INPUTS];    // unscrambled Sobol coordinates of the pair's even sample, one per random input
  uint32_t key[EBN_MAX_INPUTS];  // scrambling key per (stream, dimension)
  uint32_t index;                // sample index the state stands at (even)
};

__device__ __forceinline__ uint32_t* sobol_smem() {
  __shared__ uint32_t s_v[EBN_SOBOL_DIMS * 32];
  return s_v;
}
__device__ __forceinline__ void sobol_init(int tid, int nthreads) {
  uint32_t* t = sobol_smem();
  const uint32_t* g = &kSobolV[0][0];
  for (int i = tid; i < EBN_SOBOL_DIMS * 32; i += nthreads) t[i] = g[i];
}

// Owen scrambling by hashing: reversed bits turn "permute each dyadic sub-interval" into "each bit may
// flip depending on the bits below it", which the multiply-xor rounds provide.
__device__ __forceinline__ uint32_t owen_scramble(uint32_t x, uint32_t key) {
  x = __brev(x);
  x += key;
  x ^= x * 0x6c50b47cu;
  x ^= x * 0xb82f1e52u;
  x ^= x * 0xc7afe638u;
  x ^= x * 0x8d22f6e6u;
  return __brev(x);
}

// (w + 1/2) 2^-32 in (0,1), the bits in order (the stratification lives in the high bits).
__device__ __forceinline__ double qmc_unit(uint32_t w) {
  return __hiloint2double((int)(0x3FF00000u | (w >> 12)), (int)(w << 20)) - (1.0 - 0x1p-33);
}

// Point of sample `index` (even) in the first nd dimensions, from scratch.
__device__ __forceinline__ void qmc_seek(QmcState& st, uint32_t index, int nd) {
  const uint32_t* V = sobol_smem();
  uint32_t g = index ^ (index >> 1);
#pragma unroll 1
  for (int r = 0; r < nd; ++r) st.v[r] = 0u;
  while (g) {
    const int k = __ffs((int)g) - 1;
    g &= g - 1u;
#pragma unroll 1
    for (int r = 0; r < nd; ++r) st.v[r] ^= V[r * 32 + k];
  }
  st.index = index;
}

// Moves the state to another (even) sample index: only the gray-code bits that differ are touched.
__device__ __forceinline__ void qmc_move(QmcState& st, uint32_t index, int nd) {
  const uint32_t* V = sobol_smem();
  uint32_t g = (index ^ (index >> 1)) ^ (st.index ^ (st.index >> 1));
  while (g) {
    const int k = __ffs((int)g) - 1;
    g &= g - 1u;
#pragma unroll 1
    for (int r = 0; r < nd; ++r) st.v[r] ^= V[r * 32 + k];
  }
  st.index = index;
}

// Scrambling keys of a scenario: one Philox word per dimension, counter = (dimension, 0, stream, 'SOBL').
__device__ __forceinline__ void qmc_keys(QmcState& st, uint32_t stream, uint32_t k0, uint32_t k1, int nd) {
#pragma unroll 1
  for (int r = 0; r < nd; ++r) st.key[r] = philox4x32_10((uint32_t)r, 0u, stream, 0x534F424Cu, k0, k1).w0;
}

}  // namespace ebn
