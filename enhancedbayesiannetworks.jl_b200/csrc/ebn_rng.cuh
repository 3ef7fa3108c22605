// Counter-based RNG and uniform/normal primitives (device side).
//
// key = (seed lo32, seed hi32); counter = (pair lo32, pair hi32, stream id of the
// scenario, call index); pair = global sample index >> 1. How the words are
// consumed is defined in ebn_sampling.cuh (stream layout v3).
// The replacement for `rand(dist, n)` inside UQ.jl's `sample(inputs, n)`
// (called from reference src/networks/ebn/evaluation/evaluate_node.jl:24,83).
#pragma once
#include "ebn_rtc_compat.h"

namespace ebn {

struct Philox4 {
  uint32_t w0, w1, w2, w3;
};

// Philox4x32-10 (Salmon et al., SC'11). 2 IMAD.WIDE + 2 LOP3 per round; the
// key schedule is uniform across the grid so it lives in uniform registers.
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                 uint32_t c3, uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#ifndef EBN_PHILOX_ROUNDS
#define EBN_PHILOX_ROUNDS 10   // measurement-only switch (profiles/r02_*): the shipped stream is Philox4x32-10
#endif
#pragma unroll
  for (int r = 0; r < EBN_PHILOX_ROUNDS; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0;
    const uint64_t p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += W0;
    k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

}  // namespace ebn
