// Component timing of the vehicle kernel's inner loop (sm_100a): cycles per sample-pair per
// warp per SM sub-partition for (A) the 3 Philox calls, (B) + the three Box-Muller transforms,
// (C) + the sign-only limit state, at several occupancies.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../enhancedbayesiannetworks.jl_b200/csrc -o components components.cu
#include <cstdio>
#include "ebn_limit_states.cuh"
#include "ebn_sampling.cuh"

using namespace ebn;

template <int STAGE>
__global__ void __launch_bounds__(256) comp(unsigned long long* out, const double* consts, long long pairs_per_thread,
                                            unsigned seed) {
  __shared__ FastMathSmem s_fm;
  fastmath_init(s_fm, threadIdx.x, 256);
  __syncthreads();
  using LS = VehicleSuspension<FastOps>;
  typename LS::Ctx ctx;
  LS::setup(ctx, consts, 5, 6, &s_fm);
  unsigned cnt = 0;
  const long long q0 = ((long long)blockIdx.x * 256 + threadIdx.x) * pairs_per_thread;
  for (long long q = q0; q < q0 + pairs_per_thread; ++q) {
    Philox4 ph[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) ph[c] = philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), 7u, c, seed, 99u);
    if (STAGE == 0) {
      cnt += (ph[0].w0 ^ ph[1].w1 ^ ph[2].w2 ^ ph[0].w3 ^ ph[1].w0 ^ ph[2].w1) & 1u;
      continue;
    }
    double x0[6], x1[6];
    x0[0] = x1[0] = 0.8; x0[1] = x1[1] = 0.27;
    x0[2] = fma(u32w(ph[0].w0), 1.5, 7.0);
    x1[2] = fma(u32w(ph[0].w1), 1.5, 7.0);
    double z0, z1;
    box_muller_words(ph[0].w2, ph[0].w3, ph[1].w0, s_fm, z0, z1);
    x0[3] = fma(z0, 10.0, 431.7221); x1[3] = fma(z1, 10.0, 431.7221);
    box_muller_words(ph[1].w1, ph[1].w2, ph[1].w3, s_fm, z0, z1);
    x0[4] = fma(z0, 10.0, 1475.5503); x1[4] = fma(z1, 10.0, 1475.5503);
    box_muller_words(ph[2].w0, ph[2].w1, ph[2].w2, s_fm, z0, z1);
    x0[5] = fma(z0, 10.0, 55.0406); x1[5] = fma(z1, 10.0, 55.0406);
    if (STAGE == 1) {
      cnt += (unsigned)((x0[2] + x0[3] + x0[4] + x0[5]) < 1969.0) + (unsigned)((x1[2] + x1[3] + x1[4] + x1[5]) < 1969.0);
      continue;
    }
    cnt += (unsigned)LS::fails(ctx, x0) + (unsigned)LS::fails(ctx, x1);
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(out, (unsigned long long)cnt);
}

template <int STAGE>
void run(const char* name, int blocks_per_sm, const double* dconsts, unsigned long long* dout) {
  const long long ppt = 4000;
  const int blocks = 148 * blocks_per_sm;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  comp<STAGE><<<blocks, 256>>>(dout, dconsts, 100, 1u);
  cudaMemset(dout, 0, 8);
  cudaEventRecord(e0);
  comp<STAGE><<<blocks, 256>>>(dout, dconsts, ppt, 1u);
  cudaEventRecord(e1);
  unsigned long long h; cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double pair_warps_per_smsp = (double)ppt * blocks_per_sm * 8 / 4;
  printf("%-28s blocks/SM=%d  %.1f cycles per pair-warp per SMSP (@1.965GHz)  %.3e samples/s  count=%llu (%s)\n", name,
         blocks_per_sm, ms * 1e-3 * 1.965e9 / pair_warps_per_smsp, 2.0 * ppt * blocks * 256 / (ms * 1e-3), h,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const double M = 3.2633, m = 0.8158, g = 981.0;
  double k[5] = {M, m, g, pow(M * g, -1.5), pow(g * (M + m), 0.877)};
  double* dk; unsigned long long* dout;
  cudaMalloc(&dk, sizeof k); cudaMalloc(&dout, 8);
  cudaMemcpy(dk, k, sizeof k, cudaMemcpyHostToDevice);
  for (int b : {2, 3, 4, 6}) {
    if (b == 2) { run<0>("A: 3 x Philox", b, dk, dout); run<1>("B: A + 3 x Box-Muller", b, dk, dout); run<2>("C: B + fails() x 2", b, dk, dout); }
    if (b == 3) { run<0>("A: 3 x Philox", b, dk, dout); run<1>("B: A + 3 x Box-Muller", b, dk, dout); run<2>("C: B + fails() x 2", b, dk, dout); }
    if (b == 4) { run<0>("A: 3 x Philox", b, dk, dout); run<1>("B: A + 3 x Box-Muller", b, dk, dout); run<2>("C: B + fails() x 2", b, dk, dout); }
    if (b == 6) { run<0>("A: 3 x Philox", b, dk, dout); run<1>("B: A + 3 x Box-Muller", b, dk, dout); run<2>("C: B + fails() x 2", b, dk, dout); }
  }
  return 0;
}
