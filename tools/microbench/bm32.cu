// Cost of a Box-Muller normal pair (sm_100a), cycles per warp and SM sub-partition:
//   (A) the production fp64 transform (ebn_fastmath.cuh),
//   (B) fp32 transform, radius through I2F + MUFU.LG2 + MUFU.SQRT, angle through MUFU.SIN/COS,
//   (C) as B with the radius uniform placed by masks (23-bit, no I2F),
// each followed by the fp64 affine map x = mu + sigma z (the limit state stays in fp64).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../enhancedbayesiannetworks.jl_b200/csrc -o bm32 bm32.cu
#include <cstdio>
#include "ebn_sampling.cuh"

using namespace ebn;

template <int VAR>
__device__ __forceinline__ void bm32(uint32_t w0, uint32_t w1, double& z0, double& z1) {
  float u;
  if (VAR == 1) u = fmaf(__uint2float_rz(w0), 0x1p-32f, 0x1p-33f);
  else u = __uint_as_float(0x3F800000u | (w0 >> 9)) - (1.0f - 0x1p-24f);
  const float t = -1.3862943611198906f * __log2f(u);          // -2 ln u
  const float r = __fsqrt_rn(t);
  const float v = __uint_as_float(0x3F800000u | ((w1 >> 7) & 0x007FFFF8u)) - (1.0f - 0x1p-21f);  // (A+1/2) 2^-20
  float s, c;
  __sincosf(1.5707963267948966f * v, &s, &c);
  const float a = r * s, b = r * c;
  z0 = (double)__uint_as_float(__float_as_uint(a) ^ (w1 & 0x80000000u));
  z1 = (double)__uint_as_float(__float_as_uint(b) ^ ((w1 << 1) & 0x80000000u));
}

template <int VAR>
__global__ void __launch_bounds__(128) k(unsigned long long* out, long long ppt, unsigned seed) {
  __shared__ FastMathSmem s_fm;
  fastmath_init(s_fm, threadIdx.x, 128);
  __syncthreads();
  unsigned cnt = 0;
  uint32_t h = (blockIdx.x * 128 + threadIdx.x) * 2654435761u + seed;
  for (long long q = 0; q < ppt; ++q) {
    double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      h = h * 1664525u + 1013904223u;                 // cheap word source: the RNG is not measured here
      const uint32_t w0 = h ^ (h >> 15);
      h = h * 1664525u + 1013904223u;
      const uint32_t w1 = h ^ (h >> 13);
      double z0, z1;
      if (VAR == 0) box_muller_words(w0, w1, s_fm, z0, z1);
      else bm32<VAR>(w0, w1, z0, z1);
      acc0 = fma(z0, 10.0 + j, acc0 + 431.7);
      acc1 = fma(z1, 10.0 + j, acc1 + 431.7);
    }
    cnt += (unsigned)(acc0 < 1290.0) + (unsigned)(acc1 < 1290.0);
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(out, (unsigned long long)cnt);
}

template <int VAR>
void run(const char* name, unsigned long long* dout) {
  const long long ppt = 20000;
  const int blocks = 148 * 4;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<VAR><<<blocks, 128>>>(dout, 100, 1u);
  cudaMemset(dout, 0, 8);
  cudaEventRecord(e0);
  k<VAR><<<blocks, 128>>>(dout, ppt, 1u);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long c; cudaMemcpy(&c, dout, 8, cudaMemcpyDeviceToHost);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double pair_warps_per_smsp = (double)blocks * 4 /*warps*/ * ppt / (148.0 * 4.0);
  const double cycles = ms * 1e-3 * clk * 1e3 / pair_warps_per_smsp;
  printf("%-52s %8.3f ms  %7.1f cycles per pair-warp (3 normal pairs)  frac<1290 = %.5f\n", name, ms, cycles,
         (double)c / ((double)blocks * 128 * ppt * 2));
}

int main() {
  unsigned long long* dout; cudaMalloc(&dout, 8);
  run<0>("A fp64 Box-Muller (production)", dout);
  run<1>("B fp32, radius via I2F + LG2 + SQRT, SIN/COS", dout);
  run<2>("C fp32, radius by masks (23 bit) + LG2 + SQRT, SIN/COS", dout);
  return 0;
}
