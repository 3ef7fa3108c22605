// Issue-rate microbenchmarks for the instruction mix of the Monte Carlo kernel (sm_100a).
// For each instruction kind, and for interleaved pairs of kinds, prints the measured
// cycles per warp-instruction per SM sub-partition. Results: profiles/r01_pipes_microbench.txt
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP8(x) x x x x x x x x

__constant__ double kc[8] = {1.0000001, 0.9999999, 1.0000002, 0.9999998, 1.0000003, 0.9999997, 1.0000004, 0.9999996};

enum { DFMA_R, DFMA_C, DMUL_R, DADD_R, DSETP_R, IMADW, LOP3_R, IMAD32, SHF_R, FFMA_R, MUFU_RSQ64, LDS_64, N_KINDS };
static const char* kNames[] = {"DFMA r,r,r", "DFMA r,c[],r", "DMUL", "DADD", "DSETP+SEL", "IMAD.WIDE.U32", "LOP3", "IMAD(32)", "SHF", "FFMA", "MUFU.RSQ64H", "LDS.64"};

template <int K>
__device__ __forceinline__ void op(double (&d)[8], uint32_t (&u)[8], uint64_t (&w)[8], float (&f)[8], int i,
                                   const double* sm, double cm, double ca, uint32_t ua, uint32_t ub, float fa, float fb) {
  if (K == DFMA_R) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(cm), "d"(ca));
  if (K == DFMA_C) d[i] = fma(d[i], kc[i], 1e-9);
  if (K == DMUL_R) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(cm));
  if (K == DADD_R) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(ca));
  if (K == DSETP_R) asm volatile("{ .reg .pred p; setp.lt.f64 p, %1, %2; selp.u32 %0, %0, %3, p; }" : "+r"(u[i]) : "d"(d[i]), "d"(d[(i + 1) & 7]), "r"(u[(i + 1) & 7]));
  if (K == IMADW) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(ua), "r"(ub));
  if (K == LOP3_R) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(ua), "r"(ub));
  if (K == IMAD32) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(ua), "r"(ub));
  if (K == SHF_R) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(u[i]) : "r"(ua));
  if (K == FFMA_R) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb));
  if (K == MUFU_RSQ64) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(d[i]));
  if (K == LDS_64) d[i] += sm[(threadIdx.x * 7 + i * 33 + (int)u[i]) & 255];
}

template <int A, int B, int NA, int NB>
__global__ void __launch_bounds__(256) bench(long long* cycles, double* sink, int iters) {
  __shared__ double sm[256];
  sm[threadIdx.x] = threadIdx.x * 1e-3;
  __syncthreads();
  double d[8]; uint32_t u[8]; uint64_t w[8]; float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { d[i] = 1.0 + 1e-3 * (threadIdx.x + i); u[i] = threadIdx.x * 2654435761u + i; w[i] = u[i]; f[i] = 1.0f + i; }
  const double cm = 0.9999999 + 1e-12 * threadIdx.x, ca = 1e-9 * (1 + threadIdx.x);
  const uint32_t ua = 0x9E3779B9u * (threadIdx.x + 1), ub = 0x85EBCA6Bu + threadIdx.x;
  const float fa = 0.999999f, fb = 1e-6f * threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int a = 0; a < NA; ++a) op<A>(d, u, w, f, i, sm, cm, ca, ua, ub, fa, fb);
#pragma unroll
        for (int b = 0; b < NB; ++b) op<B>(d, u, w, f, (i + 3) & 7, sm, cm, ca, ua, ub, fa, fb);
      }
    }
  }
  long long t1 = clock64();
  double s = 0; uint32_t x = 0; uint64_t y = 0; float g = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s += d[i]; x ^= u[i]; y ^= w[i]; g += f[i]; }
  if (s == 1.2345 || x == 77u || y == 99ull || g == 3.25f) sink[0] = s + x + y + g;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int A, int B, int NA, int NB>
void run(const char* label, int warps_per_smsp) {
  long long* cyc; double* sink;
  cudaMalloc(&cyc, 8); cudaMalloc(&sink, 8);
  const int iters = 2000;
  const int blocks_per_sm = warps_per_smsp * 4 * 32 / 256;
  bench<A, B, NA, NB><<<148 * blocks_per_sm, 256>>>(cyc, sink, 10);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bench<A, B, NA, NB><<<148 * blocks_per_sm, 256>>>(cyc, sink, iters);
  cudaEventRecord(e1);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double n_inst = (double)iters * 64 * (NA + NB) * warps_per_smsp;  // per SMSP
  printf("%-44s w/SMSP=%d  clock64: %.3f cyc/warp-inst/SMSP   events@1.965GHz: %.3f  (%s)\n", label, warps_per_smsp,
         h / n_inst, ms * 1e-3 * 1.965e9 / n_inst, cudaGetErrorString(e));
  cudaFree(cyc); cudaFree(sink);
}

#define SINGLE(K) { char b[96]; snprintf(b, 96, "%s", kNames[K]); run<K, K, 1, 0>(b, 8); }
#define PAIR(K1, K2, N1, N2) { char b[96]; snprintf(b, 96, "%d x %s + %d x %s", N1, kNames[K1], N2, kNames[K2]); run<K1, K2, N1, N2>(b, 8); }

template <int A, int B, int C3, int NA, int NB, int NC>
__global__ void __launch_bounds__(256) bench3(double* sink, int iters) {
  __shared__ double sm[256];
  sm[threadIdx.x] = threadIdx.x * 1e-3;
  __syncthreads();
  double d[8]; uint32_t u[8]; uint64_t w[8]; float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { d[i] = 1.0 + 1e-3 * (threadIdx.x + i); u[i] = threadIdx.x * 2654435761u + i; w[i] = u[i]; f[i] = 1.0f + i; }
  const double cm = 0.9999999 + 1e-12 * threadIdx.x, ca = 1e-9 * (1 + threadIdx.x);
  const uint32_t ua = 0x9E3779B9u * (threadIdx.x + 1), ub = 0x85EBCA6Bu + threadIdx.x;
  const float fa = 0.999999f, fb = 1e-6f * threadIdx.x;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int a = 0; a < NA; ++a) op<A>(d, u, w, f, i, sm, cm, ca, ua, ub, fa, fb);
#pragma unroll
        for (int b = 0; b < NB; ++b) op<B>(d, u, w, f, (i + 3) & 7, sm, cm, ca, ua, ub, fa, fb);
#pragma unroll
        for (int c = 0; c < NC; ++c) op<C3>(d, u, w, f, (i + 5) & 7, sm, cm, ca, ua, ub, fa, fb);
      }
    }
  }
  double s = 0; uint32_t x = 0; uint64_t y = 0; float g = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s += d[i]; x ^= u[i]; y ^= w[i]; g += f[i]; }
  if (s == 1.2345 || x == 77u || y == 99ull || g == 3.25f) sink[0] = s + x + y + g;
}

template <int A, int B, int C3, int NA, int NB, int NC>
void run3(int warps_per_smsp = 8) {
  double* sink; cudaMalloc(&sink, 8);
  const int iters = 4000;
  const int blocks_per_sm = warps_per_smsp * 4 * 32 / 256;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench3<A, B, C3, NA, NB, NC><<<148 * blocks_per_sm, 256>>>(sink, 10);
  cudaEventRecord(e0);
  bench3<A, B, C3, NA, NB, NC><<<148 * blocks_per_sm, 256>>>(sink, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double groups = (double)iters * 32 * warps_per_smsp;  // per SMSP
  printf("%d x %-14s + %d x %-14s + %d x %-14s : %.2f cycles per group per SMSP (%d instr)\n", NA, kNames[A], NB, kNames[B],
         NC, kNames[C3], ms * 1e-3 * 1.965e9 / groups, NA + NB + NC);
  cudaFree(sink);
}

int main() {
  run3<IMADW, IMADW, IMADW, 1, 0, 0>();
  run3<IMADW, LOP3_R, IMADW, 1, 1, 0>();
  run3<IMADW, DFMA_R, IMADW, 1, 1, 0>();
  run3<DFMA_R, LOP3_R, IMAD32, 1, 1, 0>();
  run3<DFMA_R, LOP3_R, IMAD32, 1, 1, 1>();
  run3<DFMA_R, LOP3_R, IMADW, 1, 1, 1>();
  run3<DFMA_R, LOP3_R, IMADW, 2, 1, 1>();
  run3<DFMA_R, LOP3_R, IMADW, 2, 2, 1>();
  run3<DFMA_R, LOP3_R, FFMA_R, 1, 1, 1>();
  run3<DFMA_R, LOP3_R, FFMA_R, 1, 1, 2>();
  run3<DFMA_C, LOP3_R, IMAD32, 1, 1, 1>();
  run3<DFMA_R, DSETP_R, IMAD32, 1, 1, 0>();
  run3<DSETP_R, DSETP_R, IMAD32, 1, 0, 0>();
  run3<LOP3_R, IMAD32, FFMA_R, 1, 1, 0>();
  run3<LOP3_R, SHF_R, FFMA_R, 1, 1, 0>();
  run3<LOP3_R, IMAD32, FFMA_R, 1, 1, 1>();
  return 0;
}
