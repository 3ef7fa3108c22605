// Registry of the per-limit-state launch tables plus the two kernels that do
// not depend on a limit state (moment reduction, FP64 issue-rate probe).
#include "ebn_kernels.h"

namespace ebn {

constexpr int kThreads = EBN_THREADS;

// Fixed-order reduction of the per-item moment partials: one thread per
// scenario walks its items in order, so the result does not depend on which
// block processed which item.
__global__ void moments_kernel(const double* partial, int64_t local_items,
                               int64_t items_per_scenario, int shard_rank, int shard_world,
                               int S, double* sum, double* sumsq) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  double a = 0.0, b = 0.0;
  // items of scenario s: w in [s*ips, (s+1)*ips), local iff w % world == rank
  const int64_t w0 = (int64_t)s * items_per_scenario, w1 = w0 + items_per_scenario;
  int64_t w = w0 + ((shard_rank - (w0 % shard_world)) + shard_world) % shard_world;
  for (; w < w1; w += shard_world) {
    const int64_t li = (w - shard_rank) / shard_world;
    if (li < local_items) {
      a += partial[2 * li];
      b += partial[2 * li + 1];
    }
  }
  sum[s] = a;
  sumsq[s] = b;
}

// FP64 issue-rate probe: 8 independent DFMA chains per thread.
__global__ void __launch_bounds__(kThreads) dfma_kernel(double* out, int iters, double seed) {
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5,
         a6 = seed + 6, a7 = seed + 7;
  const double m = 0.9999999, c = 1e-9 * threadIdx.x;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (r == 123.456) out[0] = r;  // keeps the chains alive, practically never taken
}


int threads_per_block() { return kThreads; }

LsVTable vt_vehicle();
LsVTable vt_straub();
LsVTable vt_hydrogen();
LsVTable vt_polynomial();
LsVTable vt_min_affine();
LsVTable vt_straub_r45();
LsVTable vt_straub_capacity();

const LsVTable* ls_vtable(int limit_state) {
  static const LsVTable v1 = vt_vehicle(), v2 = vt_straub(), v3 = vt_hydrogen(),
                        v4 = vt_polynomial(), v5 = vt_min_affine(), v6 = vt_straub_r45(),
                        v7 = vt_straub_capacity();
  switch (limit_state) {
    case EBN_LS_VEHICLE_SUSPENSION: return &v1;
    case EBN_LS_STRAUB_FRAME: return &v2;
    case EBN_LS_HYDROGEN_RADIUS: return &v3;
    case EBN_LS_POLYNOMIAL: return &v4;
    case EBN_LS_MIN_AFFINE: return &v5;
    case EBN_LS_STRAUB_FRAME_R45: return &v6;
    case EBN_LS_STRAUB_CAPACITY: return &v7;
  }
  return nullptr;
}

int static_plan_matches(int limit_state, const DevInput* plan, int S, int d) {
  const LsVTable* v = ls_vtable(limit_state);
  if (!v) return -1;
  for (int which = 0; which < v->n_plans; ++which) {
    if (v->n_static[which] != d) continue;
    bool ok = true;
    for (int s = 0; s < S && ok; ++s)
      for (int j = 0; j < d; ++j)
        if (plan[(int64_t)s * d + j].kind != v->static_kinds[which][j]) { ok = false; break; }
    if (ok) return which;
  }
  return -1;
}

cudaError_t moments_launch(const double* partial, int64_t local_items, int64_t items_per_scenario,
                           int shard_rank, int shard_world, int S, double* sum, double* sumsq,
                           cudaStream_t st) {
  moments_kernel<<<(S + 127) / 128, 128, 0, st>>>(partial, local_items, items_per_scenario,
                                                  shard_rank, shard_world, S, sum, sumsq);
  return cudaGetLastError();
}

cudaError_t dfma_launch(double* out, int iters, int blocks, cudaStream_t st) {
  dfma_kernel<<<blocks, kThreads, 0, st>>>(out, iters, 1.0);
  return cudaGetLastError();
}

}  // namespace ebn
