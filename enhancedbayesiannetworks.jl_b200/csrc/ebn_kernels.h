// Host-callable launchers of ebn_kernels.cu.
#pragma once
#include "ebn_rtc_compat.h"
#include "../../include/ebncuda.h"
#include "ebn_device.cuh"

namespace ebn {

enum PlanClass { PLAN_DYNAMIC = 0, PLAN_STATIC = 1, PLAN_COPULA = 2, PLAN_STATIC2 = 3 };

#ifndef __CUDACC_RTC__
int threads_per_block();
// 0: no compile-time plan matches every scenario; 1 / 2: the limit state's first / second plan does.
int static_plan_matches(int limit_state, const DevInput* plan, int S, int d);

// Launchers of one limit state (filled by its translation unit).
struct LsVTable {
  cudaError_t (*occupancy)(int strict, int plan_class, int mode, size_t dyn_smem, int* blocks_per_sm);
  cudaError_t (*launch)(int strict, int plan_class, int mode, const McParams& p, int blocks,
                        size_t dyn_smem, cudaStream_t st);
  cudaError_t (*replay)(int strict, int plan_class, const McParams& p, int scenario, int64_t n,
                        double* X, double* g_out, int blocks, cudaStream_t st);
  cudaError_t (*matrix)(int strict, const double* consts, int n_consts, const double* X, int64_t n,
                        int d, int64_t ld, unsigned long long* count, double* g_out, int blocks,
                        cudaStream_t st);
  const int* static_kinds;  // compile-time sampling plan, if the limit state has one
  int n_static;
  const int* static_kinds2;  // an optional second one
  int n_static2;
};
const LsVTable* ls_vtable(int limit_state);

cudaError_t moments_launch(const double* partial, int64_t local_items, int64_t items_per_scenario,
                           int shard_rank, int shard_world, int S, double* sum, double* sumsq,
                           cudaStream_t st);
cudaError_t dfma_launch(double* out, int iters, int blocks, cudaStream_t st);
#endif  // !__CUDACC_RTC__

}  // namespace ebn
