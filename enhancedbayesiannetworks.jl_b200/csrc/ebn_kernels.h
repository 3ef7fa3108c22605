// Host-callable launchers of ebn_kernels.cu.
#pragma once
#include "ebn_rtc_compat.h"
#include "../../include/ebncuda.h"
#include "ebn_device.cuh"

namespace ebn {

// PLAN_STATIC0 + i: the i-th compile-time plan of the limit state (or, for a run-time compiled kernel,
// the job's own kinds)
enum PlanClass { PLAN_DYNAMIC = 0, PLAN_COPULA = 1, PLAN_STATIC0 = 2, PLAN_QMC = 100 };  // PLAN_STATIC0 + i, i < kMaxStaticPlans
constexpr int kMaxStaticPlans = 12;

#ifndef __CUDACC_RTC__
int threads_per_block();
// -1: no compile-time plan matches every scenario; i >= 0: the limit state's i-th plan does.
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
  const int* static_kinds[kMaxStaticPlans];  // compile-time sampling plans of the limit state
  int n_static[kMaxStaticPlans];
  int n_plans;
};
const LsVTable* ls_vtable(int limit_state);

cudaError_t moments_launch(const double* partial, int64_t local_items, int64_t items_per_scenario,
                           int shard_rank, int shard_world, int S, double* sum, double* sumsq,
                           cudaStream_t st);
cudaError_t dfma_launch(double* out, int iters, int blocks, cudaStream_t st);
#endif  // !__CUDACC_RTC__

}  // namespace ebn
