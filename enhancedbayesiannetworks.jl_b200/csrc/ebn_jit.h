// Run-time compilation of kernels (NVRTC, loaded with dlopen) for
//   * expression limit states (EBN_LS_EXPRESSION): the functor is generated from the caller's program;
//   * sampling plans no compiled-in StaticPlan covers: the job's marginal kinds become a StaticPlan,
//     which removes the per-input dispatch of the DynamicPlan (3-4x on the truncated-normal jobs).
// The kernels are the same templates as the compiled-in ones (ebn_mc_impl.cuh, embedded in the
// library as text by the build); only the template arguments are chosen at run time.
#pragma once
#include <cstddef>
#include <string>
#include <vector>

#include <cuda_runtime.h>

namespace ebn {

enum JitRole { JIT_COUNT = 0, JIT_HIST = 1, JIT_REPLAY = 2, JIT_MATRIX = 3 };

struct JitSpec {
  int limit_state = 0;
  int d = 0;
  bool strict = false;
  bool copula = false;
  bool qmc = false;                // EBN_JOB_QMC_SOBOL: QmcPlan
  std::vector<int> kinds;          // d DevKinds -> StaticPlan<kinds...>; empty -> DynamicPlan / CopulaPlan
  const double* consts = nullptr;  // expression program (EBN_LS_EXPRESSION only)
  int n_consts = 0;
};

struct JitKernel;  // cubin + entry name + the functions loaded per device

// Is libnvrtc loadable AND able to generate sm_100a code (one-off probe compile)? No GPU needed.
bool jit_available(std::string* why);

// Compiles the kernel, or finds it in the in-memory cache or (EBN_JIT_CACHE=dir) on disk. No GPU needed.
// Returns false with *err set (compiler log included) on failure. *compiled = NVRTC ran.
bool jit_get(const JitSpec& spec, JitRole role, JitKernel** out, bool* compiled, std::string* err);
size_t jit_cubin_bytes(const JitKernel* k);

// Device side (driver API through cudaGetDriverEntryPoint; the current device is `device`).
bool jit_occupancy(JitKernel* k, int device, int threads, size_t dyn_smem, int* blocks_per_sm, std::string* err);
bool jit_launch(JitKernel* k, int device, int blocks, int threads, size_t dyn_smem, cudaStream_t st,
                void** args, std::string* err);
// Unloads every module (ebn_shutdown); compiled code stays cached.
void jit_unload_all();

}  // namespace ebn
