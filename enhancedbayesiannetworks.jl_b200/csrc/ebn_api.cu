// C ABI of libebncuda.so (see include/ebncuda.h). Host side only: argument
// validation, canonicalisation of marginals into device sampling plans, device
// buffers, launches, the NCCL all-reduce of the count vectors, result copy.
//
// The reference call sites this file stands behind:
//   ebn_pf_batch    <- probability_of_failure(...) per scenario,
//                      src/networks/ebn/evaluation/evaluate_node.jl:79-95
//   ebn_hist_batch  <- UQ.sample / UQ.evaluate! / EmpiricalDistribution,
//                      src/networks/ebn/evaluation/evaluate_node.jl:22-33
//   ebn_eval_matrix <- `performance(df) .< 0` summed (UQ.jl, called from :83)
#include <dlfcn.h>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <vector>

#include "ebn_expr.h"
#include "ebn_jit.h"
#include "ebn_kernels.h"
#include "ebn_build_id.inc"

static_assert(sizeof(ebn_input_t) == 56, "ebn_input_t must be 56 bytes");
static_assert(sizeof(ebn_stats_t) == 48, "ebn_stats_t must be 48 bytes");
static_assert(offsetof(ebn_job_t, consts) == 8, "ebn_job_t layout");
static_assert(offsetof(ebn_job_t, inputs) == 24, "ebn_job_t layout");
static_assert(offsetof(ebn_job_t, n_per_scenario) == 32, "ebn_job_t layout");
static_assert(offsetof(ebn_job_t, corr_chol) == 56, "ebn_job_t layout");
static_assert(sizeof(ebn_job_t) == 104, "ebn_job_t must be 104 bytes");

namespace {

using namespace ebn;

thread_local std::string g_err;
// Every entry point that touches the library's state (devices, staging buffers, communicators, statistics)
// takes this lock: calls from several host threads are serialised, never interleaved. Recursive because
// entry points call one another (ebn_init -> ebn_shutdown, ebn_sample_matrix -> ebn_sample_matrix_ex).
std::recursive_mutex g_api_mu;
#define API_LOCK std::lock_guard<std::recursive_mutex> api_lock_(g_api_mu)

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                       \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess)                                                            \
      return fail(EBN_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),  \
                  __FILE__, __LINE__);                                                 \
  } while (0)

// ---- NCCL through dlopen: no link-time dependency, and inside a process that
// already loaded a libnccl.so.2 (torch) the same copy is reused. -------------
typedef struct ncclComm* ncclComm_t;
struct NcclId { char b[128]; };  // ncclUniqueId: 128 opaque bytes, passed by value
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclId, int) = nullptr;
  int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
} g_nccl;
constexpr int kNcclInt64 = 4, kNcclFloat64 = 8, kNcclSum = 0;

int nccl_load() {
  if (g_nccl.handle) return EBN_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return fail(EBN_ENCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                       \
  *(void**)(&g_nccl.field) = dlsym(h, name);                                   \
  if (!g_nccl.field) return fail(EBN_ENCCL, "NCCL symbol %s not found", name);
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitAll, "ncclCommInitAll")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllReduce, "ncclAllReduce")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
  SYM(CommInitRank, "ncclCommInitRank")
#undef SYM
  g_nccl.handle = h;
  return EBN_OK;
}
#define NC(call)                                                                  \
  do {                                                                            \
    int r__ = (call);                                                             \
    if (r__ != 0)                                                                 \
      return fail(EBN_ENCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r__)); \
  } while (0)

// ---- per-device state --------------------------------------------------------
struct Buf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return EBN_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes < 4096 ? 4096 : bytes;
    if (cudaMalloc(&p, want) != cudaSuccess) {
      cudaGetLastError();
      return fail(EBN_ENOMEM, "cudaMalloc(%zu) failed", want);
    }
    cap = want;
    return EBN_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct Dev {
  int id = 0;
  int sms = 0;
  cudaStream_t own = nullptr;
  cudaStream_t st = nullptr;  // stream in use (own or caller's)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  Buf plan, consts, tables, chol, streams, counts, edges, partial, sums, mat, gout;
  ncclComm_t comm = nullptr;  // single-process multi-device communicator
};

struct Ctx {
  bool ready = false;
  std::vector<Dev> devs;
  ncclComm_t rank_comm = nullptr;  // one-process-per-GPU communicator
  int rank = 0, world = 1;
  ebn_stats_t stats{};
} g;

// ---- host-side distribution helpers (truncation windows -> probability) -----
const double kInf = std::numeric_limits<double>::infinity();

double norm_cdf(double z) { return 0.5 * erfc(-z * 0.70710678118654752440); }

// Maps one ABI input to the device recipe. Returns EBN_OK or an error.
int canonicalise(const ebn_input_t& in, bool copula, bool fp32, DevInput& o, int s, int j,
                 int64_t n_tables) {
  o = DevInput{};
  const double lo = in.lo, hi = in.hi;
  const bool trunc = (lo > -kInf) || (hi < kInf);
  if (in.flags != 0) return fail(EBN_EINVAL, "input (%d,%d): flags must be 0", s, j);
  if (std::isnan(lo) || std::isnan(hi) || !(lo < hi)) {
    if (in.kind != EBN_IN_PARAM)
      return fail(EBN_EMARGINAL, "input (%d,%d): truncation window [%g,%g] is empty", s, j, lo, hi);
  }
  double Flo = 0.0, Fhi = 1.0;
  switch (in.kind) {
    case EBN_IN_PARAM:
      o.kind = DK_CONST;
      o.a = in.p[0];
      return EBN_OK;
    case EBN_IN_UNIFORM: {
      double a = in.p[0], b = in.p[1];
      if (!(a < b)) return fail(EBN_EMARGINAL, "input (%d,%d): Uniform needs a < b", s, j);
      if (lo > a) a = lo;
      if (hi < b) b = hi;
      if (!(a < b))
        return fail(EBN_EMARGINAL, "input (%d,%d): truncation leaves no Uniform support", s, j);
      o.kind = DK_UNIFORM;
      o.a = a;
      o.b = b - a;
      o.c = a - (b - a) * (1.0 - 0x1p-33);  // x = fma(d, b, c) for the raw [1,2) double d
      o.d = 1.0;
      return EBN_OK;
    }
    case EBN_IN_NORMAL:
    case EBN_IN_LOGNORMAL: {
      const double mu = in.p[0], sd = in.p[1];
      if (!(sd > 0.0)) return fail(EBN_EMARGINAL, "input (%d,%d): sigma must be > 0", s, j);
      const bool logn = in.kind == EBN_IN_LOGNORMAL;
      o.a = mu;
      o.b = sd;
      if (!trunc) {
        o.kind = (fp32 && !copula) ? (logn ? DK_LOGNORMAL_BM32 : DK_NORMAL_BM32)
                                   : (logn ? DK_LOGNORMAL_BM : DK_NORMAL_BM);
        o.c = 0.0;
        o.d = 1.0;
        return EBN_OK;
      }
      double zl = -kInf, zh = kInf;
      if (logn) {
        if (lo > 0.0) zl = (log(lo) - mu) / sd;
        if (hi < kInf) { if (!(hi > 0.0)) return fail(EBN_EMARGINAL, "input (%d,%d): LogNormal window below 0", s, j); zh = (log(hi) - mu) / sd; }
      } else {
        if (lo > -kInf) zl = (lo - mu) / sd;
        if (hi < kInf) zh = (hi - mu) / sd;
      }
      Flo = zl == -kInf ? 0.0 : norm_cdf(zl);
      Fhi = zh == kInf ? 1.0 : norm_cdf(zh);
      o.kind = logn ? DK_LOGNORMAL_ICDF : DK_NORMAL_ICDF;
      break;
    }
    case EBN_IN_GUMBEL: {
      const double mu = in.p[0], beta = in.p[1];
      if (!(beta > 0.0)) return fail(EBN_EMARGINAL, "input (%d,%d): beta must be > 0", s, j);
      o.kind = DK_GUMBEL;
      o.a = mu;
      o.b = beta;
      if (lo > -kInf) Flo = exp(-exp(-(lo - mu) / beta));
      if (hi < kInf) Fhi = exp(-exp(-(hi - mu) / beta));
      break;
    }
    case EBN_IN_EXPONENTIAL_AFFINE: {
      const double theta = in.p[0], sign = in.p[1], shift = in.p[2];
      if (!(theta > 0.0) || !(sign == 1.0 || sign == -1.0))
        return fail(EBN_EMARGINAL, "input (%d,%d): Exponential needs theta > 0, sign +-1", s, j);
      o.kind = DK_EXP_AFFINE;
      o.a = shift;
      o.b = sign * theta;
      // window in the exponential's own variable e = (x - shift)/sign >= 0
      double el = 0.0, eh = kInf;
      if (sign > 0) { if (lo > shift) el = lo - shift; if (hi < kInf) eh = hi - shift; }
      else { if (hi < shift) el = shift - hi; if (lo > -kInf) eh = shift - lo; }
      if (!(eh > el)) return fail(EBN_EMARGINAL, "input (%d,%d): window outside the support", s, j);
      Flo = -expm1(-el / theta);
      Fhi = eh == kInf ? 1.0 : -expm1(-eh / theta);
      break;
    }
    case EBN_IN_GAMMA: {
      if (!(in.p[0] > 0.0) || !(in.p[1] > 0.0))
        return fail(EBN_EMARGINAL, "input (%d,%d): Gamma needs shape, scale > 0", s, j);
      if (trunc || copula)
        return fail(EBN_EMARGINAL,
                    "input (%d,%d): truncated / copula Gamma must be passed as EBN_IN_TABULATED", s, j);
      o.kind = DK_GAMMA_MT;
      o.a = in.p[0];
      o.b = in.p[1];
      {  // Marsaglia-Tsang constants: d = shape' - 1/3, c = 1/sqrt(9 d), shape' = shape (+1 if < 1)
        const double sh = in.p[0] < 1.0 ? in.p[0] + 1.0 : in.p[0];
        o.c = sh - 1.0 / 3.0;
        o.d = 1.0 / sqrt(9.0 * o.c);
      }
      return EBN_OK;
    }
    case EBN_IN_TABULATED: {
      const double off = in.p[0], npts = in.p[1];
      if (!(npts >= 2.0) || off < 0.0 || off + npts > (double)n_tables)
        return fail(EBN_EINVAL, "input (%d,%d): table [%g,+%g) outside the pool of %lld", s, j, off,
                    npts, (long long)n_tables);
      if (trunc)
        return fail(EBN_EMARGINAL, "input (%d,%d): truncate a tabulated marginal in the table", s, j);
      o.kind = DK_TABLE;
      o.toff = (int32_t)off;
      o.npts = (int32_t)npts;
      o.c = 0.0;
      o.d = 1.0;
      return EBN_OK;
    }
    default:
      return fail(EBN_EMARGINAL, "input (%d,%d): unknown marginal kind %d", s, j, in.kind);
  }
  if (!(Fhi > Flo))
    return fail(EBN_EMARGINAL, "input (%d,%d): truncation window has zero probability", s, j);
  o.c = Flo;
  o.d = Fhi - Flo;
  return EBN_OK;
}

struct LsInfo { int id; const char* name; int d; int nc; };
const LsInfo kLs[] = {
    {EBN_LS_VEHICLE_SUSPENSION, "vehicle_suspension", 6, 5},
    {EBN_LS_STRAUB_FRAME, "straub_frame", 8, 3},
    {EBN_LS_HYDROGEN_RADIUS, "hydrogen_radius", 4, 0},
    {EBN_LS_POLYNOMIAL, "polynomial", -1, -1},
    {EBN_LS_MIN_AFFINE, "min_affine", -1, -1},
    {EBN_LS_STRAUB_FRAME_R45, "straub_frame_r45", 8, 3},
    {EBN_LS_STRAUB_CAPACITY, "straub_capacity", 2, 3},
    {EBN_LS_EXPRESSION, "expression", -1, -1},
};
const LsInfo* ls_info(int id) {
  for (const auto& l : kLs)
    if (l.id == id) return &l;
  return nullptr;
}

// Validates limit state, arity and the consts program.
int check_limit_state(int id, const double* consts, int n_consts, int d) {
  const LsInfo* li = ls_info(id);
  if (!li) return fail(EBN_ELIMITSTATE, "unknown limit state id %d (no CPU fallback exists)", id);
  if (d < 1 || d > EBN_MAX_INPUTS)
    return fail(EBN_EINVAL, "n_inputs %d outside [1,%d]", d, EBN_MAX_INPUTS);
  if (n_consts < 0 || (n_consts > 0 && !consts)) return fail(EBN_EINVAL, "consts is NULL");
  if (li->d >= 0 && li->d != d)
    return fail(EBN_ELIMITSTATE, "limit state %s takes %d inputs, got %d", li->name, li->d, d);
  if (li->nc >= 0 && li->nc != n_consts)
    return fail(EBN_ELIMITSTATE, "limit state %s takes %d consts, got %d", li->name, li->nc, n_consts);
  if (id == EBN_LS_POLYNOMIAL) {
    if (n_consts < 1) return fail(EBN_ELIMITSTATE, "polynomial program is empty");
    const int T = (int)consts[0];
    if (T < 1 || d + T > EBN_MAX_INPUTS + 1)
      return fail(EBN_ELIMITSTATE, "polynomial program: %d stages on %d inputs exceed %d columns", T, d,
                  EBN_MAX_INPUTS);
    int pos = 1;
    for (int t = 0; t < T; ++t) {
      if (pos >= n_consts) return fail(EBN_ELIMITSTATE, "polynomial program truncated (stage %d)", t);
      const int nt = (int)consts[pos++];
      if (nt < 1) return fail(EBN_ELIMITSTATE, "polynomial stage %d has no terms", t);
      for (int k = 0; k < nt; ++k) {
        pos++;  // coefficient
        if (pos >= n_consts)
          return fail(EBN_ELIMITSTATE, "polynomial program truncated (stage %d term %d)", t, k);
        const int nf = (int)consts[pos++];
        if (nf < 0 || pos + nf > n_consts)
          return fail(EBN_ELIMITSTATE, "polynomial program truncated (stage %d term %d)", t, k);
        for (int f = 0; f < nf; ++f) {
          const int col = (int)consts[pos++];
          if (col < 0 || col >= d + t)
            return fail(EBN_ELIMITSTATE, "polynomial stage %d reads column %d (only %d exist)", t, col,
                        d + t);
        }
      }
    }
    if (pos != n_consts)
      return fail(EBN_ELIMITSTATE, "polynomial program has %d trailing consts", n_consts - pos);
  } else if (id == EBN_LS_EXPRESSION) {
    const std::string bad = expr_check(consts, n_consts, d, nullptr);
    if (!bad.empty()) return fail(EBN_ELIMITSTATE, "%s", bad.c_str());
  } else if (id == EBN_LS_MIN_AFFINE) {
    if (n_consts < 1) return fail(EBN_ELIMITSTATE, "min_affine needs consts");
    const int K = (int)consts[0];
    if (K < 1 || n_consts != 1 + K * (1 + d))
      return fail(EBN_ELIMITSTATE, "min_affine: expected %d consts for K=%d, d=%d, got %d",
                  1 + K * (1 + d), K, d, n_consts);
  }
  return EBN_OK;
}

int require_ready() {
  if (!g.ready) return fail(EBN_ENODEV, "ebn_init has not been called (or found no CUDA device)");
  return EBN_OK;
}

// A fully prepared batch: validated job + canonical plan + partition.
struct Prepared {
  std::vector<DevInput> plan;
  int plan_class = PLAN_DYNAMIC;
  int64_t first = 0, last = 0, pair_base = 0, pairs_per_item = 0, items_per_scenario = 0;
  int world = 1;  // job world * devices
  // run-time compiled kernels (expression limit states, sampling plans without a compiled-in plan)
  // the kernel dispatches marginal kinds at run time, or samples under a copula, and some scenario draws a
  // truncated normal: the launch carries the Phi^-1 cell table in dynamic shared memory (McParams::icdf_dyn)
  bool icdf_dyn = false;
  bool jit = false;
  JitSpec jit_spec;
  JitKernel* jit_kernel[4] = {nullptr, nullptr, nullptr, nullptr};  // by JitRole
  int jit_compiles = 0;
};

// Jobs below this many samples keep the runtime-dispatched plan. Measured on B200 (profiles/
// r01_k_workloads.txt): 3.6e10 samples/s dispatched vs 7.9e10 compiled, 1-3 s to compile a plan the
// first time it is seen (then cached in memory, and on disk when EBN_JIT_CACHE names a directory):
// break-even near 1e11 samples.
constexpr double kJitPlanMinSamples = 1e11;
// Under a Gaussian copula the looping plan keeps scores, columns and words in local memory and runs at
// 1.5e10 samples/s against 3.3e10 for the unrolled plan (profiles/r02_af_workloads.txt): break-even near 5e10,
// below BASELINE config 3's 6.4e10.
constexpr double kJitCopulaMinSamples = 5e10;

int jit_kernel_of(Prepared& P, JitRole role, JitKernel** out) {
  if (!P.jit_kernel[role]) {
    bool compiled = false;
    std::string err;
    if (!jit_get(P.jit_spec, role, &P.jit_kernel[role], &compiled, &err))
      return fail(EBN_EJIT, "%s", err.c_str());
    if (compiled) ++P.jit_compiles;
  }
  *out = P.jit_kernel[role];
  return EBN_OK;
}

// The four kernel entry points of a job, compiled-in or run-time compiled.
int k_occupancy(Dev& dv, const ebn_job_t* job, Prepared& P, int mode, size_t dyn, int* per_sm) {
  if (!P.jit) {
    CU(ls_vtable(job->limit_state)->occupancy(job->strict ? 1 : 0, P.plan_class, mode, dyn, per_sm));
    return EBN_OK;
  }
  JitKernel* k;
  int rc = jit_kernel_of(P, mode == 0 ? JIT_COUNT : JIT_HIST, &k);
  if (rc) return rc;
  std::string err;
  if (!jit_occupancy(k, dv.id, threads_per_block(), dyn, per_sm, &err)) return fail(EBN_ECUDA, "%s", err.c_str());
  return EBN_OK;
}

int k_launch(Dev& dv, const ebn_job_t* job, Prepared& P, int mode, const McParams& mp, int blocks, size_t dyn) {
  if (!P.jit) {
    CU(ls_vtable(job->limit_state)->launch(job->strict ? 1 : 0, P.plan_class, mode, mp, blocks, dyn, dv.st));
    return EBN_OK;
  }
  JitKernel* k;
  int rc = jit_kernel_of(P, mode == 0 ? JIT_COUNT : JIT_HIST, &k);
  if (rc) return rc;
  McParams arg = mp;
  void* args[] = {&arg};
  std::string err;
  if (!jit_launch(k, dv.id, blocks, threads_per_block(), dyn, dv.st, args, &err)) return fail(EBN_ECUDA, "%s", err.c_str());
  return EBN_OK;
}

int k_replay(Dev& dv, const ebn_job_t* job, Prepared& P, const McParams& mp, int scenario, int64_t n, double* X,
             double* g_out, int blocks) {
  if (!P.jit) {
    CU(ls_vtable(job->limit_state)->replay(job->strict ? 1 : 0, P.plan_class, mp, scenario, n, X, g_out, blocks, dv.st));
    return EBN_OK;
  }
  JitKernel* k;
  int rc = jit_kernel_of(P, JIT_REPLAY, &k);
  if (rc) return rc;
  McParams arg = mp;
  void* args[] = {&arg, &scenario, &n, &X, &g_out};
  std::string err;
  if (!jit_launch(k, dv.id, blocks, threads_per_block(), 0, dv.st, args, &err)) return fail(EBN_ECUDA, "%s", err.c_str());
  return EBN_OK;
}

int prepare(const ebn_job_t* job, int mode, Prepared& P, int ndev_override = 0) {
  if (!job) return fail(EBN_EINVAL, "job is NULL");
  if (job->n_scenarios < 1) return fail(EBN_EINVAL, "n_scenarios must be >= 1");
  if (!job->inputs) return fail(EBN_EINVAL, "inputs is NULL");
  if (job->n_per_scenario < 1) return fail(EBN_EINVAL, "n_per_scenario must be >= 1");
  if (job->sample_offset < 0) return fail(EBN_EINVAL, "sample_offset must be >= 0");
  if (job->flags & ~(EBN_JOB_COPULA_SHARED | EBN_JOB_PARTIAL | EBN_JOB_JIT_PLAN | EBN_JOB_NO_JIT_PLAN | EBN_JOB_FP32_SAMPLING | EBN_JOB_QMC_SOBOL)) return fail(EBN_EINVAL, "unknown job flags 0x%x", job->flags);
  int rc = check_limit_state(job->limit_state, job->consts, job->n_consts, job->n_inputs);
  if (rc) return rc;
  const int jw = job->shard_world <= 1 ? 1 : job->shard_world;
  const int jr = job->shard_world <= 1 ? 0 : job->shard_rank;
  if (jr < 0 || jr >= jw) return fail(EBN_EINVAL, "shard_rank %d outside [0,%d)", jr, jw);
  if (job->n_tables < 0 || (job->n_tables > 0 && !job->tables)) return fail(EBN_EINVAL, "tables is NULL");
  const int S = job->n_scenarios, d = job->n_inputs;
  const bool copula = job->corr_chol != nullptr;
  const bool qmc = (job->flags & EBN_JOB_QMC_SOBOL) != 0;
  if (qmc) {
    if (copula) return fail(EBN_EINVAL, "EBN_JOB_QMC_SOBOL: a Gaussian copula is not supported with the Sobol sequence");
    if ((double)job->sample_offset + (double)job->n_per_scenario > 4294967296.0)
      return fail(EBN_EINVAL, "EBN_JOB_QMC_SOBOL: the sequence holds 2^32 points per scenario");
  }
  P.plan.resize((size_t)S * d);
  for (int s = 0; s < S; ++s)
    for (int j = 0; j < d; ++j) {
      rc = canonicalise(job->inputs[(size_t)s * d + j], copula, (job->flags & EBN_JOB_FP32_SAMPLING) != 0,
                        P.plan[(size_t)s * d + j], s, j, job->n_tables);
      if (rc) return rc;
      if (qmc && P.plan[(size_t)s * d + j].kind == DK_GAMMA_MT)
        return fail(EBN_EMARGINAL, "input (%d,%d): Gamma has no closed-form quantile on the device; under "
                    "EBN_JOB_QMC_SOBOL pass it as an inverse-CDF table (EBN_IN_TABULATED)", s, j);
    }
  // word offsets in the pair's Philox stream (layout v3): inputs consume words in order;
  // under a copula every random input takes the 2 words of its Box-Muller pair
  for (int s = 0; s < S; ++s) {
    int off = 0;
    for (int j = 0; j < d; ++j) {
      DevInput& di = P.plan[(size_t)s * d + j];
      di.woff = off;
      off += copula ? (di.kind == DK_CONST ? 0 : 2) : words_of_kind(di.kind);
    }
  }
  if (copula) {
    const int nblk = (job->flags & EBN_JOB_COPULA_SHARED) ? 1 : S;
    for (int b = 0; b < nblk; ++b) {
      const double* L = job->corr_chol + (size_t)b * d * d;
      for (int r = 0; r < d; ++r) {
        double ss = 0.0;
        for (int c = 0; c < d; ++c) {
          if (c > r && L[r * d + c] != 0.0)
            return fail(EBN_EINVAL, "corr_chol block %d is not lower triangular", b);
          ss += L[r * d + c] * L[r * d + c];
        }
        if (fabs(ss - 1.0) > 1e-9)
          return fail(EBN_EINVAL, "corr_chol block %d row %d: L*L' has diagonal %g, expected 1", b, r, ss);
      }
    }
    P.plan_class = PLAN_COPULA;
  } else if (qmc) {
    P.plan_class = PLAN_QMC;
  } else {
    const int which = static_plan_matches(job->limit_state, P.plan.data(), S, d);
    if (which >= 0) P.plan_class = PLAN_STATIC0 + which;
  }
  // Run-time compilation: always for an expression limit state; for the others when no compiled-in
  // plan fits, every scenario has the same marginal kinds, and the job is large (or asks for it).
  {
    const bool expr = job->limit_state == EBN_LS_EXPRESSION;
    bool uniform = true;
    for (int s = 1; s < S && uniform; ++s)
      for (int j = 0; j < d; ++j)
        if (P.plan[(size_t)s * d + j].kind != P.plan[j].kind) { uniform = false; break; }
    const bool never = (job->flags & EBN_JOB_NO_JIT_PLAN) != 0;
    const bool forced = (job->flags & EBN_JOB_JIT_PLAN) != 0;
    if (never && forced) return fail(EBN_EINVAL, "EBN_JOB_JIT_PLAN and EBN_JOB_NO_JIT_PLAN are exclusive");
    bool plan_jit = false;
    if (!expr && !qmc && (P.plan_class == PLAN_DYNAMIC || P.plan_class == PLAN_COPULA) && uniform && !never) {
      const char* env = getenv("EBN_JIT_PLAN");  // "0": never, "1": always (experiments)
      const bool big = (double)job->n_per_scenario * S >= (copula ? kJitCopulaMinSamples : kJitPlanMinSamples);
      if (forced || (env && env[0] == '1')) {
        std::string why;
        if (!jit_available(&why)) return fail(EBN_EJIT, "EBN_JOB_JIT_PLAN: %s", why.c_str());
        plan_jit = true;
      } else if (big && !(env && env[0] == '0')) {
        plan_jit = jit_available(nullptr);  // optional speed-up: without NVRTC the dynamic plan runs
      }
    }
    if (expr || plan_jit) {
      P.jit = true;
      P.jit_spec.limit_state = job->limit_state;
      P.jit_spec.d = d;
      P.jit_spec.strict = job->strict != 0;
      P.jit_spec.copula = copula;
      P.jit_spec.qmc = qmc;
      P.jit_spec.consts = job->consts;
      P.jit_spec.n_consts = job->n_consts;
      if (uniform && !(expr && never) && !qmc)
        for (int j = 0; j < d; ++j) P.jit_spec.kinds.push_back(P.plan[j].kind);
      if (!P.jit_spec.kinds.empty() && !copula) P.plan_class = PLAN_STATIC0;
    }
  }
  if (P.plan_class == PLAN_DYNAMIC || P.plan_class == PLAN_COPULA) {
    for (const DevInput& in : P.plan)
      if (in.kind == DK_NORMAL_ICDF || in.kind == DK_LOGNORMAL_ICDF) { P.icdf_dyn = true; break; }
  }
  P.world = jw * (ndev_override > 0 ? ndev_override : (int)g.devs.size());
  P.first = job->sample_offset;
  P.last = job->sample_offset + job->n_per_scenario;
  P.pair_base = P.first >> 1;
  const int64_t pairs = ((P.last + 1) >> 1) - P.pair_base;
  // Partition depends on the job only (never on the device), so every rank
  // derives the same item grid.
  const int T = threads_per_block();
  const double target_items = 9472.0 * P.world;  // 148 SMs * 4 * 16
  int64_t ppi = (int64_t)ceil((double)pairs * S / target_items);
  ppi = ((ppi + T - 1) / T) * T;
  const int64_t max_ppi = mode == 0 ? (1 << 15) : (1 << 22);
  if (ppi < T) ppi = T;
  if (ppi > max_ppi) ppi = max_ppi;
  P.pairs_per_item = ppi;
  P.items_per_scenario = (pairs + ppi - 1) / ppi;
  return EBN_OK;
}

int64_t local_items_of(const Prepared& P, int S, int rank) {
  const int64_t total = P.items_per_scenario * S;
  return total > rank ? (total - rank + P.world - 1) / P.world : 0;
}

// Uploads the job to one device and fills the kernel parameters.
int stage(Dev& dv, const ebn_job_t* job, const Prepared& P, int rank, McParams& mp, double& h2d) {
  const int S = job->n_scenarios, d = job->n_inputs;
  CU(cudaSetDevice(dv.id));
  int rc;
  const size_t plan_b = P.plan.size() * sizeof(DevInput);
  if ((rc = dv.plan.ensure(plan_b))) return rc;
  CU(cudaMemcpyAsync(dv.plan.p, P.plan.data(), plan_b, cudaMemcpyHostToDevice, dv.st));
  h2d += plan_b;
  const size_t consts_b = (size_t)job->n_consts * sizeof(double);
  if ((rc = dv.consts.ensure(consts_b))) return rc;
  if (consts_b) {
    CU(cudaMemcpyAsync(dv.consts.p, job->consts, consts_b, cudaMemcpyHostToDevice, dv.st));
    h2d += consts_b;
  }
  mp = McParams{};
  mp.plan = (const DevInput*)dv.plan.p;
  mp.consts = (const double*)dv.consts.p;
  if (job->n_tables > 0) {
    const size_t b = (size_t)job->n_tables * sizeof(double);
    if ((rc = dv.tables.ensure(b))) return rc;
    CU(cudaMemcpyAsync(dv.tables.p, job->tables, b, cudaMemcpyHostToDevice, dv.st));
    h2d += b;
    mp.tables = (const double*)dv.tables.p;
  }
  if (job->corr_chol) {
    const size_t b = (size_t)((job->flags & EBN_JOB_COPULA_SHARED) ? 1 : S) * d * d * sizeof(double);
    if ((rc = dv.chol.ensure(b))) return rc;
    CU(cudaMemcpyAsync(dv.chol.p, job->corr_chol, b, cudaMemcpyHostToDevice, dv.st));
    h2d += b;
    mp.chol = (const double*)dv.chol.p;
    mp.chol_shared = (job->flags & EBN_JOB_COPULA_SHARED) ? 1 : 0;
  }
  if (job->stream_id) {
    const size_t b = (size_t)S * sizeof(uint32_t);
    if ((rc = dv.streams.ensure(b))) return rc;
    CU(cudaMemcpyAsync(dv.streams.p, job->stream_id, b, cudaMemcpyHostToDevice, dv.st));
    h2d += b;
    mp.stream_id = (const uint32_t*)dv.streams.p;
  }
  mp.first = P.first;
  mp.last = P.last;
  mp.pairs_per_item = P.pairs_per_item;
  mp.items_per_scenario = P.items_per_scenario;
  mp.pair_base = P.pair_base;
  mp.shard_rank = rank;
  mp.shard_world = P.world;
  mp.local_items = local_items_of(P, S, rank);
  mp.S = S;
  mp.d = d;
  mp.n_consts = job->n_consts;
  mp.k0 = (uint32_t)(job->seed & 0xffffffffu);
  mp.k1 = (uint32_t)(job->seed >> 32);
  return EBN_OK;
}

int grid_for(Dev& dv, const ebn_job_t* job, Prepared& P, int mode, size_t dyn,
             int64_t local_items, int* blocks) {
  int per_sm = 0;
  int rc = k_occupancy(dv, job, P, mode, dyn, &per_sm);
  if (rc) return rc;
  if (per_sm < 1) return fail(EBN_ECUDA, "kernel does not fit on an SM (dyn smem %zu)", dyn);
  if (const char* cap = getenv("EBN_BLOCKS_PER_SM")) {  // tuning experiments only
    const int c = atoi(cap);
    if (c >= 1 && c < per_sm) per_sm = c;
  }
  int64_t b = (int64_t)dv.sms * per_sm;
  if (b > local_items) b = local_items;
  if (b < 1) b = 1;
  *blocks = (int)b;
  return EBN_OK;
}

// Sums `n` int64 (and optionally doubles) across devices of this process and
// across processes. Buffers are device pointers, one per device.
int allreduce(std::vector<void*>& i64, size_t n_i64, std::vector<void*>& f64, size_t n_f64,
              bool across_ranks) {
  const size_t nd = g.devs.size();
  if (nd > 1) {
    NC(g_nccl.GroupStart());
    for (size_t k = 0; k < nd; ++k) {
      Dev& dv = g.devs[k];
      NC(g_nccl.AllReduce(i64[k], i64[k], n_i64, kNcclInt64, kNcclSum, dv.comm, dv.st));
      if (n_f64) NC(g_nccl.AllReduce(f64[k], f64[k], n_f64, kNcclFloat64, kNcclSum, dv.comm, dv.st));
    }
    NC(g_nccl.GroupEnd());
  }
  if (across_ranks) {
    if (!g.rank_comm)
      return fail(EBN_ESTATE, "job is sharded over %d ranks but ebn_comm_init_rank was not called",
                  g.world);
    Dev& dv = g.devs[0];
    NC(g_nccl.GroupStart());
    NC(g_nccl.AllReduce(i64[0], i64[0], n_i64, kNcclInt64, kNcclSum, g.rank_comm, dv.st));
    if (n_f64) NC(g_nccl.AllReduce(f64[0], f64[0], n_f64, kNcclFloat64, kNcclSum, g.rank_comm, dv.st));
    NC(g_nccl.GroupEnd());
  }
  return EBN_OK;
}

// The interpreted coefficient-table functors (PolynomialProgram, MinAffine) are the road that needs no
// NVRTC, not the fast one: the same limit state as an expression program compiles into straight-line
// code with a compile-time sampling plan (C = A + B: 7.6e10 -> 4.2e11 samples/s). Large batches are
// therefore rewritten into EBN_LS_EXPRESSION here, instruction for instruction in the functor's own
// evaluation order, so strict results do not change. Returns false when the job stays as it is.
constexpr double kJitFunctorMinSamples = 5e10;

bool rewrite_as_expression(const ebn_job_t* job, std::vector<double>& out) {
  if (job->limit_state != EBN_LS_POLYNOMIAL && job->limit_state != EBN_LS_MIN_AFFINE) return false;
  if (job->flags & EBN_JOB_NO_JIT_PLAN) return false;
  const char* env = getenv("EBN_JIT_PLAN");  // "0": never, "1": always (experiments), as for sampling plans
  if (env && env[0] == '0') return false;
  const bool asked = (job->flags & EBN_JOB_JIT_PLAN) || (env && env[0] == '1');
  if ((double)job->n_per_scenario * job->n_scenarios < kJitFunctorMinSamples && !asked) return false;
  if (check_limit_state(job->limit_state, job->consts, job->n_consts, job->n_inputs) != EBN_OK) return false;
  if (!jit_available(nullptr)) return false;
  const double* k = job->consts;
  const int d = job->n_inputs;
  std::vector<double> code;
  auto emit = [&](int op, double arg) { code.push_back((double)op); code.push_back(arg); };
  if (job->limit_state == EBN_LS_POLYNOMIAL) {
    const int T = (int)k[0];
    if (d + T - 1 > EBN_MAX_INPUTS) return false;
    int pos = 1;
    for (int t = 0; t < T; ++t) {
      const int nt = (int)k[pos++];
      for (int q = 0; q < nt; ++q) {
        emit(EBN_OP_CONST, k[pos++]);                       // term = coef
        const int nf = (int)k[pos++];
        for (int f = 0; f < nf; ++f) {                      // term = term * x[col]
          emit(EBN_OP_INPUT, k[pos++]);
          emit(EBN_OP_MUL, 0.0);
        }
        if (q > 0) emit(EBN_OP_ADD, 0.0);                   // acc = acc + term
      }
      if (t + 1 < T) emit(EBN_OP_STORE, 0.0);               // a model column; the last stage is g
    }
  } else {
    const int K = (int)k[0];
    for (int r = 0; r < K; ++r) {
      const double* row = k + 1 + (size_t)r * (1 + d);
      emit(EBN_OP_CONST, row[0]);
      for (int j = 0; j < d; ++j) {
        if (row[1 + j] == 0.0 && !job->strict) continue;    // 0 * x only matters for the bit pattern
        emit(EBN_OP_CONST, row[1 + j]);
        emit(EBN_OP_INPUT, (double)j);
        emit(EBN_OP_MUL, 0.0);
        emit(EBN_OP_ADD, 0.0);
      }
      if (r > 0) emit(EBN_OP_MIN, 0.0);
    }
  }
  if (code.size() / 2 > EBN_EXPR_MAX_OPS) return false;
  out.clear();
  out.push_back(0.0);
  out.push_back((double)(code.size() / 2));
  out.insert(out.end(), code.begin(), code.end());
  return expr_check(out.data(), (int)out.size(), d, nullptr).empty();
}

int run_batch(const ebn_job_t* job_in, int mode, const double* edges, int nedges, int64_t* counts_out,
              double* sum_out, double* sumsq_out) {
  int rc = require_ready();
  if (rc) return rc;
  if (!job_in) return fail(EBN_EINVAL, "job is NULL");
  ebn_job_t rewritten;
  std::vector<double> program;
  const ebn_job_t* job = job_in;
  if (rewrite_as_expression(job_in, program)) {
    rewritten = *job_in;
    rewritten.limit_state = EBN_LS_EXPRESSION;
    rewritten.consts = program.data();
    rewritten.n_consts = (int32_t)program.size();
    job = &rewritten;
  }
  Prepared P;
  if ((rc = prepare(job, mode, P))) return rc;
  const int S = job->n_scenarios;
  const size_t nd = g.devs.size();
  const int jw = job->shard_world <= 1 ? 1 : job->shard_world;
  const int jr = job->shard_world <= 1 ? 0 : job->shard_rank;
  if (jw > 1 && g.rank_comm && jw != g.world && !(job->flags & EBN_JOB_PARTIAL))
    return fail(EBN_EINVAL, "shard_world %d differs from the communicator's world %d", jw, g.world);
  if (jw > 1 && nd > 1)
    return fail(EBN_ESTATE, "multi-device init and multi-rank sharding cannot be combined");
  const size_t ncnt = mode == 0 ? (size_t)S : (size_t)S * (nedges + 1);
  const size_t dyn = hist_dyn_bytes(mode == 0 ? 0 : nedges) + (P.icdf_dyn ? (size_t)kIcdfDynBytes : 0);
  std::vector<void*> ci(nd), cf(nd);
  std::vector<McParams> mps(nd);
  std::vector<int> blocks(nd);
  double h2d = 0.0;
  int64_t my_samples = 0;
  // stage + launch on every device before waiting on any
  for (size_t k = 0; k < nd; ++k) {
    Dev& dv = g.devs[k];
    const int rank = jr * (int)nd + (int)k;
    if ((rc = stage(dv, job, P, rank, mps[k], h2d))) return rc;
    McParams& mp = mps[k];
    mp.icdf_dyn = P.icdf_dyn ? 1 : 0;
    if ((rc = dv.counts.ensure(ncnt * sizeof(unsigned long long)))) return rc;
    CU(cudaMemsetAsync(dv.counts.p, 0, ncnt * sizeof(unsigned long long), dv.st));
    mp.counts = (unsigned long long*)dv.counts.p;
    ci[k] = dv.counts.p;
    if (mode == 1) {
      if ((rc = dv.edges.ensure((size_t)nedges * sizeof(double)))) return rc;
      CU(cudaMemcpyAsync(dv.edges.p, edges, (size_t)nedges * sizeof(double), cudaMemcpyHostToDevice, dv.st));
      h2d += (double)nedges * sizeof(double);
      const size_t pb = (size_t)(mp.local_items > 0 ? mp.local_items : 1) * 2 * sizeof(double);
      if ((rc = dv.partial.ensure(pb))) return rc;
      if ((rc = dv.sums.ensure((size_t)2 * S * sizeof(double)))) return rc;
      mp.edges = (const double*)dv.edges.p;
      mp.partial = (double*)dv.partial.p;
      mp.nedges = nedges;
      if (nedges >= 3) {  // equally spaced edges (the grid the host mirror uses) get the direct index
        const double h = (edges[nedges - 1] - edges[0]) / (nedges - 1);
        bool uni = h > 0.0 && std::isfinite(h);
        for (int i = 1; uni && i < nedges; ++i)
          if (fabs(edges[i] - (edges[0] + i * h)) > 1e-9 * h) uni = false;
        if (uni) {
          mp.hist_uniform = 1;
          mp.hist_e0 = edges[0];
          mp.hist_inv_h = 1.0 / h;
        }
      }
      cf[k] = dv.sums.p;
    }
    if ((rc = grid_for(dv, job, P, mode, dyn, mp.local_items, &blocks[k]))) return rc;
    CU(cudaEventRecord(dv.ev0, dv.st));
    if (mp.local_items > 0 && (rc = k_launch(dv, job, P, mode, mp, blocks[k], dyn))) return rc;
    if (mode == 1)
      CU(moments_launch(mp.partial, mp.local_items, mp.items_per_scenario, mp.shard_rank,
                        mp.shard_world, S, (double*)dv.sums.p, (double*)dv.sums.p + S, dv.st));
    CU(cudaEventRecord(dv.ev1, dv.st));
  }
  const bool partial = (job->flags & EBN_JOB_PARTIAL) != 0;
  if (nd > 1 || (jw > 1 && !partial)) {
    if ((rc = allreduce(ci, ncnt, cf, mode == 1 ? (size_t)2 * S : 0, jw > 1 && !partial))) return rc;
  }
  // results come from device 0 (after the all-reduce every device holds them)
  Dev& d0 = g.devs[0];
  CU(cudaSetDevice(d0.id));
  CU(cudaMemcpyAsync(counts_out, d0.counts.p, ncnt * sizeof(int64_t), cudaMemcpyDeviceToHost, d0.st));
  double d2h = (double)ncnt * sizeof(int64_t);
  std::vector<double> sums;
  if (mode == 1) {
    sums.resize((size_t)2 * S);
    CU(cudaMemcpyAsync(sums.data(), d0.sums.p, sums.size() * sizeof(double), cudaMemcpyDeviceToHost, d0.st));
    d2h += (double)sums.size() * sizeof(double);
  }
  double kms = 0.0;
  for (size_t k = 0; k < nd; ++k) {
    Dev& dv = g.devs[k];
    CU(cudaSetDevice(dv.id));
    CU(cudaStreamSynchronize(dv.st));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, dv.ev0, dv.ev1));
    if (ms > kms) kms = ms;
  }
  CU(cudaSetDevice(d0.id));
  if (mode == 1) {
    for (int s = 0; s < S; ++s) {
      if (sum_out) sum_out[s] = sums[s];
      if (sumsq_out) sumsq_out[s] = sums[(size_t)S + s];
    }
  }
  // samples processed by this process: its share of the work items
  {
    const int64_t total_items = P.items_per_scenario * S;
    int64_t mine = 0;
    for (size_t k = 0; k < nd; ++k) mine += mps[k].local_items;
    my_samples = (int64_t)((double)job->n_per_scenario * S * ((double)mine / (double)total_items));
  }
  g.stats = ebn_stats_t{};
  g.stats.kernel_ms = kms;
  g.stats.h2d_bytes = h2d;
  g.stats.d2h_bytes = d2h;
  g.stats.samples = my_samples;
  g.stats.launches = (int)nd * (mode == 1 ? 2 : 1);
  g.stats.devices = (int)nd;
  g.stats.specialised = (P.plan_class >= PLAN_STATIC0 && P.plan_class != PLAN_QMC) || !P.jit_spec.kinds.empty();
  g.stats.jit_compiles = P.jit_compiles;
  if (getenv("EBN_TRACE")) {  // one line per call on stderr: which kernel path ran, how long
    const LsInfo* li = ls_info(job->limit_state);
    fprintf(stderr,
            "[ebncuda] %s %s S=%d d=%d n=%lld plan=%s%s%s strict=%d devices=%zu blocks=%d items/scenario=%lld "
            "kernel=%.3f ms compiled=%d\n",
            mode == 0 ? "pf_batch" : "hist_batch", li ? li->name : "?", S, job->n_inputs, (long long)job->n_per_scenario,
            P.plan_class == PLAN_COPULA ? "copula" : P.plan_class >= PLAN_STATIC0 ? "compile-time" : "runtime-dispatched",
            P.jit ? (P.jit_spec.kinds.empty() ? " (run-time compiled functor)" : " (run-time compiled)") : "",
            (job->flags & EBN_JOB_FP32_SAMPLING) ? " fp32-sampling" : "", job->strict ? 1 : 0, nd, blocks[0],
            (long long)P.items_per_scenario, kms, P.jit_compiles);
  }
  return EBN_OK;
}

}  // namespace

// ===========================================================================
// exported functions
// ===========================================================================
extern "C" {

int ebn_abi_version(void) { return EBN_ABI_VERSION; }
const char* ebn_build_id(void) { return EBN_BUILD_ID; }

const char* ebn_last_error(void) { return g_err.c_str(); }

int ebn_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int ebn_shutdown(void) {
  API_LOCK;
  if (g.rank_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(g.rank_comm);
  g.rank_comm = nullptr;
  g.rank = 0;
  g.world = 1;
  for (Dev& dv : g.devs) {
    cudaSetDevice(dv.id);
    if (dv.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(dv.comm);
    for (Buf* b : {&dv.plan, &dv.consts, &dv.tables, &dv.chol, &dv.streams, &dv.counts, &dv.edges,
                   &dv.partial, &dv.sums, &dv.mat, &dv.gout})
      b->release();
    if (dv.ev0) cudaEventDestroy(dv.ev0);
    if (dv.ev1) cudaEventDestroy(dv.ev1);
    if (dv.own) cudaStreamDestroy(dv.own);
  }
  jit_unload_all();
  g.devs.clear();
  g.ready = false;
  return EBN_OK;
}

// Brings the devices up; on any failure the half-built state (streams, events, communicators) is torn
// down again through ebn_shutdown, so a later ebn_init starts clean.
static int init_devices(const std::vector<int>& ids) {
  g.devs.resize(ids.size());
  for (size_t k = 0; k < ids.size(); ++k) {
    Dev& dv = g.devs[k];
    dv.id = ids[k];
    CU(cudaSetDevice(dv.id));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, dv.id));
    if (prop.major < 10)
      return fail(EBN_ENODEV, "device %d is sm_%d%d; libebncuda is built for sm_100a only", dv.id,
                  prop.major, prop.minor);
    dv.sms = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&dv.own, cudaStreamNonBlocking));
    dv.st = dv.own;
    CU(cudaEventCreate(&dv.ev0));
    CU(cudaEventCreate(&dv.ev1));
  }
  if (ids.size() > 1) {
    int rc = nccl_load();
    if (rc) return rc;
    std::vector<ncclComm_t> comms(ids.size());
    NC(g_nccl.CommInitAll(comms.data(), (int)ids.size(), ids.data()));
    for (size_t k = 0; k < ids.size(); ++k) g.devs[k].comm = comms[k];
  }
  CU(cudaSetDevice(g.devs[0].id));
  return EBN_OK;
}

int ebn_init(const int* devices, int ndev) {
  API_LOCK;
  if (g.ready || !g.devs.empty()) ebn_shutdown();
  if (ndev < 0 || (ndev > 0 && !devices)) return fail(EBN_EINVAL, "bad device list");
  const int avail = ebn_device_count();
  if (avail < 1) return fail(EBN_ENODEV, "no CUDA device is visible; libebncuda has no CPU fallback");
  std::vector<int> ids;
  if (ndev == 0) ids.push_back(0);
  for (int i = 0; i < ndev; ++i) {
    if (devices[i] < 0 || devices[i] >= avail)
      return fail(EBN_ENODEV, "device %d requested but only %d visible", devices[i], avail);
    for (int k = 0; k < i; ++k)
      if (devices[k] == devices[i]) return fail(EBN_EINVAL, "device %d listed twice", devices[i]);
    ids.push_back(devices[i]);
  }
  const int rc = init_devices(ids);
  if (rc) {
    const std::string why = ebn_last_error();   // ebn_shutdown must not hide the reason
    ebn_shutdown();
    return fail(rc, "%s", why.c_str());
  }
  g.ready = true;
  return EBN_OK;
}

int ebn_set_stream(void* cuda_stream) {
  API_LOCK;
  int rc = require_ready();
  if (rc) return rc;
  if (g.devs.size() != 1) return fail(EBN_ESTATE, "ebn_set_stream needs single-device mode");
  g.devs[0].st = cuda_stream ? (cudaStream_t)cuda_stream : g.devs[0].own;
  return EBN_OK;
}

int ebn_nccl_unique_id(void* id128) {
  API_LOCK;
  if (!id128) return fail(EBN_EINVAL, "id128 is NULL");
  int rc = nccl_load();
  if (rc) return rc;
  NC(g_nccl.GetUniqueId(id128));
  return EBN_OK;
}

int ebn_comm_init_rank(const void* id128, int rank, int world) {
  API_LOCK;
  int rc = require_ready();
  if (rc) return rc;
  if (!id128 || world < 1 || rank < 0 || rank >= world) return fail(EBN_EINVAL, "bad rank/world");
  if (g.devs.size() != 1) return fail(EBN_ESTATE, "one process per GPU: init with exactly one device");
  if ((rc = nccl_load())) return rc;
  if (g.rank_comm) { g_nccl.CommDestroy(g.rank_comm); g.rank_comm = nullptr; }
  CU(cudaSetDevice(g.devs[0].id));
  NcclId id;
  memcpy(id.b, id128, 128);
  NC(g_nccl.CommInitRank(&g.rank_comm, world, id, rank));
  g.rank = rank;
  g.world = world;
  return EBN_OK;
}

int ebn_comm_destroy(void) {
  API_LOCK;
  if (g.rank_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(g.rank_comm);
  g.rank_comm = nullptr;
  g.rank = 0;
  g.world = 1;
  return EBN_OK;
}

// Scenarios whose marginal KINDS differ cannot share a compile-time sampling plan: the Uniform / Exponential-tail
// rows `_approximate` puts next to each other (reference discretize.jl:83-91), a parent that is truncated in some
// states only. Instead of sending the whole job through the runtime-dispatched plan (2-5x slower), a large job is
// split by kind signature: every group is an ordinary job of its own (its scenarios keep their Philox streams, so
// the counts are bit-identical to the unsplit call) and finds a compiled-in plan, or gets one compiled at run time.
constexpr double kSplitMinSamples = 2e9;   // below this the extra launches are not worth it
constexpr size_t kSplitMaxGroups = 16;

int run_batch_split(const ebn_job_t* job, int mode, const double* edges, int nedges, int64_t* counts_out,
                    double* sum_out, double* sumsq_out) {
  const auto whole = [&]() { return run_batch(job, mode, edges, nedges, counts_out, sum_out, sumsq_out); };
  // count mode only: the moments of histogram mode are summed per work item in item order, and the item grid
  // depends on the number of scenarios, so a split would change their last bits
  if (mode != 0 || !job || !job->inputs || job->n_scenarios < 2 || job->n_inputs < 1 || job->corr_chol ||
      (job->flags & (EBN_JOB_QMC_SOBOL | EBN_JOB_NO_JIT_PLAN)) || job->limit_state == EBN_LS_EXPRESSION ||
      (double)job->n_per_scenario * job->n_scenarios < kSplitMinSamples)
    return whole();
  if (const char* env = getenv("EBN_SPLIT_KINDS"))   // "0": never (experiments)
    if (env[0] == '0') return whole();
  const int S = job->n_scenarios, d = job->n_inputs;
  std::vector<DevInput> plan((size_t)S * d);
  for (int s = 0; s < S; ++s)
    for (int j = 0; j < d; ++j)
      if (canonicalise(job->inputs[(size_t)s * d + j], false, (job->flags & EBN_JOB_FP32_SAMPLING) != 0,
                       plan[(size_t)s * d + j], s, j, job->n_tables))
        return whole();   // the ordinary path reports the error
  std::vector<std::vector<int>> groups;
  for (int s = 0; s < S; ++s) {
    size_t g = 0;
    for (; g < groups.size(); ++g) {
      bool same = true;
      for (int j = 0; j < d && same; ++j) same = plan[(size_t)groups[g][0] * d + j].kind == plan[(size_t)s * d + j].kind;
      if (same) break;
    }
    if (g == groups.size()) {
      if (groups.size() == kSplitMaxGroups) return whole();
      groups.emplace_back();
    }
    groups[g].push_back(s);
  }
  if (groups.size() < 2) return whole();
  // worth it only if some group ends up on a compile-time plan
  bool gain = false;
  for (const auto& grp : groups) {
    std::vector<DevInput> rows;
    for (int s : grp) rows.insert(rows.end(), plan.begin() + (size_t)s * d, plan.begin() + (size_t)(s + 1) * d);
    if (static_plan_matches(job->limit_state, rows.data(), (int)grp.size(), d) >= 0 ||
        ((double)job->n_per_scenario * grp.size() >= kJitPlanMinSamples && jit_available(nullptr)))
      gain = true;
  }
  if (!gain) return whole();
  const int nb = mode == 0 ? 1 : nedges + 1;
  ebn_stats_t total{};
  total.specialised = 1;
  for (const auto& grp : groups) {
    const int n = (int)grp.size();
    std::vector<ebn_input_t> rows;
    std::vector<uint32_t> streams(n);
    for (int i = 0; i < n; ++i) {
      const int s = grp[i];
      rows.insert(rows.end(), job->inputs + (size_t)s * d, job->inputs + (size_t)(s + 1) * d);
      streams[i] = job->stream_id ? job->stream_id[s] : (uint32_t)s;
    }
    ebn_job_t sub = *job;
    sub.n_scenarios = n;
    sub.inputs = rows.data();
    sub.stream_id = streams.data();
    std::vector<int64_t> cnt((size_t)n * nb);
    std::vector<double> s1(n), s2(n);
    const int rc = run_batch(&sub, mode, edges, nedges, cnt.data(), mode == 1 ? s1.data() : nullptr,
                             mode == 1 ? s2.data() : nullptr);
    if (rc) return rc;
    for (int i = 0; i < n; ++i) {
      for (int b = 0; b < nb; ++b) counts_out[(size_t)grp[i] * nb + b] = cnt[(size_t)i * nb + b];
      if (mode == 1 && sum_out) sum_out[grp[i]] = s1[i];
      if (mode == 1 && sumsq_out) sumsq_out[grp[i]] = s2[i];
    }
    total.kernel_ms += g.stats.kernel_ms;
    total.h2d_bytes += g.stats.h2d_bytes;
    total.d2h_bytes += g.stats.d2h_bytes;
    total.samples += g.stats.samples;
    total.launches += g.stats.launches;
    total.devices = g.stats.devices;
    total.specialised &= g.stats.specialised;
    total.jit_compiles += g.stats.jit_compiles;
  }
  g.stats = total;
  return EBN_OK;
}

int ebn_pf_batch(const ebn_job_t* job, int64_t* fail_count, double* pf, double* var) {
  API_LOCK;
  if (!fail_count) return fail(EBN_EINVAL, "fail_count is NULL");
  int rc = run_batch_split(job, 0, nullptr, 0, fail_count, nullptr, nullptr);
  if (rc) return rc;
  const double n = (double)job->n_per_scenario;
  for (int s = 0; s < job->n_scenarios; ++s) {
    const double p = (double)fail_count[s] / n;  // sum(Bool)/n
    if (pf) pf[s] = p;
    if (var) var[s] = (p - p * p) / n;
  }
  return EBN_OK;
}

int ebn_hist_batch(const ebn_job_t* job, const double* edges, int nedges, int64_t* counts,
                   double* sum, double* sumsq) {
  API_LOCK;
  if (!counts) return fail(EBN_EINVAL, "counts is NULL");
  if (!edges || nedges < 1 || nedges > EBN_MAX_EDGES)
    return fail(EBN_EINVAL, "need 1..%d histogram edges", EBN_MAX_EDGES);
  for (int i = 0; i < nedges; ++i) {
    if (std::isnan(edges[i])) return fail(EBN_EINVAL, "edge %d is NaN", i);
    if (i && !(edges[i] > edges[i - 1])) return fail(EBN_EINVAL, "edges must be strictly increasing");
  }
  return run_batch_split(job, 1, edges, nedges, counts, sum, sumsq);
}

int ebn_eval_matrix(int limit_state, const double* consts, int n_consts, const double* X, int64_t n,
                    int d, int64_t ld, int strict, int64_t* fail_count, double* g_out) {
  API_LOCK;
  int rc = require_ready();
  if (rc) return rc;
  if (!fail_count) return fail(EBN_EINVAL, "fail_count is NULL");
  if (n < 0 || (n > 0 && !X) || ld < n) return fail(EBN_EINVAL, "bad matrix shape n=%lld ld=%lld", (long long)n, (long long)ld);
  if ((rc = check_limit_state(limit_state, consts, n_consts, d))) return rc;
  *fail_count = 0;
  if (n == 0) return EBN_OK;
  Dev& dv = g.devs[0];
  CU(cudaSetDevice(dv.id));
  const size_t xb = (size_t)ld * d * sizeof(double);
  if ((rc = dv.mat.ensure(xb))) return rc;
  if ((rc = dv.consts.ensure((size_t)n_consts * sizeof(double)))) return rc;
  if ((rc = dv.counts.ensure(sizeof(unsigned long long)))) return rc;
  if (g_out && (rc = dv.gout.ensure((size_t)n * sizeof(double)))) return rc;
  // a column-major matrix with ld > n only guarantees (d-1)*ld + n elements on the host side
  const size_t xcopy = d > 0 ? ((size_t)(d - 1) * ld + (size_t)n) * sizeof(double) : 0;
  if (xcopy) CU(cudaMemcpyAsync(dv.mat.p, X, xcopy, cudaMemcpyHostToDevice, dv.st));
  if (n_consts) CU(cudaMemcpyAsync(dv.consts.p, consts, (size_t)n_consts * sizeof(double), cudaMemcpyHostToDevice, dv.st));
  CU(cudaMemsetAsync(dv.counts.p, 0, sizeof(unsigned long long), dv.st));
  int64_t blocks = (n + threads_per_block() - 1) / threads_per_block();
  const int64_t cap = (int64_t)dv.sms * 8;
  if (blocks > cap) blocks = cap;
  CU(cudaEventRecord(dv.ev0, dv.st));
  if (limit_state == EBN_LS_EXPRESSION) {
    JitSpec spec;
    spec.limit_state = limit_state;
    spec.d = d;
    spec.strict = strict != 0;
    spec.consts = consts;
    spec.n_consts = n_consts;
    JitKernel* k = nullptr;
    bool compiled = false;
    std::string err;
    if (!jit_get(spec, JIT_MATRIX, &k, &compiled, &err)) return fail(EBN_EJIT, "%s", err.c_str());
    const double* a_consts = (const double*)dv.consts.p;
    const double* a_X = (const double*)dv.mat.p;
    unsigned long long* a_count = (unsigned long long*)dv.counts.p;
    double* a_g = g_out ? (double*)dv.gout.p : nullptr;
    void* args[] = {&a_consts, &n_consts, &a_X, &n, &d, &ld, &a_count, &a_g};
    if (!jit_launch(k, dv.id, (int)blocks, threads_per_block(), 0, dv.st, args, &err))
      return fail(EBN_ECUDA, "%s", err.c_str());
  } else {
    CU(ls_vtable(limit_state)->matrix(strict ? 1 : 0, (const double*)dv.consts.p, n_consts,
                     (const double*)dv.mat.p, n, d, ld, (unsigned long long*)dv.counts.p,
                     g_out ? (double*)dv.gout.p : nullptr, (int)blocks, dv.st));
  }
  CU(cudaEventRecord(dv.ev1, dv.st));
  CU(cudaMemcpyAsync(fail_count, dv.counts.p, sizeof(int64_t), cudaMemcpyDeviceToHost, dv.st));
  if (g_out) CU(cudaMemcpyAsync(g_out, dv.gout.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, dv.st));
  CU(cudaStreamSynchronize(dv.st));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, dv.ev0, dv.ev1));
  g.stats = ebn_stats_t{};
  g.stats.kernel_ms = ms;
  g.stats.h2d_bytes = (double)xb + (double)n_consts * sizeof(double);
  g.stats.d2h_bytes = 8.0 + (g_out ? (double)n * 8.0 : 0.0);
  g.stats.samples = n;
  g.stats.launches = 1;
  g.stats.devices = 1;
  return EBN_OK;
}

int ebn_sample_matrix_ex(const ebn_job_t* job, int scenario, int64_t first, int64_t n, int ncols,
                         double* X, double* g_out) {
  API_LOCK;
  int rc = require_ready();
  if (rc) return rc;
  if (!X) return fail(EBN_EINVAL, "X is NULL");
  if (n < 1 || first < 0) return fail(EBN_EINVAL, "bad sample range");
  Prepared P;
  if ((rc = prepare(job, 0, P))) return rc;
  if (scenario < 0 || scenario >= job->n_scenarios) return fail(EBN_EINVAL, "scenario %d out of range", scenario);
  Dev& dv = g.devs[0];
  McParams mp;
  double h2d = 0.0;
  if ((rc = stage(dv, job, P, 0, mp, h2d))) return rc;
  mp.first = first;
  mp.last = first + n;
  mp.pair_base = first >> 1;
  const int d = job->n_inputs;
  if (ncols < d || ncols > EBN_MAX_INPUTS) return fail(EBN_EINVAL, "ncols %d outside [%d,%d]", ncols, d, EBN_MAX_INPUTS);
  if (ncols > d) {
    int stages = job->limit_state == EBN_LS_POLYNOMIAL ? (int)job->consts[0] : 0;
    if (job->limit_state == EBN_LS_EXPRESSION) {
      ExprInfo info;
      expr_check(job->consts, job->n_consts, d, &info);
      stages = info.n_store;
    }
    if (ncols > d + stages)
      return fail(EBN_EINVAL, "ncols %d: the limit state defines only %d model columns after the %d inputs", ncols, stages, d);
  }
  mp.ncols = ncols;
  if ((rc = dv.mat.ensure((size_t)n * ncols * sizeof(double)))) return rc;
  if (g_out && (rc = dv.gout.ensure((size_t)n * sizeof(double)))) return rc;
  int64_t blocks = ((n + 1) / 2 + threads_per_block()) / threads_per_block();
  if (blocks > (int64_t)dv.sms * 8) blocks = (int64_t)dv.sms * 8;
  if ((rc = k_replay(dv, job, P, mp, scenario, n, (double*)dv.mat.p, g_out ? (double*)dv.gout.p : nullptr,
                     (int)blocks)))
    return rc;
  CU(cudaMemcpyAsync(X, dv.mat.p, (size_t)n * ncols * sizeof(double), cudaMemcpyDeviceToHost, dv.st));
  if (g_out) CU(cudaMemcpyAsync(g_out, dv.gout.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, dv.st));
  CU(cudaStreamSynchronize(dv.st));
  return EBN_OK;
}

int ebn_sample_matrix(const ebn_job_t* job, int scenario, int64_t first, int64_t n, double* X,
                      double* g_out) {
  API_LOCK;
  if (!job) return fail(EBN_EINVAL, "job is NULL");
  return ebn_sample_matrix_ex(job, scenario, first, n, job->n_inputs, X, g_out);
}

int ebn_peak_fp64(double seconds, double* dfma_per_s) {
  API_LOCK;
  int rc = require_ready();
  if (rc) return rc;
  if (!dfma_per_s) return fail(EBN_EINVAL, "dfma_per_s is NULL");
  Dev& dv = g.devs[0];
  CU(cudaSetDevice(dv.id));
  if ((rc = dv.gout.ensure(64))) return rc;
  const int blocks = dv.sms * 8;  // 2048 threads per SM
  const int T = threads_per_block();
  int iters = 4096;
  double best = 0.0, spent = 0.0;
  if (!(seconds > 0.0)) seconds = 0.2;
  // warm-up + calibrate, then repeat until `seconds` of device time is spent
  for (int rep = 0; rep < 1000 && spent < seconds; ++rep) {
    CU(cudaEventRecord(dv.ev0, dv.st));
    CU(dfma_launch((double*)dv.gout.p, iters, blocks, dv.st));
    CU(cudaEventRecord(dv.ev1, dv.st));
    CU(cudaStreamSynchronize(dv.st));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, dv.ev0, dv.ev1));
    const double dfma = (double)blocks * T * (double)iters * 64.0;
    if (rep > 0) {
      spent += ms * 1e-3;
      const double rate = dfma / (ms * 1e-3);
      if (rate > best) best = rate;
    }
    if (ms < 20.f && iters < (1 << 24)) iters *= 2;
  }
  *dfma_per_s = best;
  return EBN_OK;
}

int ebn_partition(const ebn_job_t* job, int histogram, int64_t out[4]) {
  if (!out) return fail(EBN_EINVAL, "out is NULL");
  Prepared P;
  int rc = prepare(job, histogram ? 1 : 0, P, 1);  // host logic only: needs no device
  if (rc) return rc;
  const int jr = job->shard_world <= 1 ? 0 : job->shard_rank;
  out[0] = P.pairs_per_item;
  out[1] = P.items_per_scenario;
  out[2] = local_items_of(P, job->n_scenarios, jr);
  out[3] = P.pair_base;
  return EBN_OK;
}

int ebn_jit_available(void) {
  std::string why;
  if (jit_available(&why)) return 1;
  fail(EBN_EJIT, "%s", why.c_str());
  return 0;
}

int ebn_expression_check(const double* consts, int n_consts, int d, int* n_columns) {
  if (d < 1 || d > EBN_MAX_INPUTS) return fail(EBN_EINVAL, "n_inputs %d outside [1,%d]", d, EBN_MAX_INPUTS);
  ExprInfo info;
  const std::string bad = expr_check(consts, n_consts, d, &info);
  if (!bad.empty()) return fail(EBN_ELIMITSTATE, "%s", bad.c_str());
  if (n_columns) *n_columns = d + info.n_store;
  return EBN_OK;
}

int64_t ebn_expression_source(const double* consts, int n_consts, int d, char* buf, int64_t cap) {
  int rc = ebn_expression_check(consts, n_consts, d, nullptr);
  if (rc) return rc;
  const std::string src = expr_codegen(consts, n_consts, d);
  if (buf && cap > 0) {
    const size_t m = src.size() < (size_t)cap - 1 ? src.size() : (size_t)cap - 1;
    memcpy(buf, src.data(), m);
    buf[m] = '\0';
  }
  return (int64_t)src.size();
}

int ebn_jit_compile(const ebn_job_t* job, int role, int64_t* cubin_bytes) {
  API_LOCK;
  if (role < 0 || role > 2) return fail(EBN_EINVAL, "role must be 0 (count), 1 (histogram) or 2 (replay)");
  if (!job) return fail(EBN_EINVAL, "job is NULL");
  ebn_job_t rewritten;
  std::vector<double> program;
  if (role != 2 && rewrite_as_expression(job, program)) {  // what ebn_pf_batch / ebn_hist_batch would run
    rewritten = *job;
    rewritten.limit_state = EBN_LS_EXPRESSION;
    rewritten.consts = program.data();
    rewritten.n_consts = (int32_t)program.size();
    job = &rewritten;
  }
  Prepared P;
  int rc = prepare(job, role == 1 ? 1 : 0, P, 1);  // host logic only
  if (rc) return rc;
  if (cubin_bytes) *cubin_bytes = 0;
  if (!P.jit) return EBN_OK;  // this job runs a compiled-in kernel
  JitKernel* k = nullptr;
  if ((rc = jit_kernel_of(P, (JitRole)role, &k))) return rc;
  if (cubin_bytes) *cubin_bytes = (int64_t)jit_cubin_bytes(k);
  g.stats = ebn_stats_t{};
  g.stats.jit_compiles = P.jit_compiles;
  return EBN_OK;
}

int ebn_last_stats(ebn_stats_t* out) {
  API_LOCK;
  if (!out) return fail(EBN_EINVAL, "out is NULL");
  *out = g.stats;
  return EBN_OK;
}

int ebn_limit_state_id(const char* name) {
  if (!name) return EBN_EINVAL;
  for (const auto& l : kLs)
    if (strcmp(l.name, name) == 0) return l.id;
  return fail(EBN_ELIMITSTATE, "unknown limit state '%s'", name);
}

const char* ebn_limit_state_name(int id) {
  const LsInfo* li = ls_info(id);
  return li ? li->name : nullptr;
}

int ebn_limit_state_arity(int id, int* n_inputs, int* n_consts) {
  const LsInfo* li = ls_info(id);
  if (!li) return fail(EBN_ELIMITSTATE, "unknown limit state id %d", id);
  if (n_inputs) *n_inputs = li->d;
  if (n_consts) *n_consts = li->nc;
  return EBN_OK;
}

}  // extern "C"
