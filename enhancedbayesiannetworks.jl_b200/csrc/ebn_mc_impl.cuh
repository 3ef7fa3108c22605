// sm_100a kernels of libebncuda: fused sample -> limit state -> count.
//
// Replaces, for ALL scenarios of a functional node at once, the body of the
// loop at reference src/networks/ebn/evaluation/evaluate_node.jl:79-119
// (UQ.jl `sample` + `evaluate!` + `sum(performance(df) .< 0)`), and for
// ContinuousFunctionalNode the body at :22-33 (histogram instead of a KDE).
//
// Work decomposition: item = (scenario, chunk of sample pairs). Persistent
// blocks stride over this rank's items; threads stride over the pairs of an
// item. No global memory is touched per sample: the only traffic is the plan
// (48 B per input per item, L2 resident) in and one 64-bit atomic per item out.
#pragma once
#ifndef __CUDACC_RTC__
#include <cstdio>
#endif
#include "ebn_kernels.h"
#include "ebn_limit_states.cuh"
#include "ebn_sampling.cuh"

namespace ebn {

constexpr int kThreads = EBN_THREADS;
constexpr int kWarps = kThreads / 32;
#ifndef EBN_UNROLL
#define EBN_UNROLL 1
#endif
constexpr int kPairUnroll = EBN_UNROLL;  // pairs per trip of the static count loop
#ifndef EBN_PINGPONG
#define EBN_PINGPONG 0
#endif

// Compile-time sampling plans ------------------------------------------------
template <int... K>
struct StaticPlan {
  static constexpr bool is_static = true;
  static constexpr bool is_copula = false;
  static constexpr bool is_static_copula = false;
  static constexpr int n = sizeof...(K);
  static constexpr int kinds[sizeof...(K) ? sizeof...(K) : 1] = {K...};
};
struct DynamicPlan {
  static constexpr bool is_static = false;
  static constexpr bool is_copula = false;
  static constexpr bool is_static_copula = false;
};
struct CopulaPlan {
  static constexpr bool is_static = false;
  static constexpr bool is_copula = true;
  static constexpr bool is_static_copula = false;
};
// Quasi-Monte Carlo: scrambled Sobol points instead of the Philox words, every marginal by inverse CDF
// (ebn_qmc.cuh). Kinds are dispatched at run time.
struct QmcPlan {
  static constexpr bool is_static = false;
  static constexpr bool is_copula = false;
  static constexpr bool is_static_copula = false;
  static constexpr bool is_qmc = true;
};
template <class Plan, class = void>
struct plan_is_qmc { static constexpr bool value = false; };
template <class Plan>
struct plan_is_qmc<Plan, decltype((void)Plan::is_qmc)> { static constexpr bool value = Plan::is_qmc; };
// Gaussian copula with the marginal kinds known at compile time (instantiated at run time only,
// ebn_jit.cu): the eps draws, the triangular product z = L eps and the marginal transforms are
// unrolled over registers instead of looping over local-memory arrays.
template <int... K>
struct StaticCopulaPlan {
  static constexpr bool is_static = false;  // uses the generic sample loop
  static constexpr bool is_copula = true;
  static constexpr bool is_static_copula = true;
  static constexpr int n = sizeof...(K);
  static constexpr int kinds[sizeof...(K) ? sizeof...(K) : 1] = {K...};
};

// Can the plan draw a truncated (inverse-CDF) normal? Only then does the kernel hold the Phi^-1 cell table.
template <class Plan>
__host__ __device__ constexpr bool plan_uses_icdf() {
  if constexpr (Plan::is_static) {
    for (int i = 0; i < Plan::n; ++i)
      if (Plan::kinds[i] == DK_NORMAL_ICDF || Plan::kinds[i] == DK_LOGNORMAL_ICDF) return true;
    return false;
  } else {
    // kinds known at run time only, and every copula plan (whose kernels already hold the Phi cells in static
    // shared memory: a second static table would pass the 48 KB limit): the cells are read through the
    // block's IcdfDyn handle (see icdf<>). Under quasi-Monte Carlo every normal goes through Phi^-1.
    return plan_is_qmc<Plan>::value;
  }
}

// Plans whose kinds are known at run time only (dynamic, looping copula): the pair's Philox words come from
// a buffer (pair_words).
template <class Plan>
__host__ __device__ constexpr bool plan_is_runtime_dispatched() {
  return !Plan::is_static && !Plan::is_static_copula && !plan_is_qmc<Plan>::value;
}
// Plans that read the Phi^-1 cells through the block's IcdfDyn handle: those, and the unrolled copula plans.
template <class Plan>
__host__ __device__ constexpr bool plan_reads_icdf_dyn() {
  return !Plan::is_static && !plan_is_qmc<Plan>::value;
}

// Word offset of input J in a static plan, and the number of Philox calls a pair needs.
template <class Plan, int J>
__host__ __device__ constexpr int static_word_offset() {
  int off = 0;
  for (int i = 0; i < J; ++i) off += words_of_kind(Plan::kinds[i]);
  return off;
}
template <class Plan>
__host__ __device__ constexpr int static_calls() {
  return (static_word_offset<Plan, Plan::n>() + 3) / 4;
}

// Under a copula every random input owns the 2 words of its Box-Muller pair.
template <class Plan, int J>
__host__ __device__ constexpr int copula_word_offset() {
  int off = 0;
  for (int i = 0; i < J; ++i) off += Plan::kinds[i] == DK_CONST ? 0 : 2;
  return off;
}
template <class Plan, int J>
struct CopulaEps {
  __device__ __forceinline__ static void run(const Philox4* ph, const FastMathSmem& sm, double* e0, double* e1) {
    if constexpr (J < Plan::n) {
      if constexpr (Plan::kinds[J] != DK_CONST) {
        StaticWords<copula_word_offset<Plan, J>()> ws{ph};
        box_muller_words(ws.template w<0>(), ws.template w<1>(), sm, e0[J], e1[J]);
      } else {
        e0[J] = 0.0;
        e1[J] = 0.0;
      }
      CopulaEps<Plan, J + 1>::run(ph, sm, e0, e1);
    }
  }
};
template <class Plan, int J>
struct CopulaMarginals {
  __device__ __forceinline__ static void run(const DevInput* sp, const double* L, const double* tables,
                                             const FastMathSmem& sm, const double* e0, const double* e1,
                                             double* x0, double* x1) {
    if constexpr (J < Plan::n) {
      constexpr int KIND = Plan::kinds[J];
      const DevInput in = sp[J];
      if constexpr (KIND == DK_CONST) {
        x0[J] = in.a;
        x1[J] = in.a;
      } else {
        // same accumulation order as draw_pair_copula (a Parameter column contributes exactly 0)
        double z0 = 0.0, z1 = 0.0;
#pragma unroll
        for (int k = 0; k <= J; ++k) {
          if (Plan::kinds[k] != DK_CONST) {
            const double l = L[J * Plan::n + k];
            z0 = fma(l, e0[k], z0);
            z1 = fma(l, e1[k], z1);
          }
        }
        if constexpr (KIND == DK_NORMAL_BM) {
          x0[J] = fma(z0, in.b, in.a);
          x1[J] = fma(z1, in.b, in.a);
        } else if constexpr (KIND == DK_LOGNORMAL_BM) {
          x0[J] = exp_fast(fma(z0, in.b, in.a), sm);
          x1[J] = exp_fast(fma(z1, in.b, in.a), sm);
        } else if constexpr (KIND == DK_UNIFORM) {
          x0[J] = fma(phi_clamped(z0, sm), in.b, in.a);
          x1[J] = fma(phi_clamped(z1, sm), in.b, in.a);
        } else {
          x0[J] = icdf<KIND, plan_uses_icdf<Plan>()>(in, tables, sm, phi_clamped(z0, sm));
          x1[J] = icdf<KIND, plan_uses_icdf<Plan>()>(in, tables, sm, phi_clamped(z1, sm));
        }
      }
      CopulaMarginals<Plan, J + 1>::run(sp, L, tables, sm, e0, e1, x0, x1);
    }
  }
};

template <class Plan, int XCAP, int J>
struct DrawAll {
  __device__ __forceinline__ static void run(const DevInput* sp, const Philox4* ph,
                                             const PairCtr& pc, const double* tables,
                                             const FastMathSmem& sm, double* x0, double* x1) {
    if constexpr (J < Plan::n) {
      StaticWords<static_word_offset<Plan, J>()> ws{ph};
      draw_pair<Plan::kinds[J], StaticWords<static_word_offset<Plan, J>()>, plan_uses_icdf<Plan>()>(sp[J], ws, pc, (uint32_t)J, tables, sm, x0[J], x1[J]);
      DrawAll<Plan, XCAP, J + 1>::run(sp, ph, pc, tables, sm, x0, x1);
    }
  }
};

// sp: the scenario's plan (registers for a static plan, shared memory otherwise).
template <class Plan, int XCAP, bool LOOPED = false>
__device__ __forceinline__ void draw_inputs(const DevInput* sp, const double* sL, int d,
                                            const PairCtr& pc, const double* tables,
                                            const FastMathSmem& sm, double* x0, double* x1) {
  if constexpr (Plan::is_static) {
    constexpr int NC = static_calls<Plan>();
    Philox4 ph[NC > 0 ? NC : 1];
#pragma unroll
    for (int c = 0; c < NC; ++c)
      ph[c] = philox4x32_10(pc.qlo, pc.qhi, pc.stream, (uint32_t)c, pc.k0, pc.k1);
    DrawAll<Plan, XCAP, 0>::run(sp, ph, pc, tables, sm, x0, x1);
  } else if constexpr (Plan::is_static_copula) {
    constexpr int NC = (copula_word_offset<Plan, Plan::n>() + 3) / 4;
    Philox4 ph[NC > 0 ? NC : 1];
#pragma unroll
    for (int c = 0; c < NC; ++c)
      ph[c] = philox4x32_10(pc.qlo, pc.qhi, pc.stream, (uint32_t)c, pc.k0, pc.k1);
    double e0[Plan::n], e1[Plan::n];
    CopulaEps<Plan, 0>::run(ph, sm, e0, e1);
    CopulaMarginals<Plan, 0>::run(sp, sL, tables, sm, e0, e1, x0, x1);
  } else if constexpr (Plan::is_copula) {
    draw_pair_copula(sp, sL, d, pc, tables, sm, x0, x1);
  } else {
    __align__(16) uint32_t wbuf[kMaxPairWords];
    pair_words(pc, wbuf);
    if constexpr (LOOPED) {
      // the fused kernels: ONE loop over the d inputs with the marginal transforms inlined into it and the
      // columns in local memory (measured, profiles/r02_tune_generic.txt: against a 16-way predicated unroll
      // with an out-of-line call per input and, for the variadic limit states that index x[] at run time,
      // compare trees behind every x[k]); the replay kernels keep the unrolled form
#pragma unroll 1
      for (int j = 0; j < d; ++j) draw_pair_dyn_body(sp[j], pc, wbuf, (uint32_t)j, tables, sm, x0[j], x1[j]);
    } else {
#pragma unroll
      for (int j = 0; j < XCAP; ++j) {
        if (j < d) draw_pair_dyn(sp[j], pc, wbuf, (uint32_t)j, tables, sm, x0[j], x1[j]);
      }
    }
  }
}

// Quasi-Monte Carlo draw of one pair: `st` stands at the pair's even sample; Parameters take no dimension.
template <int XCAP, bool LOOPED = false>
__device__ __forceinline__ void draw_inputs_qmc(const DevInput* sp, int d, const QmcState& st, const double* tables,
                                                const FastMathSmem& sm, double* x0, double* x1) {
  const uint32_t* V = sobol_smem();
  int r = 0;
  auto one = [&](int j) {
    if (sp[j].kind == DK_CONST) {
      x0[j] = sp[j].a;
      x1[j] = sp[j].a;
    } else {
      draw_pair_qmc(sp[j], st.v[r], V[r * 32], st.key[r], tables, sm, x0[j], x1[j]);
      ++r;
    }
  };
  if constexpr (LOOPED) {   // see draw_inputs
#pragma unroll 1
    for (int j = 0; j < d; ++j) one(j);
  } else {
#pragma unroll
    for (int j = 0; j < XCAP; ++j) {
      if (j < d) one(j);
    }
  }
}
// number of random (non-Parameter) inputs of a scenario = Sobol dimensions it uses
__device__ __forceinline__ int qmc_dims(const DevInput* sp, int d) {
  int nd = 0;
  for (int j = 0; j < d; ++j) nd += sp[j].kind != DK_CONST;
  return nd;
}

// All Philox calls of one pair for a static plan.
template <class Plan>
__device__ __forceinline__ void static_gen(Philox4* ph, uint32_t qlo, uint32_t qhi,
                                           const PairCtr& pc) {
  constexpr int NC = static_calls<Plan>();
#pragma unroll
  for (int c = 0; c < NC; ++c) ph[c] = philox4x32_10(qlo, qhi, pc.stream, (uint32_t)c, pc.k0, pc.k1);
}

// Failure indicator of one sample. In the fast (non-strict) count mode a limit
// state may provide a sign-only test that skips the divisions and square roots
// whose value does not matter for `g < 0`.
template <class LS>
__device__ __forceinline__ bool sample_fails(const typename LS::Ctx& ctx, double* x) {
  if constexpr (LS::HAS_FAILS) {
    return LS::fails(ctx, x);
  } else {
    return LS::eval(ctx, x) < 0.0;
  }
}

// Limit states that decode a program per evaluation may offer eval_pair (both samples of a pair in one pass).
template <class LS, class = void>
struct ls_has_eval_pair { static constexpr bool value = false; };
template <class LS>
struct ls_has_eval_pair<LS, decltype((void)LS::HAS_EVAL_PAIR)> { static constexpr bool value = LS::HAS_EVAL_PAIR; };
template <class LS>
__device__ __forceinline__ void eval_two(const typename LS::Ctx& ctx, double* x0, double* x1, double& g0, double& g1) {
  if constexpr (ls_has_eval_pair<LS>::value) {
    LS::eval_pair(ctx, x0, x1, g0, g1);
  } else {
    g0 = LS::eval(ctx, x0);
    g1 = LS::eval(ctx, x1);
  }
}

// Front / back halves of a static plan (see InState).
template <class Plan, int J>
struct FrontAll {
  __device__ __forceinline__ static void run(const Philox4* ph, const FastMathSmem& sm, InState* st) {
    if constexpr (J < Plan::n) {
      StaticWords<static_word_offset<Plan, J>()> ws{ph};
      st[J] = draw_front<Plan::kinds[J]>(ws, sm);
      FrontAll<Plan, J + 1>::run(ph, sm, st);
    }
  }
};
template <class Plan, int J>
struct BackAll {
  __device__ __forceinline__ static void run(const DevInput* sp, const InState* st, const PairCtr& pc,
                                             const double* tables, const FastMathSmem& sm,
                                             double* x0, double* x1) {
    if constexpr (J < Plan::n) {
      draw_back<Plan::kinds[J], plan_uses_icdf<Plan>()>(sp[J], st[J], pc, (uint32_t)J, tables, sm, x0[J], x1[J]);
      BackAll<Plan, J + 1>::run(sp, st, pc, tables, sm, x0, x1);
    }
  }
};

// Count loop of one work item for a static plan, software-pipelined: the Philox words of the
// NEXT pair are generated and pushed through the front end of every input (bit placement,
// table loads) while the current pair goes through the fp64 back end and the limit state.
// The integer stream (ALU / IMAD.WIDE) and the fp64 stream then sit in the same basic block and
// ptxas interleaves them; without this, warps convoy through an all-integer phase and an
// all-fp64 phase (profiles/r01_b_*, tools/microbench), and the shared-memory latency of the
// log table is exposed.
// CHECK = the item touches the [first, last) boundary and samples must be masked.
// One pair of the pipelined count loop: generate and pre-process pair `qn` into `nxt` while pair `q`
// (already in `cur`) goes through the back end and the limit state.
template <class LS, class Plan, bool CHECK>
__device__ __forceinline__ unsigned static_count_step(const typename LS::Ctx& ctx, const DevInput* r_plan,
                                                      PairCtr& pc, int64_t q, int64_t qn,
                                                      const InState* cur, InState* nxt,
                                                      const McParams& p, const FastMathSmem& sm) {
  constexpr int XCAP = LS::XCAP;
  constexpr int NC = static_calls<Plan>() > 0 ? static_calls<Plan>() : 1;
  {  // prefetch (one surplus generation per item and thread)
    Philox4 ph[NC];
    static_gen<Plan>(ph, (uint32_t)qn, (uint32_t)((uint64_t)qn >> 32), pc);
    FrontAll<Plan, 0>::run(ph, sm, nxt);
  }
  pc.qlo = (uint32_t)q;
  pc.qhi = (uint32_t)((uint64_t)q >> 32);
  double x0[XCAP], x1[XCAP];
  BackAll<Plan, 0>::run(r_plan, cur, pc, p.tables, sm, x0, x1);
  const unsigned f0 = sample_fails<LS>(ctx, x0) ? 1u : 0u;
  const unsigned f1 = sample_fails<LS>(ctx, x1) ? 1u : 0u;
  if (CHECK) {
    const int64_t i0 = q << 1;
    return ((i0 >= p.first) ? f0 : 0u) + (((i0 + 1) < p.last) ? f1 : 0u);
  }
  return f0 + f1;
}

// Plans with more random inputs than this keep front and back end of a pair in the same trip: the state that
// would have to be carried (and copied) between trips grows with the number of inputs, and at eight inputs
// (Straub frame) the copies cost more than the overlap gains (profiles/r02_tune_w2.txt).
#ifndef EBN_PIPE_MAX_INPUTS
#define EBN_PIPE_MAX_INPUTS 7
#endif

template <class LS, class Plan, bool CHECK>
__device__ __forceinline__ unsigned static_count_loop(const typename LS::Ctx& ctx,
                                                      const DevInput* r_plan, PairCtr pc,
                                                      int64_t q_begin, int64_t q_end, int tid,
                                                      const McParams& p, const FastMathSmem& sm) {
  constexpr int NC = static_calls<Plan>() > 0 ? static_calls<Plan>() : 1;
  unsigned cnt = 0;
  int64_t q = q_begin + tid;
  if constexpr (Plan::n > EBN_PIPE_MAX_INPUTS) {
    constexpr int XCAP = LS::XCAP;
    for (; q < q_end; q += kThreads) {
      pc.qlo = (uint32_t)q;
      pc.qhi = (uint32_t)((uint64_t)q >> 32);
      double x0[XCAP], x1[XCAP];
      draw_inputs<Plan, XCAP>(r_plan, nullptr, p.d, pc, p.tables, sm, x0, x1);
      const unsigned f0 = sample_fails<LS>(ctx, x0) ? 1u : 0u;
      const unsigned f1 = sample_fails<LS>(ctx, x1) ? 1u : 0u;
      if (CHECK) {
        const int64_t i0 = q << 1;
        cnt += ((i0 >= p.first) ? f0 : 0u) + (((i0 + 1) < p.last) ? f1 : 0u);
      } else {
        cnt += f0 + f1;
      }
    }
    return cnt;
  }
  InState nxt[Plan::n];
  {
    Philox4 ph[NC];
    static_gen<Plan>(ph, (uint32_t)q, (uint32_t)((uint64_t)q >> 32), pc);
    FrontAll<Plan, 0>::run(ph, sm, nxt);
  }
#if EBN_PINGPONG
  // Two pairs per trip, the two pipeline stages swapping buffers: no register copies between trips.
  InState alt[Plan::n];
#pragma unroll 1
  for (; q + kThreads < q_end; q += 2 * kThreads) {
    cnt += static_count_step<LS, Plan, CHECK>(ctx, r_plan, pc, q, q + kThreads, nxt, alt, p, sm);
    cnt += static_count_step<LS, Plan, CHECK>(ctx, r_plan, pc, q + kThreads, q + 2 * kThreads, alt, nxt, p, sm);
  }
  if (q < q_end) cnt += static_count_step<LS, Plan, CHECK>(ctx, r_plan, pc, q, q + kThreads, nxt, alt, p, sm);
#else
#pragma unroll kPairUnroll
  for (; q < q_end; q += kThreads) {
    InState cur[Plan::n];
#pragma unroll
    for (int j = 0; j < Plan::n; ++j) cur[j] = nxt[j];
    cnt += static_count_step<LS, Plan, CHECK>(ctx, r_plan, pc, q, q + kThreads, cur, nxt, p, sm);
  }
#endif
  return cnt;
}

// Equally spaced edges (to 1e-9 of a step, checked by the host): the direct index floor((y - e0)/h) is
// within one cell of the truth, so the number of edges <= y is k + [edges[k] <= y] + [edges[k+1] <= y]
// with k the clamped estimate — two independent shared-memory loads and two comparisons, no loop and no
// branch, identical to the binary search (NaN -> nedges).
__device__ __forceinline__ int find_bin_uniform(const double* edges, int nedges, double y,
                                                double e0, double inv_h) {
  int k = __double2int_rd((y - e0) * inv_h);    // saturating; NaN -> 0
  k = max(0, min(k, nedges - 1));
  const int k1 = min(k + 1, nedges - 1);
  const double lo = edges[k], hi = edges[k1];
  int b = k + (int)(lo <= y) + (int)((k1 > k) & (hi <= y));
  return (y == y) ? b : nedges;
}

__device__ __forceinline__ int find_bin(const double* edges, int nedges, double y) {
  // number of edges <= y  (NaN -> nedges)
  if (!(y == y)) return nedges;
  int lo = 0, hi = nedges;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (edges[mid] <= y) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// MODE 0: failure count. MODE 1: histogram + moments of the output.
template <class LS, class Plan, int MODE>
__global__ void __launch_bounds__(kThreads, LS::MIN_BLOCKS) mc_kernel(const McParams p) {
  constexpr int XCAP = LS::XCAP;
  __shared__ FastMathSmem s_fm;
  __shared__ DevInput s_plan[EBN_MAX_INPUTS];
  __shared__ double s_chol[Plan::is_copula ? EBN_MAX_INPUTS * EBN_MAX_INPUTS : 1];
  __shared__ unsigned int s_red[kWarps];
  __shared__ double s_sum[MODE == 1 ? 2 * kWarps : 1];
  // dynamic shared memory: MODE 1: edges (double) then bins (u32); then, 16-byte aligned, the Phi^-1 cells of a
  // run-time dispatched plan that draws truncated normals (p.icdf_dyn)
  extern __shared__ __align__(16) unsigned char s_dyn[];
  double* s_edges = reinterpret_cast<double*>(s_dyn);
  unsigned int* s_bins = reinterpret_cast<unsigned int*>(s_dyn + sizeof(double) * (MODE == 1 ? p.nedges : 0));

  if (blockDim.x != kThreads) __trap();  // launched with a block shape the kernel was not compiled for

  typename LS::Ctx ctx;
  LS::setup(ctx, p.consts, p.n_consts, p.d, &s_fm);

  const int tid = threadIdx.x;
  fastmath_init(s_fm, tid, kThreads);
  if constexpr (plan_uses_icdf<Plan>()) icdf_init(tid, kThreads);
  if constexpr (plan_reads_icdf_dyn<Plan>())
    icdf_dyn_init(tid, kThreads, p.icdf_dyn ? s_dyn + hist_dyn_bytes(MODE == 1 ? p.nedges : 0) : nullptr);
  if constexpr (plan_is_qmc<Plan>::value) sobol_init(tid, kThreads);
  if constexpr (Plan::is_copula) phi_init(tid, kThreads);
  if (MODE == 1) {
    for (int i = tid; i < p.nedges; i += kThreads) s_edges[i] = p.edges[i];
  }

  for (int64_t li = blockIdx.x; li < p.local_items; li += gridDim.x) {
    const int64_t w = li * p.shard_world + p.shard_rank;  // global item
    const int s = (int)(w / p.items_per_scenario);
    const int64_t c = w - (int64_t)s * p.items_per_scenario;

    __syncthreads();  // previous item's readers are done with s_plan / s_bins
    if (tid < p.d) s_plan[tid] = p.plan[(int64_t)s * p.d + tid];
    if (Plan::is_copula) {
      const double* L = p.chol + (p.chol_shared ? 0 : (int64_t)s * p.d * p.d);
      for (int i = tid; i < p.d * p.d; i += kThreads) s_chol[i] = L[i];
    }
    if (MODE == 1) {
      for (int i = tid; i <= p.nedges; i += kThreads) s_bins[i] = 0u;
    }
    __syncthreads();

    PairCtr pc;
    pc.stream = p.stream_id ? p.stream_id[s] : (uint32_t)s;
    pc.k0 = p.k0;
    pc.k1 = p.k1;
    pc.ncalls = plan_is_runtime_dispatched<Plan>() ? pair_calls(s_plan, p.d, Plan::is_copula) : 0;

    const int64_t q_begin = p.pair_base + c * p.pairs_per_item;
    int64_t q_end = q_begin + p.pairs_per_item;
    const int64_t q_last = (p.last + 1) >> 1;  // one past the last pair
    if (q_end > q_last) q_end = q_last;

    // a static plan keeps the scenario's recipe in registers for the whole item
    DevInput r_plan[Plan::is_static ? XCAP : 1];
    if constexpr (Plan::is_static) {
#pragma unroll
      for (int j = 0; j < XCAP; ++j) r_plan[j] = s_plan[j];
    }
    const DevInput* plan_ptr = Plan::is_static ? r_plan : s_plan;

    unsigned int cnt = 0;
    double sum = 0.0, sumsq = 0.0;
    if constexpr (Plan::is_static && MODE == 0) {
      const bool interior = (q_begin << 1) >= p.first && (q_end << 1) <= p.last;
      cnt = interior ? static_count_loop<LS, Plan, false>(ctx, r_plan, pc, q_begin, q_end, tid, p, s_fm)
                     : static_count_loop<LS, Plan, true>(ctx, r_plan, pc, q_begin, q_end, tid, p, s_fm);
    } else {
    QmcState qs;
    int qmc_nd = 0;
    if constexpr (plan_is_qmc<Plan>::value) {
      qmc_nd = qmc_dims(s_plan, p.d);
      qmc_keys(qs, pc.stream, pc.k0, pc.k1, qmc_nd);
      qmc_seek(qs, (uint32_t)((q_begin + tid) << 1), qmc_nd);
    }
    for (int64_t q = q_begin + tid; q < q_end; q += kThreads) {
      pc.qlo = (uint32_t)q;
      pc.qhi = (uint32_t)((uint64_t)q >> 32);
      double x0[XCAP], x1[XCAP];
      if constexpr (plan_is_qmc<Plan>::value) {
        qmc_move(qs, (uint32_t)(q << 1), qmc_nd);
        draw_inputs_qmc<XCAP, true>(s_plan, p.d, qs, p.tables, s_fm, x0, x1);
      } else {
        draw_inputs<Plan, XCAP, true>(plan_ptr, s_chol, p.d, pc, p.tables, s_fm, x0, x1);
      }
      const int64_t i0 = q << 1;
      const bool v0 = i0 >= p.first;           // i0 < last holds by construction
      const bool v1 = (i0 + 1) < p.last;       // i0+1 >= first holds by construction
      if (MODE == 0) {
        bool f0, f1;
        if constexpr (ls_has_eval_pair<LS>::value && !LS::HAS_FAILS) {
          double g0, g1;
          eval_two<LS>(ctx, x0, x1, g0, g1);
          f0 = g0 < 0.0;
          f1 = g1 < 0.0;
        } else {
          f0 = sample_fails<LS>(ctx, x0);
          f1 = sample_fails<LS>(ctx, x1);
        }
        cnt += (unsigned)(v0 && f0) + (unsigned)(v1 && f1);
      } else {
        double g0, g1;
        eval_two<LS>(ctx, x0, x1, g0, g1);
        const int b0 = p.hist_uniform ? find_bin_uniform(s_edges, p.nedges, g0, p.hist_e0, p.hist_inv_h)
                                      : find_bin(s_edges, p.nedges, g0);
        const int b1 = p.hist_uniform ? find_bin_uniform(s_edges, p.nedges, g1, p.hist_e0, p.hist_inv_h)
                                      : find_bin(s_edges, p.nedges, g1);
        if (v0) {
          atomicAdd(&s_bins[b0], 1u);
          if (g0 == g0) { sum += g0; sumsq = fma(g0, g0, sumsq); }
        }
        if (v1) {
          atomicAdd(&s_bins[b1], 1u);
          if (g1 == g1) { sum += g1; sumsq = fma(g1, g1, sumsq); }
        }
      }
    }
    }

    if (MODE == 0) {
      // warp shuffle -> shared memory -> one 64-bit atomic per item
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if ((tid & 31) == 0) s_red[tid >> 5] = cnt;
      __syncthreads();
      if (tid == 0) {
        unsigned long long tot = 0;
#pragma unroll
        for (int i = 0; i < kWarps; ++i) tot += s_red[i];
        if (tot) atomicAdd(&p.counts[s], tot);
      }
    } else {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_down_sync(0xffffffffu, sum, o);
        sumsq += __shfl_down_sync(0xffffffffu, sumsq, o);
      }
      if ((tid & 31) == 0) {
        s_sum[2 * (tid >> 5)] = sum;
        s_sum[2 * (tid >> 5) + 1] = sumsq;
      }
      __syncthreads();
      for (int i = tid; i <= p.nedges; i += kThreads) {
        const unsigned int b = s_bins[i];
        if (b) atomicAdd(&p.counts[(int64_t)s * (p.nedges + 1) + i], (unsigned long long)b);
      }
      if (tid == 0) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int i = 0; i < kWarps; ++i) {
          a += s_sum[2 * i];
          b += s_sum[2 * i + 1];
        }
        // per-item partials, combined in item order afterwards (deterministic)
        p.partial[2 * li] = a;
        p.partial[2 * li + 1] = b;
      }
    }
  }
}

// Sampler replay: the n samples [first, first+n) of one scenario, exactly as
// mc_kernel draws them, written column-major with ld = n.
template <class LS, class Plan>
__global__ void __launch_bounds__(kThreads) replay_kernel(const McParams p, int scenario,
                                                          int64_t n, double* X, double* g_out) {
  constexpr int XCAP = LS::XCAP;
  __shared__ FastMathSmem s_fm;
  __shared__ DevInput s_plan[EBN_MAX_INPUTS];
  __shared__ double s_chol[Plan::is_copula ? EBN_MAX_INPUTS * EBN_MAX_INPUTS : 1];
  if (blockDim.x != kThreads) __trap();
  typename LS::Ctx ctx;
  LS::setup(ctx, p.consts, p.n_consts, p.d, &s_fm);
  fastmath_init(s_fm, threadIdx.x, kThreads);
  if constexpr (plan_uses_icdf<Plan>()) icdf_init(threadIdx.x, kThreads);
  if constexpr (plan_reads_icdf_dyn<Plan>()) icdf_dyn_init(threadIdx.x, kThreads, nullptr);
  if constexpr (plan_is_qmc<Plan>::value) sobol_init(threadIdx.x, kThreads);
  if constexpr (Plan::is_copula) phi_init(threadIdx.x, kThreads);
  if (threadIdx.x < p.d) s_plan[threadIdx.x] = p.plan[(int64_t)scenario * p.d + threadIdx.x];
  if (Plan::is_copula) {
    const double* L = p.chol + (p.chol_shared ? 0 : (int64_t)scenario * p.d * p.d);
    for (int i = threadIdx.x; i < p.d * p.d; i += kThreads) s_chol[i] = L[i];
  }
  __syncthreads();
  PairCtr pc;
  pc.stream = p.stream_id ? p.stream_id[scenario] : (uint32_t)scenario;
  pc.k0 = p.k0;
  pc.k1 = p.k1;
  pc.ncalls = plan_is_runtime_dispatched<Plan>() ? pair_calls(s_plan, p.d, Plan::is_copula) : 0;
  const int64_t q_last = (p.last + 1) >> 1;
  for (int64_t q = p.pair_base + (int64_t)blockIdx.x * kThreads + threadIdx.x; q < q_last;
       q += (int64_t)gridDim.x * kThreads) {
    pc.qlo = (uint32_t)q;
    pc.qhi = (uint32_t)((uint64_t)q >> 32);
    double x0[XCAP], x1[XCAP];
    if constexpr (plan_is_qmc<Plan>::value) {
      QmcState qs;
      const int nd = qmc_dims(s_plan, p.d);
      qmc_keys(qs, pc.stream, pc.k0, pc.k1, nd);
      qmc_seek(qs, (uint32_t)(q << 1), nd);
      draw_inputs_qmc<XCAP>(s_plan, p.d, qs, p.tables, s_fm, x0, x1);
    } else {
      draw_inputs<Plan, XCAP>(s_plan, s_chol, p.d, pc, p.tables, s_fm, x0, x1);
    }
    const int64_t i0 = q << 1;
    double y0[XCAP], y1[XCAP];
#pragma unroll
    for (int j = 0; j < XCAP; ++j) { y0[j] = x0[j]; y1[j] = x1[j]; }
    const double g0 = LS::eval(ctx, y0);
    const double g1 = LS::eval(ctx, y1);
    if (i0 >= p.first) {
      const int64_t r = i0 - p.first;
#pragma unroll
      for (int j = 0; j < XCAP; ++j)
        if (j < p.ncols) X[(int64_t)j * n + r] = y0[j];
      if (g_out) g_out[r] = g0;
    }
    if (i0 + 1 < p.last) {
      const int64_t r = i0 + 1 - p.first;
#pragma unroll
      for (int j = 0; j < XCAP; ++j)
        if (j < p.ncols) X[(int64_t)j * n + r] = y1[j];
      if (g_out) g_out[r] = g1;
    }
  }
}

// Limit state over a resident sample matrix (column-major). HBM-bound:
// d*8 bytes read per row (+8 written when g_out is requested).
template <class LS>
__global__ void __launch_bounds__(kThreads) matrix_kernel(const double* consts, int n_consts,
                                                          const double* X, int64_t n, int d,
                                                          int64_t ld, unsigned long long* count,
                                                          double* g_out) {
  constexpr int XCAP = LS::XCAP;
  __shared__ unsigned int s_red[kWarps];
  __shared__ FastMathSmem s_fm;
  if (blockDim.x != kThreads) __trap();
  typename LS::Ctx ctx;
  LS::setup(ctx, consts, n_consts, d, &s_fm);
  fastmath_init(s_fm, threadIdx.x, kThreads);
  __syncthreads();
  unsigned int cnt = 0;
  for (int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x; r < n;
       r += (int64_t)gridDim.x * kThreads) {
    double x[XCAP];
#pragma unroll
    for (int j = 0; j < XCAP; ++j) x[j] = (j < d) ? __ldg(&X[(int64_t)j * ld + r]) : 0.0;
    const double g = LS::eval(ctx, x);
    if (g_out) g_out[r] = g;
    cnt += (unsigned)(g < 0.0);
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long tot = 0;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) tot += s_red[i];
    if (tot) atomicAdd(count, tot);
  }
}

#ifndef __CUDACC_RTC__
// ---------------------------------------------------------------------------
// per-limit-state table of launchers (one translation unit per limit state so
// the build parallelises)
// ---------------------------------------------------------------------------
// LST: the limit-state functor template; SPlans: its compile-time sampling plans, tried in order.
// plan_class = PLAN_DYNAMIC, PLAN_COPULA, or PLAN_STATIC0 + index into SPlans.
template <template <class> class LST, class... SPlans>
struct LsLaunch {
  template <class LS, class Plan>
  static cudaError_t occ2(int mode, size_t dyn, int* out) {
    if (mode == 0) {
      if (dyn) cudaFuncSetAttribute(mc_kernel<LS, Plan, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
      return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, mc_kernel<LS, Plan, 0>, kThreads, dyn);
    }
    // histogram mode: static tables + dynamic edges and bins may pass 48 KB together: opt in
    cudaFuncSetAttribute(mc_kernel<LS, Plan, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, mc_kernel<LS, Plan, 1>, kThreads, dyn);
  }
  template <class LS, class Plan>
  static cudaError_t launch2(int mode, const McParams& p, int blocks, size_t dyn, cudaStream_t st) {
    if (mode == 0) {
      if (dyn) cudaFuncSetAttribute(mc_kernel<LS, Plan, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
      mc_kernel<LS, Plan, 0><<<blocks, kThreads, dyn, st>>>(p);
    } else {
      cudaFuncSetAttribute(mc_kernel<LS, Plan, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
      mc_kernel<LS, Plan, 1><<<blocks, kThreads, dyn, st>>>(p);
    }
    return cudaGetLastError();
  }
  template <class LS, class Plan>
  static cudaError_t replay2(const McParams& p, int scenario, int64_t n, double* X, double* g,
                             int blocks, cudaStream_t st) {
    replay_kernel<LS, Plan><<<blocks, kThreads, 0, st>>>(p, scenario, n, X, g);
    return cudaGetLastError();
  }

  // Calls f.template operator()<Plan>() for the plan plan_class names.
  template <class F, class First, class... Rest>
  static cudaError_t pick_static(int index, F&& f) {
    if (index == 0) return f.template operator()<First>();
    if constexpr (sizeof...(Rest) > 0) return pick_static<F, Rest...>(index - 1, static_cast<F&&>(f));
    return f.template operator()<DynamicPlan>();
  }
  template <class F>
  static cudaError_t with_plan(int plan_class, F&& f) {
    if (plan_class == PLAN_COPULA) return f.template operator()<CopulaPlan>();
    if (plan_class == PLAN_QMC) return f.template operator()<QmcPlan>();
    if constexpr (sizeof...(SPlans) > 0) {
      if (plan_class >= PLAN_STATIC0) return pick_static<F, SPlans...>(plan_class - PLAN_STATIC0, static_cast<F&&>(f));
    }
    return f.template operator()<DynamicPlan>();
  }

  template <class LS>
  struct Occ {
    int mode; size_t dyn; int* out;
    template <class Plan> cudaError_t operator()() const { return occ2<LS, Plan>(mode, dyn, out); }
  };
  template <class LS>
  struct Launch {
    int mode; const McParams& p; int blocks; size_t dyn; cudaStream_t st;
    template <class Plan> cudaError_t operator()() const { return launch2<LS, Plan>(mode, p, blocks, dyn, st); }
  };
  template <class LS>
  struct Replay {
    const McParams& p; int scenario; int64_t n; double* X; double* g; int blocks; cudaStream_t st;
    template <class Plan> cudaError_t operator()() const { return replay2<LS, Plan>(p, scenario, n, X, g, blocks, st); }
  };

  static cudaError_t occupancy(int strict, int plan_class, int mode, size_t dyn, int* out) {
    return strict ? with_plan(plan_class, Occ<LST<StrictOps>>{mode, dyn, out})
                  : with_plan(plan_class, Occ<LST<FastOps>>{mode, dyn, out});
  }
  static cudaError_t launch(int strict, int plan_class, int mode, const McParams& p, int blocks,
                            size_t dyn, cudaStream_t st) {
    return strict ? with_plan(plan_class, Launch<LST<StrictOps>>{mode, p, blocks, dyn, st})
                  : with_plan(plan_class, Launch<LST<FastOps>>{mode, p, blocks, dyn, st});
  }
  static cudaError_t replay(int strict, int plan_class, const McParams& p, int scenario, int64_t n,
                            double* X, double* g, int blocks, cudaStream_t st) {
    return strict ? with_plan(plan_class, Replay<LST<StrictOps>>{p, scenario, n, X, g, blocks, st})
                  : with_plan(plan_class, Replay<LST<FastOps>>{p, scenario, n, X, g, blocks, st});
  }
  static cudaError_t matrix(int strict, const double* consts, int n_consts, const double* X,
                            int64_t n, int d, int64_t ld, unsigned long long* count, double* g_out,
                            int blocks, cudaStream_t st) {
    if (strict)
      matrix_kernel<LST<StrictOps>><<<blocks, kThreads, 0, st>>>(consts, n_consts, X, n, d, ld, count, g_out);
    else
      matrix_kernel<LST<FastOps>><<<blocks, kThreads, 0, st>>>(consts, n_consts, X, n, d, ld, count, g_out);
    return cudaGetLastError();
  }
  static LsVTable vtable() {
    static_assert(sizeof...(SPlans) <= kMaxStaticPlans, "raise kMaxStaticPlans");
    LsVTable v{};
    v.occupancy = &occupancy;
    v.launch = &launch;
    v.replay = &replay;
    v.matrix = &matrix;
    const int* kinds[] = {SPlans::kinds..., nullptr};
    const int ns[] = {SPlans::n..., 0};
    v.n_plans = (int)sizeof...(SPlans);
    for (int i = 0; i < v.n_plans; ++i) {
      v.static_kinds[i] = kinds[i];
      v.n_static[i] = ns[i];
    }
    return v;
  }
};

#endif  // !__CUDACC_RTC__

}  // namespace ebn
