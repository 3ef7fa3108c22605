/*
 * ebncuda.h — C ABI of libebncuda.so
 *
 * B200 (sm_100a) Monte Carlo evaluation of EnhancedBayesianNetworks.jl
 * functional nodes. The reference has no FFI today; the seam this library
 * plugs into is Julia dispatch on a functional node's `simulation` field
 * (reference: src/networks/ebn/evaluation/evaluate_node.jl:83 for
 * DiscreteFunctionalNode, :24-26 for ContinuousFunctionalNode). Each entry
 * point below names the reference call site it replaces.
 *
 * Conventions
 *   - Every function returns 0 (EBN_OK) or a negative EBN_E* code; the text of
 *     the last error on the calling thread is available via ebn_last_error().
 *   - The caller owns every host buffer passed in or out. The library owns all
 *     device memory, streams, events and NCCL communicators.
 *   - Plain-old-data structs, fixed-width fields, natural alignment, no
 *     implicit padding (checked by static_asserts in the implementation).
 *   - No CPU fallback: without a CUDA device every compute call fails with
 *     EBN_ENODEV.
 *   - Calls are blocking. The library is safe to call from several host threads: entry points that touch
 *     its state take one process-wide lock, so concurrent calls are serialised (one batch in flight per
 *     process; a batch already uses every device the library was initialised with). ebn_last_error() is
 *     per thread; ebn_last_stats() describes the most recent call of any thread.
 */
#ifndef EBNCUDA_H
#define EBNCUDA_H

#ifndef __CUDACC_RTC__ /* the device code is also compiled at run time, without host headers */
#include <stddef.h>
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define EBN_ABI_VERSION 2
#define EBN_MAX_INPUTS 16 /* columns a limit state can read (random + Parameter) */
#define EBN_MAX_EDGES 1025 /* histogram edges per ebn_hist_batch call */

/* ---- error codes ------------------------------------------------------- */
enum {
  EBN_OK = 0,
  EBN_EINVAL = -1,       /* bad argument (null pointer, negative size, ...)   */
  EBN_ELIMITSTATE = -2,  /* unknown limit state id or wrong arity/consts       */
  EBN_EMARGINAL = -3,    /* unsupported marginal / truncation combination      */
  EBN_ECUDA = -4,        /* CUDA runtime error (text in ebn_last_error)        */
  EBN_ENCCL = -5,        /* NCCL missing or NCCL error                         */
  EBN_ENODEV = -6,       /* no CUDA device / library not initialised           */
  EBN_ENOMEM = -7,       /* host or device allocation failed                   */
  EBN_ESTATE = -8,       /* call not valid in the current state                */
  EBN_EJIT = -9          /* run-time compilation unavailable (no libnvrtc) or failed;
                            the compiler log is in ebn_last_error                */
};

/* ---- limit states compiled into the kernels as device functors ---------- */
enum {
  /* demo/vehicle_suspension.jl:17-25. inputs (A, b0, V, C, Ck, K);
   * consts (M, m, g, (M*g)^-1.5, (g*(M+m))^0.877) — the two powers are
   * computed by the CALLER so they carry the caller's pow() rounding.        */
  EBN_LS_VEHICLE_SUSPENSION = 1,
  /* demo/straub_example.jl:16-39 after _transfer_continuous! fused R1..R5
   * into the frame node. inputs (Ur, V, H, e1..e5) where e1..e5 are the five
   * in-model standard-normal draws (`rand(Normal(mu, sd))` inside the model
   * closure); consts (lambda, zeta, rho).                                     */
  EBN_LS_STRAUB_FRAME = 2,
  /* demo/Hydrogen_new/ebn.jl:98-111 (only the part that is in the repo):
   * inputs (scenario, r_operators, burn_r, explosion_r); no consts.          */
  EBN_LS_HYDROGEN_RADIUS = 3,
  /* Generic chain of polynomial models followed by a polynomial performance
   * (the toy limit states of test/networks/ebn/evaluate/evaluate_node.jl:96-98
   * and friends). consts encode the program, see "polynomial program" below. */
  EBN_LS_POLYNOMIAL = 4,
  /* min over several affine forms  g = min_k (c_k0 + sum_j c_kj x_j).
   * consts (K, then K rows of 1+d coefficients). d = n_inputs of the job.     */
  EBN_LS_MIN_AFFINE = 5,
  /* Straub frame with R4 and R5 observed (test/inference/variableselimination.jl:207-284: R4, R5
   * are discretised continuous functional nodes, so the frame reads their columns directly).
   * inputs (Ur, V, H, R4, R5, e1, e2, e3); consts (lambda, zeta, rho).                     */
  EBN_LS_STRAUB_FRAME_R45 = 6,
  /* One plastic moment capacity r = exp(lambda + rho*zeta*u_r + sqrt((1-rho^2) zeta^2) * e)
   * (demo/straub_example.jl:16-26) — the model of a ContinuousFunctionalNode R_i that is
   * evaluated on its own (histogram path). inputs (Ur, e); consts (lambda, zeta, rho).
   * As a limit state its value is r itself.                                              */
  EBN_LS_STRAUB_CAPACITY = 7,
  /* Any chain of algebraic models and a performance function, given as a stack program in
   * consts (see "expression program" below) and compiled for the GPU at run time (NVRTC).
   * This is the route for user-written `Model(df -> ..., :name)` closures
   * (src/nodes/functionalnodes.jl:27-52). d = n_inputs of the job.                          */
  EBN_LS_EXPRESSION = 8
};

/*
 * Polynomial program (EBN_LS_POLYNOMIAL), all numbers stored as doubles:
 *   consts[0]            = number of stages T (models + 1 performance stage)
 *   per stage:  n_terms, then per term: coefficient, n_factors,
 *               then n_factors column indices (a column may repeat: x*x).
 * Stage t appends column d+t (0-based) = sum of its terms evaluated left to
 * right: term = ((coef * x_a) * x_b) ...; acc = ((term_0 + term_1) + ...).
 * The last stage is the performance function; its value is g. A term with
 * zero factors is the constant `coef`. At most EBN_MAX_INPUTS columns total.
 */

/*
 * Expression program (EBN_LS_EXPRESSION), all numbers stored as doubles:
 *   consts[0] = P, the number of runtime constants; consts[1] = L, the number of instructions;
 *   consts[2 .. 2+P) = the runtime constants k[0..P);
 *   then L instructions of two doubles each: (opcode, operand).
 * The program runs on an evaluation stack, once per sample. Columns 0..d-1 are the job's inputs;
 * every EBN_OP_STORE appends the next column (a model output, readable by later instructions and
 * returned by ebn_sample_matrix_ex). After the last instruction the stack holds exactly one value,
 * g; the sample fails iff g < 0. Kernels are cached by (instructions, literals, d, sampling plan,
 * strict): changing only the runtime constants reuses the compiled kernel.
 * Strict mode executes one IEEE operation per instruction in program order, which is Julia's
 * evaluation order when the wrapper emits the expression tree in post-order; sqrt/log of a
 * negative number gives NaN (a safe sample) where Julia would throw a DomainError.
 * Outside strict mode the failure count of ebn_pf_batch may come from a sign-only predicate derived from
 * the program (min -> or, square roots against constants -> comparisons of the radicand, quotients ->
 * comparisons of numerator and denominator; a sample with a zero or infinite divisor is evaluated in full):
 * it agrees with `g < 0` except on ties, as long as products of the formula's numerators and denominators
 * stay finite. ebn_hist_batch, ebn_eval_matrix and ebn_sample_matrix_ex always evaluate g.
 * exp, log, sin, cos and pow are within 1-2 ulp of Julia's, not bit-identical. min / max of two
 * zeros of opposite sign return the first operand (Julia orders -0.0 below 0.0); `g < 0` cannot tell.
 */
enum {
  EBN_OP_INPUT = 0,   /* push column `operand` (an input or an already stored column)        */
  EBN_OP_PARAM = 1,   /* push runtime constant k[operand]                                    */
  EBN_OP_CONST = 2,   /* push the literal `operand`                                          */
  EBN_OP_ADD = 3,     /* a b -> a + b                                                        */
  EBN_OP_SUB = 4,     /* a b -> a - b                                                        */
  EBN_OP_MUL = 5,     /* a b -> a * b                                                        */
  EBN_OP_DIV = 6,     /* a b -> a / b                                                        */
  EBN_OP_NEG = 7,     /* a -> -a                                                             */
  EBN_OP_ABS = 8,     /* a -> |a|                                                            */
  EBN_OP_SQRT = 9,    /* a -> sqrt(a)                                                        */
  EBN_OP_EXP = 10,    /* a -> exp(a)                                                         */
  EBN_OP_LOG = 11,    /* a -> log(a)                                                         */
  EBN_OP_SIN = 12,    /* a -> sin(a)                                                         */
  EBN_OP_COS = 13,    /* a -> cos(a)                                                         */
  EBN_OP_MIN = 14,    /* a b -> min(a, b), NaN-propagating like Julia's min                  */
  EBN_OP_MAX = 15,    /* a b -> max(a, b), NaN-propagating                                   */
  EBN_OP_POW = 16,    /* a b -> a ^ b                                                        */
  EBN_OP_POWI = 17,   /* a -> a ^ operand for a literal integer: 2 -> a*a, 3 -> (a*a)*a as
                         Julia's literal_pow does; other exponents go through pow()          */
  EBN_OP_LT = 18,     /* a b -> 1.0 if a < b else 0.0                                        */
  EBN_OP_SELECT = 19, /* c a b -> a if c != 0 else b                                         */
  EBN_OP_STORE = 20,  /* a -> (nothing); appends `a` as the next column                      */
  EBN_OP_NOPS = 21
};
#define EBN_EXPR_MAX_OPS 4096
#define EBN_EXPR_MAX_PARAMS 64
#define EBN_EXPR_MAX_DEPTH 64

/* ---- marginals ---------------------------------------------------------- */
enum {
  EBN_IN_PARAM = 0,       /* UQ Parameter: constant p[0] (discretenodes.jl:43-49) */
  EBN_IN_NORMAL = 1,      /* Normal(mu=p[0], sigma=p[1])                       */
  EBN_IN_LOGNORMAL = 2,   /* LogNormal(mu=p[0], sigma=p[1]) (log-space params) */
  EBN_IN_UNIFORM = 3,     /* Uniform(a=p[0], b=p[1])                           */
  EBN_IN_GUMBEL = 4,      /* Gumbel(mu=p[0], beta=p[1]) (maximum type)         */
  EBN_IN_GAMMA = 5,       /* Gamma(shape=p[0], scale=p[1])                     */
  EBN_IN_EXPONENTIAL_AFFINE = 6, /* p[2] + p[1]*Exponential(theta=p[0]), p[1]=+-1
                                    (discretize.jl:83-91 `_approximate`)        */
  EBN_IN_TABULATED = 7    /* inverse CDF table: x = Q(u), Q piecewise linear on a
                             uniform u grid; p[0] = offset into job->tables,
                             p[1] = number of points (>= 2)                     */
};

/* One input of one scenario. `lo`/`hi` is the truncation window of
 * `truncated(dist, lo, hi)` (continuousnodes.jl:71-73); +-INFINITY = none. */
typedef struct ebn_input {
  int32_t kind;
  int32_t flags; /* reserved, must be 0 */
  double p[4];
  double lo;
  double hi;
} ebn_input_t; /* 56 bytes */

/* job->flags */
#define EBN_JOB_COPULA_SHARED 1u /* corr_chol is one [d*d] block for all scenarios */
#define EBN_JOB_JIT_PLAN 4u      /* compile the job's sampling plan into the kernel at run time even for
                                    a small job, and run the polynomial / min-affine functors as
                                    compiled expression programs instead of interpreting their
                                    coefficient tables (default: only when the job is large enough to
                                    repay ~1 s of compilation: 5e10 samples for the two functors,
                                    1e11 for a sampling plan alone). Strict results do not change.  */
#define EBN_JOB_NO_JIT_PLAN 8u   /* never do that: runtime-dispatched plan, interpreted functors     */
#define EBN_JOB_FP32_SAMPLING 16u /* "FP32 where tolerated": untruncated (Log)Normal inputs are drawn with
                                    the Box-Muller transform evaluated in fp32 (same random words; each
                                    normal agrees with the fp64 draw to ~1e-6 sigma, tail to 6.7 sigma);
                                    the limit state stays in fp64. About 1.3x faster on the vehicle
                                    workload. Ignored under a copula.                               */
#define EBN_JOB_QMC_SOBOL 32u    /* quasi-Monte Carlo (UQ.jl SobolSampling, an AbstractMonteCarlo): sample i of a
                                    scenario is point i of an Owen-scrambled Sobol sequence (Joe-Kuo direction
                                    numbers, hash-based nested uniform scrambling keyed by seed and stream id),
                                    every marginal through its inverse CDF. At most 2^32 points per scenario,
                                    no Gaussian copula, Gamma as EBN_IN_TABULATED. pf = count/n as before;
                                    `var` is still (pf - pf^2)/n, an upper bound in practice                */
#define EBN_JOB_PARTIAL 2u       /* sharded job: return THIS shard's counts, no all-reduce
                                    (the caller combines shards; also used to emulate
                                    several ranks on one GPU in tests)                  */

/*
 * One batch = every scenario of one functional node (the whole loop at
 * evaluate_node.jl:79-119 in a single call).
 */
typedef struct ebn_job {
  int32_t limit_state;        /* EBN_LS_*                                        */
  int32_t n_consts;
  const double* consts;       /* [n_consts]                                      */
  int32_t n_scenarios;        /* S                                               */
  int32_t n_inputs;           /* d: functor argument count, Parameters included  */
  const ebn_input_t* inputs;  /* [S*d], scenario-major                           */
  int64_t n_per_scenario;     /* n: samples per scenario (MonteCarlo(n).n)       */
  uint64_t seed;              /* Philox key                                      */
  int32_t strict;             /* 1: Julia evaluation order, no FMA contraction   */
  uint32_t flags;
  const double* corr_chol;    /* nullable. Gaussian copula: lower-triangular
                                 Cholesky factor(s) of the correlation matrix,
                                 row-major [S*d*d] (or [d*d] with
                                 EBN_JOB_COPULA_SHARED). Rows/cols of Parameter
                                 inputs must be identity.                         */
  const uint32_t* stream_id;  /* nullable [S]. Philox stream of each scenario;
                                 default = scenario index. Equal ids give common
                                 random numbers (imprecise outer loop).           */
  const double* tables;       /* nullable. Pool of inverse-CDF tables             */
  int64_t n_tables;           /* doubles in `tables`                              */
  int64_t sample_offset;      /* first sample index of this call (resume)         */
  int32_t shard_rank;         /* this process' share of the (scenario, chunk)     */
  int32_t shard_world;        /* work items; 0/1 = everything                     */
} ebn_job_t;

typedef struct ebn_stats {
  double kernel_ms;        /* device time of the last call's kernels (CUDA events
                              on the launching stream)                           */
  double h2d_bytes;        /* descriptor bytes copied host->device               */
  double d2h_bytes;        /* result bytes copied device->host                   */
  int64_t samples;         /* limit-state evaluations done by THIS process       */
  int32_t launches;        /* kernels launched by the last call                  */
  int32_t devices;         /* devices that took part                             */
  int32_t specialised;     /* 1 if a compile-time sampling plan was used         */
  int32_t jit_compiles;    /* kernels compiled at run time by the last call (0 when
                              they came from the cache)                          */
} ebn_stats_t;

#ifndef __CUDACC_RTC__ /* host entry points */
/* ---- lifecycle ---------------------------------------------------------- */
int ebn_abi_version(void);
/* Identity of the device code in this library: a hash of the kernel sources and tuning flags it was
 * built from. Measurements frozen under profiles/ name the build they belong to. No GPU needed.   */
const char* ebn_build_id(void);
int ebn_device_count(void);
/* devices == NULL, ndev == 0: device 0 only. ndev > 1: single process driving
 * several GPUs; counts are combined with ncclAllReduce (ncclCommInitAll).    */
int ebn_init(const int* devices, int ndev);
int ebn_shutdown(void);
/* Launch on a caller-owned cudaStream_t (single-device mode); NULL restores the
 * library's own stream. */
int ebn_set_stream(void* cuda_stream);

/* One process per GPU (torchrun): rank 0 makes an id, the host framework
 * broadcasts the 128 bytes, every rank joins. After that ebn_pf_batch /
 * ebn_hist_batch with shard_world == world all-reduce their int64 vectors.   */
int ebn_nccl_unique_id(void* id128);
int ebn_comm_init_rank(const void* id128, int rank, int world);
int ebn_comm_destroy(void);

/* ---- the hot path ------------------------------------------------------- */
/*
 * Replaces the S calls of `probability_of_failure(models, performance, inputs,
 * sim::MonteCarlo)` at evaluate_node.jl:83. fail_count[s] = #{g < 0} (strict
 * `<`, NaN counts as safe), pf[s] = fail_count[s]/n, var[s] = (pf - pf^2)/n.
 * pf and var may be NULL.
 */
int ebn_pf_batch(const ebn_job_t* job, int64_t* fail_count, double* pf, double* var);

/*
 * Replaces `UQ.sample` + `UQ.evaluate!` + `EmpiricalDistribution` at
 * evaluate_node.jl:24-26 for ContinuousFunctionalNode: per scenario a
 * histogram of the last model's output over `edges` (counts[s*(nedges+1)+b];
 * bin 0 = y < edges[0], bin b = edges[b-1] <= y < edges[b], bin nedges =
 * y >= edges[nedges-1] or NaN), plus sum and sum of squares of y. For
 * EBN_LS_POLYNOMIAL the output is the last stage; for the other limit states
 * it is g.
 */
int ebn_hist_batch(const ebn_job_t* job, const double* edges, int nedges,
                   int64_t* counts, double* sum, double* sumsq);

/*
 * Limit state over a caller-supplied sample matrix (column-major, leading
 * dimension ld >= n): the bit-exact parity entry point for matrices exported
 * from the reference (`performance(df) .< 0` summed). g_out nullable [n].
 */
int ebn_eval_matrix(int limit_state, const double* consts, int n_consts,
                    const double* X, int64_t n, int d, int64_t ld, int strict,
                    int64_t* fail_count, double* g_out);

/*
 * Replays the sampler: writes the n samples [first, first+n) of scenario
 * `scenario` exactly as ebn_pf_batch draws them (column-major X[n*d], ld = n)
 * and, if g_out != NULL, the limit state on them. This is what
 * `additional_info[...][:samples]` (evaluate_node.jl:93-95) is filled from.
 */
int ebn_sample_matrix(const ebn_job_t* job, int scenario, int64_t first,
                      int64_t n, double* X, double* g_out);
/* Same, with ncols >= d columns per sample: columns d.. are the model outputs the
 * limit state defines (the stages of a polynomial program), i.e. the DataFrame
 * UQ.jl holds after `evaluate!(models, samples)`. X is [n*ncols], ld = n. */
int ebn_sample_matrix_ex(const ebn_job_t* job, int scenario, int64_t first,
                         int64_t n, int ncols, double* X, double* g_out);

/*
 * The static work partition of a job (host logic, needs no device): a work item
 * is (scenario, chunk of sample pairs); item w = scenario*items_per_scenario +
 * chunk belongs to rank w % shard_world. out = {pairs_per_item,
 * items_per_scenario, items of job->shard_rank, first pair index}. The
 * partition depends on the job only, so all ranks derive the same grid, and
 * counts never depend on it (the RNG is a function of seed, stream, sample).
 */
int ebn_partition(const ebn_job_t* job, int histogram, int64_t out[4]);

/* ---- measurement -------------------------------------------------------- */
/* Independent DFMA chains at full occupancy for about `seconds`; DFMA/s.    */
int ebn_peak_fp64(double seconds, double* dfma_per_s);
int ebn_last_stats(ebn_stats_t* out);

/* ---- introspection ------------------------------------------------------ */
const char* ebn_last_error(void);
int ebn_limit_state_id(const char* name);       /* <0 if unknown */
const char* ebn_limit_state_name(int id);       /* NULL if unknown */
/* n_inputs / n_consts = -1 when the limit state is variadic. */
int ebn_limit_state_arity(int id, int* n_inputs, int* n_consts);

/* ---- run-time compilation (expression limit states, sampling plans) ------ */
/* 1 if libnvrtc could be loaded, else 0 (reason in ebn_last_error). No GPU needed. */
int ebn_jit_available(void);
/* Validates an expression program for d inputs without running anything (no GPU needed).
 * On success *n_columns (nullable) = d + the number of stored columns.                    */
int ebn_expression_check(const double* consts, int n_consts, int d, int* n_columns);
/* Writes the CUDA source generated for the program into buf (NUL-terminated, truncated to cap);
 * returns the full length. For inspection and tests; no GPU needed.                        */
int64_t ebn_expression_source(const double* consts, int n_consts, int d, char* buf, int64_t cap);
/* Compiles (or finds in the cache) the kernel `job` would run, without running it; no GPU needed.
 * role 0 = count (ebn_pf_batch), 1 = histogram (ebn_hist_batch), 2 = replay (ebn_sample_matrix).
 * *cubin_bytes (nullable) = size of the machine code, 0 when the job uses a compiled-in kernel.
 * Lets a caller pay the compilation ahead of time and check a program compiles.            */
int ebn_jit_compile(const ebn_job_t* job, int role, int64_t* cubin_bytes);
#endif /* !__CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif /* EBNCUDA_H */
