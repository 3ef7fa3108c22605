/*
 * ebn_oracle.c — CPU oracle. TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this file's library; libebncuda never does.
 *
 * A plain-C restatement of the reference's Monte Carlo path for functional
 * nodes, written from the reference sources and (for the arithmetic that lives
 * in UncertaintyQuantification.jl 0.11, which is NOT under /root/reference and
 * cannot run here — no Julia) from that package's published semantics as
 * recorded in SURVEY.md §8c A1/A2:
 *
 *   samples = sample(inputs, n)          one n-vector per input
 *   evaluate!(models, samples)           one n-vector per model
 *   pf = sum(performance(samples) .< 0) / n ; var = (pf - pf^2)/n
 *
 * PARITY STATUS: "parity unpinned" at the bit level — the reference holds no
 * seeded vectors, sample matrices or bit-level goldens for this path
 * (SURVEY.md §8c); its only anchors are statistical (test/networks/ebn/
 * evaluate/evaluate_node.jl:110,137; demo/straub_example.jl:71) and analytic
 * (Phi(-1), Phi(0)). tests/test_oracle.py pins the oracle to every one of them.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: one IEEE operation per
 * source operation, like Julia, which never contracts a*b+c).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/ebncuda.h"

/* ------------------------------------------------------------------------ */
/* limit states, Julia evaluation order                                      */
/* ------------------------------------------------------------------------ */

/* Julia's min(x, y) for Float64 propagates NaN. */
static double jl_min(double a, double b) {
  if (isnan(a)) return a;
  if (isnan(b)) return b;
  return b < a ? b : a;
}

/* demo/vehicle_suspension.jl:17-23 (composite_model)
 *   g1 = 1 .- (pi .* m .* V .* A) ./ (b0 .* K .* g .^ 2) .* [(Ck ./ (m .+ M) .- (C ./ M)) .^ 2
 *        .+ C .^ 2 ./ (m .* M) .+ Ck .* K .^ 2 ./ (m .* M .^ 2)]
 *   g2 = 4000 .* C .* (M .* g) .^ (-1.5) .- 8.6394
 *   g3 = 2 .* sqrt(M .* g .* (K .^ 2 .* Ck ./ (C .* (m .+ M)) .+ C)) .- 1
 *   g4 = Ck .- [g .* (M .+ m)] .^ 0.877
 *   minimum([g1[1], g2, g3, g4[1]])
 * x = (A, b0, V, C, Ck, K); k = (M, m, g, (M*g)^-1.5, (g*(M+m))^0.877).       */
static double g_vehicle(const double *x, const double *k) {
  const double A = x[0], b0 = x[1], V = x[2], C = x[3], Ck = x[4], K = x[5];
  const double M = k[0], m = k[1], g = k[2], pow_a = k[3], pow_b = k[4];
  const double pi = 3.141592653589793; /* Float64(pi) */
  double lhs = pi * m;
  lhs = lhs * V;
  lhs = lhs * A;
  double den = b0 * K;
  den = den * (g * g);
  const double frac = lhs / den;
  const double d1 = Ck / (m + M) - C / M;
  double br = d1 * d1;
  br = br + (C * C) / (m * M);
  br = br + (Ck * (K * K)) / (m * (M * M));
  const double g1 = 1.0 - frac * br;
  const double g2 = (4000.0 * C) * pow_a - 8.6394;
  const double in3 = ((K * K) * Ck) / (C * (m + M)) + C;
  const double g3 = 2.0 * sqrt((M * g) * in3) - 1.0;
  const double g4 = Ck - pow_b;
  return jl_min(jl_min(jl_min(g1, g2), g3), g4);
}

/* demo/straub_example.jl:16-39: plastic_moment_capacities (x5) + frame_model.
 * x = (Ur, V, H, e1..e5), e_i the randn() behind rand(Normal(mu, sd));
 * k = (lambda, zeta, rho).                                                    */
static double g_straub(const double *x, const double *k) {
  const double ur = x[0], v = x[1], h = x[2];
  const double lambda = k[0], zeta = k[1], rho = k[2];
  const double normal_mu = lambda + (rho * zeta) * ur;
  const double normal_std = sqrt((1.0 - rho * rho) * (zeta * zeta));
  double r[5];
  for (int i = 0; i < 5; ++i) r[i] = exp(normal_mu + normal_std * x[3 + i]);
  const double g1 = (((r[0] + r[1]) + r[3]) + r[4]) - 5.0 * h;
  const double g2 = ((r[1] + 2.0 * r[2]) + r[3]) - 5.0 * v;
  const double g3 = ((((r[0] + 2.0 * r[2]) + 2.0 * r[3]) + r[4]) - 5.0 * h) - 5.0 * v;
  return jl_min(jl_min(g1, g2), g3);
}

/* test/inference/variableselimination.jl:207-217: the frame with R4, R5 read as columns.
 * x = (Ur, V, H, R4, R5, e1, e2, e3). */
static double g_straub_r45(const double *x, const double *k) {
  const double ur = x[0], v = x[1], h = x[2];
  const double lambda = k[0], zeta = k[1], rho = k[2];
  const double normal_mu = lambda + (rho * zeta) * ur;
  const double normal_std = sqrt((1.0 - rho * rho) * (zeta * zeta));
  double r[5];
  for (int i = 0; i < 3; ++i) r[i] = exp(normal_mu + normal_std * x[5 + i]);
  r[3] = x[3];
  r[4] = x[4];
  const double g1 = (((r[0] + r[1]) + r[3]) + r[4]) - 5.0 * h;
  const double g2 = ((r[1] + 2.0 * r[2]) + r[3]) - 5.0 * v;
  const double g3 = ((((r[0] + 2.0 * r[2]) + 2.0 * r[3]) + r[4]) - 5.0 * h) - 5.0 * v;
  return jl_min(jl_min(g1, g2), g3);
}

/* demo/straub_example.jl:16-26 (plastic_moment_capacities): x = (Ur, e). */
static double g_straub_capacity(const double *x, const double *k) {
  const double lambda = k[0], zeta = k[1], rho = k[2];
  const double normal_mu = lambda + (rho * zeta) * x[0];
  const double normal_std = sqrt((1.0 - rho * rho) * (zeta * zeta));
  return exp(normal_mu + normal_std * x[1]);
}

/* demo/Hydrogen_new/ebn.jl:98-108 (hyram_performance). */
static double g_hydrogen(const double *x) {
  if (x[0] == 1.0) return x[1] - x[2];
  if (x[0] == 2.0) return x[1] - x[3];
  return 1.0;
}

/* Polynomial program (include/ebncuda.h): the Model/performance closures of
 * test/networks/ebn/evaluate/evaluate_node.jl:96-98 etc. cols has room for
 * EBN_MAX_INPUTS + 1 entries. */
static double g_polynomial(double *cols, int d, const double *p) {
  const int T = (int)p[0];
  int pos = 1;
  double acc = 0.0;
  for (int t = 0; t < T; ++t) {
    const int nterms = (int)p[pos++];
    for (int q = 0; q < nterms; ++q) {
      double term = p[pos++];
      const int nf = (int)p[pos++];
      for (int f = 0; f < nf; ++f) term = term * cols[(int)p[pos++]];
      acc = q == 0 ? term : acc + term;
    }
    if (d + t < EBN_MAX_INPUTS) cols[d + t] = acc;
  }
  return acc;
}

static double g_min_affine(const double *x, int d, const double *k) {
  const int K = (int)k[0];
  double g = 0.0;
  for (int r = 0; r < K; ++r) {
    const double *row = k + 1 + r * (1 + d);
    double acc = row[0];
    for (int j = 0; j < d; ++j) acc = acc + row[1 + j] * x[j];
    g = r == 0 ? acc : jl_min(g, acc);
  }
  return g;
}

static double jl_max(double a, double b) {
  if (isnan(a)) return a;
  if (isnan(b)) return b;
  return b > a ? b : a;
}

/* Expression program (include/ebncuda.h "expression program"): a stack machine, one IEEE
 * operation per instruction in program order — the evaluation order of the Julia expression the
 * wrapper compiled (Model closures + performance, src/nodes/functionalnodes.jl:27-52).
 * x^2 = x*x and x^3 = (x*x)*x as Julia's literal_pow; other powers through pow(). */
static double g_expression(double *cols, int d, const double *p) {
  const int P = (int)p[0], L = (int)p[1];
  const double *k = p + 2, *code = p + 2 + P;
  double st[EBN_EXPR_MAX_DEPTH];
  int sp = 0, col = d;
  for (int i = 0; i < L; ++i) {
    const int op = (int)code[2 * i];
    const double arg = code[2 * i + 1];
    switch (op) {
      case EBN_OP_INPUT: st[sp++] = cols[(int)arg]; break;
      case EBN_OP_PARAM: st[sp++] = k[(int)arg]; break;
      case EBN_OP_CONST: st[sp++] = arg; break;
      case EBN_OP_ADD: sp--; st[sp - 1] = st[sp - 1] + st[sp]; break;
      case EBN_OP_SUB: sp--; st[sp - 1] = st[sp - 1] - st[sp]; break;
      case EBN_OP_MUL: sp--; st[sp - 1] = st[sp - 1] * st[sp]; break;
      case EBN_OP_DIV: sp--; st[sp - 1] = st[sp - 1] / st[sp]; break;
      case EBN_OP_NEG: st[sp - 1] = -st[sp - 1]; break;
      case EBN_OP_ABS: st[sp - 1] = fabs(st[sp - 1]); break;
      case EBN_OP_SQRT: st[sp - 1] = sqrt(st[sp - 1]); break;
      case EBN_OP_EXP: st[sp - 1] = exp(st[sp - 1]); break;
      case EBN_OP_LOG: st[sp - 1] = log(st[sp - 1]); break;
      case EBN_OP_SIN: st[sp - 1] = sin(st[sp - 1]); break;
      case EBN_OP_COS: st[sp - 1] = cos(st[sp - 1]); break;
      case EBN_OP_MIN: sp--; st[sp - 1] = jl_min(st[sp - 1], st[sp]); break;
      case EBN_OP_MAX: sp--; st[sp - 1] = jl_max(st[sp - 1], st[sp]); break;
      case EBN_OP_POW: sp--; st[sp - 1] = pow(st[sp - 1], st[sp]); break;
      case EBN_OP_POWI: {
        const int n = (int)arg;
        const double a = st[sp - 1];
        st[sp - 1] = n == 0 ? 1.0 : n == 1 ? a : n == 2 ? a * a : n == 3 ? (a * a) * a : pow(a, (double)n);
        break;
      }
      case EBN_OP_LT: sp--; st[sp - 1] = st[sp - 1] < st[sp] ? 1.0 : 0.0; break;
      case EBN_OP_SELECT: sp -= 2; st[sp - 1] = st[sp - 1] != 0.0 ? st[sp] : st[sp + 1]; break;
      case EBN_OP_STORE: cols[col++] = st[--sp]; break;
    }
  }
  return st[0];
}

static double limit_state(int ls, const double *consts, double *x, int d) {
  switch (ls) {
    case EBN_LS_EXPRESSION: return g_expression(x, d, consts);
    case EBN_LS_VEHICLE_SUSPENSION: return g_vehicle(x, consts);
    case EBN_LS_STRAUB_FRAME: return g_straub(x, consts);
    case EBN_LS_HYDROGEN_RADIUS: return g_hydrogen(x);
    case EBN_LS_POLYNOMIAL: return g_polynomial(x, d, consts);
    case EBN_LS_MIN_AFFINE: return g_min_affine(x, d, consts);
    case EBN_LS_STRAUB_FRAME_R45: return g_straub_r45(x, consts);
    case EBN_LS_STRAUB_CAPACITY: return g_straub_capacity(x, consts);
  }
  return NAN;
}

/* `sum(performance(df) .< 0)` over a column-major sample matrix. */
int64_t oracle_eval_matrix(int ls, const double *consts, const double *X, int64_t n, int d,
                           int64_t ld, double *g_out) {
  int64_t count = 0;
  double x[EBN_MAX_INPUTS + 1];
  for (int64_t r = 0; r < n; ++r) {
    for (int j = 0; j < d; ++j) x[j] = X[(int64_t)j * ld + r];
    const double g = limit_state(ls, consts, x, d);
    if (g_out) g_out[r] = g;
    count += g < 0.0; /* strict <, NaN is safe */
  }
  return count;
}

/* ------------------------------------------------------------------------ */
/* structural Monte Carlo: materialise columns like UQ.jl's sample()          */
/* ------------------------------------------------------------------------ */

/* xoshiro256++ (Julia's default generator family), seeded by splitmix64. */
typedef struct { uint64_t s[4]; } rng_t;
static uint64_t splitmix64(uint64_t *x) {
  uint64_t z = (*x += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
static void rng_seed(rng_t *r, uint64_t seed) {
  for (int i = 0; i < 4; ++i) r->s[i] = splitmix64(&seed);
}
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t rng_next(rng_t *r) {
  uint64_t *s = r->s;
  const uint64_t result = rotl(s[0] + s[3], 23) + s[0];
  const uint64_t t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
  s[2] ^= t;
  s[3] = rotl(s[3], 45);
  return result;
}
static inline double rng_u01(rng_t *r) { /* (0,1) */
  return ((double)(rng_next(r) >> 11) + 0.5) * 0x1p-53;
}
static double rng_normal(rng_t *r) { /* Marsaglia polar */
  for (;;) {
    const double a = 2.0 * rng_u01(r) - 1.0, b = 2.0 * rng_u01(r) - 1.0;
    const double s = a * a + b * b;
    if (s > 0.0 && s < 1.0) return a * sqrt(-2.0 * log(s) / s);
  }
}
static double rng_gamma(rng_t *r, double shape, double scale) { /* Marsaglia-Tsang */
  const double a = shape < 1.0 ? shape + 1.0 : shape;
  const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  double v;
  for (;;) {
    const double z = rng_normal(r);
    v = 1.0 + c * z;
    if (v <= 0.0) continue;
    v = v * v * v;
    const double u = rng_u01(r);
    if (u < 1.0 - 0.0331 * z * z * z * z) break;
    if (log(u) < 0.5 * z * z + d * (1.0 - v + log(v))) break;
  }
  double x = d * v;
  if (shape < 1.0) x *= pow(rng_u01(r), 1.0 / shape);
  return x * scale;
}

static double phi_cdf(double z) { return 0.5 * erfc(-z * M_SQRT1_2); }
/* Phi^-1 by Halley iterations on erfc, started from A&S 26.2.23. */
static double phi_inv(double p) {
  if (p <= 0.0) return -INFINITY;
  if (p >= 1.0) return INFINITY;
  const int upper = p > 0.5;
  const double q = upper ? 1.0 - p : p;
  const double t = sqrt(-2.0 * log(q));
  double x = -(t - (2.515517 + t * (0.802853 + t * 0.010328)) /
                       (1.0 + t * (1.432788 + t * (0.189269 + t * 0.001308))));
  for (int it = 0; it < 4; ++it) {
    const double e = phi_cdf(x) - q;
    const double pdf = exp(-0.5 * x * x) * 0.3989422804014327;
    const double u = e / pdf;
    x = x - u / (1.0 + 0.5 * x * u);
  }
  return upper ? -x : x;
}

/* One column = rand(dist, n) (or a constant fill for a Parameter). Truncated
 * marginals are drawn by inversion on the truncated probability window. */
static int fill_column(const ebn_input_t *in, rng_t *r, double *col, int64_t n) {
  const int trunc = in->lo > -INFINITY || in->hi < INFINITY;
  const double lo = in->lo, hi = in->hi;
  switch (in->kind) {
    case EBN_IN_PARAM:
      for (int64_t i = 0; i < n; ++i) col[i] = in->p[0];
      return 0;
    case EBN_IN_UNIFORM: {
      double a = in->p[0], b = in->p[1];
      if (lo > a) a = lo;
      if (hi < b) b = hi;
      for (int64_t i = 0; i < n; ++i) col[i] = a + (b - a) * rng_u01(r);
      return 0;
    }
    case EBN_IN_NORMAL:
    case EBN_IN_LOGNORMAL: {
      const double mu = in->p[0], sd = in->p[1];
      const int logn = in->kind == EBN_IN_LOGNORMAL;
      if (!trunc) {
        for (int64_t i = 0; i < n; ++i) {
          const double v = mu + sd * rng_normal(r);
          col[i] = logn ? exp(v) : v;
        }
        return 0;
      }
      const double zl = logn ? (lo > 0.0 ? (log(lo) - mu) / sd : -INFINITY) : (lo - mu) / sd;
      const double zh = logn ? (hi < INFINITY ? (log(hi) - mu) / sd : INFINITY) : (hi - mu) / sd;
      const double Fl = phi_cdf(zl), Fh = phi_cdf(zh);
      for (int64_t i = 0; i < n; ++i) {
        const double v = mu + sd * phi_inv(Fl + (Fh - Fl) * rng_u01(r));
        col[i] = logn ? exp(v) : v;
      }
      return 0;
    }
    case EBN_IN_GUMBEL: {
      const double mu = in->p[0], beta = in->p[1];
      const double Fl = lo > -INFINITY ? exp(-exp(-(lo - mu) / beta)) : 0.0;
      const double Fh = hi < INFINITY ? exp(-exp(-(hi - mu) / beta)) : 1.0;
      for (int64_t i = 0; i < n; ++i)
        col[i] = mu - beta * log(-log(Fl + (Fh - Fl) * rng_u01(r)));
      return 0;
    }
    case EBN_IN_GAMMA:
      if (trunc) return -1;
      for (int64_t i = 0; i < n; ++i) col[i] = rng_gamma(r, in->p[0], in->p[1]);
      return 0;
    case EBN_IN_EXPONENTIAL_AFFINE: {
      const double theta = in->p[0], sign = in->p[1], shift = in->p[2];
      double el = 0.0, eh = INFINITY;
      if (sign > 0) { if (lo > shift) el = lo - shift; if (hi < INFINITY) eh = hi - shift; }
      else { if (hi < shift) el = shift - hi; if (lo > -INFINITY) eh = shift - lo; }
      const double Fl = -expm1(-el / theta), Fh = eh == INFINITY ? 1.0 : -expm1(-eh / theta);
      for (int64_t i = 0; i < n; ++i)
        col[i] = shift + sign * (-theta * log1p(-(Fl + (Fh - Fl) * rng_u01(r))));
      return 0;
    }
  }
  return -1;
}

/*
 * probability_of_failure(models, performance, inputs, MonteCarlo(n)) for ONE
 * scenario, with the reference's structure: materialise the d columns, evaluate
 * the model into a new column, compare, sum. `block` bounds the rows held at
 * once (the reference holds all n; the result does not depend on it).
 * Returns the fail count, or -1 for an unsupported marginal.
 */
int64_t oracle_pf_scenario(int ls, const double *consts, const ebn_input_t *inputs, int d,
                           int64_t n, uint64_t seed, int64_t block) {
  if (block <= 0 || block > n) block = n;
  double *X = (double *)malloc(sizeof(double) * (size_t)block * (size_t)(d + 1));
  if (!X) return -1;
  rng_t r;
  rng_seed(&r, seed);
  int64_t count = 0;
  for (int64_t done = 0; done < n; done += block) {
    const int64_t m = n - done < block ? n - done : block;
    for (int j = 0; j < d; ++j)
      if (fill_column(&inputs[j], &r, X + (size_t)j * block, m)) { free(X); return -1; }
    count += oracle_eval_matrix(ls, consts, X, m, d, block, X + (size_t)d * block);
  }
  free(X);
  return count;
}

/* All scenarios; `threads` pthreads pull scenarios from a shared counter (the
 * reference itself is single-threaded and scenario-serial,
 * evaluate_node.jl:79). Scenario s uses seed + s; block = rows materialised at
 * once (0 = all n, the reference's behaviour). */
typedef struct {
  int ls, S, d;
  const double *consts;
  const ebn_input_t *inputs;
  int64_t n, block;
  uint64_t seed;
  int64_t *fail_count;
  int next, bad;
  pthread_mutex_t mu;
} batch_t;

static void *batch_worker(void *arg) {
  batch_t *b = (batch_t *)arg;
  for (;;) {
    pthread_mutex_lock(&b->mu);
    const int s = b->next++;
    pthread_mutex_unlock(&b->mu);
    if (s >= b->S) return NULL;
    const int64_t c = oracle_pf_scenario(b->ls, b->consts, b->inputs + (size_t)s * b->d, b->d, b->n,
                                         b->seed + (uint64_t)s, b->block);
    if (c < 0) b->bad = 1;
    b->fail_count[s] = c;
  }
}

int oracle_pf_batch(int ls, const double *consts, const ebn_input_t *inputs, int S, int d,
                    int64_t n, uint64_t seed, int threads, int64_t block, int64_t *fail_count) {
  batch_t b = {ls, S, d, consts, inputs, n, block, seed, fail_count, 0, 0, PTHREAD_MUTEX_INITIALIZER};
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  pthread_t tid[256];
  for (int t = 1; t < threads; ++t) pthread_create(&tid[t], NULL, batch_worker, &b);
  batch_worker(&b);
  for (int t = 1; t < threads; ++t) pthread_join(tid[t], NULL);
  return b.bad ? -1 : 0;
}

double oracle_phi_inv(double p) { return phi_inv(p); }
