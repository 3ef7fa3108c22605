// Lean fp64 primitives for the sampler's hot path.
//
// The CUDA math library's log/sincospi/sqrt are general (denormals, infinities,
// huge arguments) and ptxas materialises their 64-bit coefficients with UMOV
// pairs: in the first profile of the vehicle kernel (profiles/r01_a_*) 66 % of
// the issued instructions were NOT fp64, which capped the fp64 pipe at 45 %.
// The sampler only ever needs
//     -ln(u)        u in (0,1), normal
//     sqrt(t)       t in [2^-53, 37]
//     sin, cos      of a first-quadrant angle
// so each is written for exactly that domain: no special cases, no branches,
// coefficients in __constant__ memory (consumed as c[bank][off] operands).
// Accuracy: a few ulp (measured against NumPy/mpmath in tests/test_gpu_parity.py).
#pragma once
#include "ebn_rtc_compat.h"

namespace ebn {

#include "ebn_fastmath_tables.inc"

struct FastMathConst {
  double ln1p[5];   // -1/2, 1/3, -1/4, 1/5, -1/6  (negated for -ln)
  double ln2;
  double expc[4];   // 1/2, 1/6, 1/24, 1/120
  double inv_ln2_64, ln2_64_hi, ln2_64_lo;
  double sind[EBN_SC_TAB_BITS < 10 ? 3 : 2];   // sin(delta) = t*(d0 + d1 t^2 [+ d2 t^4]), delta = (pi/2) t inside a table cell
  double cosd[2];   // cos(delta) - 1 = t^2*(e0 + e1 t^2)
  uint32_t mant20, one_hi;   // 0x000fffff, 0x3ff00000: LOP3 operands the compiler must not fold (see hi_of_one_two)
  double td_off;             // -(2^22 + half cell - 2^-21): addend of the in-cell offset (bm_trig)
  double c375;               // 3/8 of the square-root correction
};
__constant__ FastMathConst kFm = {
    {0.5, -0.33333333333333331483, 0.25, -0.2000000000000000111, 0.16666666666666665741},
    0.69314718055994528623,
    {0.5, 0.16666666666666665741, 0.041666666666666664354, 0.0083333333333333332177},
    92.332482616893656877,          // 64/ln2
    0x1.62e42fef00000p-7,           // ln2/64, high 33 bits (n * hi is exact for |n| < 2^20)
    0x1.473de6af278edp-40,          // ln2/64 - high part
    {EBN_SIND_COEFS},
    {EBN_COSD_COEFS},
    0x000FFFFFu, 0x3FF00000u,
    -(4194304.0 + 1.0 / (double)(1 << (EBN_SC_TAB_BITS + 1)) - 0x1p-21),
    0.375,
};

// Shared-memory copy of the log table (per-lane random index: shared memory,
// not the constant cache, which serialises divergent addresses).
struct FastMathSmem {
  double2 logtab[1 << EBN_LOG_TAB_BITS];
  double2 sctab[4 << EBN_SC_TAB_BITS];   // sqrt(2)*(sin, cos) at the cell centres of the full circle
  double exptab[64];
};

__device__ __forceinline__ void fastmath_init(FastMathSmem& sm, int tid, int nthreads) {
  for (int i = tid; i < (1 << EBN_LOG_TAB_BITS); i += nthreads) sm.logtab[i] = kLogTab[i];
  for (int i = tid; i < (4 << EBN_SC_TAB_BITS); i += nthreads) sm.sctab[i] = kSinCosTab[i];
  for (int i = tid; i < 64; i += nthreads) sm.exptab[i] = kExpTab[i];
}

// Tuning switches of the Box-Muller transform (bit mask; profiles/r02_*):
//   1  -ln u through one Horner chain (two operations fewer than r^2 p + (base - r))
//   2  x = mu + (sigma R) * trig: the scale is folded into the radius (one multiplication fewer per pair)
//   4  bit placement with single LOP3s: (w & 0xfffff) | 0x3ff00000 with both masks as register operands,
//      and the low mantissa word as one bit-select
//   8  exponent -> double through I2F (XU pipe, idle) instead of the 2^52 magic (MOV + DADD)
//  32  polynomial top coefficients as immediates, the remaining literals from the constant bank
//  16  r = u * (rc 2^e) - 1 with the exponent added to the table entry's high word (one IMAD) instead of
//      rebuilding the mantissa m = u 2^e as a new register pair (LOP3 + MOV); the same value bit for bit
#ifndef EBN_BM_OPT
#define EBN_BM_OPT 39   // measured best on B200 with stream layout v4 (profiles/r02_tune_w1.txt): 8 (I2F on the XU pipe)
                        // helped layout v3 by 1 % and costs 6 % now, 16 costs 3 %
#endif

// High word of the double in [1,2) whose top 20 mantissa bits are the low 20 bits of w:
// (w & 0x000fffff) | 0x3ff00000. Written as ONE three-input LOP3 whose two masks are operands
// (constant bank / registers): with two literals ptxas needs two instructions.
__device__ __forceinline__ uint32_t hi_of_one_two(uint32_t w) {
  if (EBN_BM_OPT & 4) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(w), "r"(kFm.mant20), "r"(kFm.one_hi));
    return d;
  }
  return (w & 0x000FFFFFu) | 0x3FF00000u;
}
// (a & mask) | (b & ~mask) in one LOP3.
template <uint32_t MASK>
__device__ __forceinline__ uint32_t bit_select(uint32_t a, uint32_t b) {
  if (EBN_BM_OPT & 4) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "n"(MASK));
    return d;
  }
  return (a & MASK) | (b & ~MASK);
}

// Index of the log-table cell from the high word of a double, and the polynomial
// p(r) with log1p(r) = r - r^2 p(r): 1/2 - r/3 + r^2/4 [- r^3/5 + r^4/6 when the cells are wide].
__device__ __forceinline__ int log_cell(int hi) { return (hi >> (20 - EBN_LOG_TAB_BITS)) & ((1 << EBN_LOG_TAB_BITS) - 1); }
__device__ __forceinline__ double log1p_poly(double r) {
  if (EBN_LOG_TAB_BITS >= 10) {      // |r| <= 2^-11: the r^5/5 term is below 6e-18
    // 1/4 and 1/2 as literals: their low 32 bits are zero, so they are instruction immediates and
    // only -1/3 comes from the constant bank (two bank operands in one DFMA would cost an LDC)
    return fma(r, fma(r, 0.25, kFm.ln1p[1]), 0.5);
  } else {
    double p = fma(r, kFm.ln1p[4], kFm.ln1p[3]);
    p = fma(r, p, kFm.ln1p[2]);
    p = fma(r, p, kFm.ln1p[1]);
    return fma(r, p, kFm.ln1p[0]);
  }
}

// -ln(u) for a normal double u in (0,1).
//   u = 2^e * m, m in [1,2);  cell i of 128 (top 7 mantissa bits) gives rc ~ 1/m (24 bits)
//   and ln(rc):   -ln(u) = -e*ln2 + ln(rc) - log1p(r),  r = m*rc - 1,  |r| <= 2^-8.
__device__ __forceinline__ double neg_log01(double u, const FastMathSmem& sm) {
  const int hi = __double2hiint(u);
  const int lo = __double2loint(u);
  // m in [1,2), e = exponent: four integer instructions fewer than re-centring m on
  // [0.75,1.5). The price is cancellation as u -> 1 (-ln u = ln 2 - ln m): an ABSOLUTE error of ~1e-16 in
  // -ln u, i.e. ~1e-15 absolute in the radius where the radius itself is < 0.05. x = mu +
  // sigma*z only ever needs z to absolute accuracy, so nothing is lost.
  const int ne = 1023 - (hi >> 20);                                   // -e in [1, 53]
  const int mhi = (hi & 0x000FFFFF) | 0x3FF00000;                     // m in [1,2)
  const int idx = log_cell(hi);                                       // top mantissa bits
  const double m = __hiloint2double(mhi, lo);
  const double2 t = sm.logtab[idx];
  const double r = fma(m, t.x, -1.0);
  // int -> double without a conversion instruction: 2^52+2^51 magic (ne >= 0)
  const double ned = __hiloint2double(0x43380000, ne) - 6755399441055744.0;
  // log1p(r) = r - r^2*(1/2 - r*(1/3 - r*(1/4 - r*(1/5 - r/6))))
  const double p = log1p_poly(r);      // 1/2 - r/3 + r^2/4 ...
  const double r2 = r * r;
  const double base = fma(ned, kFm.ln2, t.y);   // -e*ln2 + ln(rc)
  return fma(r2, p, base - r);                   // ... - r + r^2*(...)
}

// -ln(t) for any positive normal double t (the exponent may have either sign).
__device__ __forceinline__ double neg_log_pos(double t, const FastMathSmem& sm) {
  const int hi = __double2hiint(t);
  const int lo = __double2loint(t);
  const int ne = 1023 - (hi >> 20);
  const int mhi = (hi & 0x000FFFFF) | 0x3FF00000;
  const int idx = log_cell(hi);
  const double m = __hiloint2double(mhi, lo);
  const double2 tb = sm.logtab[idx];
  const double r = fma(m, tb.x, -1.0);
  // signed int -> double: bias by 2^31 inside the 2^52+2^51 magic
  const double ned = __hiloint2double(0x43380000, ne ^ (int)0x80000000) - 6755401588539392.0;
  const double p = log1p_poly(r);
  const double r2 = r * r;
  const double base = fma(ned, kFm.ln2, tb.y);
  return fma(r2, p, base - r);
}

// exp(x) for |x| < 700 (no overflow/underflow handling: the callers exponentiate
// log-normal scores). x = (64 k + j) ln2/64 + r, |r| <= ln2/128:
//   exp(x) = 2^k * 2^(j/64) * (1 + r + r^2/2 + ... + r^5/120),  ~1 ulp.
__device__ __forceinline__ double exp_fast(double x, const FastMathSmem& sm) {
  const double t = fma(x, kFm.inv_ln2_64, 6755399441055744.0);   // round to nearest integer n
  const int n = __double2loint(t);
  const double nd = t - 6755399441055744.0;
  double r = fma(-nd, kFm.ln2_64_hi, x);
  r = fma(-nd, kFm.ln2_64_lo, r);
  const double s = sm.exptab[n & 63];
  double p = fma(r, kFm.expc[3], kFm.expc[2]);
  p = fma(r, p, kFm.expc[1]);
  p = fma(r, p, kFm.expc[0]);
  p = fma(r * r, p, r);                 // e^r - 1
  const double y = fma(s, p, s);
  return __hiloint2double(__double2hiint(y) + ((n >> 6) << 20), __double2loint(y));
}

// exp / log for callers that cannot promise the lean versions' domain (generated functors in fast
// mode): the lean path inside it, the CUDA library outside (rare, warp-divergent only there).
__device__ __forceinline__ double exp_guarded(double x, const FastMathSmem& sm) {
  return fabs(x) < 700.0 ? exp_fast(x, sm) : exp(x);
}
__device__ __forceinline__ double log_guarded(double x, const FastMathSmem& sm) {
  return (x > 0x1p-1000 && x < 0x1p1000) ? -neg_log_pos(x, sm) : log(x);
}

// Phi^-1(p) for p in (0, 1) without logarithm or square root: q = min(p, 1 - p) = 2^-e m picks one of
// 16 cells per octave from its exponent and top four mantissa bits (one shift of the high word), and the
// cell carries a degree-7 polynomial in (m - centre): 3e-16 relative to max(1, |z|)
// (tools/gen_fastmath_tables.py). 11 fp64 operations against ~60 plus two branches of CUDA's
// normcdfinv. The table (16 octaves, 16 KB) lives in shared memory: each lane reads its own row, which
// the banks serve in parallel (behind L1 the same loads cost ~25 cycles each, measured). It is a
// function-scope __shared__ array, so only kernels whose sampling plan can draw a truncated normal
// allocate it; they fill it with icdf_init(). The lookup is branch-free (the row is clamped) so that the
// loads schedule early; `rare` reports q < 2^-17 (3e-5 of the unit interval) or q = 0, for which the
// caller substitutes the library routine afterwards.
static __device__ __noinline__ double normcdfinv_rare(double p) { return normcdfinv(p); }

// A row is 64 bytes; with that stride every lane's first 16-byte load would fall on one of two bank
// groups. Rows are spaced 80 bytes apart instead, which spreads random rows over all eight groups.
constexpr int kIcdfStride = 5;   // double2 per row in shared memory (4 used)
__device__ __forceinline__ double2* icdf_smem() {
  __shared__ double2 s_icdf[EBN_ICDF_ROWS * kIcdfStride];
  return s_icdf;
}
__device__ __forceinline__ void icdf_init(int tid, int nthreads) {
  double2* t = icdf_smem();
  const double2* g = reinterpret_cast<const double2*>(kIcdfTab);
  for (int i = tid; i < EBN_ICDF_ROWS * 4; i += nthreads) t[(i >> 2) * kIcdfStride + (i & 3)] = g[i];
}

// Run-time dispatched plans: where the cells are read from is decided per launch. A job whose plan holds a
// truncated normal gets a copy of the table in DYNAMIC shared memory (the host adds kIcdfDynBytes to the
// launch, McParams::icdf_dyn); every other job reads the global table behind L1 and spends no shared
// memory on it. Behind L1 the 32 lanes' rows are 32 cache lines per load (one tag stage each): W1 with two
// truncated normals ran at 2.6e10 samples/s that way against 1.0e11 from shared memory.
struct IcdfDyn {
  const double2* base;
  int stride;   // double2 per row
};
static_assert(kIcdfDynBytes == EBN_ICDF_ROWS * kIcdfStride * 16, "ebn_device.cuh: kIcdfDynBytes");
__device__ __forceinline__ IcdfDyn& icdf_dyn() {
  __shared__ IcdfDyn s_icdf_dyn;
  return s_icdf_dyn;
}
// Called by every thread of the block before a __syncthreads(); `smem` = 16-byte aligned room for the copy
// or nullptr.
__device__ __forceinline__ void icdf_dyn_init(int tid, int nthreads, void* smem) {
  const double2* g = reinterpret_cast<const double2*>(kIcdfTab);
  if (smem) {
    double2* t = reinterpret_cast<double2*>(smem);
    for (int i = tid; i < EBN_ICDF_ROWS * 4; i += nthreads) t[(i >> 2) * kIcdfStride + (i & 3)] = g[i];
    if (tid == 0) icdf_dyn() = IcdfDyn{t, kIcdfStride};
  } else if (tid == 0) {
    icdf_dyn() = IcdfDyn{g, 4};
  }
}

// SMEM: read the cell from the static shared-memory copy (compile-time sampling plans), else through the
// block's IcdfDyn handle (run-time dispatched plans). The arithmetic is the same, so all give the same bits.
template <bool SMEM>
__device__ __forceinline__ double normcdfinv_cells(double p, bool& rare) {
  const bool lower = p < 0.5;
  const double q = lower ? p : 1.0 - p;
  const int hi = __double2hiint(q);
  unsigned row = (unsigned)((1022 * 16) - (hi >> 16));   // rows run with decreasing q; q = 1/2 is row 0
  rare = row >= (unsigned)EBN_ICDF_ROWS;
  row = rare ? EBN_ICDF_ROWS - 1 : row;
  const int mhi = (int)hi_of_one_two((uint32_t)hi);
  const double m = __hiloint2double(mhi, __double2loint(q));
  const double c = __hiloint2double((mhi & (int)0xFFFF0000) | 0x00008000, 0);   // centre of the cell
  const double dm = m - c;
  double2 c01, c23, c45, c67;
  if constexpr (SMEM) {
    const double2* co = icdf_smem() + kIcdfStride * row;
    c01 = co[0], c23 = co[1], c45 = co[2], c67 = co[3];
  } else {
    const IcdfDyn t = icdf_dyn();
    const double2* co = t.base + t.stride * (int)row;
    c01 = co[0], c23 = co[1], c45 = co[2], c67 = co[3];
  }
  double z = fma(dm, c67.y, c67.x);
  z = fma(dm, z, c45.y);
  z = fma(dm, z, c45.x);
  z = fma(dm, z, c23.y);
  z = fma(dm, z, c23.x);
  z = fma(dm, z, c01.y);
  z = fma(dm, z, c01.x);
  return lower ? z : -z;
}

template <bool SMEM>
__device__ __forceinline__ double normcdfinv_fast(double p) {
  bool rare;
  const double z = normcdfinv_cells<SMEM>(p, rare);
  return rare ? normcdfinv_rare(p) : z;
}

// Both samples of a pair: the two lookups interleave, one (never taken) branch for the pair.
template <bool SMEM>
__device__ __forceinline__ void normcdfinv_fast2(double p0, double p1, double& z0, double& z1) {
  bool r0, r1;
  z0 = normcdfinv_cells<SMEM>(p0, r0);
  z1 = normcdfinv_cells<SMEM>(p1, r1);
  if (r0 | r1) {
    if (r0) z0 = normcdfinv_rare(p0);
    if (r1) z1 = normcdfinv_rare(p1);
  }
}

// Phi(z), clamped to [2^-60, 1 - 2^-53] so that quantile functions of it stay finite — the step between the
// correlated normal scores of a Gaussian copula and a non-normal marginal. a = min(|z|, 9.5) picks the cell
// k = round(16 a) centred on c = k/16; the cell's degree-7 polynomial in dm = a - c gives
// Phi(-(c + dm)) exp(c dm), and exp(-c dm) with |c dm| <= 0.3 comes from exp_fast: 4.4e-16 relative to
// Phi(-a) over the whole range, tail included (tools/gen_fastmath_tables.py). 25 fp64 operations, no branch,
// against ~55 plus two branches and ~60 constant moves of CUDA's normcdf (which W3's kernel calls four
// times per sample pair). The 154-row table (12 KB at the bank-spreading 80-byte stride, see kIcdfStride)
// is a function-scope __shared__ array: only copula kernels allocate it; they fill it with phi_init().
__device__ __forceinline__ double2* phi_smem() {
  __shared__ double2 s_phi[EBN_PHI_ROWS * kIcdfStride];
  return s_phi;
}
__device__ __forceinline__ void phi_init(int tid, int nthreads) {
  double2* t = phi_smem();
  const double2* g = reinterpret_cast<const double2*>(kPhiTab);
  for (int i = tid; i < EBN_PHI_ROWS * 4; i += nthreads) t[(i >> 2) * kIcdfStride + (i & 3)] = g[i];
}
__device__ __forceinline__ double phi_clamped(double z, const FastMathSmem& sm) {
  const double a = fmin(fabs(z), EBN_PHI_AMAX);
  const double t = fma(a, 16.0, 6755399441055744.0);   // k = round(16 a) in the low word
  const int k = __double2loint(t);
  const double kf = t - 6755399441055744.0;
  const double dm = fma(kf, -0.0625, a);               // exact
  const double2* co = phi_smem() + kIcdfStride * k;
  const double2 c01 = co[0], c23 = co[1], c45 = co[2], c67 = co[3];
  double p = fma(dm, c67.y, c67.x);
  p = fma(dm, p, c45.y);
  p = fma(dm, p, c45.x);
  p = fma(dm, p, c23.y);
  p = fma(dm, p, c23.x);
  p = fma(dm, p, c01.y);
  p = fma(dm, p, c01.x);
  const double q = p * exp_fast((kf * dm) * -0.0625, sm);   // Phi(-|z|)
  const double r = z < 0.0 ? q : 1.0 - q;
  return fmin(fmax(r, 0x1p-60), 0x1.fffffffffffffp-1);
}

// sqrt(t) for t in [2^-60, 2^60]: MUFU.RSQ64H seed y ~ t^-1/2 (relative error e0 ~ 2^-20),
// then one third-order correction: with a = t*y and e = 1 - a*y,
//   sqrt(t) = a / sqrt(1 - e) = a * (1 + e/2 + 3 e^2/8 + O(e^3)),   truncation 5 e^3/16 < 2^-58.
// Five fp64 operations (two coupled Newton steps need seven).
__device__ __forceinline__ double sqrt_pos(double t) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(t));
  const double a = t * y;
  const double e = fma(-a, y, 1.0);
  // 32: literals that are not instruction immediates come from the constant bank (one c[] operand per
  // instruction is free; a literal costs two MOVs per use inside the loop)
  const double c = (EBN_BM_OPT & 32) ? fma(e, kFm.c375, 0.5) : fma(e, 0.375, 0.5);
  return fma(a * e, c, a);
}

// Box-Muller for the sample pair from two 32-bit words (64 bits = 42 + 22):
//   radius  u = (M + 1/2) 2^-42 in (0,1), M = 42 bits: all of w0 (low 20 bits most significant)
//           followed by bits 0..9 of w1; reaches 7.6 sigma, P(lost tail) = 2^-43 per draw
//   angle   Phi = (pi/2)(A + 1/2) 2^-20 in (0, 2 pi), A = bits 10..31 of w1 (22 bits: quadrant and 20-bit offset)
// z0 = R*sin(Phi), z1 = R*cos(Phi) with R = sqrt(-2 ln u), Phi uniform on the circle: a uniformly random
// direction, so (z0, z1) are independent standard normals. (Layout v3 took a first-quadrant angle and two
// sign bits: the same distribution, three integer instructions more per pair.)
//
// Split in two so that the count loop can software-pipeline it: the FRONT end (integer bit
// placement, the table load and the two fp64 operations that consume it) runs one pair ahead,
// next to the Philox rounds, and hides the shared-memory latency; the BACK end is pure fp64.
struct BmFront {
  double r;       // m*rc - 1 of the -ln(u) reduction
  double base;    // -e*ln2 + ln(rc)
  uint32_t w1;    // angle and sign bits
};

__device__ __forceinline__ BmFront box_muller_front(uint32_t w0, uint32_t w1, const FastMathSmem& sm) {
  // Bit placement (the integer ALU is the scarce pipe next to the fp64 pipe): the low 20 bits of w0
  // become the high mantissa bits, its top 12 bits stay in place in the low word, bits 0..9 of w1
  // land right below them; the low 10 mantissa bits are 0.
  const double u = __hiloint2double((int)hi_of_one_two(w0), (int)bit_select<0xFFF00000u>(w0, w1 << 10)) -
                   (1.0 - 0x1p-43);
  const int hi = __double2hiint(u);
  const int ne = 1023 - (hi >> 20);
  const double2 t = sm.logtab[log_cell(hi)];
  BmFront f;
  if (EBN_BM_OPT & 16) {
    f.r = fma(u, __hiloint2double(__double2hiint(t.x) + (ne << 20), __double2loint(t.x)), -1.0);
  } else {
    const double m = __hiloint2double((int)hi_of_one_two((uint32_t)hi), __double2loint(u));
    f.r = fma(m, t.x, -1.0);
  }
  const double ned = (EBN_BM_OPT & 8) ? __int2double_rn(ne) : __hiloint2double(0x43380000, ne) - 6755399441055744.0;
  f.base = fma(ned, kFm.ln2, t.y);
  f.w1 = w1;
  return f;
}

// -ln u from the front end's reduction.
__device__ __forceinline__ double bm_neg_log(const BmFront& f) {
  const double r = f.r;
  const double p = log1p_poly(r);
  if (EBN_BM_OPT & 1) return fma(r, fma(r, p, -1.0), f.base);   // base - r + r^2 p as one Horner chain
  return fma(r * r, p, f.base - r);
}

// sqrt(2)*(sin, cos) of the pair's angle.
// Angle Phi = (pi/2)(A + 1/2) 2^-20 over the FULL circle, A = bits 10..31 of w1 (22 bits): the top B + 2 bits
// (B = EBN_SC_TAB_BITS) pick one of 4*2^B cells whose centre carries sqrt(2)*(sin, cos) in shared memory —
// a plain shift of w1, no mask, and no sign handling afterwards; the low 20-B bits give the offset inside the
// cell, delta = (pi/2) t with |t| <= 2^-(B+1), short enough for two 2-3 term polynomials (10-11 fp64
// operations instead of the 20 of a full-quadrant pair). The offset bits stay where they are in the low
// mantissa word: d = 1 + r 2^-42, t = (r + 1/2) 2^-20 - 2^-(B+1) = d 2^22 - (2^22 + 2^-(B+1) - 2^-21), exact.
__device__ __forceinline__ void bm_trig(uint32_t w1, const FastMathSmem& sm, double& cps, double& cms) {
  constexpr int B = EBN_SC_TAB_BITS;
  constexpr uint32_t kOffsetMask = ((1u << (20 - B)) - 1u) << 10;
  constexpr double kHalfCell = 1.0 / (double)(1 << (B + 1));
  const double2 sc = sm.sctab[w1 >> (30 - B)];
  double td, sd, cm;
  if (EBN_BM_OPT & 32) {
    // the top coefficients are immediates (low word zero, tools/gen_fastmath_tables.py), every other
    // constant is a constant-bank operand: no coefficient lives in a register
    td = fma(__hiloint2double(0x3FF00000, (int)(w1 & kOffsetMask)), 4194304.0, kFm.td_off);
    const double w = td * td;
    if (B < 10) sd = td * fma(w, fma(w, EBN_SIND_TOP, kFm.sind[1]), kFm.sind[0]);
    else sd = td * fma(w, EBN_SIND_TOP, kFm.sind[0]);
    cm = w * fma(w, EBN_COSD_TOP, kFm.cosd[0]);
  } else {
    td = fma(__hiloint2double(0x3FF00000, (int)(w1 & kOffsetMask)), 4194304.0,
             -(4194304.0 + kHalfCell - 0x1p-21));
    const double w = td * td;
    if (B < 10) sd = td * fma(w, fma(w, kFm.sind[B < 10 ? 2 : 1], kFm.sind[1]), kFm.sind[0]);  // sin(delta)
    else sd = td * fma(w, kFm.sind[1], kFm.sind[0]);
    cm = w * fma(w, kFm.cosd[1], kFm.cosd[0]);                        // cos(delta) - 1
  }
  cps = fma(sc.y, sd, fma(sc.x, cm, sc.x));    // sqrt(2) sin(phi)
  cms = fma(-sc.x, sd, fma(sc.y, cm, sc.y));   // sqrt(2) cos(phi)
}

__device__ __forceinline__ void box_muller_back(const BmFront& f, const FastMathSmem& sm, double& z0, double& z1) {
  double cps, cms;
  bm_trig(f.w1, sm, cps, cms);
  const double rr = sqrt_pos(bm_neg_log(f));           // sqrt(-ln u) = R / sqrt(2)
  z0 = rr * cps;        // R sin(Phi)
  z1 = rr * cms;        // R cos(Phi)
}

// x = mu + sigma z for both samples of the pair.
__device__ __forceinline__ void box_muller_back_affine(const BmFront& f, const FastMathSmem& sm, double mu,
                                                       double sigma, double& x0, double& x1) {
  if (EBN_BM_OPT & 2) {
    double cps, cms;
    bm_trig(f.w1, sm, cps, cms);
    const double rs = sqrt_pos(bm_neg_log(f)) * sigma;   // sigma R / sqrt(2)
    x0 = fma(rs, cps, mu);
    x1 = fma(rs, cms, mu);
  } else {
    double z0, z1;
    box_muller_back(f, sm, z0, z1);
    x0 = fma(z0, sigma, mu);
    x1 = fma(z1, sigma, mu);
  }
}

// The same transform in fp32 ("FP32 where tolerated"): the radius uniform keeps the 32 bits of w0
// (the 10 extra bits of w1 are below fp32 resolution), angle and signs are the same bit fields, so
// z agrees with the fp64 transform to ~1e-6 (MUFU.LG2/SIN/COS approximations, fp32 rounding) and
// the tail reaches 6.7 sigma. 5 MUFU + 2 conversions instead of 31 fp64 operations per pair:
// 221 against 355 cycles per three pairs (profiles/r01_microbench_bm32.txt). Opt-in per job.
__device__ __forceinline__ void box_muller_words_f32(uint32_t w0, uint32_t w1, double& z0, double& z1) {
  const float u = fmaf(__uint2float_rz(w0 << 12 | w0 >> 20), 0x1p-32f, 0x1p-33f);   // rotl(w0, 12): the bit order of the fp64 path
  const float r = __fsqrt_rn(-1.3862943611198906f * __log2f(u));                     // sqrt(-2 ln u)
  // fraction of the circle f = (A + 1/2) 2^-22, A = bits 10..31 of w1, exact in fp32; the hardware sine is
  // most accurate on [-pi, pi]: evaluate at 2 pi (f - 1/2) = Phi - pi, where sin and cos are minus those of Phi
  const float f = __uint_as_float(0x3F800000u | ((w1 >> 9) | 1u)) - 1.5f;
  float s, c;
  __sincosf(6.283185307179586f * f, &s, &c);
  z0 = (double)(-(r * s));
  z1 = (double)(-(r * c));
}

__device__ __forceinline__ void box_muller_words(uint32_t w0, uint32_t w1, const FastMathSmem& sm,
                                                 double& z0, double& z1) {
  box_muller_back(box_muller_front(w0, w1, sm), sm, z0, z1);
}

}  // namespace ebn
