"""csrc/ebn_fastmath.cuh — the lean fp64 routines every sample goes through — compiled for the HOST and measured
against mpmath and the NumPy twin. No GPU. The header is taken as it is, with two substitutions the host needs:
the MUFU.RSQ64H seed of sqrt_pos (the one inline-PTX statement that is not behind a tuning switch) becomes a
20-bit host seed, and the build uses EBN_BM_OPT = 35 (the shipped 39 without bit 4, whose single-LOP3 bit placement
is PTX; the header states both forms give the same value bit for bit). Everything else — the log / sin-cos / exp
tables, the Phi^-1 and Phi cell tables and their polynomials, the Box-Muller transform from 64 random bits — is
the text the kernels compile.

  * Phi cells (copula path): 4.4e-16 relative down the lower tail, exact clamps at 2^-60 and 1 - 2^-53,
  * Phi^-1 cells (truncated normals): 4e-16 relative to max(1, |z|), rare path included,
  * exp_fast, neg_log_pos, sqrt_pos: a few 1e-16,
  * box_muller_words on 2e5 word pairs: equal to the NumPy twin (oracle/replay.py) to 5e-12, unit moments.
The GPU suite measures the same routines as the device executes them (tests/test_gpu_parity.py)."""
import ctypes as C
import os
import subprocess

import mpmath as mp
import numpy as np
import pytest

from oracle import replay

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "enhancedbayesiannetworks.jl_b200", "csrc")
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
HARNESS = r'''
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>
#include <stdint.h>
#undef __device__
#undef __forceinline__
#undef __noinline__
#undef __shared__
#undef __constant__
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __shared__ static
#define __constant__ static
static inline double __hiloint2double(int hi, int lo) { uint64_t v = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double d; memcpy(&d, &v, 8); return d; }
static inline int __double2hiint(double d) { uint64_t v; memcpy(&v, &d, 8); return (int)(v >> 32); }
static inline int __double2loint(double d) { uint64_t v; memcpy(&v, &d, 8); return (int)(uint32_t)v; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __uint2float_rz(unsigned u) { return (float)(u & 0xFFFFFF00u) ; }
static inline void host_sincosf(float x, float* s, float* c) { *s = sinf(x); *c = cosf(x); }
#define __sincosf host_sincosf
static inline float host_log2f(float x) { return log2f(x); }
#define __log2f host_log2f
static inline double __int2double_rn(int i) { return (double)i; }
static inline float __fsqrt_rn(float x) { return sqrtf(x); }
static inline void __syncthreads() {}
// MUFU.RSQ64H: reciprocal square root from the high word only, about 20 good bits
static inline double host_rsqrt_seed(double t) { float f = (float)t; double y = 1.0 / sqrt((double)f); float g = (float)y; return (double)g; }
static double normcdfinv(double p) {   // library routine of the rare path: Newton on the lower tail, < 1e-15
  if (p > 0.5) return -normcdfinv(1.0 - p);
  double z = 0.0;
  for (int i = 0; i < 200; ++i) {
    const double f = 0.5 * erfc(-z / sqrt(2.0)) - p;
    const double d = exp(-0.5 * z * z) / sqrt(2.0 * M_PI);
    double step = f / d;
    if (step > 0.5) step = 0.5;
    if (step < -0.5) step = -0.5;
    z -= step;
    if (fabs(step) < 1e-17 * fmax(1.0, fabs(z))) break;
  }
  return z;
}
#include "ebn_device.cuh"
#include "ebn_fastmath_host.cuh"
using namespace ebn;
static FastMathSmem g_sm;
extern "C" void init() { fastmath_init(g_sm, 0, 1); icdf_init(0, 1); phi_init(0, 1); }
extern "C" void phi(const double* z, long n, double* out) { for (long i = 0; i < n; ++i) out[i] = phi_clamped(z[i], g_sm); }
extern "C" void icdf(const double* p, long n, double* out) { for (long i = 0; i < n; ++i) out[i] = normcdfinv_fast<true>(p[i]); }
extern "C" void lean_exp(const double* x, long n, double* out) { for (long i = 0; i < n; ++i) out[i] = exp_fast(x[i], g_sm); }
extern "C" void neglog(const double* x, long n, double* out) { for (long i = 0; i < n; ++i) out[i] = neg_log_pos(x[i], g_sm); }
extern "C" void sqrtp(const double* x, long n, double* out) { for (long i = 0; i < n; ++i) out[i] = sqrt_pos(x[i]); }
extern "C" void bm(const unsigned* w0, const unsigned* w1, long n, double* z0, double* z1) { for (long i = 0; i < n; ++i) box_muller_words(w0[i], w1[i], g_sm, z0[i], z1[i]); }
'''


@pytest.fixture(scope="module")
def lean(tmp_path_factory):
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("no CUDA headers on this machine")
    tmp = tmp_path_factory.mktemp("fastmath")
    src = open(os.path.join(CSRC, "ebn_fastmath.cuh")).read()
    seed = '  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(t));'
    assert src.count(seed) == 1
    (tmp / "ebn_fastmath_host.cuh").write_text(src.replace(seed, "  y = host_rsqrt_seed(t);"))
    (tmp / "harness.cpp").write_text(HARNESS)
    subprocess.run(["g++", "-O1", "-std=c++17", "-DEBN_BM_OPT=35", "-I" + CUDA_INC, "-I" + CSRC, "-I" + str(tmp), "-shared", "-fPIC",
                    str(tmp / "harness.cpp"), "-o", str(tmp / "fm.so")], check=True, stderr=subprocess.DEVNULL)
    lib = C.CDLL(str(tmp / "fm.so"))
    lib.init()
    dp = C.POINTER(C.c_double)

    def call(fn, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty_like(x)
        getattr(lib, fn)(x.ctypes.data_as(dp), C.c_long(len(x)), out.ctypes.data_as(dp))
        return out
    call.lib = lib
    return call


def test_lean_math_on_the_host(lean):
    mp.mp.prec = 120
    rng = np.random.default_rng(1)
    # Phi: relative accuracy down the lower tail, absolute on the upper side, clamps outside
    z = np.concatenate([rng.normal(0, 2.5, 3000), np.linspace(-8.7, 8.7, 3001)])
    z = z[np.abs(z) <= 8.7]
    got = lean("phi", z)
    ref = np.array([float(mp.ncdf(mp.mpf(float(v)))) for v in z])
    lo = z < 0
    assert np.max(np.abs(got[lo] - ref[lo]) / ref[lo]) < 1e-15
    assert np.max(np.abs(got[~lo] - ref[~lo])) < 2.5e-16
    assert np.array_equal(lean("phi", np.array([-9.6, -30.0, 9.6, 30.0])), [2.0 ** -60, 2.0 ** -60, 1 - 2.0 ** -53, 1 - 2.0 ** -53])
    # Phi^-1: cells for min(p, 1-p) >= 2^-17, the library routine beyond
    p = np.concatenate([rng.uniform(0, 1, 3000), 10.0 ** rng.uniform(-17, -0.4, 1500), 1 - 10.0 ** rng.uniform(-15, -0.4, 800)])
    p = p[(p > 0) & (p < 1)]
    got = lean("icdf", p)
    ref = np.array([float(mp.sqrt(2) * mp.erfinv(2 * mp.mpf(float(v)) - 1)) for v in p])
    assert np.max(np.abs(got - ref) / np.maximum(1, np.abs(ref))) < 1e-15
    # exp, -ln, sqrt on the sampler's domains
    x = rng.uniform(-40, 40, 3000)
    ref = np.array([float(mp.exp(mp.mpf(float(v)))) for v in x])
    assert np.max(np.abs(lean("lean_exp", x) - ref) / ref) < 5e-16
    u = np.concatenate([rng.uniform(0, 1, 2000), 10.0 ** rng.uniform(-12, 0, 1000)])
    u = u[(u > 0) & (u < 1)]
    ref = np.array([float(-mp.log(mp.mpf(float(v)))) for v in u])
    assert np.max(np.abs(lean("neglog", u) - ref) / np.maximum(ref, 1e-3)) < 5e-13
    t = 10.0 ** rng.uniform(-15, 1.6, 20000)
    assert np.max(np.abs(lean("sqrtp", t) - np.sqrt(t)) / np.sqrt(t)) < 4e-16
    # Box-Muller from 64 random bits (stream layout v4) against the twin
    n = 200_000
    w0 = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    w1 = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    z0, z1 = np.empty(n), np.empty(n)
    dp, up = C.POINTER(C.c_double), C.POINTER(C.c_uint)
    lean.lib.bm(w0.ctypes.data_as(up), w1.ctypes.data_as(up), C.c_long(n), z0.ctypes.data_as(dp), z1.ctypes.data_as(dp))
    r0, r1 = replay._box_muller(w0, w1)
    assert np.max(np.abs(z0 - r0)) < 5e-12 and np.max(np.abs(z1 - r1)) < 5e-12
    assert abs(z0.mean()) < 0.01 and abs(z0.std() - 1) < 0.01 and abs(z1.std() - 1) < 0.01 and abs(np.corrcoef(z0, z1)[0, 1]) < 0.01
