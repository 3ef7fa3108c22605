"""NumPy twin of the device sampler (test infrastructure, see oracle/__init__.py).

Restates, independently of the CUDA sources, the stream layout of DESIGN.md
("Random streams, layout v4") so that the uniforms the GPU consumed can be
regenerated on the CPU:

    key     = (seed & 0xffffffff, seed >> 32)
    pair    = global sample index >> 1
    W[4c .. 4c+3] = Philox4x32-10(counter = (pair lo, pair hi, stream id, c), key)

    inputs consume words of W in input order:
      Parameter 0 | Uniform 2 | untruncated Normal/LogNormal 2 | Gamma 0 | other marginals 4
    U52(lo, hi) = ((((hi & 0xfffff) << 32) | lo) + 0.5) * 2**-52    U32(w) = (rotl(w, 12) + 0.5) * 2**-32

    Uniform            : sample 2q uses U32(W[o]), sample 2q+1 uses U32(W[o+1])
    inverse-CDF kinds  : sample 2q uses U52(W[o], W[o+1]), sample 2q+1 uses U52(W[o+2], W[o+3])
    (Log)Normal        : Box-Muller across the pair from 64 bits (w0, w1) = (W[o], W[o+1]):
                         R = sqrt(-2 ln u), u = (M + 1/2) 2^-42, M = rotl(w0, 12) * 2^10 + (w1 & 0x3ff);
                         Phi = (pi/2) (A + 1/2) 2^-20 in (0, 2 pi), A = w1 >> 10 (22 bits);
                         z(2q) = R sin(Phi), z(2q+1) = R cos(Phi)
    Gamma (Marsaglia-Tsang): the first attempt of the pair's two samples is ONE Philox call, counter
                         word 3 = 0x80000000 | j << 16: Box-Muller of (w0, w1) gives the normal of
                         sample 2q and of sample 2q+1, acceptance uniforms U32(w2), U32(w3); retry t >= 1
                         of sample e reads call 0x80000000 | j << 16 | e << 15 | t (normal = first
                         Box-Muller output, uniform U32(w3)); the shape < 1 boost uniform is
                         U52(w0, w1) of call e << 15 | 0x7fff
    Gaussian copula    : every random input takes 2 words (its Box-Muller pair), z = L eps

Philox4x32-10 itself is pinned by the Random123 known-answer vectors in
tests/test_oracle.py.
"""
from __future__ import annotations

import numpy as np
from scipy import special

from .cpu import (IN_EXP_AFFINE, IN_GAMMA, IN_GUMBEL, IN_LOGNORMAL, IN_NORMAL, IN_PARAM,
                  IN_TABULATED, IN_UNIFORM, LS_HYDROGEN, LS_MIN_AFFINE, LS_POLYNOMIAL, LS_STRAUB,
                  LS_VEHICLE)

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Vectorised Philox4x32-10; counters are uint32 arrays (broadcastable), key two ints."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 &= 0xFFFFFFFF
    k1 &= 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        c1 = p1 & _MASK
        c3 = p0 & _MASK
        c0, c2 = n0, n2
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _u52(lo, hi):
    x = ((hi & np.uint64(0xFFFFF)) << np.uint64(32)) | lo
    return (x.astype(np.float64) + 0.5) * 2.0**-52


def _u32(w):
    rot = ((w << np.uint64(12)) & _MASK) | (w >> np.uint64(20))  # rotate left by 12
    return (rot.astype(np.float64) + 0.5) * 2.0**-32


def _call(pairs, stream: int, c: int, seed: int):
    pairs = np.asarray(pairs, dtype=np.uint64)
    return philox4x32_10(pairs & _MASK, pairs >> np.uint64(32), np.uint64(stream), np.uint64(c),
                         seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


class _Words:
    """The word stream W of a set of pairs, calls made on demand."""

    def __init__(self, pairs, stream, seed):
        self.pairs, self.stream, self.seed, self.calls = pairs, stream, seed, {}

    def __getitem__(self, k: int):
        c = k >> 2
        if c not in self.calls:
            self.calls[c] = _call(self.pairs, self.stream, c, self.seed)
        return self.calls[c][k & 3]


def _box_muller(w0, w1):
    rot = ((w0 << np.uint64(12)) & _MASK) | (w0 >> np.uint64(20))  # rotate left by 12
    M = (rot << np.uint64(10)) | (w1 & np.uint64(0x3FF))             # 42 bits
    u = (M.astype(np.float64) + 0.5) * 2.0**-42
    A = w1 >> np.uint64(10)                                          # 22 bits: the full circle
    R = np.sqrt(-2.0 * np.log(u))
    phi = (np.pi / 2) * ((A.astype(np.float64) + 0.5) * 2.0**-20)    # (0, 2 pi)
    return R * np.sin(phi), R * np.cos(phi)


def _box_muller_f32(w0, w1):
    """EBN_JOB_FP32_SAMPLING: the same bit fields, the transform in float32 (radius from the 32 bits
    of w0 only). NumPy's float32 log/sqrt/sin/cos are correctly rounded to ~1 ulp, the device uses the
    MUFU approximations: agreement ~1e-6 in z."""
    f32 = np.float32
    rot = ((w0 << np.uint64(12)) & _MASK) | (w0 >> np.uint64(20))
    u = rot.astype(f32) * f32(2.0**-32) + f32(2.0**-33)
    R = np.sqrt(f32(-2.0) * np.log(u))
    A = w1 >> np.uint64(10)
    f = ((A.astype(np.float64) + 0.5) * 2.0**-22 - 0.5).astype(f32)      # exact in float32
    ang = f32(2 * np.pi) * f                                              # Phi - pi
    return (-(R * np.sin(ang))).astype(np.float64), (-(R * np.cos(ang))).astype(np.float64)


def words_of(inp) -> int:
    kind = int(inp["kind"])
    trunc = np.isfinite(float(inp["lo"])) or np.isfinite(float(inp["hi"]))
    if kind == IN_PARAM or kind == IN_GAMMA:
        return 0
    if kind == IN_UNIFORM:
        return 2
    if kind in (IN_NORMAL, IN_LOGNORMAL) and not trunc:
        return 2
    return 4


def _gamma_mt_pair(pairs, shape: float, scale: float, stream: int, j: int, seed: int):
    """Both samples of every pair (layout v4): the first attempt shares one Philox call — Box-Muller of
    (w0, w1) gives z of sample 0 and z of sample 1, acceptance uniforms w2 and w3; rejected samples retry on
    their own counters base | e<<15 | t (normal from the first Box-Muller output, uniform from w3)."""
    a = shape + 1.0 if shape < 1.0 else shape
    dd = a - 1.0 / 3.0
    cc = 1.0 / np.sqrt(9.0 * dd)
    base = 0x80000000 | (j << 16)

    def accept(z, u):
        t1 = 1.0 + cc * z
        v = t1 * t1 * t1
        z2 = z * z
        with np.errstate(invalid="ignore", divide="ignore"):
            ok = (t1 > 0.0) & ((u < 1.0 - 0.0331 * z2 * z2) | (np.log(u) < 0.5 * z2 + dd * (1.0 - v + np.log(v))))
        return ok, v

    w0, w1, w2, w3 = _call(pairs, stream, base, seed)
    z0, z1 = _box_muller(w0, w1)
    outs = []
    for e, (z, u) in enumerate(((z0, _u32(w2)), (z1, _u32(w3)))):
        ok, v = accept(z, u)
        out = np.where(ok, dd * v, np.nan)
        todo = np.flatnonzero(~ok)
        t = 1
        while len(todo):
            r0, r1, _, r3 = _call(pairs[todo], stream, base | (e << 15) | t, seed)
            z, _ = _box_muller(r0, r1)
            ok, v = accept(z, _u32(r3))
            out[todo[ok]] = dd * v[ok]
            todo = todo[~ok]
            t += 1
            assert t < 64
        if shape < 1.0:
            b0, b1, _, _ = _call(pairs, stream, base | (e << 15) | 0x7FFF, seed)
            out = out * _u52(b0, b1) ** (1.0 / shape)
        outs.append(out * scale)
    return outs[0], outs[1]


def _window(inp):
    """(F(lo), F(hi)) of the truncation window in probability space, per marginal."""
    kind, p, lo, hi = int(inp["kind"]), inp["p"], float(inp["lo"]), float(inp["hi"])
    if kind in (IN_NORMAL, IN_LOGNORMAL):
        mu, sd = p[0], p[1]
        if kind == IN_LOGNORMAL:
            zl = (np.log(lo) - mu) / sd if lo > 0 else -np.inf
            zh = (np.log(hi) - mu) / sd if np.isfinite(hi) else np.inf
        else:
            zl, zh = (lo - mu) / sd, (hi - mu) / sd
        return special.ndtr(zl), special.ndtr(zh)
    if kind == IN_GUMBEL:
        mu, beta = p[0], p[1]
        Fl = np.exp(-np.exp(-(lo - mu) / beta)) if np.isfinite(lo) else 0.0
        Fh = np.exp(-np.exp(-(hi - mu) / beta)) if np.isfinite(hi) else 1.0
        return Fl, Fh
    if kind == IN_EXP_AFFINE:
        theta, sign, shift = p[0], p[1], p[2]
        if sign > 0:
            el, eh = max(lo - shift, 0.0), hi - shift
        else:
            el, eh = max(shift - hi, 0.0), shift - lo
        return -np.expm1(-el / theta), (1.0 if not np.isfinite(eh) else -np.expm1(-eh / theta))
    return 0.0, 1.0


def _quantile(inp, u, tables):
    """x = Q(F(lo) + u (F(hi) - F(lo))) for the inverse-CDF marginals."""
    kind, p = int(inp["kind"]), inp["p"]
    Fl, Fh = _window(inp)
    up = Fl + u * (Fh - Fl)
    if kind == IN_NORMAL:
        return p[0] + p[1] * special.ndtri(up)
    if kind == IN_LOGNORMAL:
        return np.exp(p[0] + p[1] * special.ndtri(up))
    if kind == IN_GUMBEL:
        return p[0] - p[1] * np.log(-np.log(up))
    if kind == IN_EXP_AFFINE:
        return p[2] + p[1] * (-p[0] * np.log1p(-up))
    if kind == IN_TABULATED:
        off, npts = int(p[0]), int(p[1])
        q = np.asarray(tables)[off:off + npts]
        pos = up * (npts - 1)
        i = np.clip(pos.astype(np.int64), 0, npts - 2)
        return q[i] + (pos - i) * (q[i + 1] - q[i])
    raise ValueError(f"unknown marginal kind {kind}")


def _pairs_of(first, n):
    idx = np.arange(first, first + n, dtype=np.uint64)
    pairs = idx >> np.uint64(1)
    odd = (idx & np.uint64(1)).astype(bool)
    upairs, inv = np.unique(pairs, return_inverse=True)
    return upairs, inv, odd


def sample_matrix(inputs_row, stream: int, seed: int, first: int, n: int, tables=None, fp32: bool = False) -> np.ndarray:
    """The n samples [first, first+n) of one scenario, (n, d), as the device draws them
    (fp32: with EBN_JOB_FP32_SAMPLING)."""
    d = len(inputs_row)
    upairs, inv, odd = _pairs_of(first, n)
    W = _Words(upairs, stream, seed)
    X = np.empty((n, d))
    off = 0
    for j in range(d):
        inp = inputs_row[j]
        kind, p = int(inp["kind"]), inp["p"]
        lo, hi = float(inp["lo"]), float(inp["hi"])
        nw = words_of(inp)
        if kind == IN_PARAM:
            X[:, j] = p[0]
        elif kind == IN_UNIFORM:
            a, b = max(p[0], lo), min(p[1], hi)
            u = np.where(odd, _u32(W[off + 1])[inv], _u32(W[off])[inv])
            X[:, j] = a + u * (b - a)
        elif kind in (IN_NORMAL, IN_LOGNORMAL) and nw == 2:
            z0, z1 = (_box_muller_f32 if fp32 else _box_muller)(W[off], W[off + 1])
            v = p[0] + p[1] * np.where(odd, z1[inv], z0[inv])
            X[:, j] = np.exp(v) if kind == IN_LOGNORMAL else v
        elif kind == IN_GAMMA:
            assert not (np.isfinite(lo) or np.isfinite(hi))
            g0, g1 = _gamma_mt_pair(upairs, p[0], p[1], stream, j, seed)
            X[:, j] = np.where(odd, g1[inv], g0[inv])
        else:
            u = np.where(odd, _u52(W[off + 2], W[off + 3])[inv], _u52(W[off], W[off + 1])[inv])
            X[:, j] = _quantile(inp, u, tables)
        off += nw
    return X


def copula_sample_matrix(inputs_row, L, stream: int, seed: int, first: int, n: int, tables=None):
    """Gaussian-copula variant: eps by Box-Muller per random input (2 words each), z = L eps,
    x = a + b z for the (log)normal family, x = Q(Phi(z)) otherwise."""
    d = len(inputs_row)
    L = np.asarray(L, dtype=np.float64).reshape(d, d)
    upairs, inv, odd = _pairs_of(first, n)
    W = _Words(upairs, stream, seed)
    eps = np.zeros((n, d))
    off = 0
    for j in range(d):
        if int(inputs_row[j]["kind"]) == IN_PARAM:
            continue
        z0, z1 = _box_muller(W[off], W[off + 1])
        eps[:, j] = np.where(odd, z1[inv], z0[inv])
        off += 2
    Z = eps @ L.T
    X = np.empty((n, d))
    for j in range(d):
        inp = inputs_row[j]
        kind, p = int(inp["kind"]), inp["p"]
        lo, hi = float(inp["lo"]), float(inp["hi"])
        trunc = np.isfinite(lo) or np.isfinite(hi)
        z = Z[:, j]
        if kind == IN_PARAM:
            X[:, j] = p[0]
        elif kind in (IN_NORMAL, IN_LOGNORMAL) and not trunc:
            v = p[0] + p[1] * z
            X[:, j] = np.exp(v) if kind == IN_LOGNORMAL else v
        elif kind == IN_UNIFORM:
            a, b = max(p[0], lo), min(p[1], hi)
            X[:, j] = a + special.ndtr(z) * (b - a)
        elif kind == IN_GAMMA:
            raise ValueError("Gamma under a copula must be tabulated")
        else:
            X[:, j] = _quantile(inp, special.ndtr(z), tables)
    return X


# ---- quasi-Monte Carlo: the Sobol twin (csrc/ebn_qmc.cuh) --------------------------------------------
def sobol_direction_numbers():
    """Joe & Kuo direction numbers of the first 16 dimensions, recomputed from the primitive polynomials
    (tools/gen_sobol_tables.py holds the table; tests check both against scipy.stats.qmc.Sobol)."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "gen_sobol_tables.py")
    spec = importlib.util.spec_from_file_location("gen_sobol_tables", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return np.array(mod.direction_numbers(), dtype=np.uint64)


def sobol_points(first: int, n: int, nd: int, V=None):
    """Unscrambled 32-bit Sobol coordinates of samples first..first+n-1 (gray-code order), shape (n, nd)."""
    V = sobol_direction_numbers() if V is None else V
    idx = np.arange(first, first + n, dtype=np.uint64)
    g = idx ^ (idx >> np.uint64(1))
    out = np.zeros((n, nd), dtype=np.uint64)
    for k in range(32):
        bit = ((g >> np.uint64(k)) & np.uint64(1)).astype(bool)
        out[bit] ^= V[:nd, k]
    return out


def _brev32(x):
    x = np.asarray(x, dtype=np.uint64)
    r = np.zeros_like(x)
    for k in range(32):
        r |= ((x >> np.uint64(k)) & np.uint64(1)) << np.uint64(31 - k)
    return r


def owen_scramble(x, key: int):
    """Hash-based nested uniform scrambling (Laine & Karras 2011 / Burley 2020), uint32 arithmetic."""
    m = np.uint64(0xFFFFFFFF)
    x = _brev32(x)
    x = (x + np.uint64(key)) & m
    for c in (0x6C50B47C, 0xB82F1E52, 0xC7AFE638, 0x8D22F6E6):
        x = x ^ ((x * np.uint64(c)) & m)
    return _brev32(x)


def qmc_sample_matrix(inputs_row, stream: int, seed: int, first: int, n: int, tables=None, scramble: bool = True):
    """EBN_JOB_QMC_SOBOL: sample i is point i of the scrambled Sobol sequence; random input number r (Parameters
    take no dimension) uses dimension r, key = Philox(counter (r, 0, stream, 'SOBL'), seed).w0; every
    marginal through its quantile on the truncation window (SciPy)."""
    d = len(inputs_row)
    random = [j for j in range(d) if int(inputs_row[j]["kind"]) != IN_PARAM]
    P = sobol_points(first, n, len(random))
    X = np.empty((n, d))
    for j in range(d):
        inp = inputs_row[j]
        kind, p = int(inp["kind"]), inp["p"]
        lo, hi = float(inp["lo"]), float(inp["hi"])
        if kind == IN_PARAM:
            X[:, j] = p[0]
            continue
        r = random.index(j)
        w = P[:, r]
        if scramble:
            key = int(philox4x32_10(np.uint64(r), np.uint64(0), np.uint64(stream), np.uint64(0x534F424C),
                                    seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)[0])
            w = owen_scramble(w, key)
        u = (w.astype(np.float64) + 0.5) * 2.0**-32
        if kind == IN_UNIFORM:
            a, b = max(p[0], lo), min(p[1], hi)
            X[:, j] = a + u * (b - a)
        elif kind == IN_GAMMA:
            raise ValueError("Gamma under quasi-Monte Carlo must be tabulated")
        elif kind in (IN_NORMAL, IN_LOGNORMAL) and not (np.isfinite(lo) or np.isfinite(hi)):
            v = p[0] + p[1] * special.ndtri(u)
            X[:, j] = np.exp(v) if kind == IN_LOGNORMAL else v
        else:
            X[:, j] = _quantile(inp, u, tables)
    return X


# ---- limit states in NumPy, Julia evaluation order (second restatement) -----
def _jl_min(a, b):
    return np.where(np.isnan(a), a, np.where(np.isnan(b), b, np.minimum(a, b)))


def g_vehicle(X, k):
    """demo/vehicle_suspension.jl:17-23."""
    A, b0, V, C, Ck, K = (X[:, i] for i in range(6))
    M, m, g, pow_a, pow_b = k
    g1 = 1.0 - (((np.pi * m) * V) * A) / ((b0 * K) * (g * g)) * (
        ((Ck / (m + M) - C / M) * (Ck / (m + M) - C / M) + (C * C) / (m * M)) + (Ck * (K * K)) / (m * (M * M)))
    g2 = (4000.0 * C) * pow_a - 8.6394
    g3 = 2.0 * np.sqrt((M * g) * (((K * K) * Ck) / (C * (m + M)) + C)) - 1.0
    g4 = Ck - pow_b
    return _jl_min(_jl_min(_jl_min(g1, g2), g3), g4)


def g_straub(X, k):
    """demo/straub_example.jl:16-39."""
    lam, zeta, rho = k
    ur, v, h = X[:, 0], X[:, 1], X[:, 2]
    mu = lam + (rho * zeta) * ur
    sd = np.sqrt((1.0 - rho * rho) * (zeta * zeta))
    r = [np.exp(mu + sd * X[:, 3 + i]) for i in range(5)]
    g1 = (((r[0] + r[1]) + r[3]) + r[4]) - 5.0 * h
    g2 = ((r[1] + 2.0 * r[2]) + r[3]) - 5.0 * v
    g3 = ((((r[0] + 2.0 * r[2]) + 2.0 * r[3]) + r[4]) - 5.0 * h) - 5.0 * v
    return _jl_min(_jl_min(g1, g2), g3)


def g_hydrogen(X, k=None):
    """demo/Hydrogen_new/ebn.jl:98-108."""
    return np.where(X[:, 0] == 1.0, X[:, 1] - X[:, 2], np.where(X[:, 0] == 2.0, X[:, 1] - X[:, 3], 1.0))


def g_polynomial(X, prog):
    prog = list(prog)
    cols = [X[:, j] for j in range(X.shape[1])]
    T = int(prog[0])
    pos = 1
    acc = None
    for _ in range(T):
        nterms = int(prog[pos]); pos += 1
        acc = None
        for _k in range(nterms):
            term = np.full(X.shape[0], prog[pos]); pos += 1
            nf = int(prog[pos]); pos += 1
            for _f in range(nf):
                term = term * cols[int(prog[pos])]; pos += 1
            acc = term if acc is None else acc + term
        cols.append(acc)
    return acc


def g_min_affine(X, k):
    K = int(k[0])
    d = X.shape[1]
    g = None
    for r in range(K):
        row = k[1 + r * (1 + d):1 + (r + 1) * (1 + d)]
        acc = np.full(X.shape[0], row[0])
        for j in range(d):
            acc = acc + row[1 + j] * X[:, j]
        g = acc if g is None else _jl_min(g, acc)
    return g


def _jl_max(a, b):
    return np.where(np.isnan(a), a, np.where(np.isnan(b), b, np.maximum(a, b)))


def g_expression(X, prog, want_columns: bool = False):
    """Expression program (include/ebncuda.h), vectorised over the rows of X: second restatement of
    oracle/ebn_oracle.c g_expression. Opcode numbers are the EBN_OP_* values."""
    prog = np.asarray(prog, dtype=np.float64)
    n = X.shape[0]
    P, L = int(prog[0]), int(prog[1])
    k, code = prog[2:2 + P], prog[2 + P:]
    cols = [X[:, j] for j in range(X.shape[1])]
    st = []
    full = lambda v: np.full(n, v)  # noqa: E731
    with np.errstate(all="ignore"):
        for i in range(L):
            op, arg = int(code[2 * i]), code[2 * i + 1]
            if op == 0: st.append(cols[int(arg)])
            elif op == 1: st.append(full(k[int(arg)]))
            elif op == 2: st.append(full(arg))
            elif op in (3, 4, 5, 6, 14, 15, 16, 18):
                b = st.pop(); a = st.pop()
                st.append({3: lambda: a + b, 4: lambda: a - b, 5: lambda: a * b, 6: lambda: a / b,
                           14: lambda: _jl_min(a, b), 15: lambda: _jl_max(a, b), 16: lambda: np.power(a, b),
                           18: lambda: (a < b).astype(np.float64)}[op]())
            elif op in (7, 8, 9, 10, 11, 12, 13):
                a = st.pop()
                st.append({7: np.negative, 8: np.abs, 9: np.sqrt, 10: np.exp, 11: np.log, 12: np.sin, 13: np.cos}[op](a))
            elif op == 17:
                a = st.pop(); e = int(arg)
                st.append(np.ones(n) if e == 0 else a if e == 1 else a * a if e == 2 else (a * a) * a if e == 3
                          else np.power(a, float(e)))
            elif op == 19:
                b = st.pop(); a = st.pop(); c = st.pop()
                st.append(np.where(c != 0.0, a, b))
            elif op == 20:
                cols.append(st.pop())
            else:
                raise ValueError(f"opcode {op}")
    assert len(st) == 1
    return (st[0], np.column_stack(cols)) if want_columns else st[0]


LS_EXPRESSION = 8
LIMIT_STATES = {LS_EXPRESSION: g_expression, LS_VEHICLE: g_vehicle, LS_STRAUB: g_straub, LS_HYDROGEN: g_hydrogen,
                LS_POLYNOMIAL: g_polynomial, LS_MIN_AFFINE: g_min_affine}


def limit_state(ls: int, consts, X):
    return LIMIT_STATES[ls](np.asarray(X, dtype=np.float64), np.asarray(consts, dtype=np.float64))
