"""NumPy twin of the device sampler (test infrastructure, see oracle/__init__.py).

Restates, independently of the CUDA sources, the stream layout of DESIGN.md
("Random streams, layout v1") so that the uniforms the GPU consumed can be
regenerated on the CPU:

    key     = (seed & 0xffffffff, seed >> 32)
    counter = (pair & 0xffffffff, pair >> 32, stream id of the scenario, slot)
    pair    = global sample index >> 1,   slot = input index j
    (w0,w1,w2,w3) = Philox4x32-10(counter, key)
    A = w1:w0, B = w3:w2,  U(x) = ((x >> 12) + 0.5) * 2**-52

    inverse-CDF marginals : sample 2q uses U(A), sample 2q+1 uses U(B)
    untruncated (log)normal: r = sqrt(-2 ln U(A)), t = 2*pi*U(B);
                             z(2q) = r cos t, z(2q+1) = r sin t      (Box-Muller)
    Gamma (Marsaglia-Tsang): attempt t of sample e reads slot j + 16*(1+2t+e):
                             radius U(w1:w0), angle (w2+0.5)*2**-32 (cos branch),
                             acceptance uniform (w3+0.5)*2**-32

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
    x = (hi << np.uint64(32)) | lo
    return ((x >> np.uint64(12)).astype(np.float64) + 0.5) * 2.0**-52


def _u32(w):
    return (w.astype(np.float64) + 0.5) * 2.0**-32


def _pair_words(pairs, stream: int, slot: int, seed: int):
    pairs = np.asarray(pairs, dtype=np.uint64)
    return philox4x32_10(pairs & _MASK, pairs >> np.uint64(32), np.uint64(stream), np.uint64(slot),
                         seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


def _gamma_mt(pairs, e: int, shape: float, scale: float, stream: int, j: int, seed: int, uboost):
    a = shape + 1.0 if shape < 1.0 else shape
    dd = a - 1.0 / 3.0
    cc = 1.0 / np.sqrt(9.0 * dd)
    out = np.empty(len(pairs))
    todo = np.arange(len(pairs))
    t = 0
    while len(todo):
        w0, w1, w2, w3 = _pair_words(pairs[todo], stream, j + 16 * (1 + 2 * t + e), seed)
        r = np.sqrt(-2.0 * np.log(_u52(w0, w1)))
        z = r * np.cos(2.0 * np.pi * _u32(w2))
        t1 = 1.0 + cc * z
        v = t1 * t1 * t1
        u = _u32(w3)
        z2 = z * z
        with np.errstate(invalid="ignore", divide="ignore"):
            ok = (t1 > 0.0) & ((u < 1.0 - 0.0331 * z2 * z2) | (np.log(u) < 0.5 * z2 + dd * (1.0 - v + np.log(v))))
        out[todo[ok]] = dd * v[ok]
        todo = todo[~ok]
        t += 1
        assert t < 64
    if shape < 1.0:
        out = out * uboost ** (1.0 / shape)
    return out * scale


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
        el, eh = 0.0, np.inf
        if sign > 0:
            el = max(lo - shift, 0.0)
            eh = hi - shift
        else:
            el = max(shift - hi, 0.0)
            eh = shift - lo
        return -np.expm1(-el / theta), (1.0 if not np.isfinite(eh) else -np.expm1(-eh / theta))
    return 0.0, 1.0


def sample_matrix(inputs_row, stream: int, seed: int, first: int, n: int, tables=None) -> np.ndarray:
    """The n samples [first, first+n) of one scenario, (n, d), as the device draws them."""
    d = len(inputs_row)
    idx = np.arange(first, first + n, dtype=np.uint64)
    pairs = idx >> np.uint64(1)
    odd = (idx & np.uint64(1)).astype(bool)
    upairs, inv = np.unique(pairs, return_inverse=True)
    X = np.empty((n, d))
    for j in range(d):
        inp = inputs_row[j]
        kind, p = int(inp["kind"]), inp["p"]
        lo, hi = float(inp["lo"]), float(inp["hi"])
        trunc = np.isfinite(lo) or np.isfinite(hi)
        if kind == IN_PARAM:
            X[:, j] = p[0]
            continue
        w0, w1, w2, w3 = _pair_words(upairs, stream, j, seed)
        ua, ub = _u52(w0, w1), _u52(w2, w3)
        u = np.where(odd, ub[inv], ua[inv])
        if kind == IN_UNIFORM:
            a, b = max(p[0], lo), min(p[1], hi)
            X[:, j] = a + u * (b - a)
        elif kind in (IN_NORMAL, IN_LOGNORMAL):
            if not trunc:
                r = np.sqrt(-2.0 * np.log(ua))
                z = np.where(odd, (r * np.sin(2.0 * np.pi * ub))[inv], (r * np.cos(2.0 * np.pi * ub))[inv])
            else:
                Fl, Fh = _window(inp)
                z = special.ndtri(Fl + u * (Fh - Fl))
            v = p[0] + p[1] * z
            X[:, j] = np.exp(v) if kind == IN_LOGNORMAL else v
        elif kind == IN_GUMBEL:
            Fl, Fh = _window(inp)
            X[:, j] = p[0] - p[1] * np.log(-np.log(Fl + u * (Fh - Fl)))
        elif kind == IN_EXP_AFFINE:
            Fl, Fh = _window(inp)
            X[:, j] = p[2] + p[1] * (-p[0] * np.log1p(-(Fl + u * (Fh - Fl))))
        elif kind == IN_GAMMA:
            assert not trunc
            g0 = _gamma_mt(upairs, 0, p[0], p[1], stream, j, seed, ua)
            g1 = _gamma_mt(upairs, 1, p[0], p[1], stream, j, seed, ub)
            X[:, j] = np.where(odd, g1[inv], g0[inv])
        elif kind == IN_TABULATED:
            off, npts = int(p[0]), int(p[1])
            q = np.asarray(tables)[off:off + npts]
            pos = u * (npts - 1)
            i = np.clip(pos.astype(np.int64), 0, npts - 2)
            X[:, j] = q[i] + (pos - i) * (q[i + 1] - q[i])
        else:
            raise ValueError(f"unknown marginal kind {kind}")
    return X


def copula_sample_matrix(inputs_row, L, stream: int, seed: int, first: int, n: int, tables=None):
    """Gaussian-copula variant: eps by Box-Muller per input, z = L eps, x = Q(Phi(z))."""
    d = len(inputs_row)
    L = np.asarray(L, dtype=np.float64).reshape(d, d)
    idx = np.arange(first, first + n, dtype=np.uint64)
    pairs = idx >> np.uint64(1)
    odd = (idx & np.uint64(1)).astype(bool)
    upairs, inv = np.unique(pairs, return_inverse=True)
    eps = np.zeros((n, d))
    for j in range(d):
        if int(inputs_row[j]["kind"]) == IN_PARAM:
            continue
        w0, w1, w2, w3 = _pair_words(upairs, stream, j, seed)
        ua, ub = _u52(w0, w1), _u52(w2, w3)
        r = np.sqrt(-2.0 * np.log(ua))
        eps[:, j] = np.where(odd, (r * np.sin(2.0 * np.pi * ub))[inv], (r * np.cos(2.0 * np.pi * ub))[inv])
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
        elif kind in (IN_NORMAL, IN_LOGNORMAL):
            if trunc:
                Fl, Fh = _window(inp)
                z = special.ndtri(Fl + special.ndtr(z) * (Fh - Fl))
            v = p[0] + p[1] * z
            X[:, j] = np.exp(v) if kind == IN_LOGNORMAL else v
        elif kind == IN_UNIFORM:
            a, b = max(p[0], lo), min(p[1], hi)
            X[:, j] = a + special.ndtr(z) * (b - a)
        elif kind == IN_GUMBEL:
            Fl, Fh = _window(inp)
            X[:, j] = p[0] - p[1] * np.log(-np.log(Fl + special.ndtr(z) * (Fh - Fl)))
        elif kind == IN_EXP_AFFINE:
            Fl, Fh = _window(inp)
            X[:, j] = p[2] + p[1] * (-p[0] * np.log1p(-(Fl + special.ndtr(z) * (Fh - Fl))))
        elif kind == IN_TABULATED:
            off, npts = int(p[0]), int(p[1])
            q = np.asarray(tables)[off:off + npts]
            pos = special.ndtr(z) * (npts - 1)
            i = np.clip(pos.astype(np.int64), 0, npts - 2)
            X[:, j] = q[i] + (pos - i) * (q[i + 1] - q[i])
        else:
            raise ValueError(f"marginal kind {kind} not supported under a copula")
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


LIMIT_STATES = {LS_VEHICLE: g_vehicle, LS_STRAUB: g_straub, LS_HYDROGEN: g_hydrogen,
                LS_POLYNOMIAL: g_polynomial, LS_MIN_AFFINE: g_min_affine}


def limit_state(ls: int, consts, X):
    return LIMIT_STATES[ls](np.asarray(X, dtype=np.float64), np.asarray(consts, dtype=np.float64))
