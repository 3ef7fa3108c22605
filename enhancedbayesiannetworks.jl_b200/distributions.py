"""Univariate marginals the reference's continuous nodes carry (Distributions.jl objects on the
Julia side) and the UQ.jl input types `_uq_inputs` emits (reference src/nodes/continuousnodes.jl:39-50,
src/nodes/discretenodes.jl:43-49). Host-side only: each knows its cdf/quantile/support (used by
the discretisation pre-step, reference src/networks/ebn/discretization/discretize.jl:68-81) and how
to describe itself as an `ebn_input_t` row for the C ABI. Sampling never happens here.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Sequence, Tuple

import numpy as np
from scipy import special, stats

from . import _lib as L

INF = math.inf


class Distribution:
    """Base class; subclasses define cdf, quantile, support and abi()."""

    def cdf(self, x):
        raise NotImplementedError

    def quantile(self, u):
        raise NotImplementedError

    def support(self) -> Tuple[float, float]:
        raise NotImplementedError

    def abi(self, lo: float = -INF, hi: float = INF):
        """(kind, p0, p1, p2, p3, lo, hi) — an `ebn_input_t` row, truncated to [lo, hi]."""
        raise NotImplementedError

    def mean(self):
        raise NotImplementedError


@dataclass(frozen=True)
class Normal(Distribution):
    mu: float = 0.0
    sigma: float = 1.0

    def cdf(self, x):
        return special.ndtr((np.asarray(x, dtype=float) - self.mu) / self.sigma)

    def quantile(self, u):
        return self.mu + self.sigma * special.ndtri(u)

    def support(self):
        return (-INF, INF)

    def abi(self, lo=-INF, hi=INF):
        return (L.IN_NORMAL, self.mu, self.sigma, 0.0, 0.0, lo, hi)

    def mean(self):
        return self.mu


@dataclass(frozen=True)
class LogNormal(Distribution):
    mu: float = 0.0
    sigma: float = 1.0

    def cdf(self, x):
        x = np.asarray(x, dtype=float)
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.where(x > 0, special.ndtr((np.log(np.maximum(x, 1e-300)) - self.mu) / self.sigma), 0.0)

    def quantile(self, u):
        return np.exp(self.mu + self.sigma * special.ndtri(u))

    def support(self):
        return (0.0, INF)

    def abi(self, lo=-INF, hi=INF):
        return (L.IN_LOGNORMAL, self.mu, self.sigma, 0.0, 0.0, lo, hi)

    def mean(self):
        return math.exp(self.mu + 0.5 * self.sigma**2)


@dataclass(frozen=True)
class Uniform(Distribution):
    a: float = 0.0
    b: float = 1.0

    def cdf(self, x):
        return np.clip((np.asarray(x, dtype=float) - self.a) / (self.b - self.a), 0.0, 1.0)

    def quantile(self, u):
        return self.a + np.asarray(u) * (self.b - self.a)

    def support(self):
        return (self.a, self.b)

    def abi(self, lo=-INF, hi=INF):
        return (L.IN_UNIFORM, self.a, self.b, 0.0, 0.0, lo, hi)

    def mean(self):
        return 0.5 * (self.a + self.b)


@dataclass(frozen=True)
class Gumbel(Distribution):
    """Maximum-type Gumbel(mu, beta) as in Distributions.jl."""
    mu: float = 0.0
    beta: float = 1.0

    def cdf(self, x):
        return np.exp(-np.exp(-(np.asarray(x, dtype=float) - self.mu) / self.beta))

    def quantile(self, u):
        return self.mu - self.beta * np.log(-np.log(u))

    def support(self):
        return (-INF, INF)

    def abi(self, lo=-INF, hi=INF):
        return (L.IN_GUMBEL, self.mu, self.beta, 0.0, 0.0, lo, hi)

    def mean(self):
        return self.mu + 0.5772156649015329 * self.beta


@dataclass(frozen=True)
class Gamma(Distribution):
    """Gamma(shape alpha, scale theta)."""
    alpha: float = 1.0
    theta: float = 1.0

    def cdf(self, x):
        return stats.gamma.cdf(x, self.alpha, scale=self.theta)

    def quantile(self, u):
        return stats.gamma.ppf(u, self.alpha, scale=self.theta)

    def support(self):
        return (0.0, INF)

    def abi(self, lo=-INF, hi=INF):
        if lo > 0.0 or hi < INF:
            # the device draws Gamma by Marsaglia-Tsang (no truncation): a truncated Gamma is
            # shipped as an inverse-CDF table (see Tabulated / abi_with_tables)
            raise NeedsTable(self, lo, hi)
        return (L.IN_GAMMA, self.alpha, self.theta, 0.0, 0.0, -INF, INF)

    def mean(self):
        return self.alpha * self.theta


@dataclass(frozen=True)
class Exponential(Distribution):
    """`sign*Exponential(theta) + shift` — the tails `_approximate` builds
    (reference src/networks/ebn/discretization/discretize.jl:83-91)."""
    theta: float = 1.0
    sign: float = 1.0
    shift: float = 0.0

    def cdf(self, x):
        x = np.asarray(x, dtype=float)
        e = (x - self.shift) * self.sign
        base = np.where(e > 0, -np.expm1(-np.maximum(e, 0.0) / self.theta), 0.0)
        return base if self.sign > 0 else np.where(e > 0, np.exp(-np.maximum(e, 0) / self.theta), 1.0)

    def quantile(self, u):
        u = np.asarray(u, dtype=float)
        if self.sign > 0:
            return self.shift - self.theta * np.log1p(-u)
        return self.shift + self.theta * np.log(u)

    def support(self):
        return (self.shift, INF) if self.sign > 0 else (-INF, self.shift)

    def abi(self, lo=-INF, hi=INF):
        return (L.IN_EXPONENTIAL_AFFINE, self.theta, self.sign, self.shift, 0.0, lo, hi)

    def mean(self):
        return self.shift + self.sign * self.theta

    def __neg__(self):
        return Exponential(self.theta, -self.sign, -self.shift)

    def __add__(self, c):
        return Exponential(self.theta, self.sign, self.shift + c)

    __radd__ = __add__


class NeedsTable(Exception):
    """A marginal the device can only sample through an inverse-CDF table."""

    def __init__(self, dist, lo, hi):
        super().__init__(f"{dist} truncated to [{lo}, {hi}] needs an inverse-CDF table")
        self.dist, self.lo, self.hi = dist, lo, hi


@dataclass(frozen=True)
class Truncated(Distribution):
    """`truncated(dist, lo, hi)` (reference src/nodes/continuousnodes.jl:71-73)."""
    dist: Distribution
    lo: float
    hi: float

    def _F(self):
        return float(self.dist.cdf(self.lo)), float(self.dist.cdf(self.hi))

    def cdf(self, x):
        Fl, Fh = self._F()
        return np.clip((self.dist.cdf(np.clip(x, self.lo, self.hi)) - Fl) / (Fh - Fl), 0.0, 1.0)

    def quantile(self, u):
        Fl, Fh = self._F()
        return self.dist.quantile(Fl + np.asarray(u) * (Fh - Fl))

    def support(self):
        a, b = self.dist.support()
        return (max(a, self.lo), min(b, self.hi))

    def abi(self, lo=-INF, hi=INF):
        return self.dist.abi(max(lo, self.lo), min(hi, self.hi))

    def mean(self):
        u = (np.arange(4096) + 0.5) / 4096
        return float(np.mean(self.quantile(u)))


def truncated(dist: Distribution, lo: float, hi: float) -> Truncated:
    return Truncated(dist, float(lo), float(hi))


@dataclass(frozen=True)
class Tabulated(Distribution):
    """Inverse-CDF table on a uniform u grid (EBN_IN_TABULATED): x = Q(u) piecewise linear."""
    q: Tuple[float, ...]

    def quantile(self, u):
        n = len(self.q)
        pos = np.asarray(u, dtype=float) * (n - 1)
        i = np.clip(pos.astype(np.int64), 0, n - 2)
        q = np.asarray(self.q)
        return q[i] + (pos - i) * (q[i + 1] - q[i])

    def cdf(self, x):
        q = np.asarray(self.q)
        return np.interp(x, q, np.linspace(0.0, 1.0, len(q)), left=0.0, right=1.0)

    def support(self):
        return (self.q[0], self.q[-1])

    def mean(self):
        q = np.asarray(self.q)
        return float(np.mean(0.5 * (q[1:] + q[:-1])))

    def abi(self, lo=-INF, hi=INF):
        raise NeedsTable(self, lo, hi)


class EmpiricalDistribution(Distribution):
    """What a ContinuousFunctionalNode evaluates to. The reference fits a KDE on the n model
    outputs (`EmpiricalDistribution(df[:, name])`, evaluate_node.jl:26); the device path keeps a
    per-scenario HISTOGRAM over a fine edge grid instead — the CDF is
    piecewise linear between the edges, which is all the discretisation step consumes
    (bin probabilities and `truncated(dist, a, b)` samplers)."""

    def __init__(self, edges: Sequence[float], counts: Sequence[int], total: int, s1: float = 0.0, s2: float = 0.0):
        self.edges = np.asarray(edges, dtype=float)
        c = np.asarray(counts, dtype=float)
        assert c.shape == (len(self.edges) + 1,)
        self.counts, self.total = c, float(total)
        self.cum = np.concatenate([[0.0], np.cumsum(c)]) / self.total  # F at -inf, edges..., +inf
        self.sum, self.sumsq = s1, s2

    def cdf(self, x):
        x = np.asarray(x, dtype=float)
        F_edges = self.cum[1:-1]
        out = np.interp(x, self.edges, F_edges, left=np.nan, right=np.nan)
        out = np.where(x < self.edges[0], 0.0, out)      # mass outside the grid sits at the ends
        out = np.where(x >= self.edges[-1], 1.0, out)
        return out

    def quantile(self, u):
        F = self.cum[1:-1].copy()
        F[0], F[-1] = 0.0, 1.0
        keep = np.concatenate([[True], np.diff(F) > 0])
        return np.interp(u, F[keep], self.edges[keep])

    def support(self):
        return (float(self.edges[0]), float(self.edges[-1]))

    def mean(self):
        return self.sum / self.total

    def std(self):
        m = self.mean()
        return math.sqrt(max(self.sumsq / self.total - m * m, 0.0))

    def abi(self, lo=-INF, hi=INF):
        raise NeedsTable(self, lo, hi)

    def __eq__(self, other):
        return isinstance(other, EmpiricalDistribution) and np.array_equal(self.edges, other.edges) \
            and np.array_equal(self.counts, other.counts)

    def __hash__(self):
        return hash((self.edges.tobytes(), self.counts.tobytes()))

    def __repr__(self):
        return f"EmpiricalDistribution(n={int(self.total)}, mean={self.mean():.6g}, std={self.std():.6g})"


def table_for(dist: Distribution, lo: float, hi: float, npts: int = 4097) -> np.ndarray:
    """Quantiles of `truncated(dist, lo, hi)` on a uniform u grid of npts points."""
    a, b = dist.support()
    lo, hi = max(lo, a), min(hi, b)
    Fl, Fh = float(dist.cdf(lo)), float(dist.cdf(hi))
    if not Fh > Fl:
        # A window the (histogram) distribution gives no mass to, e.g. an outer bin no sample of a
        # ContinuousFunctionalNode fell into. The reference's KDE has a positive tail there; the bin has
        # probability 0 in the discretised parent, so any proper law on the window serves: uniform.
        if not (np.isfinite(lo) and np.isfinite(hi)):
            raise ValueError(f"{dist} has no mass in [{lo}, {hi}]")
        return np.linspace(lo, hi, npts)
    eps = 1e-12
    u = np.linspace(0.0, 1.0, npts)
    p = np.clip(Fl + u * (Fh - Fl), eps if not np.isfinite(lo) else Fl, 1 - eps if not np.isfinite(hi) else Fh)
    q = np.asarray(dist.quantile(p), dtype=float)
    if np.isfinite(lo):
        q[0] = lo
    if np.isfinite(hi):
        q[-1] = hi
    return np.maximum.accumulate(q)


def distribution_parameters(mean: float, std: float, kind) -> Tuple[float, float]:
    """UQ.jl's `distribution_parameters(mean, std, Dist)` for the three families the demos use
    (demo/straub_example.jl:9,14,21)."""
    name = kind if isinstance(kind, str) else kind.__name__
    if name == "LogNormal":
        z2 = math.log(1.0 + (std / mean) ** 2)
        return math.log(mean) - 0.5 * z2, math.sqrt(z2)
    if name == "Gamma":
        return (mean / std) ** 2, std * std / mean
    if name == "Gumbel":
        beta = std * math.sqrt(6.0) / math.pi
        return mean - 0.5772156649015329 * beta, beta
    raise ValueError(f"distribution_parameters: unsupported family {name}")


# ---- UQ.jl input types ---------------------------------------------------------------------
@dataclass(frozen=True)
class Parameter:
    value: float
    name: str


@dataclass(frozen=True)
class RandomVariable:
    dist: Distribution
    name: str


@dataclass(frozen=True)
class Interval:
    lb: float
    ub: float
    name: str


@dataclass(frozen=True)
class ProbabilityBox:
    """`ProbabilityBox{T}(parameters, name, lb, ub)`: a family T(p...) with interval parameters."""
    family: type
    parameters: Tuple[Interval, ...]
    name: str
    lb: float = -INF
    ub: float = INF

    def member(self, values: Sequence[float]) -> Distribution:
        d = self.family(*values)
        if self.lb > -INF or self.ub < INF:
            return Truncated(d, self.lb, self.ub)
        return d

    def cdf_bounds(self, x) -> Tuple[float, float]:
        """Envelope of the family's cdf at x over the vertices of the parameter box."""
        import itertools
        vals = [float(self.member(v).cdf(x)) for v in itertools.product(*[(p.lb, p.ub) for p in self.parameters])]
        return min(vals), max(vals)


@dataclass(frozen=True)
class GaussianCopula:
    """UQ.jl `GaussianCopula(R)`: R is the correlation matrix of the normal scores."""
    correlation: Tuple[Tuple[float, ...], ...]

    def __init__(self, correlation):
        R = np.asarray(correlation, dtype=float)
        if R.ndim != 2 or R.shape[0] != R.shape[1] or not np.allclose(R, R.T) or not np.allclose(np.diag(R), 1.0):
            raise ValueError("GaussianCopula needs a symmetric correlation matrix with unit diagonal")
        np.linalg.cholesky(R)  # raises if not positive definite
        object.__setattr__(self, "correlation", tuple(map(tuple, R.tolist())))


@dataclass(frozen=True)
class JointDistribution:
    """UQ.jl `JointDistribution(marginals, copula)`: z = L eps, u = Phi(z),
    x_j = quantile(marginal_j, u_j). The reference's node API never builds one; it is
    accepted wherever a list of UQ inputs is (probability_of_failure)."""
    marginals: Tuple[RandomVariable, ...]
    copula: GaussianCopula

    def __init__(self, marginals, copula):
        marginals = tuple(marginals)
        if len(marginals) != len(copula.correlation):
            raise ValueError("one marginal per row of the correlation matrix")
        object.__setattr__(self, "marginals", marginals)
        object.__setattr__(self, "copula", copula)

    @property
    def name(self):
        return tuple(m.name for m in self.marginals)
