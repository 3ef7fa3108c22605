"""The BASELINE.json configurations expressed as C-ABI jobs.

Parameter values come from the reference demos:
  vehicle suspension  demo/vehicle_suspension.jl:3-14 (M, m, g, A, b0, V, C, Ck, K)
  Straub frame        demo/straub_example.jl:4-24 (Ur, V~Gamma, H~Gumbel, rho, lambda, zeta)
"""
from __future__ import annotations

import math

import numpy as np

from . import _lib as L

INF = math.inf

# demo/vehicle_suspension.jl:3-5
VEH_M, VEH_m, VEH_g = 3.2633, 0.8158, 981.0


def vehicle_consts() -> np.ndarray:
    """(M, m, g, (M*g)^-1.5, (g*(M+m))^0.877): the two powers are evaluated by the caller."""
    return np.array([VEH_M, VEH_m, VEH_g, (VEH_M * VEH_g) ** (-1.5), (VEH_g * (VEH_M + VEH_m)) ** 0.877])


def _veh_row(A, b0, v_lo, v_hi, c_mu=431.7221):
    return [
        (L.IN_PARAM, A, 0, 0, 0, -INF, INF),
        (L.IN_PARAM, b0, 0, 0, 0, -INF, INF),
        (L.IN_UNIFORM, 7.0, 12.0, 0, 0, v_lo, v_hi),       # truncated(Uniform(7,12), lo, hi)
        (L.IN_NORMAL, c_mu, 10.0, 0, 0, -INF, INF),        # C
        (L.IN_NORMAL, 1475.5503, 10.0, 0, 0, -INF, INF),   # Ck
        (L.IN_NORMAL, 55.0406, 10.0, 0, 0, -INF, INF),     # K
    ]


VEH_BINS = [(7.0, 8.5), (8.5, 9.5), (9.5, 10.5), (10.5, 11.5), (11.5, 12.0)]


def vehicle_scenarios():
    """20 scenarios in the reference's order: states sorted lexicographically per
    ancestor (A: offroad < road, b0: normal_load < over_load, V_d bins), product
    with the first ancestor varying fastest then sorted (evaluate_node.jl:55)."""
    rows, labels = [], []
    for A_name, A in (("offroad", 0.8), ("road", 0.15915)):
        for b_name, b0 in (("normal_load", 0.27), ("over_load", 0.5)):
            for lo, hi in VEH_BINS:
                rows.append(_veh_row(A, b0, lo, hi))
                labels.append((A_name, b_name, f"[{lo}, {hi}]"))
    return rows, labels


def w1_vehicle(n: int = 10**8, seed: int = 0x5EED0001, strict: bool = False, fp32_sampling: bool = False) -> L.Job:
    """W1 = BASELINE config 2: vehicle suspension, S=20, n per scenario."""
    rows, _ = vehicle_scenarios()
    return L.Job(L.LS_VEHICLE_SUSPENSION, vehicle_consts(), L.make_inputs(rows), n, seed, strict,
                 fp32_sampling=fp32_sampling)


def w5_vehicle_grid(n: int = 10**10, seed: int = 0x5EED0005, side: int = 32, strict: bool = False) -> L.Job:
    """W5 = BASELINE config 5: S = side^2 synthetic scenarios, A x b0 grid, V~U(7,12)."""
    rows = []
    for A in np.linspace(0.1, 1.0, side):
        for b0 in np.linspace(0.2, 0.6, side):
            rows.append(_veh_row(float(A), float(b0), -INF, INF))
    return L.Job(L.LS_VEHICLE_SUSPENSION, vehicle_consts(), L.make_inputs(rows), n, seed, strict)


def w4_vehicle_fixed_v(v_values, n: int = 10**6, seed: int = 0x5EED0004, strict: bool = False,
                       crn: bool = True) -> L.Job:
    """W4 = BASELINE config 4, the inner estimator of the imprecise vehicle net
    (demo/vehicle_suspension_imprecise.jl): every (A, b0) state x every outer-loop value of V, with V
    entering as a Parameter (evaluate_node.jl:104-117 -> UQ.jl DoubleLoop). `crn` gives all V values
    of one (A, b0) state the same stream (common random numbers across the outer loop)."""
    rows, streams = [], []
    k = 0
    for A in (0.8, 0.15915):
        for b0 in (0.27, 0.5):
            for v in v_values:
                row = _veh_row(A, b0, -INF, INF)
                row[2] = (L.IN_PARAM, float(v), 0, 0, 0, -INF, INF)
                rows.append(row)
                streams.append(k if crn else len(streams))
            k += 1
    return L.Job(L.LS_VEHICLE_SUSPENSION, vehicle_consts(), L.make_inputs(rows), n, seed, strict,
                 stream_id=np.asarray(streams, dtype=np.uint32))


# demo/straub_example.jl:5-24 — distribution_parameters(mean, std, Dist)
STRAUB_RHO = 0.5477


def lognormal_params(mean: float, std: float):
    zeta2 = math.log(1.0 + (std / mean) ** 2)
    return math.log(mean) - 0.5 * zeta2, math.sqrt(zeta2)


def gamma_params(mean: float, std: float):
    return (mean / std) ** 2, std * std / mean  # shape, scale


def gumbel_params(mean: float, std: float):
    beta = std * math.sqrt(6.0) / math.pi
    return mean - 0.5772156649015329 * beta, beta


def straub_consts() -> np.ndarray:
    lam, zeta = lognormal_params(150.0, 30.0)
    return np.array([lam, zeta, STRAUB_RHO])


def straub_row():
    a, th = gamma_params(60.0, 12.0)
    mu, beta = gumbel_params(50.0, 20.0)
    row = [
        (L.IN_NORMAL, 0.0, 1.0, 0, 0, -INF, INF),   # Ur
        (L.IN_GAMMA, a, th, 0, 0, -INF, INF),       # V
        (L.IN_GUMBEL, mu, beta, 0, 0, -INF, INF),   # H
    ]
    row += [(L.IN_NORMAL, 0.0, 1.0, 0, 0, -INF, INF)] * 5  # in-model randn() draws
    return row


def w2_straub(n: int = 10**6, seed: int = 0x5EED0002, strict: bool = False) -> L.Job:
    """W2 = BASELINE config 1: Straub frame, S=1."""
    return L.Job(L.LS_STRAUB_FRAME, straub_consts(), L.make_inputs([straub_row()]), n, seed, strict)


def w3_straub_copula(n: int = 10**9, seed: int = 0x5EED0003, bits: int = 6, strict: bool = False) -> L.Job:
    """W3 = BASELINE config 3: synthetic Straub-style EBN, 2^bits discrete-parent scenarios, the seven
    continuous inputs (r1..r5 ~ LogNormal(lambda, zeta), V ~ Gamma, H ~ Gumbel) joined by a Gaussian
    copula with equicorrelation rho^2 = 0.3 among the r's (which is exactly the dependence the
    common factor u_r induces in demo/straub_example.jl:16-26, so scenario 0 reproduces Straub's pf).
    Each binary parent shifts the Gumbel location (odd bits) or the Gamma mean (even bits) by 2 %.
    Limit state: the frame's three mechanisms as a min of affine forms of (r1..r5, V, H)."""
    from .distributions import Gamma, table_for
    lam, zeta = lognormal_params(150.0, 30.0)
    d = 7
    R = np.eye(d)
    R[:5, :5] = STRAUB_RHO**2
    np.fill_diagonal(R, 1.0)
    chol = np.linalg.cholesky(R)
    rows, tables, off = [], [], 0
    for s in range(1 << bits):
        mean_v, mean_h = 60.0, 50.0
        for b in range(bits):
            if (s >> b) & 1:
                if b % 2:
                    mean_h *= 1.02
                else:
                    mean_v *= 1.02
        a, th = gamma_params(mean_v, 0.2 * mean_v)
        mu, beta = gumbel_params(mean_h, 0.4 * mean_h)
        q = table_for(Gamma(a, th), -INF, INF, 8193)   # Gamma under a copula is shipped tabulated
        tables.append(q)
        row = [(L.IN_LOGNORMAL, lam, zeta, 0, 0, -INF, INF)] * 5
        row.append((L.IN_TABULATED, float(off), float(len(q)), 0, 0, -INF, INF))
        row.append((L.IN_GUMBEL, mu, beta, 0, 0, -INF, INF))
        off += len(q)
        rows.append(row)
    #            c0  r1 r2 r3 r4 r5  V   H
    k = np.array([3.0,
                  0, 1, 1, 0, 1, 1, 0, -5,      # g1 = r1 + r2 + r4 + r5 - 5h
                  0, 0, 1, 2, 1, 0, -5, 0,      # g2 = r2 + 2 r3 + r4 - 5v
                  0, 1, 0, 2, 2, 1, -5, -5.0])  # g3 = r1 + 2 r3 + 2 r4 + r5 - 5h - 5v
    return L.Job(L.LS_MIN_AFFINE, k, L.make_inputs(rows), n, seed, strict, corr_chol=chol,
                 tables=np.concatenate(tables))
