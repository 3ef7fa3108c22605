"""CPU tests: the oracle against the golden fixtures, the reference's own statistical/analytic
anchors, and the Random123 known-answer vectors for Philox4x32-10."""
import math
import os

import numpy as np
from scipy import special, stats

from oracle import cpu as ocpu
from oracle import replay

GOLD = os.path.join(os.path.dirname(__file__), "golden")
INF = math.inf


def test_philox4x32_10_known_answers():
    """Random123 kat_vectors (philox4x32 10 rounds)."""
    def kat(c, k):
        r = replay.philox4x32_10(*[np.array([x], dtype=np.uint64) for x in c], k[0], k[1])
        return [int(x[0]) for x in r]
    assert kat([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert kat([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert kat([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_golden_limit_states():
    z = np.load(os.path.join(GOLD, "limit_states.npz"))
    c, g = ocpu.eval_matrix(ocpu.LS_VEHICLE, z["vehicle_k"], z["vehicle_X"], want_g=True)
    assert c == int(z["vehicle_count"]) and np.array_equal(g, z["vehicle_g"])
    assert np.array_equal(replay.g_vehicle(z["vehicle_X"], z["vehicle_k"]), z["vehicle_g"])
    c, g = ocpu.eval_matrix(ocpu.LS_STRAUB, z["straub_k"], z["straub_X"], want_g=True)
    assert c == int(z["straub_count"]) and np.array_equal(g, z["straub_g"])
    assert np.allclose(replay.g_straub(z["straub_X"], z["straub_k"]), z["straub_g"], rtol=0, atol=1e-10)


def test_golden_expression_programs():
    """The committed programs (the demos' closures compiled from text) reproduce the golden g of the
    hand-written limit states through both stack machines, and today's text compiler still emits
    exactly these programs."""
    import sys
    sys.path.insert(0, GOLD)
    import make_golden as mg
    from ebn_b200 import expr
    z = np.load(os.path.join(GOLD, "limit_states.npz"))
    e = np.load(os.path.join(GOLD, "expression.npz"))
    c, g = ocpu.eval_matrix(ocpu.LS_EXPRESSION, e["vehicle_program"], z["vehicle_X"], want_g=True)
    assert c == int(z["vehicle_count"]) and np.array_equal(g, z["vehicle_g"])
    assert np.array_equal(replay.g_expression(z["vehicle_X"], e["vehicle_program"]), z["vehicle_g"])
    c, g = ocpu.eval_matrix(ocpu.LS_EXPRESSION, e["straub_program"], z["straub_X"], want_g=True)
    assert c == int(z["straub_count"]) and np.array_equal(g, z["straub_g"])
    k = z["vehicle_k"]
    veh = expr.compile_program(["A", "b₀", "V", "C", "Cₖ", "K"], mg.VEHICLE_TEXT, "df.y",
                               params={"M": k[0], "m": k[1], "g": k[2], "Mg15": k[3], "gMm": k[4]})
    assert np.array_equal(veh.consts, e["vehicle_program"])
    sk = z["straub_k"]
    stb = expr.compile_program(["Uᵣ", "V", "H"], [(f"r{i}", mg.CAPACITY_TEXT) for i in range(1, 6)] + [("G", mg.FRAME_TEXT)],
                               "df.G", params={"λ": sk[0], "ζ": sk[1], "ρ": sk[2]})
    assert np.array_equal(stb.consts, e["straub_program"])


def test_golden_stream_layout():
    z = np.load(os.path.join(GOLD, "stream_v4.npz"))
    inputs = z["inputs"].view(ocpu.INPUT_DTYPE).reshape(1, -1)
    for first in (0, 7, 2**33 + 5):
        X = replay.sample_matrix(inputs[0], int(z["stream"]), int(z["seed"]), first, 64, tables=z["table"])
        assert np.allclose(X, z[f"first_{first}"], rtol=1e-13, atol=0)


def test_vehicle_constants_of_the_survey():
    """4000*(M*g)^-1.5 and (g*(M+m))^0.877."""
    M, m, g = 3.2633, 0.8158, 981.0
    assert math.isclose(4000 * (M * g) ** -1.5, 0.022083656272679515, rel_tol=1e-15)
    assert math.isclose((g * (M + m)) ** 0.877, 1442.642281188081, rel_tol=1e-15)


def test_c_and_numpy_restatements_agree_bitwise():
    rng = np.random.default_rng(7)
    X = rng.normal(0.3, 1.2, (5000, 3))
    prog = np.array([2, 2, 1.0, 2, 0, 0, 1.0, 2, 2, 2,   # stage 0: x0^2 + x2^2
                     2, 0.5, 1, 3, -0.7, 0], dtype=float)  # stage 1: 0.5*col3 - 0.7
    c, g = ocpu.eval_matrix(ocpu.LS_POLYNOMIAL, prog, X, want_g=True)
    assert np.array_equal(g, 0.5 * (X[:, 0] * X[:, 0] + X[:, 2] * X[:, 2]) + (-0.7))
    assert np.array_equal(g, replay.g_polynomial(X, prog)) and c == int(np.sum(g < 0))
    k = np.array([2.0, 1.0, -1.0, 0.5, 0.25, 2.0, -0.75, 0.1, 0.2])
    c, g = ocpu.eval_matrix(ocpu.LS_MIN_AFFINE, k, X, want_g=True)
    assert np.array_equal(g, replay.g_min_affine(X, k))
    Xh = np.column_stack([rng.integers(0, 4, 1000).astype(float), rng.uniform(0, 9, (1000, 3))])
    c, g = ocpu.eval_matrix(ocpu.LS_HYDROGEN, np.zeros(0), Xh, want_g=True)
    assert np.array_equal(g, replay.g_hydrogen(Xh))
    # NaN is safe (`NaN < 0` is false) and propagates through Julia's min
    Xn = np.array([[np.nan, 1.0, 2.0], [1.0, 1.0, 1.0]])
    c, g = ocpu.eval_matrix(ocpu.LS_MIN_AFFINE, k, Xn, want_g=True)
    assert np.isnan(g[0]) and c == int(g[1] < 0)


def test_phi_inv():
    for p in (1e-300, 1e-20, 1e-9, 0.001, 0.3, 0.5, 0.9, 1 - 1e-12):
        assert math.isclose(ocpu.phi_inv(p), special.ndtri(p), rel_tol=1e-13, abs_tol=1e-15)


POLY_2_MINUS_A_PLUS_B = np.array([2, 2, 1.0, 1, 0, 1.0, 1, 1, 2, 2.0, 0, -1.0, 1, 2], dtype=float)


def test_reference_anchor_toy_limit_state():
    """test/networks/ebn/evaluate/evaluate_node.jl:96-110: C = A + B, performance 2 - C,
    A in {1, 2}, B ~ N(0,1), n = 100_000 → [0.16, 0.84, 0.5, 0.5] atol 0.02; :137 → [0.1587, 0.8413]."""
    rows = [[(0, a, 0, 0, 0, -INF, INF), (1, 0.0, 1.0, 0, 0, -INF, INF)] for a in (1.0, 2.0)]
    n = 100_000
    cnt = ocpu.pf_batch(ocpu.LS_POLYNOMIAL, POLY_2_MINUS_A_PLUS_B, ocpu.make_inputs(rows), n, seed=1)
    pf = cnt / n
    assert np.allclose([pf[0], 1 - pf[0], pf[1], 1 - pf[1]], [0.16, 0.84, 0.5, 0.5], atol=0.02)
    assert np.allclose([pf[0], 1 - pf[0]], [0.1587, 0.8413], atol=0.1)
    exact = np.array([special.ndtr(-1.0), 0.5])
    assert np.all(np.abs(pf - exact) < 4 * np.sqrt(exact * (1 - exact) / n))


def test_reference_anchor_straub():
    """demo/straub_example.jl:71 (== commented assertion test/inference/variableselimination.jl:205):
    pf ≈ [0.026129, 0.973871] atol 0.01 at n = 10^6."""
    lam, zeta = 4.991024937519615, 0.1980422004353651
    row = [(1, 0.0, 1.0, 0, 0, -INF, INF), (5, 25.0, 2.4, 0, 0, -INF, INF),
           (4, 40.99893584908611, 15.593936024673523, 0, 0, -INF, INF)] + [(1, 0.0, 1.0, 0, 0, -INF, INF)] * 5
    n = 1_000_000
    cnt = ocpu.pf_batch(ocpu.LS_STRAUB, np.array([lam, zeta, 0.5477]), ocpu.make_inputs([row]), n, seed=3)
    pf = cnt[0] / n
    assert abs(pf - 0.026129) < 0.01 and abs((1 - pf) - 0.973871) < 0.01
    assert abs(pf - 0.0258) < 0.001  # SURVEY §6 sanity value


def test_vehicle_expected_values():
    """pf of the (A=0.8, b0=0.27) V-bins and the g4-only scenarios."""
    M, m, g = 3.2633, 0.8158, 981.0
    k = np.array([M, m, g, (M * g) ** -1.5, (g * (M + m)) ** 0.877])
    bins = [(7.0, 8.5), (8.5, 9.5), (9.5, 10.5), (10.5, 11.5), (11.5, 12.0)]
    rows = [[(0, 0.8, 0, 0, 0, -INF, INF), (0, 0.27, 0, 0, 0, -INF, INF), (3, 7.0, 12.0, 0, 0, lo, hi),
             (1, 431.7221, 10.0, 0, 0, -INF, INF), (1, 1475.5503, 10.0, 0, 0, -INF, INF),
             (1, 55.0406, 10.0, 0, 0, -INF, INF)] for lo, hi in bins]
    rows.append([(0, 0.15915, 0, 0, 0, -INF, INF), (0, 0.5, 0, 0, 0, -INF, INF)] + rows[0][2:])
    n = 1_000_000
    pf = ocpu.pf_batch(ocpu.LS_VEHICLE, k, ocpu.make_inputs(rows), n, seed=5, threads=4) / n
    want = np.array([0.0023, 0.0431, 0.2151, 0.5164, 0.7373, 5.2e-4])
    assert np.all(np.abs(pf - want) < 5 * np.sqrt(want * (1 - want) / n) + 2e-4), pf
    # the g4-only probability is analytic: P(Ck < (g(M+m))^0.877)
    assert abs(pf[5] - stats.norm(1475.5503, 10).cdf(k[4])) < 1.5e-4


def test_structural_sampler_marginals():
    """Column-wise `rand(dist, n)` restatement: first moments of every marginal incl. truncation."""
    row = [(1, 1.0, 2.0, 0, 0, 0.5, 4.0), (2, 0.3, 0.25, 0, 0, 1.0, 2.0), (4, 41.0, 15.6, 0, 0, 30.0, 80.0),
           (5, 25.0, 2.4, 0, 0, -INF, INF), (6, 2.0, -1.0, 3.0, 0, 0.0, INF), (3, 7.0, 12.0, 0, 0, 8.5, 9.5)]
    exp = [stats.truncnorm((0.5 - 1) / 2, (4 - 1) / 2, 1, 2).mean(), None, None, 60.0, None, 9.0]
    n = 400_000
    # P(x_j < t) through the min_affine limit state g = x_j - t
    for j, (t, dist) in enumerate([(2.0, stats.truncnorm((0.5 - 1) / 2, (4 - 1) / 2, 1, 2)), (1.4, None),
                                   (50.0, None), (55.0, stats.gamma(25.0, scale=2.4)), (2.0, None),
                                   (9.2, stats.uniform(8.5, 1.0))]):
        k = np.zeros(2 + 6)
        k[0], k[1], k[2 + j] = 1.0, -t, 1.0
        pf = ocpu.pf_batch(ocpu.LS_MIN_AFFINE, k, ocpu.make_inputs([row]), n, seed=11)[0] / n
        if dist is not None:
            p = dist.cdf(t)
        elif j == 1:
            Fl, Fh = stats.lognorm(0.25, scale=math.exp(0.3)).cdf([1.0, 2.0])
            p = (stats.lognorm(0.25, scale=math.exp(0.3)).cdf(t) - Fl) / (Fh - Fl)
        elif j == 2:
            G = stats.gumbel_r(41.0, 15.6)
            p = (G.cdf(t) - G.cdf(30)) / (G.cdf(80) - G.cdf(30))
        else:  # 3 - Exp(2) truncated to [0, 3]: x < 2 <=> e > 1
            E = stats.expon(0, 2.0)
            p = (E.cdf(3.0) - E.cdf(1.0)) / E.cdf(3.0)
        assert abs(pf - p) < 4 * math.sqrt(p * (1 - p) / n), (j, pf, p)


def test_sobol_direction_numbers_and_scrambling():
    """The device's direction-number table (csrc/ebn_sobol_tables.inc) equals the Joe-Kuo construction and
    scipy.stats.qmc.Sobol point for point; the hash-based Owen scrambling keeps every one-dimensional
    projection of 2^m points stratified (one point per interval of length 2^-m)."""
    import re
    from scipy.stats import qmc
    txt = open(os.path.join(os.path.dirname(GOLD), "..", "enhancedbayesiannetworks.jl_b200", "csrc", "ebn_sobol_tables.inc")).read()
    rows = re.findall(r"\{([^{}]+)\}", txt[txt.index("kSobolV[16]"):])
    V = np.array([[int(x.strip().rstrip("u"), 16) for x in r.split(",")] for r in rows], dtype=np.uint64)
    assert V.shape == (16, 32) and np.array_equal(V, replay.sobol_direction_numbers())
    P = replay.sobol_points(0, 2048, 16, V)
    assert np.array_equal(P.astype(np.float64) * 2.0**-32, qmc.Sobol(d=16, scramble=False, bits=32).random(2048))
    assert np.array_equal(replay.sobol_points(1000, 48, 16, V), P[1000:1048])
    for dim, key in ((0, 1), (5, 0xDEADBEEF), (15, 12345)):
        w = replay.owen_scramble(P[:, dim], key)
        assert np.array_equal(np.bincount((w >> np.uint64(21)).astype(np.int64), minlength=2048), np.ones(2048, dtype=np.int64))
        assert not np.array_equal(w, P[:, dim])
