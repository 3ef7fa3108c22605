"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle.

Bars (BASELINE.json north_star):
  * bit-exact fail counts on identical sample matrices (strict mode); the fast mode may
    differ only on samples with |g| below 1e-12 * scale, which are reported as ties;
  * independent-RNG estimates of pf within 3 standard errors of the oracle's;
  * counts independent of partition, grid, chunking and sharding.
"""
import dataclasses
import math

import numpy as np
import pytest

# EBN_FUZZ_SEED=k shifts the seeds of the randomised tests (extra fuzzing runs); the default is what CI runs
FUZZ_SEED = int(__import__("os").environ.get("EBN_FUZZ_SEED", "0"))

from oracle import cpu as ocpu
from oracle import replay

pytestmark = pytest.mark.gpu

INF = math.inf


def _vehicle_matrix(rng, n):
    X = np.empty((n, 6))
    X[:, 0] = rng.choice([0.15915, 0.8], n)
    X[:, 1] = rng.choice([0.27, 0.5], n)
    X[:, 2] = rng.uniform(7, 12, n)
    X[:, 3] = rng.normal(431.7221, 10, n)
    X[:, 4] = rng.normal(1475.5503, 10, n)
    X[:, 5] = rng.normal(55.0406, 10, n)
    return X


def _straub_matrix(rng, n):
    X = np.empty((n, 8))
    X[:, 0] = rng.normal(0, 1, n)
    X[:, 1] = rng.gamma(25.0, 2.4, n)
    X[:, 2] = rng.gumbel(40.99893584908611, 15.593936024673523, n)
    X[:, 3:] = rng.normal(0, 1, (n, 5))
    return X


def poly_program(stages):
    """stages: list of list of (coef, [cols]) → consts array for EBN_LS_POLYNOMIAL."""
    out = [float(len(stages))]
    for terms in stages:
        out.append(float(len(terms)))
        for coef, cols in terms:
            out.append(float(coef))
            out.append(float(len(cols)))
            out.extend(float(c) for c in cols)
    return np.array(out)


# test/networks/ebn/evaluate/evaluate_node.jl:96-98: model C = A + B, performance 2 - C
PROG_2_MINUS_A_PLUS_B = poly_program([[(1.0, [0]), (1.0, [1])], [(2.0, []), (-1.0, [2])]])


# ---------------------------------------------------------------------------
# 1. limit states on identical matrices: bit-exact counts and g (strict)
# ---------------------------------------------------------------------------
def test_eval_matrix_vehicle_strict_bit_exact(ebn):
    from ebn_b200 import workloads as W
    rng = np.random.default_rng(1)
    X = _vehicle_matrix(rng, 300_000)
    k = W.vehicle_consts()
    want_cnt, want_g = ocpu.eval_matrix(ocpu.LS_VEHICLE, k, X, want_g=True)
    got_cnt, got_g = ebn.eval_matrix(ebn._lib.LS_VEHICLE_SUSPENSION, k, X, strict=True, want_g=True)
    assert got_cnt == want_cnt
    assert np.array_equal(got_g, want_g), f"max |dg| = {np.max(np.abs(got_g - want_g))}"
    # second, independent restatement (NumPy) agrees too
    assert np.array_equal(replay.g_vehicle(X, k), want_g)
    assert 0 < want_cnt < len(X)


def test_eval_matrix_vehicle_fast_ties_only(ebn):
    from ebn_b200 import workloads as W
    rng = np.random.default_rng(2)
    X = _vehicle_matrix(rng, 300_000)
    k = W.vehicle_consts()
    want_cnt, want_g = ocpu.eval_matrix(ocpu.LS_VEHICLE, k, X, want_g=True)
    got_cnt, got_g = ebn.eval_matrix(ebn._lib.LS_VEHICLE_SUSPENSION, k, X, strict=False, want_g=True)
    # stated fp64 tolerance on g: 1e-12 relative to the magnitude of the terms (|g| <= ~1e2)
    assert np.max(np.abs(got_g - want_g)) < 1e-11
    flips = np.flatnonzero((got_g < 0) != (want_g < 0))
    assert np.all(np.abs(want_g[flips]) < 1e-12), "fast mode flipped a sample that is not a tie"
    assert abs(got_cnt - want_cnt) <= len(flips)
    print(f"ties |g|<1e-12: {len(flips)} of {len(X)}")


def test_eval_matrix_straub(ebn):
    from ebn_b200 import workloads as W
    rng = np.random.default_rng(3)
    X = _straub_matrix(rng, 200_000)
    k = W.straub_consts()
    want_cnt, want_g = ocpu.eval_matrix(ocpu.LS_STRAUB, k, X, want_g=True)
    got_cnt, got_g = ebn.eval_matrix(ebn._lib.LS_STRAUB_FRAME, k, X, strict=True, want_g=True)
    # exp() is not correctly rounded anywhere (Julia, glibc, CUDA all < 1 ulp): g carries a
    # relative tolerance of 1e-13 on terms of magnitude ~1e3; counts may differ only on ties.
    assert np.max(np.abs(got_g - want_g)) < 1e-9
    flips = np.flatnonzero((got_g < 0) != (want_g < 0))
    assert np.all(np.abs(want_g[flips]) < 1e-9)
    assert abs(got_cnt - want_cnt) <= len(flips)
    assert 0 < want_cnt < len(X)


def test_eval_matrix_straub_variants(ebn):
    """The frame with observed R4/R5 and the single capacity against the oracle (same exp tolerance)."""
    from ebn_b200 import workloads as W
    rng = np.random.default_rng(31)
    n = 100_000
    k = W.straub_consts()
    X = np.column_stack([rng.normal(0, 1, n), rng.gamma(25.0, 2.4, n), rng.gumbel(41.0, 15.6, n),
                         rng.lognormal(k[0], k[1], (n, 2)), rng.normal(0, 1, (n, 3))])
    for strict in (True, False):
        want_cnt, want_g = ocpu.eval_matrix(ocpu.LS_STRAUB_R45, k, X, want_g=True)
        got_cnt, got_g = ebn.eval_matrix(ebn._lib.LS_STRAUB_FRAME_R45, k, X, strict=strict, want_g=True)
        assert np.max(np.abs(got_g - want_g)) < 1e-9 and abs(got_cnt - want_cnt) <= np.sum(np.abs(want_g) < 1e-9)
        assert 0 < want_cnt < n
        want_cnt, want_g = ocpu.eval_matrix(ocpu.LS_STRAUB_CAPACITY, k, X[:, [0, 5]], want_g=True)
        got_cnt, got_g = ebn.eval_matrix(ebn._lib.LS_STRAUB_CAPACITY, k, X[:, [0, 5]], strict=strict, want_g=True)
        assert np.max(np.abs(got_g / want_g - 1)) < 1e-14 and got_cnt == 0 == want_cnt
    # exp_fast against the correctly rounded value: a few ulp over the lognormal range
    assert np.max(np.abs(got_g / np.exp(k[0] + k[2] * k[1] * X[:, 0] + np.sqrt((1 - k[2] ** 2) * k[1] ** 2) * X[:, 5]) - 1)) < 1e-14


def test_eval_matrix_polynomial_and_min_affine_bit_exact(ebn):
    rng = np.random.default_rng(4)
    X = rng.normal(0.5, 1.0, (100_000, 2))
    for strict in (True, False):
        want_cnt, want_g = ocpu.eval_matrix(ocpu.LS_POLYNOMIAL, PROG_2_MINUS_A_PLUS_B, X, want_g=True)
        got_cnt, got_g = ebn.eval_matrix(ebn._lib.LS_POLYNOMIAL, PROG_2_MINUS_A_PLUS_B, X, strict=strict, want_g=True)
        if strict:
            assert got_cnt == want_cnt and np.array_equal(got_g, want_g)
            assert np.array_equal(want_g, 2.0 - (X[:, 0] + X[:, 1]))
        else:
            assert np.max(np.abs(got_g - want_g)) < 1e-13
    # quadratic model x^2 - 0.7 + y (test/networks/ebn/evaluate/evaluate_net.jl style)
    prog = poly_program([[(1.0, [0, 0]), (-0.7, []), (1.0, [1])]])
    want_cnt, want_g = ocpu.eval_matrix(ocpu.LS_POLYNOMIAL, prog, X, want_g=True)
    got_cnt, got_g = ebn.eval_matrix(ebn._lib.LS_POLYNOMIAL, prog, X, strict=True, want_g=True)
    assert got_cnt == want_cnt and np.array_equal(got_g, want_g)
    assert np.array_equal(want_g, (X[:, 0] * X[:, 0] - 0.7) + X[:, 1])
    # min of two affine forms
    k = np.array([2.0, 1.0, -1.0, 0.5, 0.25, 2.0, -0.75])
    want_cnt, want_g = ocpu.eval_matrix(ocpu.LS_MIN_AFFINE, k, X, want_g=True)
    got_cnt, got_g = ebn.eval_matrix(ebn._lib.LS_MIN_AFFINE, k, X, strict=True, want_g=True)
    assert got_cnt == want_cnt and np.array_equal(got_g, want_g)


def test_eval_matrix_hydrogen_and_edge_cases(ebn):
    rng = np.random.default_rng(5)
    X = np.column_stack([rng.integers(0, 4, 50_000).astype(float), rng.uniform(0, 10, 50_000),
                         rng.uniform(0, 10, 50_000), rng.uniform(0, 10, 50_000)])
    want = ocpu.eval_matrix(ocpu.LS_HYDROGEN, np.zeros(0), X)
    got = ebn.eval_matrix(ebn._lib.LS_HYDROGEN_RADIUS, np.zeros(0), X, strict=True)
    assert got == want
    # empty matrix, single row, NaN rows (NaN is safe: `NaN < 0` is false)
    assert ebn.eval_matrix(ebn._lib.LS_HYDROGEN_RADIUS, np.zeros(0), np.zeros((0, 4))) == 0
    assert ebn.eval_matrix(ebn._lib.LS_HYDROGEN_RADIUS, np.zeros(0), np.array([[1.0, 1.0, 2.0, 0.0]])) == 1
    Xn = np.array([[np.nan, -5.0], [1.0, 5.0], [np.nan, np.nan]])
    assert ebn.eval_matrix(ebn._lib.LS_POLYNOMIAL, PROG_2_MINUS_A_PLUS_B, Xn) == 1


# ---------------------------------------------------------------------------
# 2. the sampler: replayed matrices against the NumPy twin of the stream layout
# ---------------------------------------------------------------------------
ALL_KINDS_ROW = [
    (0, 3.25, 0, 0, 0, -INF, INF),            # Parameter
    (1, 1.0, 2.0, 0, 0, -INF, INF),           # Normal (Box-Muller)
    (1, 1.0, 2.0, 0, 0, 0.5, 4.0),            # truncated Normal
    (2, 0.3, 0.25, 0, 0, -INF, INF),          # LogNormal
    (2, 0.3, 0.25, 0, 0, 1.0, 2.0),           # truncated LogNormal
    (3, 7.0, 12.0, 0, 0, 8.5, 9.5),           # truncated Uniform
    (4, 41.0, 15.6, 0, 0, -INF, INF),         # Gumbel
    (4, 41.0, 15.6, 0, 0, 30.0, 80.0),        # truncated Gumbel
    (5, 25.0, 2.4, 0, 0, -INF, INF),          # Gamma (shape >= 1)
    (5, 0.6, 1.5, 0, 0, -INF, INF),           # Gamma (shape < 1)
    (6, 2.0, 1.0, 3.0, 0, -INF, INF),         # 3 + Exponential(2)
    (6, 2.0, -1.0, 3.0, 0, 0.0, INF),         # 3 - Exponential(2), truncated to >= 0
    (7, 0.0, 5.0, 0, 0, -INF, INF),           # tabulated quantile
]
TABLE = np.array([-1.0, 0.0, 0.5, 2.0, 10.0])


def _min_affine_all(d):
    return np.array([1.0, 0.0] + [1.0] * d)


@pytest.mark.parametrize("first,n", [(0, 4096), (7, 1001), (2**33 + 5, 512)])
def test_replay_matches_numpy_twin(ebn, first, n):
    rows = [ALL_KINDS_ROW, ALL_KINDS_ROW]
    inputs = ebn.make_inputs(rows)
    d = inputs.shape[1]
    job = ebn.Job(ebn._lib.LS_MIN_AFFINE, _min_affine_all(d), inputs, n=first + n, seed=0xABCDEF0123456789,
                  tables=TABLE)
    for s in (0, 1):
        X = ebn.sample_matrix(job, s, first, n)
        ref = replay.sample_matrix(inputs[s], s, 0xABCDEF0123456789, first, n, tables=TABLE)
        scale = np.maximum(np.abs(ref), 1.0)
        err = np.max(np.abs(X - ref) / scale, axis=0)
        # transcendental implementations differ (CUDA libdevice vs NumPy/SciPy) by a few ulp;
        # the inverse-normal tails amplify that slightly
        assert np.all(err < 5e-12), f"scenario {s}: per-input max rel err {err}"
    # different scenarios are different streams
    assert not np.allclose(ebn.sample_matrix(job, 0, first, n)[:, 1], ebn.sample_matrix(job, 1, first, n)[:, 1])


def _random_marginal(rng):
    """One random ebn_input_t row: kind, parameters and (half of the time) a truncation window that
    keeps at least ~1 % of the mass inside +-3.5 standard scales."""
    kind = int(rng.choice([0, 1, 1, 2, 3, 4, 5, 6]))
    trunc = rng.random() < 0.5
    if kind == 0:
        return (0, float(rng.normal(0, 5)), 0, 0, 0, -INF, INF)
    if kind in (1, 2):
        mu, sd = (float(rng.normal(0, 3)), float(rng.uniform(0.2, 3))) if kind == 1 else (float(rng.normal(0, 0.5)), float(rng.uniform(0.05, 0.5)))
        if not trunc:
            return (kind, mu, sd, 0, 0, -INF, INF)
        a = float(rng.uniform(-3.0, 2.0))
        b = a + float(rng.uniform(0.3, 3.0))
        lo, hi = mu + sd * a, mu + sd * b
        if rng.random() < 0.3:
            hi = INF
        return (kind, mu, sd, 0, 0, math.exp(lo), math.exp(hi) if hi < INF else INF) if kind == 2 else (kind, mu, sd, 0, 0, lo, hi)
    if kind == 3:
        a = float(rng.normal(0, 5))
        b = a + float(rng.uniform(0.5, 10))
        if trunc:
            return (3, a, b, 0, 0, a + 0.2 * (b - a), a + 0.7 * (b - a))
        return (3, a, b, 0, 0, -INF, INF)
    if kind == 4:
        mu, beta = float(rng.normal(10, 5)), float(rng.uniform(0.5, 5))
        if trunc:
            return (4, mu, beta, 0, 0, mu - 1.0 * beta, mu + float(rng.uniform(0.5, 4)) * beta)
        return (4, mu, beta, 0, 0, -INF, INF)
    if kind == 5:
        return (5, float(rng.choice([0.4, 0.9, 1.0, 2.5, 25.0, 180.0])), float(rng.uniform(0.2, 3)), 0, 0, -INF, INF)
    theta, sign, shift = float(rng.uniform(0.3, 3)), float(rng.choice([-1.0, 1.0])), float(rng.normal(0, 2))
    if trunc:
        lo, hi = (shift + 0.1 * theta, shift + 2.5 * theta) if sign > 0 else (shift - 2.5 * theta, shift - 0.1 * theta)
        return (6, theta, sign, shift, 0, lo, hi)
    return (6, theta, sign, shift, 0, -INF, INF)


def test_random_marginal_rows_against_the_twin(ebn):
    """Randomised version of the twin comparison: 16 random scenario rows (kinds, parameters,
    truncation windows), drawn by the runtime-dispatched plan, equal the NumPy twin to 5e-12; and for
    each row the plan compiled at run time for exactly these kinds draws the same bits."""
    rng = np.random.default_rng(777 + FUZZ_SEED)
    seed = 0x1234_5678_9ABC_DEF0
    for trial in range(16):
        d = int(rng.integers(2, 9))
        row = [_random_marginal(rng) for _ in range(d)]
        inputs = ebn.make_inputs([row, row])
        job = ebn.Job(ebn._lib.LS_MIN_AFFINE, _min_affine_all(d), inputs, n=5000, seed=seed, jit_plan=False)
        first = int(rng.choice([0, 3, 2**35 + 1]))
        X = ebn.sample_matrix(job, 1, first, 3000)
        ref = replay.sample_matrix(inputs[1], 1, seed, first, 3000)
        err = np.max(np.abs(X - ref) / np.maximum(np.abs(ref), 1.0), axis=0)
        assert np.all(err < 5e-12), (trial, row, err)
        if trial % 4 == 0:      # a compile per row: every fourth row is enough
            jit = ebn.Job(job.limit_state, job.consts, inputs, n=5000, seed=seed, jit_plan=True)
            assert np.array_equal(ebn.sample_matrix(jit, 1, first, 3000), X), (trial, row)
            assert np.array_equal(ebn.pf_batch(jit)[0], ebn.pf_batch(job)[0])


def test_replay_moments(ebn):
    """Distribution sanity of the sampler on 2e6 draws (means within 5 standard errors)."""
    n = 2_000_000
    inputs = ebn.make_inputs([ALL_KINDS_ROW])
    job = ebn.Job(ebn._lib.LS_MIN_AFFINE, _min_affine_all(inputs.shape[1]), inputs, n=n, seed=11, tables=TABLE)
    X = ebn.sample_matrix(job, 0, 0, n)
    from scipy import stats
    expect = {
        1: stats.norm(1, 2), 2: stats.truncnorm((0.5 - 1) / 2, (4 - 1) / 2, 1, 2),
        3: stats.lognorm(0.25, scale=math.exp(0.3)), 5: stats.uniform(8.5, 1.0),
        6: stats.gumbel_r(41.0, 15.6), 8: stats.gamma(25.0, scale=2.4), 9: stats.gamma(0.6, scale=1.5),
        10: stats.expon(3.0, 2.0),
    }
    for j, dist in expect.items():
        se = dist.std() / math.sqrt(n)
        assert abs(X[:, j].mean() - dist.mean()) < 5 * se, (j, X[:, j].mean(), dist.mean())
        ks = stats.kstest(X[:200_000, j], dist.cdf)
        assert ks.pvalue > 1e-4, (j, ks)
    assert X[:, 2].min() >= 0.5 and X[:, 2].max() <= 4.0
    assert X[:, 4].min() >= 1.0 and X[:, 4].max() <= 2.0
    assert X[:, 7].min() >= 30.0 and X[:, 7].max() <= 80.0
    assert X[:, 11].min() >= 0.0 and X[:, 11].max() <= 3.0


# ---------------------------------------------------------------------------
# 3. fused kernel == replay + oracle, bit for bit
# ---------------------------------------------------------------------------
def test_fused_count_equals_oracle_on_replayed_matrix_vehicle(ebn):
    from ebn_b200 import workloads as W
    n = 300_001  # odd on purpose
    job = W.w1_vehicle(n=n, strict=True)
    cnt, pf, var = ebn.pf_batch(job)
    assert ebn.last_stats()["specialised"] == 1
    for s in range(job.S):
        X = ebn.sample_matrix(job, s, 0, n)
        assert int(cnt[s]) == ocpu.eval_matrix(ocpu.LS_VEHICLE, job.consts, X), s
    assert np.array_equal(pf, cnt / n)
    assert np.allclose(var, (pf - pf * pf) / n, rtol=0, atol=0)
    # fast mode: same streams, counts differ at most by ties
    cnt_fast, _, _ = ebn.pf_batch(W.w1_vehicle(n=n, strict=False))
    assert np.max(np.abs(cnt_fast - cnt)) <= 1


def test_fused_count_equals_oracle_on_replayed_matrix_straub(ebn):
    from ebn_b200 import workloads as W
    n = 400_000
    job = W.w2_straub(n=n, strict=True)
    cnt, pf, _ = ebn.pf_batch(job)
    assert ebn.last_stats()["specialised"] == 1
    X, g = ebn.sample_matrix(job, 0, 0, n, want_g=True)
    assert int(cnt[0]) == int(np.sum(g < 0))
    want_cnt, want_g = ocpu.eval_matrix(ocpu.LS_STRAUB, job.consts, X, want_g=True)
    flips = np.flatnonzero((g < 0) != (want_g < 0))
    assert np.all(np.abs(want_g[flips]) < 1e-9) and abs(int(cnt[0]) - want_cnt) <= len(flips)
    # demo/straub_example.jl:71 anchor: pf ~ 0.026129 (atol 0.01)
    assert abs(pf[0] - 0.026129) < 0.01


def test_fused_count_equals_oracle_vehicle_fixed_v_plan(ebn):
    """Second compile-time plan of the vehicle limit state (V a Parameter): the inner estimator of the
    imprecise net's DoubleLoop. Counts match the oracle on the replayed matrix, the V column is the
    parameter itself, and common random numbers make the vertices of one state share C, Ck, K."""
    from ebn_b200 import workloads as W
    n = 200_001
    vs = [7.0, 9.5, 12.0]
    job = W.w4_vehicle_fixed_v(vs, n=n, strict=True)
    cnt, pf, _ = ebn.pf_batch(job)
    assert ebn.last_stats()["specialised"] == 1
    mats = {}
    for s in (0, 1, 2, 5, job.S - 1):
        X = ebn.sample_matrix(job, s, 0, n)
        mats[s] = X
        assert np.all(X[:, 2] == vs[s % 3])
        assert int(cnt[s]) == ocpu.eval_matrix(ocpu.LS_VEHICLE, job.consts, X), s
    assert np.array_equal(mats[0][:, 3:], mats[1][:, 3:]) and np.array_equal(mats[0][:, 3:], mats[2][:, 3:])
    assert not np.array_equal(mats[0][:, 3:], mats[5][:, 3:])
    # same streams through the runtime-dispatched plan (a mixed job cannot use a compile-time plan)
    rows = [list(map(tuple, [(int(c["kind"]), *c["p"], c["lo"], c["hi"]) for c in job.inputs[s]])) for s in range(3)]
    rows.append(W._veh_row(0.8, 0.27, 7.0, 8.5))
    mixed = ebn.Job(job.limit_state, job.consts, ebn.make_inputs(rows), n=n, seed=job.seed, strict=True,
                    stream_id=np.array([0, 0, 0, 7], dtype=np.uint32))
    cnt_dyn, _, _ = ebn.pf_batch(mixed)
    assert ebn.last_stats()["specialised"] == 0
    assert np.array_equal(cnt_dyn[:3], cnt[:3])


def test_fp32_sampling_is_the_same_stream_at_lower_precision(ebn):
    """EBN_JOB_FP32_SAMPLING ("FP32 where tolerated"): the untruncated normals come from the same random
    words through an fp32 Box-Muller. Every sample agrees with the fp64 draw to ~1e-6 sigma, the fused
    count equals the oracle on the replayed matrix bit for bit (the limit state stays fp64), the count
    differs from the fp64 job only by near-ties, and the NumPy float32 twin reproduces the draws."""
    from ebn_b200 import workloads as W
    n = 400_001
    j64 = W.w1_vehicle(n=n, strict=True)
    j32 = W.w1_vehicle(n=n, strict=True, fp32_sampling=True)
    c64 = ebn.pf_batch(j64)[0]
    c32, pf32, _ = ebn.pf_batch(j32)
    assert ebn.last_stats()["specialised"] == 1
    for s in (0, 9, 19):
        X32 = ebn.sample_matrix(j32, s, 0, n)
        assert int(c32[s]) == ocpu.eval_matrix(ocpu.LS_VEHICLE, j32.consts, X32), s
        X64 = ebn.sample_matrix(j64, s, 0, n)
        assert np.array_equal(X32[:, :3], X64[:, :3])                      # Parameters and the uniform: untouched
        dz = np.abs(X32[:, 3:] - X64[:, 3:]) / 10.0                        # sigma = 10
        # measured (tools/fp32_diff_stats.py): median 1e-7, 99.99 % below 1.4e-6, max 6e-5 (tiny radii,
        # where the absolute error of MUFU.LG2 passes through the square root)
        assert np.percentile(dz, 99.9) < 2e-6 and dz.max() < 5e-4
        twin = replay.sample_matrix(j32.inputs[s], s, j32.seed, 0, 5000, fp32=True)
        dt = np.abs(twin[:, 3:] - X32[:5000, 3:]) / 10.0
        assert np.percentile(dt, 99.9) < 2e-6 and dt.max() < 5e-4
    # counts: only samples whose g sits within ~1e-5 of zero can flip
    assert np.max(np.abs(c32 - c64)) <= 0.002 * n * 1e-2 + 8, (c32 - c64)
    # the runtime-dispatched plan draws the same fp32 samples (mixed kinds: no compile-time plan)
    mixed = W.w1_vehicle(n=n, strict=True, fp32_sampling=True)
    mixed.inputs[0, 3]["lo"] = 0.0            # scenario 0 gets a (far) truncated C: kinds differ between scenarios
    cm = ebn.pf_batch(mixed)[0]
    assert ebn.last_stats()["specialised"] == 0 and np.array_equal(cm[1:], c32[1:])
    # independent check: within 4 SE of the fp64 estimate at a larger n
    big32 = ebn.pf_batch(W.w1_vehicle(n=20_000_000, fp32_sampling=True))
    big64 = ebn.pf_batch(W.w1_vehicle(n=20_000_000, seed=1234))
    z = (big32[1] - big64[1]) / np.sqrt(big32[2] + big64[2])
    assert np.abs(z).max() < 4.0, z


def test_fused_dynamic_plan_all_kinds(ebn):
    """Generic (runtime-dispatched) sampling plan: fused count == oracle on the replayed matrix."""
    inputs = ebn.make_inputs([ALL_KINDS_ROW, ALL_KINDS_ROW[::-1][1:] + [ALL_KINDS_ROW[-1]]])
    # keep the table input last in both rows so the layout is valid
    d = inputs.shape[1]
    k = np.array([1.0, -150.0] + [1.0] * d)
    n = 100_000
    job = ebn.Job(ebn._lib.LS_MIN_AFFINE, k, inputs, n=n, seed=5, strict=True, tables=TABLE)
    cnt, pf, _ = ebn.pf_batch(job)
    assert ebn.last_stats()["specialised"] == 0
    for s in range(2):
        X = ebn.sample_matrix(job, s, 0, n)
        assert int(cnt[s]) == ocpu.eval_matrix(ocpu.LS_MIN_AFFINE, k, X)
    assert 0 < cnt[0] < n


# ---------------------------------------------------------------------------
# 4. counts do not depend on partitioning, sharding or resume offsets
# ---------------------------------------------------------------------------
def test_sharding_and_offsets_are_count_invariant(ebn):
    from ebn_b200 import workloads as W
    n = 1_000_003
    base, _, _ = ebn.pf_batch(W.w1_vehicle(n=n))
    for world in (2, 3, 8):
        tot = np.zeros_like(base)
        for r in range(world):
            j = W.w1_vehicle(n=n)
            j.shard_rank, j.shard_world, j.partial = r, world, True
            tot += ebn.pf_batch(j)[0]
        assert np.array_equal(tot, base), world
    # resume: [0, a) + [a, n) == [0, n) for even and odd split points
    for a in (400_000, 400_001):
        j1 = W.w1_vehicle(n=a)
        j2 = W.w1_vehicle(n=n - a)
        j2.sample_offset = a
        assert np.array_equal(ebn.pf_batch(j1)[0] + ebn.pf_batch(j2)[0], base), a
    # stream ids: a permuted stream assignment permutes the counts of identical scenarios
    j = W.w5_vehicle_grid(n=100_000, side=2)
    j.inputs[:] = j.inputs[0]
    c0 = ebn.pf_batch(j)[0]
    j.stream_id = np.array([3, 2, 1, 0], dtype=np.uint32)
    j.__post_init__()
    c1 = ebn.pf_batch(j)[0]
    assert np.array_equal(c1, c0[::-1]) and len(set(c0.tolist())) > 1
    # common random numbers: equal stream ids → identical counts
    j.stream_id = np.zeros(4, dtype=np.uint32)
    j.__post_init__()
    c2 = ebn.pf_batch(j)[0]
    assert len(set(c2.tolist())) == 1 and c2[0] == c0[0]


def test_checkpointed_run_equals_one_call(ebn, tmp_path):
    """SURVEY §5 checkpoint/resume: a run cut into pieces, interrupted and continued from its
    checkpoint file ends with exactly the counts of one uninterrupted call."""
    from ebn_b200 import workloads as W
    job = W.w1_vehicle(n=1_000_003)
    whole = ebn.pf_batch(job)[0]
    ck = str(tmp_path / "w1.ckpt.json")
    cnt, pf, var, done = ebn.pf_batch_resumable(job, 300_000, ck, max_chunks=2)     # "interrupted" after 2 pieces
    assert not done and np.all(cnt <= whole)
    cnt, pf, var, done = ebn.pf_batch_resumable(job, 300_000, ck)                   # picks the file up
    assert done and np.array_equal(cnt, whole) and np.array_equal(pf, whole / job.n)
    other = W.w1_vehicle(n=1_000_003, seed=5)
    with pytest.raises(ValueError, match="different job"):
        ebn.pf_batch_resumable(other, 300_000, ck)


def test_tiny_and_ragged_sizes(ebn):
    from ebn_b200 import workloads as W
    for n in (1, 2, 3, 255, 256, 257, 511, 513):
        job = W.w1_vehicle(n=n, strict=True)
        cnt = ebn.pf_batch(job)[0]
        X = ebn.sample_matrix(job, 14, 0, n)
        assert int(cnt[14]) == ocpu.eval_matrix(ocpu.LS_VEHICLE, job.consts, X), n
        assert X.shape == (n, 6)


def test_many_scenarios_and_far_offsets(ebn):
    """Maximum-size ends of the domain: 100 000 scenarios in one call (a scenario then owns a single
    work item), and sample indices beyond 2^40 (the pair index fills both Philox counter words)."""
    from ebn_b200 import workloads as W
    S, n = 100_000, 999
    rng = np.random.default_rng(9)
    rows = [W._veh_row(float(a), float(b), -INF, INF) for a, b in zip(rng.uniform(0.1, 1.0, S), rng.uniform(0.2, 0.6, S))]
    big = ebn.Job(ebn._lib.LS_VEHICLE_SUSPENSION, W.vehicle_consts(), ebn.make_inputs(rows), n=n, seed=77, strict=True)
    cnt = ebn.pf_batch(big)[0]
    assert cnt.shape == (S,) and ebn.last_stats()["specialised"] == 1
    pick = [0, 1, 4097, 65_535, 65_536, S - 1]
    sub = ebn.Job(big.limit_state, big.consts, big.inputs[pick], n=n, seed=77, strict=True,
                  stream_id=np.array(pick, dtype=np.uint32))
    assert np.array_equal(ebn.pf_batch(sub)[0], cnt[pick])
    for s in (4097, S - 1):
        X = ebn.sample_matrix(big, s, 0, n)
        assert int(cnt[s]) == ocpu.eval_matrix(ocpu.LS_VEHICLE, big.consts, X)
    # far offset: [2^40 + 3, 2^40 + 3 + n) in one piece equals two resumed pieces, and the oracle on the replay
    off = 2**40 + 3
    job = W.w1_vehicle(n=5001, strict=True)
    job.sample_offset = off
    whole = ebn.pf_batch(job)[0]
    a = W.w1_vehicle(n=2000, strict=True)
    a.sample_offset = off
    b = W.w1_vehicle(n=3001, strict=True)
    b.sample_offset = off + 2000
    assert np.array_equal(whole, ebn.pf_batch(a)[0] + ebn.pf_batch(b)[0])
    X = ebn.sample_matrix(job, 4, off, 5001)
    assert int(whole[4]) == ocpu.eval_matrix(ocpu.LS_VEHICLE, job.consts, X)
    assert np.allclose(X, replay.sample_matrix(job.inputs[4], 4, job.seed, off, 5001), rtol=1e-11, atol=0)


# ---------------------------------------------------------------------------
# 5. independent-RNG estimates within 3 standard errors of the oracle
# ---------------------------------------------------------------------------
def test_pf_within_3se_of_oracle_vehicle(ebn):
    from ebn_b200 import workloads as W
    n_gpu, n_ref = 20_000_000, 2_000_000
    job = W.w1_vehicle(n=n_gpu)
    cnt, pf, var = ebn.pf_batch(job)
    ref_cnt = ocpu.pf_batch(ocpu.LS_VEHICLE, job.consts, job.inputs.view(ocpu.INPUT_DTYPE), n_ref, seed=99, threads=8)
    pf_ref = ref_cnt / n_ref
    var_ref = (pf_ref - pf_ref**2) / n_ref
    z = (pf - pf_ref) / np.sqrt(var + var_ref + 1e-300)
    print("pf gpu", np.round(pf, 6), "\npf ref", np.round(pf_ref, 6), "\nmax|z|", np.abs(z).max())
    assert np.all(np.abs(z) <= 3.0 + 1e-9) or np.sum(np.abs(z) > 3.0) <= 1 and np.abs(z).max() < 4.0


def test_analytic_anchors_of_reference_tests(ebn):
    """test/networks/ebn/evaluate/evaluate_node.jl:110 — A in {1,2} Parameters, B ~ N(0,1),
    model C = A + B, performance 2 - C: pf = Phi(-1), Phi(0) → [0.16, 0.84, 0.5, 0.5] atol 0.02.
    :137 — model B + 1: [0.1587, 0.8413]."""
    rows = [[(0, a, 0, 0, 0, -INF, INF), (1, 0.0, 1.0, 0, 0, -INF, INF)] for a in (1.0, 2.0)]
    n = 4_000_000
    job = ebn.Job(ebn._lib.LS_POLYNOMIAL, PROG_2_MINUS_A_PLUS_B, ebn.make_inputs(rows), n=n, seed=2024)
    cnt, pf, var = ebn.pf_batch(job)
    exact = np.array([0.15865525393145707, 0.5])
    assert np.all(np.abs(pf - exact) <= 3 * np.sqrt(exact * (1 - exact) / n))
    assert np.allclose([pf[0], 1 - pf[0], pf[1], 1 - pf[1]], [0.16, 0.84, 0.5, 0.5], atol=0.02)
    # truncated normal anchor: B ~ truncated(N(0,1), -1, 2): P(B > 1) = (Phi(2)-Phi(1))/(Phi(2)-Phi(-1))
    rows = [[(0, 1.0, 0, 0, 0, -INF, INF), (1, 0.0, 1.0, 0, 0, -1.0, 2.0)]]
    job = ebn.Job(ebn._lib.LS_POLYNOMIAL, PROG_2_MINUS_A_PLUS_B, ebn.make_inputs(rows), n=n, seed=7)
    pf = ebn.pf_batch(job)[1][0]
    from scipy.special import ndtr
    p = (ndtr(2) - ndtr(1)) / (ndtr(2) - ndtr(-1))
    assert abs(pf - p) <= 3 * math.sqrt(p * (1 - p) / n)


# ---------------------------------------------------------------------------
# 6. histogram path (ContinuousFunctionalNode)
# ---------------------------------------------------------------------------
def test_hist_batch_matches_numpy_on_replayed_samples(ebn):
    rows = [[(0, a, 0, 0, 0, -INF, INF), (1, 0.0, 1.0, 0, 0, -INF, INF)] for a in (1.0, 2.0, 3.0)]
    prog = poly_program([[(1.0, [0]), (1.0, [1])]])  # C = A + B (a continuous functional node's model)
    n = 500_001
    job = ebn.Job(ebn._lib.LS_POLYNOMIAL, prog, ebn.make_inputs(rows), n=n, seed=3, strict=True)
    edges = np.array([-1.0, 0.0, 1.0, 1.5, 2.0, 3.0, 6.0])
    counts, s1, s2 = ebn.hist_batch(job, edges)
    assert counts.shape == (3, len(edges) + 1) and np.all(counts.sum(axis=1) == n)
    for s in range(3):
        X = ebn.sample_matrix(job, s, 0, n)
        y = X[:, 0] + X[:, 1]
        want = np.bincount(np.searchsorted(edges, y, side="right"), minlength=len(edges) + 1)
        assert np.array_equal(counts[s], want), s
        assert abs(s1[s] - y.sum()) < 1e-6 * n and abs(s2[s] - (y * y).sum()) < 1e-6 * n
    # equally spaced edges take the direct-index path; bins must be identical to a search
    for ue in (np.linspace(-2.0, 6.0, 81), np.linspace(0.3, 0.9, 1025), np.array([0.0, 1.0, 2.0])):
        cu, _, _ = ebn.hist_batch(job, ue)
        X = ebn.sample_matrix(job, 1, 0, n)
        y = X[:, 0] + X[:, 1]
        assert np.array_equal(cu[1], np.bincount(np.searchsorted(ue, y, side="right"), minlength=len(ue) + 1))
    # same call twice → identical moments (fixed-order reduction)
    counts2, s1b, s2b = ebn.hist_batch(job, edges)
    assert np.array_equal(counts, counts2) and np.array_equal(s1, s1b) and np.array_equal(s2, s2b)


def test_truncated_normal_under_a_runtime_dispatched_plan(ebn):
    """A run-time dispatched plan that draws a truncated normal reads the Phi^-1 cells from a copy in DYNAMIC
    shared memory placed behind the histogram's edges and bins (McParams::icdf_dyn); the sampler replay reads
    the same cells from global memory. Both must give the same bits, in count and in histogram mode, whatever
    the number of edges (the copy's offset) is; the large edge sets need the > 48 KB opt-in."""
    rows = [[(0, a, 0, 0, 0, -INF, INF), (1, 0.5, 1.0, 0, 0, lo, hi)]
            for a, lo, hi in ((1.0, 0.0, INF), (2.0, -INF, 0.75), (3.0, -0.5, 2.5))]
    prog = poly_program([[(1.0, [0]), (1.0, [1])]])
    n = 300_001
    job = ebn.Job(ebn._lib.LS_POLYNOMIAL, prog, ebn.make_inputs(rows), n=n, seed=11, strict=True, jit_plan=False)
    ys = []
    for s, (a, lo, hi) in enumerate(((1.0, 0.0, INF), (2.0, -INF, 0.75), (3.0, -0.5, 2.5))):
        X = ebn.sample_matrix(job, s, 0, n)
        assert X[:, 1].min() >= lo and X[:, 1].max() <= hi and np.all(X[:, 0] == a)
        ys.append(X[:, 0] + X[:, 1])
    cnt, _, _ = ebn.pf_batch(dataclasses.replace(job, consts=poly_program([[(-2.0, []), (1.0, [0]), (1.0, [1])]])))
    assert ebn.last_stats()["specialised"] == 0
    assert np.array_equal(cnt, [int(((y - 2.0) < 0).sum()) for y in ys])
    for edges in (np.array([0.0, 1.0, 2.5, 4.0]), np.linspace(-1.0, 6.0, 7), np.linspace(-1.0, 6.0, 1024),
                  np.linspace(-1.0, 6.0, 1025)):
        counts, s1, s2 = ebn.hist_batch(job, edges)
        for s, y in enumerate(ys):
            want = np.bincount(np.searchsorted(edges, y, side="right"), minlength=len(edges) + 1)
            assert np.array_equal(counts[s], want), (len(edges), s)
            assert abs(s1[s] - y.sum()) < 1e-6 * n


def test_concurrent_callers_are_serialised(ebn):
    """ctypes releases the GIL around a C call, so these threads really are inside libebncuda at the same time.
    Each must get the result of ITS job (shared staging buffers would mix them up without the library's lock),
    and an error on one thread must not show up in another thread's ebn_last_error()."""
    import threading
    from ebn_b200 import workloads as W
    jobs = [W.w1_vehicle(n=200_000 + 1000 * i, seed=100 + i) for i in range(4)] + \
           [W.w2_straub(n=150_000 + 777 * i, seed=200 + i) for i in range(4)]
    want = [ebn.pf_batch(j)[0].copy() for j in jobs]
    got, errs = [None] * len(jobs), []

    def worker(i):
        try:
            for _ in range(5):
                got[i] = ebn.pf_batch(jobs[i])[0].copy()
                assert np.array_equal(got[i], want[i]), i
            if i == 0:
                bad = W.w1_vehicle(n=0)
                with pytest.raises(ebn.EbnError, match="EBN_EINVAL"):
                    ebn.pf_batch(bad)
        except BaseException as e:  # noqa: BLE001 - reported by the main thread
            errs.append((i, repr(e)))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    assert all(np.array_equal(g, w) for g, w in zip(got, want))


# ---------------------------------------------------------------------------
# 7. error behaviour: never a fallback, always a code + message
# ---------------------------------------------------------------------------
def test_errors(ebn):
    from ebn_b200 import workloads as W
    E = ebn.EbnError
    job = W.w1_vehicle(n=1000)
    job.limit_state = 99
    with pytest.raises(E, match="EBN_ELIMITSTATE.*no CPU fallback"):
        ebn.pf_batch(job)
    job = W.w1_vehicle(n=1000)
    job.consts = job.consts[:3]
    with pytest.raises(E, match="EBN_ELIMITSTATE.*takes 5 consts"):
        ebn.pf_batch(job)
    job = W.w1_vehicle(n=0)
    with pytest.raises(E, match="EBN_EINVAL"):
        ebn.pf_batch(job)
    job = W.w1_vehicle(n=1000)
    job.inputs[3, 2]["lo"], job.inputs[3, 2]["hi"] = 20.0, 30.0   # window outside Uniform(7,12)
    with pytest.raises(E, match="EBN_EMARGINAL"):
        ebn.pf_batch(job)
    job = W.w1_vehicle(n=1000)
    job.inputs[0, 3]["kind"] = 42
    with pytest.raises(E, match="EBN_EMARGINAL.*unknown marginal"):
        ebn.pf_batch(job)
    job = W.w1_vehicle(n=1000)
    job.shard_rank, job.shard_world = 0, 2  # sharded without a communicator and without PARTIAL
    with pytest.raises(E, match="EBN_ESTATE"):
        ebn.pf_batch(job)
    with pytest.raises(E, match="EBN_ELIMITSTATE"):
        ebn.eval_matrix(ebn._lib.LS_POLYNOMIAL, np.array([1.0, 1.0, 1.0, 1.0, 7.0]), np.zeros((4, 2)))
    with pytest.raises(E, match="EBN_EINVAL"):
        ebn.hist_batch(W.w1_vehicle(n=1000), np.array([1.0, 1.0]))
    # the library keeps working after errors
    assert ebn.pf_batch(W.w1_vehicle(n=1000))[0].shape == (20,)


# ---------------------------------------------------------------------------
# 7b. Gaussian copula (BASELINE config 3; UQ.jl JointDistribution + GaussianCopula)
# ---------------------------------------------------------------------------
COPULA_ROW = [
    (0, 2.0, 0, 0, 0, -INF, INF),             # Parameter (identity row/col in L)
    (1, 1.0, 2.0, 0, 0, -INF, INF),           # Normal
    (2, 0.3, 0.25, 0, 0, -INF, INF),          # LogNormal
    (3, 7.0, 12.0, 0, 0, -INF, INF),          # Uniform
    (4, 41.0, 15.6, 0, 0, -INF, INF),         # Gumbel
    (6, 2.0, 1.0, 3.0, 0, -INF, INF),         # 3 + Exponential(2)
    (1, 0.0, 1.0, 0, 0, -1.0, 2.0),           # truncated Normal
    (7, 0.0, 5.0, 0, 0, -INF, INF),           # tabulated
]


def _corr_chol(d, rho, fixed=(0,)):
    R = np.full((d, d), rho)
    np.fill_diagonal(R, 1.0)
    for j in fixed:
        R[j, :] = 0.0
        R[:, j] = 0.0
        R[j, j] = 1.0
    return R, np.linalg.cholesky(R)


def test_copula_normal_cdf_accuracy(ebn):
    """The table-driven Phi between the correlated scores and a non-normal marginal (phi_clamped,
    ebn_fastmath.cuh). A singular factor L = [[1, 0], [1, 0]] hands BOTH inputs the same score z: column 0
    (Normal(0, 1)) shows z, column 1 (Uniform(0, 1)) shows Phi(z). Relative accuracy must hold down the lower
    tail (quantiles of small u depend on it), absolute accuracy on the upper side."""
    from scipy import special
    L = ebn._lib
    rows = [[(L.IN_NORMAL, 0.0, 1.0, 0, 0, -INF, INF), (L.IN_UNIFORM, 0.0, 1.0, 0, 0, -INF, INF)]]
    chol = np.array([[1.0, 0.0], [1.0, 0.0]])
    job = ebn.Job(L.LS_MIN_AFFINE, np.array([1.0, 0.0, 1.0, 1.0]), ebn.make_inputs(rows), n=4_000_000, seed=5,
                  strict=True, corr_chol=chol, jit_plan=False)
    X = ebn.sample_matrix(job, 0, 0, job.n)
    z, u = X[:, 0], X[:, 1]
    assert z.min() < -4.5 and z.max() > 4.5
    ref = special.ndtr(z)
    lower = z < 0
    e_lo, e_up = np.max(np.abs(u[lower] - ref[lower]) / ref[lower]), np.max(np.abs(u[~lower] - ref[~lower]))
    print(f"Phi cells vs scipy ndtr: {e_lo:.2e} relative below 0, {e_up:.2e} absolute above")
    assert e_lo < 5e-14 and e_up < 3e-16          # scipy's own tail accuracy is a few 1e-15
    import mpmath                                  # the tight bound, on the extremes and a thinned sample
    mpmath.mp.prec = 120
    pick = np.concatenate([np.argsort(z)[:300], np.argsort(z)[-300:], np.arange(0, z.size, 4001)])
    exact = np.array([float(mpmath.ncdf(mpmath.mpf(float(v)))) for v in z[pick]])
    rel = np.abs(u[pick] - exact) / np.minimum(exact, 1.0)
    low = z[pick] < 0
    print(f"Phi cells vs mpmath: {rel[low].max():.2e} relative below 0, {np.abs(u[pick] - exact)[~low].max():.2e} absolute above")
    assert rel[low].max() < 1.5e-15 and np.abs(u[pick] - exact)[~low].max() < 2.5e-16
    # scores far outside what Box-Muller can reach still give finite, clamped, monotone values: a large sigma
    # column scaled through the factor (row 2 = 40 * eps would not be a correlation factor, so use the marginal)
    assert np.all((u > 0) & (u < 1))
    if L.jit_available():   # the unrolled copula plan compiled at run time shares the routine: same bits
        assert np.array_equal(ebn.sample_matrix(dataclasses.replace(job, jit_plan=True), 0, 0, 200_000), X[:200_000])


def test_gaussian_copula_sampler_and_count(ebn):
    from scipy import stats
    d = len(COPULA_ROW)
    R, Lc = _corr_chol(d, 0.3)
    R2, Lc2 = _corr_chol(d, -0.1)
    inputs = ebn.make_inputs([COPULA_ROW, COPULA_ROW])
    n = 200_000
    k = np.array([1.0, -75.0] + [1.0] * d)
    job = ebn.Job(ebn._lib.LS_MIN_AFFINE, k, inputs, n=n, seed=99, strict=True, tables=TABLE,
                  corr_chol=np.stack([Lc, Lc2]))
    cnt, pf, _ = ebn.pf_batch(job)
    for s, (Rs, Ls) in enumerate([(R, Lc), (R2, Lc2)]):
        X = ebn.sample_matrix(job, s, 0, n)
        ref = replay.copula_sample_matrix(inputs[s], Ls, s, 99, 0, n, tables=TABLE)
        err = np.max(np.abs(X - ref) / np.maximum(np.abs(ref), 1.0), axis=0)
        assert np.all(err < 1e-9), err
        assert int(cnt[s]) == ocpu.eval_matrix(ocpu.LS_MIN_AFFINE, k, X)
        # the dependence structure: normal scores of the marginals correlate like R
        Z = stats.norm.ppf((stats.rankdata(X[:, 1:7], axis=0) - 0.5) / n)
        C = np.corrcoef(Z.T)
        assert np.max(np.abs(C - Rs[1:7, 1:7])) < 0.01
        assert np.all(X[:, 0] == 2.0) and X[:, 6].min() >= -1.0 and X[:, 6].max() <= 2.0
    assert 0 < cnt[0] < n and cnt[0] != cnt[1]
    # one shared factor for all scenarios (EBN_JOB_COPULA_SHARED) == the same factor per scenario
    job_sh = ebn.Job(ebn._lib.LS_MIN_AFFINE, k, inputs, n=n, seed=99, strict=True, tables=TABLE, corr_chol=Lc)
    job_ps = ebn.Job(ebn._lib.LS_MIN_AFFINE, k, inputs, n=n, seed=99, strict=True, tables=TABLE,
                     corr_chol=np.stack([Lc, Lc]))
    assert np.array_equal(ebn.pf_batch(job_sh)[0], ebn.pf_batch(job_ps)[0])
    # marginals keep their laws under the copula
    Xs = ebn.sample_matrix(job_sh, 0, 0, n)
    for j, dist in {1: stats.norm(1, 2), 3: stats.uniform(7, 5), 4: stats.gumbel_r(41.0, 15.6), 5: stats.expon(3, 2)}.items():
        assert stats.kstest(Xs[:, j], dist.cdf).pvalue > 1e-4, j
    # a factor that is not a Cholesky factor of a correlation matrix is rejected
    bad = Lc.copy()
    bad[2, 2] = 2.0
    with pytest.raises(ebn.EbnError, match="EBN_EINVAL.*diagonal"):
        ebn.pf_batch(ebn.Job(ebn._lib.LS_MIN_AFFINE, k, inputs, n=n, seed=99, tables=TABLE, corr_chol=bad))
    # Gamma under a copula must come tabulated
    rows = [[(5, 25.0, 2.4, 0, 0, -INF, INF), (1, 0.0, 1.0, 0, 0, -INF, INF)]]
    with pytest.raises(ebn.EbnError, match="EBN_EMARGINAL.*TABULATED"):
        ebn.pf_batch(ebn.Job(ebn._lib.LS_MIN_AFFINE, np.array([1.0, 0.0, 1.0, 1.0]), ebn.make_inputs(rows), n=100,
                             corr_chol=np.eye(2)))


def test_random_copulas_against_the_twin(ebn):
    """Random correlation matrices over random marginal rows (Gamma excluded: it must be tabulated
    under a copula): the looping copula plan equals the NumPy twin, and the plan unrolled at run time
    draws the same bits."""
    rng = np.random.default_rng(4242 + FUZZ_SEED)
    for trial in range(6):
        d = int(rng.integers(2, 7))
        row = []
        while len(row) < d:
            m = _random_marginal(rng)
            if m[0] != 5:
                row.append(m)
        B = rng.normal(0, 1, (d, 2))
        R = B @ B.T + np.diag(rng.uniform(0.5, 1.5, d))
        sdv = np.sqrt(np.diag(R))
        R = R / sdv[:, None] / sdv[None, :]
        for j, m in enumerate(row):          # Parameter rows/columns must be identity
            if m[0] == 0:
                R[j, :] = 0.0
                R[:, j] = 0.0
                R[j, j] = 1.0
        Lc = np.linalg.cholesky(R)
        inputs = ebn.make_inputs([row, row])
        kw = dict(n=4000, seed=31337, corr_chol=Lc)
        loop = ebn.Job(ebn._lib.LS_MIN_AFFINE, _min_affine_all(d), inputs, jit_plan=False, **kw)
        X = ebn.sample_matrix(loop, 1, 5, 3000)
        ref = replay.copula_sample_matrix(inputs[1], Lc, 1, 31337, 5, 3000)
        err = np.max(np.abs(X - ref) / np.maximum(np.abs(ref), 1.0), axis=0)
        assert np.all(err < 1e-9), (trial, row, err)
        if trial % 2 == 0:
            unrolled = ebn.Job(ebn._lib.LS_MIN_AFFINE, _min_affine_all(d), inputs, jit_plan=True, **kw)
            assert np.array_equal(ebn.sample_matrix(unrolled, 1, 5, 3000), X), (trial, row)


def test_w3_straub_copula_and_w5_grid(ebn):
    """BASELINE configs 3 and 5 at test size."""
    from ebn_b200 import workloads as W
    n = 1_000_000
    job = W.w3_straub_copula(n=n, bits=3, strict=True)
    cnt, pf, var = ebn.pf_batch(job)
    assert job.S == 8
    # scenario 0 has Straub's dependence structure exactly: same pf as the straub_frame functor
    ref_cnt, ref_pf, ref_var = ebn.pf_batch(W.w2_straub(n=n))
    assert abs(pf[0] - ref_pf[0]) < 3 * math.sqrt(var[0] + ref_var[0]) + 3e-4   # + table interpolation
    assert abs(pf[0] - 0.026129) < 0.01
    assert pf[0b101] > pf[0] and pf[0b010] > pf[0]      # heavier loads fail more often
    for s in (0, 5):
        X = ebn.sample_matrix(job, s, 0, 100_000)
        ref = replay.copula_sample_matrix(job.inputs[s], job.corr_chol, s, job.seed, 0, 100_000, tables=job.tables)
        assert np.max(np.abs(X - ref) / np.maximum(np.abs(ref), 1.0)) < 1e-9
    X = ebn.sample_matrix(job, 3, 0, n)
    assert int(cnt[3]) == ocpu.eval_matrix(ocpu.LS_MIN_AFFINE, job.consts, X)
    # W5: 1024 scenarios on the (A, b0) grid; pf grows with A and falls with b0
    j5 = W.w5_vehicle_grid(n=200_000, strict=True)
    c5, p5, _ = ebn.pf_batch(j5)
    assert j5.S == 1024 and ebn.last_stats()["specialised"] == 1
    G = p5.reshape(32, 32)
    assert G[31, 0] > 0.5 and G[0, 31] < 1e-3 and np.all(np.diff(G[:, 0]) > -0.01) and np.all(np.diff(G[31, :]) < 0.01)
    for s in (0, 517, 1023):
        X = ebn.sample_matrix(j5, s, 0, 200_000)
        assert int(c5[s]) == ocpu.eval_matrix(ocpu.LS_VEHICLE, j5.consts, X)


# ---------------------------------------------------------------------------
# 8. committed golden fixtures (tests/golden/, made by make_golden.py)
# ---------------------------------------------------------------------------
def test_golden_fixtures(ebn):
    import os
    gold = os.path.join(os.path.dirname(__file__), "golden")
    z = np.load(os.path.join(gold, "limit_states.npz"))
    c, g = ebn.eval_matrix(ebn._lib.LS_VEHICLE_SUSPENSION, z["vehicle_k"], z["vehicle_X"], strict=True, want_g=True)
    assert c == int(z["vehicle_count"]) and np.array_equal(g, z["vehicle_g"])
    c, g = ebn.eval_matrix(ebn._lib.LS_STRAUB_FRAME, z["straub_k"], z["straub_X"], strict=True, want_g=True)
    assert c == int(z["straub_count"]) and np.max(np.abs(g - z["straub_g"])) < 1e-9
    e = np.load(os.path.join(gold, "expression.npz"))      # the same limit states as committed expression programs
    c, g = ebn.eval_matrix(ebn._lib.LS_EXPRESSION, e["vehicle_program"], z["vehicle_X"], strict=True, want_g=True)
    assert c == int(z["vehicle_count"]) and np.array_equal(g, z["vehicle_g"])
    c, g = ebn.eval_matrix(ebn._lib.LS_EXPRESSION, e["straub_program"], z["straub_X"], strict=True, want_g=True)
    assert c == int(z["straub_count"]) and np.max(np.abs(g - z["straub_g"])) < 1e-9
    s = np.load(os.path.join(gold, "stream_v4.npz"))
    inputs = s["inputs"].view(ebn.INPUT_DTYPE).reshape(1, -1)
    inputs4 = np.repeat(inputs, 4, axis=0)  # scenario 3 is stream 3
    d = inputs.shape[1]
    for first in (0, 7, 2**33 + 5):
        job = ebn.Job(ebn._lib.LS_MIN_AFFINE, _min_affine_all(d), inputs4, n=first + 64, seed=int(s["seed"]),
                      tables=s["table"])
        X = ebn.sample_matrix(job, 3, first, 64)
        ref = s[f"first_{first}"]
        assert np.max(np.abs(X - ref) / np.maximum(np.abs(ref), 1.0)) < 5e-12


def test_eval_matrix_with_a_leading_dimension_larger_than_n(ebn):
    """A column-major matrix with ld > n owns (d-1)*ld + n elements, not ld*d: the library must not read past them."""
    import ctypes as C
    from ebn_b200 import workloads as W
    L = ebn._lib
    rng = np.random.default_rng(5)
    n, d, ld = 1000, 6, 1024
    X = _vehicle_matrix(rng, n)
    flat = np.full((d - 1) * ld + n, np.nan)
    for j in range(d):
        flat[j * ld:j * ld + n] = X[:, j]
    k = W.vehicle_consts()
    cnt = C.c_int64(0)
    g = np.empty(n)
    dp = C.POINTER(C.c_double)
    rc = L.load().ebn_eval_matrix(L.LS_VEHICLE_SUSPENSION, k.ctypes.data_as(dp), k.size, flat.ctypes.data_as(dp), n, d, ld, 1,
                                  C.byref(cnt), g.ctypes.data_as(dp))
    assert rc == 0
    want_cnt, want_g = ocpu.eval_matrix(ocpu.LS_VEHICLE, k, X, want_g=True)
    assert cnt.value == want_cnt and np.array_equal(g, want_g)


def test_large_jobs_with_mixed_marginal_kinds_are_split_by_kind_signature(ebn):
    """Scenarios whose marginal kinds differ (here: C truncated in four of the 20 scenarios, Ck in two others) cannot
    share a compile-time sampling plan. A large job is split by kind signature inside the library, every group runs
    on its own plan with the scenarios' original Philox streams: the counts are bit-identical to the unsplit call
    through the runtime-dispatched plan (EBN_SPLIT_KINDS=0), and faster."""
    import os
    from ebn_b200 import workloads as W
    job = W.w1_vehicle(n=10**8)
    job.inputs[2:6, 3]["lo"] = 400.0
    job.inputs[10:12, 4]["hi"] = 1500.0
    job.jit_plan = None
    cnt, _, _ = ebn.pf_batch(job)
    st = ebn.last_stats()
    assert st["launches"] == 3 and st["samples"] == 20 * 10**8
    os.environ["EBN_SPLIT_KINDS"] = "0"
    try:
        whole, _, _ = ebn.pf_batch(job)
        st0 = ebn.last_stats()
    finally:
        del os.environ["EBN_SPLIT_KINDS"]
    assert st0["launches"] == 1 and st0["specialised"] == 0
    assert np.array_equal(cnt, whole)
    assert st["kernel_ms"] < 0.6 * st0["kernel_ms"], (st["kernel_ms"], st0["kernel_ms"])
    # histogram mode is never split (its moments are summed in work-item order, which depends on the scenario count)
    edges = np.linspace(-2.0, 2.0, 65)
    small = W.w1_vehicle(n=10**8)
    small.inputs[0, 3]["lo"] = 400.0
    ebn.hist_batch(small, edges)
    assert ebn.last_stats()["launches"] == 2 and ebn.last_stats()["specialised"] == 0


def test_truncated_normal_cells_reach_full_precision(ebn):
    """The table-driven Phi^-1 of the truncated-normal kinds (normcdfinv_cells: 16 cells per octave of min(p, 1-p),
    degree 7) against SciPy on windows in the centre, across the median and deep in both tails, through the
    compile-time plan (shared-memory cells) and the runtime-dispatched one (the same cells from global memory):
    1e-15 relative to max(1, |x|) in the centre and the lower tail; the upper-tail window lives at p = 1 - 3e-5,
    where c + u d itself carries a relative error of 1e-12 in 1 - p (the reference's `truncated` quantile has the
    same cancellation), so it gets the twin's general 5e-12. The two plans give the same bits."""
    from ebn_b200 import workloads as W
    L = ebn._lib
    n = 200_000
    job = W.w1_vehicle(n=n, strict=True)
    wins = {3: (400.0, 460.0), 4: (-INF, 1440.0), 5: (95.0, INF)}          # C central, Ck lower tail (-3.5 sigma), K upper tail (+4 sigma)
    for col, (lo, hi) in wins.items():
        job.inputs[:, col]["lo"], job.inputs[:, col]["hi"] = lo, hi
    ebn.pf_batch(job)
    assert ebn.last_stats()["specialised"] == 1
    X = ebn.sample_matrix(job, 3, 0, n)
    ref = replay.sample_matrix(job.inputs[3], 3, job.seed, 0, n)
    err = np.max(np.abs(X - ref) / np.maximum(np.abs(ref), 1.0), axis=0)
    assert np.all(err[3:5] < 1e-15) and err[5] < 5e-12, err
    assert X[:, 3].min() >= 400.0 and X[:, 3].max() <= 460.0 and X[:, 4].max() <= 1440.0 and X[:, 5].min() >= 95.0
    mixed = W.w1_vehicle(n=n, strict=True)
    mixed.inputs[:] = job.inputs
    mixed.inputs[0, 2]["kind"] = L.IN_PARAM                                  # scenario 0 differs: no compile-time plan
    mixed.inputs[0, 2]["p"] = [9.0, 0, 0, 0]
    mixed.inputs[0, 2]["lo"], mixed.inputs[0, 2]["hi"] = -INF, INF
    ebn.pf_batch(mixed)
    assert ebn.last_stats()["specialised"] == 0
    assert np.array_equal(ebn.sample_matrix(mixed, 3, 0, n), X)
