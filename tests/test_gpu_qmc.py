"""Quasi-Monte Carlo (EBN_JOB_QMC_SOBOL; UQ.jl SobolSampling, an AbstractMonteCarlo — SURVEY.md 8c A4,
reference src/nodes/functionalnodes.jl:4,31): Owen-scrambled Sobol points, every marginal by inverse CDF."""
import dataclasses
import math

import numpy as np
import pytest
from scipy.special import ndtr

from oracle import cpu as ocpu
from oracle import replay

pytestmark = pytest.mark.gpu
INF = math.inf

ROW = [
    (0, 3.25, 0, 0, 0, -INF, INF), (1, 1.0, 2.0, 0, 0, -INF, INF), (1, 1.0, 2.0, 0, 0, 0.5, 4.0),
    (2, 0.3, 0.25, 0, 0, -INF, INF), (2, 0.3, 0.25, 0, 0, 1.0, 2.0), (3, 7.0, 12.0, 0, 0, 8.5, 9.5),
    (4, 41.0, 15.6, 0, 0, -INF, INF), (4, 41.0, 15.6, 0, 0, 30.0, 80.0), (6, 2.0, 1.0, 3.0, 0, -INF, INF),
    (6, 2.0, -1.0, 3.0, 0, 0.0, INF), (7, 0.0, 5.0, 0, 0, -INF, INF),
]
TABLE = np.array([-1.0, 0.0, 0.5, 2.0, 10.0])


def _min_affine_all(d):
    return np.array([1.0, 0.0] + [1.0] * d)


@pytest.mark.parametrize("first,n", [(0, 4096), (7, 1001), (2**20 + 3, 2048)])
def test_qmc_replay_matches_the_numpy_twin(ebn, first, n):
    L = ebn._lib
    inputs = L.make_inputs([ROW, ROW])
    seed = 0xC0FFEE1234
    job = L.Job(L.LS_MIN_AFFINE, _min_affine_all(inputs.shape[1]), inputs, n=first + n, seed=seed, tables=TABLE, qmc=True)
    X = ebn.sample_matrix(job, 1, first, n)
    ref = replay.qmc_sample_matrix(inputs[1], 1, seed, first, n, tables=TABLE)
    err = np.max(np.abs(X - ref) / np.maximum(np.abs(ref), 1.0), axis=0)
    assert np.all(err < 5e-12), err
    # another stream (scenario 0) is a different scrambling of the same sequence
    assert not np.allclose(ebn.sample_matrix(job, 0, first, n)[:, 1], X[:, 1])


def test_qmc_points_form_a_net_and_counts_are_exact_and_partition_invariant(ebn):
    L = ebn._lib
    from ebn_b200 import workloads as W
    n = 1 << 16
    base = W.w1_vehicle(n=n, strict=True)
    job = dataclasses.replace(base, qmc=True)
    cnt, pf, _ = ebn.pf_batch(job)
    assert ebn.last_stats()["specialised"] == 0
    for s in (0, 7, 19):
        X = ebn.sample_matrix(job, s, 0, n)
        assert int(cnt[s]) == ocpu.eval_matrix(ocpu.LS_VEHICLE, job.consts, X), s
        # every random input: exactly one of the 2^16 points in each interval of length 2^-16 of its uniform
        lo, hi = float(job.inputs[s, 2]["lo"]), float(job.inputs[s, 2]["hi"])
        for u in ((X[:, 2] - lo) / (hi - lo), ndtr((X[:, 3] - 431.7221) / 10.0), ndtr((X[:, 5] - 55.0406) / 10.0)):
            assert np.array_equal(np.bincount(np.minimum((u * n).astype(np.int64), n - 1), minlength=n), np.ones(n, dtype=np.int64))
        # dimensions 1 and 2 of the sequence (V and C here): a (0, 16, 2)-net — one point in every 2^-8 x 2^-8 box
        uv = (X[:, 2] - lo) / (hi - lo)
        uc = ndtr((X[:, 3] - 431.7221) / 10.0)
        box = (np.minimum((uv * 256).astype(int), 255) << 8) | np.minimum((uc * 256).astype(int), 255)
        assert np.array_equal(np.bincount(box, minlength=1 << 16), np.ones(1 << 16, dtype=np.int64))
    # shards and resumed pieces add up to the one-call counts
    tot = np.zeros_like(cnt)
    for r in range(3):
        j = dataclasses.replace(job, shard_rank=r, shard_world=3, partial=True)
        tot += ebn.pf_batch(j)[0]
    assert np.array_equal(tot, cnt)
    a = ebn.pf_batch(dataclasses.replace(job, n=20_001))[0]
    b = ebn.pf_batch(dataclasses.replace(job, n=n - 20_001, sample_offset=20_001))[0]
    assert np.array_equal(a + b, cnt)
    # a Monte Carlo run of the same size agrees within its own noise
    mc = ebn.pf_batch(base)[1]
    se = np.sqrt(np.maximum(mc * (1 - mc), 1e-12) / n)
    assert np.all(np.abs(pf - mc) < 5 * se + 1e-9)


def test_qmc_converges_faster_than_monte_carlo_on_an_analytic_limit_state(ebn):
    """Toy node of test/networks/ebn/evaluate/evaluate_node.jl:96-110: 2 - (A + B) < 0 with B ~ N(0,1), A = 1, 2:
    pf = Phi(-1), Phi(0). With 2^16 scrambled Sobol points the error is far below the Monte Carlo standard error."""
    L = ebn._lib
    prog = np.array([2, 2, 1.0, 1, 0, 1.0, 1, 1, 2, 2.0, 0, -1.0, 1, 2], dtype=float)   # C = A + B ; 2 - C
    rows = [[(0, a, 0, 0, 0, -INF, INF), (1, 0.0, 1.0, 0, 0, -INF, INF)] for a in (1.0, 2.0)]
    n = 1 << 16
    errs = []
    for seed in range(8):
        job = L.Job(L.LS_POLYNOMIAL, prog, L.make_inputs(rows), n=n, seed=seed, qmc=True)
        pf = ebn.pf_batch(job)[1]
        errs.append(np.abs(pf - np.array([ndtr(-1.0), 0.5])))
    rms = np.sqrt(np.mean(np.square(errs), axis=0))
    mc_se = np.sqrt(np.array([ndtr(-1.0) * ndtr(1.0), 0.25]) / n)
    assert np.all(rms < 0.1 * mc_se), (rms, mc_se)


def test_qmc_errors_formulas_and_the_node_api(ebn):
    L = ebn._lib
    E = ebn
    from ebn_b200 import workloads as W
    with pytest.raises(L.EbnError, match="EBN_EMARGINAL.*Gamma"):
        ebn.pf_batch(dataclasses.replace(W.w2_straub(n=1000), qmc=True))
    with pytest.raises(L.EbnError, match="copula"):
        ebn.pf_batch(dataclasses.replace(W.w3_straub_copula(n=1000, bits=1), qmc=True))
    with pytest.raises(L.EbnError, match="2\\^32"):
        ebn.pf_batch(dataclasses.replace(W.w1_vehicle(n=2**32 + 1), qmc=True))
    # a formula compiled at run time under the Sobol plan gives the counts of the polynomial functor
    from ebn_b200 import expr
    rows = [[(0, a, 0, 0, 0, -INF, INF), (1, 0.0, 1.0, 0, 0, -INF, INF)] for a in (1.0, 2.0)]
    prog = np.array([2, 2, 1.0, 1, 0, 1.0, 1, 1, 2, 2.0, 0, -1.0, 1, 2], dtype=float)
    pj = L.Job(L.LS_POLYNOMIAL, prog, L.make_inputs(rows), n=4096, seed=5, strict=True, qmc=True, jit_plan=False)
    if L.jit_available():
        ep = expr.compile_program(["A", "B"], [("C", "df.A .+ df.B")], "2 .- df.C")
        ej = L.Job(L.LS_EXPRESSION, ep.consts, pj.inputs, n=4096, seed=5, strict=True, qmc=True)
        assert np.array_equal(ebn.pf_batch(ej)[0], ebn.pf_batch(pj)[0])
    # node API: a DiscreteFunctionalNode and a ContinuousFunctionalNode with SobolSampling (evaluate_node.jl:22-33, 79-95)
    from ebn_b200 import evaluation as ev
    cpt = E.DiscreteConditionalProbabilityTable("A")
    cpt[("A", "a1")], cpt[("A", "a2")] = 0.5, 0.5
    a = E.DiscreteNode("A", cpt, {"a1": [E.Parameter(1, "A")], "a2": [E.Parameter(2, "A")]})
    cb = E.ContinuousConditionalProbabilityTable()
    cb[()] = E.Normal()
    b = E.ContinuousNode("B", cb)
    sim = E.SobolSampling(1 << 14)
    f = E.DiscreteFunctionalNode("F", [E.Model(E.Poly([(1, ["A"]), (1, ["B"])]), "C")], E.Poly([(2, []), (-1, ["C"])]), sim)
    net = E.EnhancedBayesianNetwork([a, b, f])
    E.add_child(net, a, f)
    E.add_child(net, b, f)
    E.order(net)
    out = ev._evaluate_node(net, f)
    assert np.allclose(out.cpt.column("Π"), [ndtr(-1), ndtr(1), 0.5, 0.5], atol=5e-4)
    c = E.ContinuousFunctionalNode("C", [E.Model(E.Poly([(1, ["A"]), (1, ["B"])]), "C")], sim)
    net = E.EnhancedBayesianNetwork([a, b, c])
    E.add_child(net, a, c)
    E.add_child(net, b, c)
    E.order(net)
    outc = ev._evaluate_node(net, c)
    d1 = outc.cpt[{"A": "a1"}]
    assert abs(d1.mean() - 1.0) < 0.01 and abs(d1.cdf(1.0) - 0.5) < 0.01
