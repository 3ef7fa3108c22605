"""GPU parity of the run-time compiled kernels (NVRTC): expression limit states and sampling plans
compiled on demand. Through the C ABI, against the C oracle's stack machine and the compiled-in kernels.
"""
import math

import numpy as np
import pytest

# EBN_FUZZ_SEED=k shifts the seeds of the randomised tests (extra fuzzing runs); the default is what CI runs
FUZZ_SEED = int(__import__("os").environ.get("EBN_FUZZ_SEED", "0"))

from oracle import cpu as ocpu
from oracle import replay

pytestmark = pytest.mark.gpu

INF = math.inf

VEHICLE_MODELS = [
    # demo/vehicle_suspension.jl:17-23, the closures' text; the two non-integer powers are runtime constants
    ("g1", "1 .- (pi .* m .* df.V .* df.A) ./ (df.b0 .* df.K .* g .^ 2) .* ((df.Ck ./ (m .+ M) .- (df.C ./ M)) .^ 2"
           " .+ df.C .^ 2 ./ (m .* M) .+ df.Ck .* df.K .^ 2 ./ (m .* M .^ 2))"),
    ("g2", "4000 .* df.C .* Mg15 .- 8.6394"),
    ("g3", "2 .* sqrt(M .* g .* (df.K .^ 2 .* df.Ck ./ (df.C .* (m .+ M)) .+ df.C)) .- 1"),
    ("g4", "df.Ck .- gMm"),
    ("y", "minimum([g1[1], g2, g3, g4[1]])"),
]


def vehicle_program():
    from ebn_b200 import expr, workloads as W
    k = W.vehicle_consts()
    return expr.compile_program(["A", "b0", "V", "C", "Ck", "K"], VEHICLE_MODELS, "y",
                                params={"M": k[0], "m": k[1], "g": k[2], "Mg15": k[3], "gMm": k[4]})


def test_expression_vehicle_equals_compiled_in_functor(ebn):
    """The vehicle limit state written as text gives, in strict mode, the same g bit for bit and the
    same fused counts as the hand-written functor; the sampler streams are identical."""
    from ebn_b200 import workloads as W
    prog = vehicle_program()
    n = 200_001
    ref = W.w1_vehicle(n=n, strict=True)
    job = ebn.Job(ebn._lib.LS_EXPRESSION, prog.consts, ref.inputs, n=n, seed=ref.seed, strict=True)
    cnt_ref, _, _ = ebn.pf_batch(ref)
    cnt, pf, var = ebn.pf_batch(job)
    st = ebn.last_stats()
    assert st["specialised"] == 1          # the job's kinds became a StaticPlan of the generated kernel
    assert np.array_equal(cnt, cnt_ref)
    assert np.array_equal(pf, cnt / n)
    # replayed matrix with the model columns g1..g4, y appended (evaluate!(models, df))
    X, g = ebn.sample_matrix(job, 3, 0, n, want_g=True, ncols=prog.ncols)
    Xr = ebn.sample_matrix(ref, 3, 0, n)
    assert np.array_equal(X[:, :6], Xr)
    want_cnt, want_g = ocpu.eval_matrix(ocpu.LS_VEHICLE, ref.consts, Xr, want_g=True)
    assert np.array_equal(g, want_g) and int(cnt[3]) == want_cnt
    g_np, cols = replay.g_expression(Xr, prog.consts, want_columns=True)
    assert np.array_equal(cols, X) and np.array_equal(g_np, g)
    # eval_matrix on a given matrix
    c2, g2 = ebn.eval_matrix(ebn._lib.LS_EXPRESSION, prog.consts, Xr, strict=True, want_g=True)
    assert c2 == want_cnt and np.array_equal(g2, want_g)
    # second call: kernels come from the cache
    ebn.pf_batch(job)
    assert ebn.last_stats()["jit_compiles"] == 0
    # new runtime constants reuse the code (a heavier car: more failures in g4? just check it runs and differs)
    job2 = ebn.Job(ebn._lib.LS_EXPRESSION, prog.with_params({"gMm": 1460.0}), ref.inputs, n=n, seed=ref.seed, strict=True)
    cnt2, _, _ = ebn.pf_batch(job2)
    assert ebn.last_stats()["jit_compiles"] == 0 and np.all(cnt2 >= cnt) and np.any(cnt2 > cnt)
    # fast mode: ties only
    jobf = ebn.Job(ebn._lib.LS_EXPRESSION, prog.consts, ref.inputs, n=n, seed=ref.seed, strict=False)
    cntf, _, _ = ebn.pf_batch(jobf)
    assert np.max(np.abs(cntf - cnt)) <= 1


def test_expression_all_opcodes_against_oracle(ebn):
    """Every opcode, on random inputs, against the C and NumPy stack machines: exact for the IEEE
    operations, a few ulp for exp/log/sin/cos/pow."""
    from ebn_b200 import expr
    models = [
        ("s", "a + b * c - a / (abs(b) + 1.5)"),
        ("t", "sqrt(abs(s)) + min(a, b, c) - max(a, -b) + (-c)^2 + b^3"),
        ("u", "ifelse(a < b, s, t) + (c if b > a else 2.0) + k0 * a"),
    ]
    exact = expr.compile_program(["a", "b", "c"], models, "u - k1", params={"k0": 0.25, "k1": 0.5})
    rng = np.random.default_rng(7)
    X = rng.normal(0, 2, (50_000, 3))
    X[:5] = [[np.nan, 1, 2], [1, np.nan, 2], [0.0, -0.0, 1], [np.inf, 1, 1], [1e-310, 1e300, -1e300]]
    want_cnt, want_g = ocpu.eval_matrix(ocpu.LS_EXPRESSION, exact.consts, X, want_g=True)
    got_cnt, got_g = ebn.eval_matrix(ebn._lib.LS_EXPRESSION, exact.consts, X, strict=True, want_g=True)
    assert got_cnt == want_cnt
    assert np.array_equal(got_g, want_g, equal_nan=True)
    assert np.array_equal(replay.g_expression(X, exact.consts), want_g, equal_nan=True)

    transc = expr.compile_program(["a", "b", "c"], [("w", "exp(a / 4) + log(abs(b) + 1) + sin(c) * cos(a)")],
                                  "w - abs(a)^1.5 - b^5 / 100 + c^-2", params={})
    X2 = rng.normal(0, 2, (50_000, 3))
    _, want = ocpu.eval_matrix(ocpu.LS_EXPRESSION, transc.consts, X2, want_g=True)
    want_t = want
    _, got = ebn.eval_matrix(ebn._lib.LS_EXPRESSION, transc.consts, X2, strict=True, want_g=True)
    err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
    assert np.max(err) < 1e-12, np.max(err)
    # fast mode: per-sample exp / log go through the lean table-driven versions (guarded outside their
    # domain), a constant operand through the library call in setup()
    wide = expr.compile_program(["a", "b", "c"], [("w", "exp(a * 200) + log(abs(b) * 1e-200 + 1e-320) + exp(k) * log(k + 2)")],
                                "w / (1 + abs(w)) - c / 8", params={"k": 0.75})
    _, want = ocpu.eval_matrix(ocpu.LS_EXPRESSION, wide.consts, X2, want_g=True)
    _, got = ebn.eval_matrix(ebn._lib.LS_EXPRESSION, wide.consts, X2, strict=False, want_g=True)
    ok = np.isfinite(want)
    assert ok.mean() > 0.9 and np.array_equal(np.isfinite(got), ok)
    assert np.max(np.abs(got[ok] - want[ok])) < 1e-12
    _, got_fast = ebn.eval_matrix(ebn._lib.LS_EXPRESSION, transc.consts, X2, strict=False, want_g=True)
    assert np.max(np.abs(got_fast - want_t) / np.maximum(1.0, np.abs(want_t))) < 1e-11


def test_expression_reference_toy_models(ebn):
    """The models of test/networks/ebn/evaluate/evaluate_node.jl:96-110 (C = A + B, performance 2 - C)
    as expressions: analytic anchors Phi(-1) and 0.5, and equality with the polynomial functor."""
    from ebn_b200 import expr
    from test_gpu_parity import PROG_2_MINUS_A_PLUS_B
    prog = expr.compile_program(["A", "B"], [("C", "df.A .+ df.B")], "2 .- df.C")
    rows = [[(0, a, 0, 0, 0, -INF, INF), (1, 0.0, 1.0, 0, 0, -INF, INF)] for a in (1.0, 2.0)]
    inputs = ebn.make_inputs(rows)
    n = 2_000_000
    je = ebn.Job(ebn._lib.LS_EXPRESSION, prog.consts, inputs, n=n, seed=11, strict=True)
    jp = ebn.Job(ebn._lib.LS_POLYNOMIAL, PROG_2_MINUS_A_PLUS_B, inputs, n=n, seed=11, strict=True)
    ce, pfe, vare = ebn.pf_batch(je)
    cp, _, _ = ebn.pf_batch(jp)
    assert np.array_equal(ce, cp)
    for pf, var, want in zip(pfe, vare, (0.158655, 0.5)):
        assert abs(pf - want) < 4 * math.sqrt(var) + 1e-9
    # histogram path (ContinuousFunctionalNode): same counts as the polynomial functor
    edges = np.linspace(-4, 4, 33)
    he, se, qe = ebn.hist_batch(je, edges)
    hp, sp, qp = ebn.hist_batch(jp, edges)
    assert np.array_equal(he, hp) and np.allclose(se, sp, rtol=1e-12) and np.allclose(qe, qp, rtol=1e-12)


def test_jit_sampling_plan_matches_dynamic_plan(ebn):
    """A job no compiled-in plan covers (C truncated, K log-normal): the run-time compiled StaticPlan gives
    exactly the counts and streams of the runtime-dispatched plan."""
    from ebn_b200 import workloads as W
    n = 300_001
    base = W.w1_vehicle(n=n, strict=True)
    base.inputs[:, 3]["lo"] = 400.0       # C truncated at 400 → inverse-CDF kind
    base.inputs[:, 5]["kind"], base.inputs[:, 5]["p"] = ebn._lib.IN_LOGNORMAL, (4.0, 0.18, 0, 0)
    base.inputs[:, 5]["hi"] = 80.0
    dyn = ebn.Job(base.limit_state, base.consts, base.inputs, n=n, seed=base.seed, strict=True, jit_plan=False)
    jit = ebn.Job(base.limit_state, base.consts, base.inputs, n=n, seed=base.seed, strict=True, jit_plan=True)
    c_dyn, _, _ = ebn.pf_batch(dyn)
    assert ebn.last_stats()["specialised"] == 0
    c_jit, _, _ = ebn.pf_batch(jit)
    st = ebn.last_stats()
    assert st["specialised"] == 1
    assert np.array_equal(c_dyn, c_jit)
    assert np.array_equal(ebn.sample_matrix(dyn, 7, 5, 4097), ebn.sample_matrix(jit, 7, 5, 4097))
    # every marginal kind through a run-time plan
    from test_gpu_parity import ALL_KINDS_ROW, TABLE
    inputs = ebn.make_inputs([ALL_KINDS_ROW, ALL_KINDS_ROW])
    d = inputs.shape[1]
    k = np.array([1.0, -150.0] + [1.0] * d)
    a = ebn.Job(ebn._lib.LS_MIN_AFFINE, k, inputs, n=100_000, seed=5, strict=True, tables=TABLE, jit_plan=False)
    b = ebn.Job(ebn._lib.LS_MIN_AFFINE, k, inputs, n=100_000, seed=5, strict=True, tables=TABLE, jit_plan=True)
    ca, _, _ = ebn.pf_batch(a)
    cb, _, _ = ebn.pf_batch(b)
    assert ebn.last_stats()["specialised"] == 1
    assert np.array_equal(ca, cb)
    assert np.array_equal(ebn.sample_matrix(a, 1, 0, 5000), ebn.sample_matrix(b, 1, 0, 5000))


def test_jit_copula_plan_matches_runtime_copula(ebn):
    """BASELINE config 3 shape (Gaussian copula over 5 lognormals, a tabulated Gamma and a Gumbel): the
    copula plan unrolled at run time draws exactly the samples of the looping one, and the three frame
    mechanisms written as a formula count like the min-affine functor."""
    from ebn_b200 import expr, workloads as W
    n = 200_001
    base = W.w3_straub_copula(n=n, bits=3, strict=True)
    kw = dict(n=n, seed=base.seed, strict=True, corr_chol=base.corr_chol, tables=base.tables)
    loop = ebn.Job(base.limit_state, base.consts, base.inputs, jit_plan=False, **kw)
    unrolled = ebn.Job(base.limit_state, base.consts, base.inputs, jit_plan=True, **kw)
    c_loop, _, _ = ebn.pf_batch(loop)
    assert ebn.last_stats()["specialised"] == 0
    c_unr, _, _ = ebn.pf_batch(unrolled)
    assert ebn.last_stats()["specialised"] == 1
    assert np.array_equal(c_loop, c_unr) and c_loop.min() > 0
    Xa, Xb = ebn.sample_matrix(loop, 5, 3, 4099), ebn.sample_matrix(unrolled, 5, 3, 4099)
    assert np.array_equal(Xa, Xb)
    prog = expr.compile_program(
        ["r1", "r2", "r3", "r4", "r5", "V", "H"], [],
        "min(r1 + r2 + r4 + r5 - 5 * H, r2 + 2 * r3 + r4 - 5 * V, r1 + 2 * r3 + 2 * r4 + r5 - 5 * H - 5 * V)")
    as_text = ebn.Job(ebn._lib.LS_EXPRESSION, prog.consts, base.inputs, **kw)
    c_txt, _, _ = ebn.pf_batch(as_text)
    assert ebn.last_stats()["specialised"] == 1
    # the affine functor accumulates 0 + c*x terms in column order, the formula in its own order:
    # g agrees to rounding, counts to ties
    assert np.max(np.abs(c_txt - c_loop)) <= 2
    assert np.array_equal(ebn.sample_matrix(as_text, 5, 3, 4099), Xa)


def _random_formula(rng, names, consts, depth):
    """A random formula over columns, named constants and literals built from the exactly rounded
    operations only (so strict mode must agree with the oracle bit for bit)."""
    if depth == 0 or rng.random() < 0.15:
        r = rng.random()
        if r < 0.55:
            return str(rng.choice(names))
        if r < 0.8:
            return str(rng.choice(consts))
        return repr(round(float(rng.normal(0, 3)), 3))
    sub = lambda: _random_formula(rng, names, consts, depth - 1)  # noqa: E731
    k = rng.integers(0, 12)
    if k < 4:
        return f"({sub()} {rng.choice(['+', '-', '*', '/'])} {sub()})"
    if k == 4:
        return f"sqrt(abs({sub()}))"
    if k == 5:
        return f"min({sub()}, {sub()}, {sub()})"
    if k == 6:
        return f"max({sub()}, {sub()})"
    if k == 7:
        return f"(-{sub()})"
    if k == 8:
        return f"abs({sub()})"
    if k == 9:
        return f"({sub()})^{rng.choice([2, 3])}"
    if k == 10:
        return f"ifelse({sub()} < {sub()}, {sub()}, {sub()})"
    return f"({sub()} if {sub()} > {sub()} else {sub()})"


def test_random_formulas_against_the_oracle(ebn):
    """Differential test of the whole chain text -> program -> generated functor -> NVRTC -> kernel:
    25 random model chains (constant sub-expressions, reciprocal hoisting, stored columns read back,
    selections) evaluated in strict mode equal the C oracle's stack machine bit for bit, NaNs included;
    the fast mode stays within rounding."""
    from ebn_b200 import expr
    rng = np.random.default_rng(20261020 + FUZZ_SEED)
    X = rng.normal(0.5, 2.0, (20_000, 4))
    X[:3] = [[0.0, -0.0, 1.0, 2.0], [1e300, 1e-300, -1e300, 3.0], [np.nan, 1.0, 2.0, np.inf]]
    for trial in range(25):
        d = int(rng.integers(1, 5))
        names = [f"x{j}" for j in range(d)]
        consts = {f"k{j}": float(rng.normal(0, 2)) for j in range(int(rng.integers(1, 4)))}
        models, avail = [], list(names)
        for t in range(int(rng.integers(0, 4))):
            models.append((f"m{t}", _random_formula(rng, avail, list(consts), int(rng.integers(1, 5)))))
            avail.append(f"m{t}")
        perf = _random_formula(rng, avail, list(consts), int(rng.integers(1, 5)))
        prog = expr.compile_program(names, models, perf, consts)
        want_cnt, want_g = ocpu.eval_matrix(ocpu.LS_EXPRESSION, prog.consts, X[:, :d], want_g=True)
        got_cnt, got_g = ebn.eval_matrix(ebn._lib.LS_EXPRESSION, prog.consts, X[:, :d], strict=True, want_g=True)
        assert got_cnt == want_cnt and np.array_equal(got_g, want_g, equal_nan=True), (trial, models, perf)
        _, fast_g = ebn.eval_matrix(ebn._lib.LS_EXPRESSION, prog.consts, X[:, :d], strict=False, want_g=True)
        ok = np.isfinite(want_g) & (np.abs(want_g) < 1e100)
        scale = np.maximum(1.0, np.abs(want_g[ok]))
        # contraction and reciprocal hoisting move g by rounding only — except across a discontinuity
        # (a comparison that flips), which a handful of rows may hit
        bad = np.abs(fast_g[ok] - want_g[ok]) / scale > 1e-6
        assert bad.mean() < 0.01, (trial, float(bad.mean()), models, perf)


def test_random_jobs_end_to_end(ebn):
    """Everything at once, randomised: random marginal rows (same kinds in every scenario or not), a
    random formula chain over them, odd sample counts, a resume offset. The fused count of every
    scenario equals the oracle's count on the replayed matrix, the shards of a 3-rank job add up to the
    whole, and the histogram of g over the whole real line holds every sample."""
    import dataclasses
    from ebn_b200 import expr
    from test_gpu_parity import _random_marginal
    rng = np.random.default_rng(9001 + FUZZ_SEED)
    for trial in range(8):
        d = int(rng.integers(2, 6))
        base = [_random_marginal(rng) for _ in range(d)]
        rows = [base, base, [_random_marginal(rng) for _ in range(d)] if trial % 2 else base]
        names = [f"x{j}" for j in range(d)]
        consts = {"k0": float(rng.normal(0, 2))}
        models = [("m0", _random_formula(rng, names, list(consts), 3))]
        perf = _random_formula(rng, names + ["m0"], list(consts), 3)
        prog = expr.compile_program(names, models, perf, consts)
        n = int(rng.integers(20_000, 60_000)) | 1
        job = ebn.Job(ebn._lib.LS_EXPRESSION, prog.consts, ebn.make_inputs(rows), n=n, seed=int(rng.integers(1, 2**60)),
                      strict=True, sample_offset=int(rng.choice([0, 7, 2**34 + 3])))
        cnt = ebn.pf_batch(job)[0]
        assert ebn.last_stats()["specialised"] == (0 if trial % 2 else 1)
        for s in range(3):
            X = ebn.sample_matrix(job, s, job.sample_offset, n)
            assert int(cnt[s]) == ocpu.eval_matrix(ocpu.LS_EXPRESSION, prog.consts, X), (trial, s, models, perf)
        parts = sum(ebn.pf_batch(dataclasses.replace(job, shard_rank=r, shard_world=3, partial=True))[0] for r in range(3))
        assert np.array_equal(parts, cnt)
        h = ebn.hist_batch(job, np.array([0.0]))[0]
        assert np.all(h.sum(axis=1) == n) and np.all(h[:, 0] <= n)
        # bin 0 = g < 0 (NaN lands in the last bin): the histogram and the count path agree
        assert np.array_equal(h[:, 0], cnt)


def test_library_rewrites_coefficient_table_functors(ebn):
    """Inside the library, polynomial and min-affine batches that are large (or ask for it) are rewritten
    into expression programs instruction for instruction: strict counts and histograms do not change."""
    from test_gpu_parity import poly_program
    rng = np.random.default_rng(5)
    rows = [[(1, float(rng.normal()), 1.0 + float(rng.random()), 0, 0, -INF, INF), (3, -1.0, 2.0, 0, 0, -INF, INF),
             (0, 0.3 * s, 0, 0, 0, -INF, INF)] for s in range(4)]
    inputs = ebn.make_inputs(rows)
    prog = poly_program([[(1.0, [0, 0]), (-0.7, []), (1.0, [1])], [(0.5, [3, 2]), (2.0, [0])], [(1.0, [4]), (-1.25, [])]])
    n = 300_001
    a = ebn.Job(ebn._lib.LS_POLYNOMIAL, prog, inputs, n=n, seed=8, strict=True, jit_plan=False)
    b = ebn.Job(ebn._lib.LS_POLYNOMIAL, prog, inputs, n=n, seed=8, strict=True, jit_plan=True)
    ca = ebn.pf_batch(a)[0]
    assert ebn.last_stats()["specialised"] == 0
    cb = ebn.pf_batch(b)[0]
    assert ebn.last_stats()["specialised"] == 1
    assert np.array_equal(ca, cb) and 0 < ca.min() and ca.max() < n
    edges = np.linspace(-6, 10, 65)
    ha, hb = ebn.hist_batch(a, edges), ebn.hist_batch(b, edges)
    assert np.array_equal(ha[0], hb[0]) and np.allclose(ha[1], hb[1], rtol=1e-12)
    k = np.array([2.0, 0.5, 1.0, 0.0, -2.0, -0.25, 0.0, 1.5, 1.0])
    c = ebn.Job(ebn._lib.LS_MIN_AFFINE, k, inputs, n=n, seed=8, strict=True, jit_plan=False)
    d = ebn.Job(ebn._lib.LS_MIN_AFFINE, k, inputs, n=n, seed=8, strict=True, jit_plan=True)
    assert np.array_equal(ebn.pf_batch(c)[0], ebn.pf_batch(d)[0])
    assert ebn.last_stats()["specialised"] == 1
    # and the automatic case: a batch above the threshold
    big = ebn.Job(ebn._lib.LS_POLYNOMIAL, prog, inputs[:1], n=6 * 10**10, seed=8)
    cnt, pf, var = ebn.pf_batch(big)
    st = ebn.last_stats()
    assert st["specialised"] == 1 and st["kernel_ms"] < 1000.0
    assert abs(pf[0] - ca[0] / n) < 5 * math.sqrt(ca[0] / n * (1 - ca[0] / n) / n)


def test_large_polynomial_jobs_take_the_compiled_road(ebn):
    """uq.pf_scenarios sends a polynomial chain through the run-time compiled path once the job is
    large enough to repay the compilation: same anchor (Phi(-1)), 5x the throughput."""
    from ebn_b200 import uq
    E = ebn
    models = [E.Model(E.Poly([(1, ["A"]), (1, ["B"])]), "C")]
    perf = E.Poly([(2, []), (-1, ["C"])])
    inputs = [E.Parameter(1.0, "A"), E.RandomVariable(E.Normal(), "B")]
    pf, sd, smp = E.probability_of_failure(models, perf, inputs, E.MonteCarlo(int(uq.JIT_POLY_MIN_SAMPLES)))
    st = ebn.last_stats()
    assert st["specialised"] == 1 and smp.job.limit_state == ebn._lib.LS_EXPRESSION
    assert abs(pf - 0.15865525393145707) < 5 * sd
    pf_small, _, smp_small = E.probability_of_failure(models, perf, inputs, E.MonteCarlo(100_000))
    assert smp_small.job.limit_state == ebn._lib.LS_POLYNOMIAL and abs(pf_small - 0.158655) < 0.006


def test_expression_errors(ebn):
    from ebn_b200 import expr
    L = ebn._lib
    rows = [[(1, 0.0, 1.0, 0, 0, -INF, INF)] * 2]
    bad = np.array([0, 1, L.OP_ADD, 0], dtype=float)
    with pytest.raises(ebn.EbnError) as e:
        ebn.pf_batch(ebn.Job(L.LS_EXPRESSION, bad, ebn.make_inputs(rows), n=10))
    assert e.value.code == -2 and "operands" in e.value.message
    prog = expr.compile_program(["a", "b"], [], "a - b")
    with pytest.raises(ebn.EbnError) as e:   # program compiled for 2 inputs, job has 1
        ebn.pf_batch(ebn.Job(L.LS_EXPRESSION, prog.consts, ebn.make_inputs([[rows[0][0]]]), n=10))
    assert e.value.code == -2
