"""CPU tests of the expression limit state: the text compiler, the program validator and code
generator behind the C ABI (no GPU needed), run-time compilation itself (NVRTC compiles without a
device), and the oracle's two stack machines against the hand-written limit states."""
import math
import os

import numpy as np
import pytest

import ebn_b200
from ebn_b200 import _lib as L
from ebn_b200 import expr, workloads as W
from oracle import cpu as ocpu
from oracle import replay

INF = math.inf

VEHICLE_MODELS = [
    ("g1", "1 .- (pi .* m .* df.V .* df.A) ./ (df.b₀ .* df.K .* g .^ 2) .* ((df.Cₖ ./ (m .+ M) .- (df.C ./ M)) .^ 2"
           " .+ df.C .^ 2 ./ (m .* M) .+ df.Cₖ .* df.K .^ 2 ./ (m .* M .^ 2))"),
    ("g2", "4000 .* df.C .* Mg15 .- 8.6394"),
    ("g3", "2 .* sqrt.(M .* g .* (df.K .^ 2 .* df.Cₖ ./ (df.C .* (m .+ M)) .+ df.C)) .- 1"),
    ("g4", "df.Cₖ .- gMm"),
    ("y", "minimum([g1[1], g2, g3, g4[1]])"),
]
CAPACITY = "exp(rand(Normal(λ + ρ * ζ * df.Uᵣ, sqrt((1 - ρ^2) * ζ^2))))"
FRAME = ("minimum([r1 + r2 + r4 + r5 - 5 * df.H, r2 + 2 * r3 + r4 - 5 * df.V, "
         "r1 + 2 * r3 + 2 * r4 + r5 - 5 * df.H - 5 * df.V])")


def _code(prog):
    P = int(prog.consts[0])
    c = prog.consts[2 + P:]
    return [(int(c[2 * i]), c[2 * i + 1]) for i in range(int(prog.consts[1]))]


def test_compiler_emits_post_order_in_julia_evaluation_order():
    # test/networks/ebn/evaluate/evaluate_node.jl:96-98: C = A + B, performance 2 - C
    p = expr.compile_program(["A", "B"], [("C", "df.A .+ df.B")], "2 .- df.C")
    assert _code(p) == [(L.OP_INPUT, 0), (L.OP_INPUT, 1), (L.OP_ADD, 0), (L.OP_STORE, 0),
                        (L.OP_CONST, 2.0), (L.OP_INPUT, 2), (L.OP_SUB, 0)]
    assert (p.d, p.ncols, p.columns) == (2, 3, ["A", "B", "C"])
    # a*b*c folds left; x^2 and x^3 are literal powers; -2.5 is a literal, -x a negation
    p = expr.compile_program(["x", "y"], [], "x*y*k + x^2 - y^3 + -2.5 - (-x) + x^0.5 + y^-1", {"k": 3.0})
    ops = [o for o, _ in _code(p)]
    assert ops[:5] == [L.OP_INPUT, L.OP_INPUT, L.OP_MUL, L.OP_PARAM, L.OP_MUL]
    assert (L.OP_POWI, 2.0) in _code(p) and (L.OP_POWI, 3.0) in _code(p) and (L.OP_POWI, -1.0) in _code(p)
    assert (L.OP_CONST, -2.5) in _code(p) and L.OP_NEG in ops and L.OP_POW in ops
    assert p.params == ["k"] and p.consts[2] == 3.0
    assert np.array_equal(p.with_params({"k": 4.0})[3:], p.consts[3:]) and p.with_params({"k": 4.0})[2] == 4.0
    # python syntax, comparisons and selection
    p = expr.compile_program(["a", "b"], [], "(a if a > b else b) - ifelse(a < b, 1.0, 2.0) + max(a, b, 0)")
    assert [o for o, _ in _code(p)].count(L.OP_SELECT) == 2
    for bad, why in [("a +", "parse"), ("foo(a)", "not supported"), ("a + zz", "unknown name"), ("a % b", "not supported"),
                     ("a <= b", "only <")]:
        with pytest.raises(expr.ExpressionError, match=why):
            expr.compile_program(["a", "b"], [], bad)


def _vehicle_matrix(rng, n):
    X = np.empty((n, 6))
    X[:, 0] = rng.choice([0.15915, 0.8], n)
    X[:, 1] = rng.choice([0.27, 0.5], n)
    X[:, 2] = rng.uniform(7, 12, n)
    X[:, 3] = rng.normal(431.7221, 10, n)
    X[:, 4] = rng.normal(1475.5503, 10, n)
    X[:, 5] = rng.normal(55.0406, 10, n)
    return X


def test_demo_closures_as_text_equal_the_hand_written_limit_states():
    """demo/vehicle_suspension.jl:17-23 and demo/straub_example.jl:16-39 typed in as the closures'
    own text reproduce the hand-written oracle functors bit for bit (both stack machines)."""
    k = W.vehicle_consts()
    prog = expr.compile_program(["A", "b₀", "V", "C", "Cₖ", "K"], VEHICLE_MODELS, "df.y",
                                params={"M": k[0], "m": k[1], "g": k[2], "Mg15": k[3], "gMm": k[4]})
    X = _vehicle_matrix(np.random.default_rng(3), 100_000)
    c1, g1 = ocpu.eval_matrix(ocpu.LS_VEHICLE, k, X, want_g=True)
    c2, g2 = ocpu.eval_matrix(ocpu.LS_EXPRESSION, prog.consts, X, want_g=True)
    g3, cols = replay.g_expression(X, prog.consts, want_columns=True)
    assert c1 == c2 and np.array_equal(g1, g2) and np.array_equal(g2, g3)
    assert cols.shape[1] == prog.ncols == 11 and np.array_equal(cols[:, -1], g2)

    lam, zeta = W.lognormal_params(150.0, 30.0)
    sp = expr.compile_program(["Uᵣ", "V", "H"], [(f"r{i}", CAPACITY) for i in range(1, 6)] + [("G", FRAME)], "df.G",
                              params={"λ": lam, "ζ": zeta, "ρ": W.STRAUB_RHO})
    assert (sp.n_hidden, sp.d, sp.ncols) == (5, 8, 14)
    rng = np.random.default_rng(4)
    n = 100_000
    Xs = np.empty((n, 8))
    Xs[:, 0] = rng.normal(0, 1, n)
    Xs[:, 1] = rng.gamma(25.0, 2.4, n)
    Xs[:, 2] = rng.gumbel(40.99893584908611, 15.593936024673523, n)
    Xs[:, 3:] = rng.normal(0, 1, (n, 5))
    c1, g1 = ocpu.eval_matrix(ocpu.LS_STRAUB, W.straub_consts(), Xs, want_g=True)
    c2, g2 = ocpu.eval_matrix(ocpu.LS_EXPRESSION, sp.consts, Xs, want_g=True)
    assert c1 == c2 and np.array_equal(g1, g2)
    # NumPy's SIMD exp differs from libm's by an ulp: the NumPy twin is exact only for IEEE operations
    assert np.allclose(replay.g_expression(Xs, sp.consts), g2, rtol=0, atol=1e-10)  # terms ~1e3 cancel
    assert abs(c1 / n - 0.026129) < 0.01          # demo/straub_example.jl:71


def test_program_validation_behind_the_abi():
    ok = [0, 3, L.OP_INPUT, 0, L.OP_INPUT, 1, L.OP_SUB, 0]
    assert L.expression_check(ok, 2) == 2
    cases = [
        ([], "empty"),
        ([0, 1, L.OP_ADD, 0], "operands"),
        ([0, 2, L.OP_INPUT, 0, L.OP_INPUT, 1], "leaves 2 values"),
        ([0, 1, L.OP_INPUT, 2], "reads column 2"),
        ([0, 1, L.OP_PARAM, 0], "runtime constant"),
        ([0, 1, 99, 0], "unknown opcode"),
        ([0, 2, L.OP_INPUT, 0], "expected"),
        ([0, 2, L.OP_INPUT, 0, L.OP_POWI, 0.5], "not an integer"),
        ([0.5, 1, L.OP_INPUT, 0], "integers"),
    ]
    for prog, why in cases:
        with pytest.raises(ebn_b200.EbnError, match=why) as e:
            L.expression_check(prog, 2)
        assert e.value.code == -2
    # more stored columns than the kernels hold
    many = []
    for _ in range(L.EBN_MAX_INPUTS):
        many += [L.OP_INPUT, 0, L.OP_STORE, 0]
    many += [L.OP_INPUT, 0]
    with pytest.raises(ebn_b200.EbnError, match="exceed"):
        L.expression_check([0, len(many) // 2] + many, 2)
    src = L.expression_source(ok, 2)
    assert "struct JitExpr" in src and "Ops::sub(x[0], x[1])" in src and "XCAP = 2" in src


def test_run_time_compilation_without_a_device():
    """NVRTC compiles for sm_100a without a GPU: the generated functor and a run-time sampling plan
    compile, the second request comes from the cache, and a compiled-in job needs nothing."""
    if not L.jit_available():
        pytest.skip("libnvrtc not loadable here")
    k = W.vehicle_consts()
    prog = expr.compile_program(["A", "b₀", "V", "C", "Cₖ", "K"], VEHICLE_MODELS, "df.y",
                                params={"M": k[0], "m": k[1], "g": k[2], "Mg15": k[3], "gMm": k[4]})
    ref = W.w1_vehicle(n=1000)
    for strict in (True, False):
        job = L.Job(L.LS_EXPRESSION, prog.consts, ref.inputs, n=1000, strict=strict)
        for role in (0, 1, 2):
            assert L.jit_compile(job, role) > 10_000
            assert L.last_stats()["jit_compiles"] == 1
        assert L.jit_compile(job, 0) > 10_000 and L.last_stats()["jit_compiles"] == 0
        # other runtime constants: same kernel
        job2 = L.Job(L.LS_EXPRESSION, prog.with_params({"gMm": 1500.0}), ref.inputs, n=1000, strict=strict)
        assert L.jit_compile(job2, 0) > 10_000 and L.last_stats()["jit_compiles"] == 0
    assert L.jit_compile(ref, 0) == 0                     # compiled-in plan
    trunc = W.w1_vehicle(n=1000)
    trunc.inputs[:, 3]["lo"] = 400.0
    assert L.jit_compile(trunc, 0) == 0                   # C truncated: a compiled-in plan (discretised root)
    trunc.inputs[:, 4]["hi"] = 1500.0                     # C and Ck truncated: compiled in as well (all 8 combinations are)
    assert L.jit_compile(trunc, 0) == 0
    trunc.inputs[:, 5]["kind"], trunc.inputs[:, 5]["p"] = L.IN_LOGNORMAL, (4.0, 0.18, 0, 0)   # K log-normal: no compiled-in plan
    assert L.jit_compile(trunc, 0) == 0                   # small job: runtime-dispatched plan
    trunc.jit_plan = True
    assert L.jit_compile(trunc, 0) > 10_000               # forced: StaticPlan compiled on demand
    big = W.w1_vehicle(n=10**10)
    big.inputs[:, 3]["lo"] = 400.0
    big.inputs[:, 5]["kind"], big.inputs[:, 5]["p"] = L.IN_LOGNORMAL, (4.0, 0.18, 0, 0)
    assert L.jit_compile(big, 0) > 10_000                 # large job: automatic
    big.jit_plan = False
    assert L.jit_compile(big, 0) == 0
    mixed = W.w1_vehicle(n=10**10)
    mixed.inputs[0, 3]["lo"] = 400.0                      # kinds differ between scenarios: no static plan
    assert L.jit_compile(mixed, 0) == 0


def test_coefficient_table_functors_compile_as_formulas_when_asked():
    """The library rewrites polynomial / min-affine batches into expression programs (large batches, or
    EBN_JOB_JIT_PLAN): visible without a GPU through ebn_jit_compile."""
    if not L.jit_available():
        pytest.skip("libnvrtc not loadable here")
    rows = [[(1, 0.0, 1.0, 0, 0, -INF, INF), (3, -1.0, 2.0, 0, 0, -INF, INF)]]
    prog = np.array([2, 3, 1.0, 2, 0, 0, -0.7, 0, 1.0, 1, 1, 2, 2.0, 0, -1.0, 1, 2], dtype=float)  # x0^2 - 0.7 + x1 ; 2 - col2
    small = L.Job(L.LS_POLYNOMIAL, prog, L.make_inputs(rows), n=1000)
    assert L.jit_compile(small, 0) == 0                         # interpreted functor, runtime-dispatched plan
    asked = L.Job(L.LS_POLYNOMIAL, prog, L.make_inputs(rows), n=1000, jit_plan=True)
    assert L.jit_compile(asked, 0) > 10_000 and L.jit_compile(asked, 1) > 10_000
    large = L.Job(L.LS_POLYNOMIAL, prog, L.make_inputs(rows), n=6 * 10**10)
    assert L.jit_compile(large, 0) > 10_000
    large.jit_plan = False
    assert L.jit_compile(large, 0) == 0
    k = np.array([2.0, 0.5, 1.0, 0.0, -2.0, -0.25, 1.5])
    assert L.jit_compile(L.Job(L.LS_MIN_AFFINE, k, L.make_inputs(rows), n=1000, jit_plan=True), 0) > 10_000


def test_lowering_of_expression_models():
    from ebn_b200 import Expression, Model, Poly
    from ebn_b200.uq import UnsupportedModel, lower
    low = lower([Model(Expression("df.A .+ df.B"), "C")], Expression("lim .- df.C", lim=2.0), ["A", "B"])
    assert low.limit_state == L.LS_EXPRESSION and low.columns == ["A", "B"] and low.out_columns == ["A", "B", "C"]
    assert L.expression_check(low.consts, 2) == 3
    # Poly models mix in, with their evaluation order; a column name is a valid performance
    low2 = lower([Model(Poly([(1, ["A"]), (1, ["B"])]), "C"), Model(Expression("sqrt(abs(df.C)) * randn()"), "D")], "D", ["A", "B"])
    assert len(low2.hidden) == 1 and low2.out_columns == ["A", "B", "__randn1", "C", "D"]
    X = np.random.default_rng(0).normal(0, 1, (1000, 3))
    g = replay.g_expression(X, low2.consts)
    assert np.array_equal(g, np.sqrt(np.abs(X[:, 0] + X[:, 1])) * X[:, 2])
    # a polynomial chain can be sent down the same road (what pf_scenarios does for large jobs)
    chain = [Model(Poly([(1, ["A"]), (1, ["B"])]), "C")]
    lp = lower(chain, Poly([(2, []), (-1, ["C"])]), ["A", "B"])
    le = lower(chain, Poly([(2, []), (-1, ["C"])]), ["A", "B"], prefer_expression=True)
    assert lp.limit_state == L.LS_POLYNOMIAL and le.limit_state == L.LS_EXPRESSION and lp.out_columns == le.out_columns
    Xp = np.random.default_rng(1).normal(0, 1, (1000, 2))
    assert np.array_equal(replay.g_expression(Xp, le.consts), replay.g_polynomial(Xp, lp.consts))
    with pytest.raises(UnsupportedModel, match="two different values"):
        lower([Model(Expression("df.A * k", k=1.0), "C")], Expression("df.C - k", k=2.0), ["A"])
    with pytest.raises(UnsupportedModel, match="unknown name"):
        lower([Model(Expression("df.Z"), "C")], "C", ["A"])
    with pytest.raises(UnsupportedModel, match="Expression"):
        lower([Model(lambda df: df, "C")], "C", ["A"])


def test_kernel_cache_on_disk_is_opt_in(tmp_path):
    """EBN_JIT_CACHE=dir persists compiled kernels across processes; without it nothing is written."""
    import subprocess
    import sys
    if not L.jit_available():
        pytest.skip("libnvrtc not loadable here")
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from ebn_b200 import workloads as W, _lib as L\n"
            "j = W.w1_vehicle(n=1000); j.inputs[:, 5]['kind'] = L.IN_LOGNORMAL; j.inputs[:, 5]['p'] = (4.0, 0.18, 0, 0); j.jit_plan = True\n"
            "assert L.jit_compile(j, 0) > 10000; print(L.last_stats()['jit_compiles'])\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, EBN_JIT_CACHE=str(tmp_path))
    runs = [subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True) for _ in range(2)]
    assert [r.stdout.strip() for r in runs] == ["1", "0"], [r.stderr[-300:] for r in runs]
    assert len(list(tmp_path.iterdir())) == 1
