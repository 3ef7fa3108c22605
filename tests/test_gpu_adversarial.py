"""Adversarial parity of the code the headline bench actually times.

`bench.py` runs W1 with strict_fp = false, i.e. the fused kernel evaluates the vehicle limit state
through `VehicleSuspension<FastOps>::fails()` (csrc/ebn_limit_states.cuh): a division-free,
sqrt-free predicate that takes `g < 0` from sign bits (XOR logic for a negative D = b0*K, a
negative C, and the NaN mask of a negative sqrt argument). With the demo's marginals K < 0, C < 0
or a negative sqrt argument never occur, so the nominal tests cannot see a wrong sign rule. Here the
marginals are chosen to hit every branch of the predicate about half of the time:

    C ~ N(0, 300)   Ck ~ N(1442.64, 300)   K ~ N(0, 40)      (reference arithmetic:
    demo/vehicle_suspension.jl:17-23 — `.sqrt` of a negative argument is NaN, `minimum` propagates
    NaN, `NaN < 0` is false: the sample is safe)

For every scenario the replayed sample matrix goes through the C oracle (Julia order, no FMA); the
fused kernel's count is compared with the oracle's indicators over `sample_offset` windows, windows
that disagree are bisected down to single samples, and every sample found that way must be a tie
(|g_oracle| < 1e-12). The same is done for the fp32-sampling plans, the plan with V a Parameter, and
for a limit state compiled from text in fast mode (`JitExpr<FastOps>`, reciprocal hoisting).
"""
import dataclasses
import math

import numpy as np
import pytest

# EBN_FUZZ_SEED=k shifts the seeds of the randomised tests (extra fuzzing runs); the default is what CI runs
FUZZ_SEED = int(__import__("os").environ.get("EBN_FUZZ_SEED", "0"))

from oracle import cpu as ocpu

import formula_fuzz

pytestmark = pytest.mark.gpu

INF = math.inf
TIE = 1e-12


def _adversarial_rows(L, fixed_v=False):
    rows = []
    for A, b0 in ((0.8, 0.27), (0.8, 0.5), (0.15915, 0.27), (0.15915, 0.5)):
        v = (L.IN_PARAM, 9.75, 0, 0, 0, -INF, INF) if fixed_v else (L.IN_UNIFORM, 7.0, 12.0, 0, 0, -INF, INF)
        rows.append([
            (L.IN_PARAM, A, 0, 0, 0, -INF, INF),
            (L.IN_PARAM, b0, 0, 0, 0, -INF, INF),
            v,
            (L.IN_NORMAL, 0.0, 300.0, 0, 0, -INF, INF),        # C: both signs
            (L.IN_NORMAL, 1442.64, 300.0, 0, 0, -INF, INF),    # Ck: straddles (g (M+m))^0.877 = 1442.642...
            (L.IN_NORMAL, 0.0, 40.0, 0, 0, -INF, INF),         # K: both signs, D = b0 K changes sign
        ])
    return rows


def _explain_by_ties(ebn, job, oracle_ls, oracle_consts, window=1 << 17, both_signs=False):
    """Localises every sample on which the fused kernel's failure indicator differs from the oracle's
    on the replayed matrix. Returns (rows examined, NaN-g fraction, list of (scenario, row, g))."""
    n, S = job.n, job.S
    want, gs, nan_rows = [], [], 0
    for s in range(S):
        X = ebn.sample_matrix(job, s, job.sample_offset, n)
        _, g = ocpu.eval_matrix(oracle_ls, oracle_consts, X, want_g=True)
        want.append(np.concatenate([[0], np.cumsum(g < 0)]))   # NaN < 0 is False: safe, like the reference
        gs.append(g)
        nan_rows += int(np.isnan(g).sum())
        if both_signs:
            assert (X[:, 3] < 0).any() and (X[:, 3] > 0).any() and (X[:, 5] < 0).any() and (X[:, 5] > 0).any()

    def counts(first, m):
        piece = dataclasses.replace(job, n=m, sample_offset=job.sample_offset + first)
        return ebn.pf_batch(piece)[0]

    bad = []

    def descend(first, m, scen):
        got = counts(first, m)
        scen = [s for s in scen if int(got[s]) != int(want[s][first + m] - want[s][first])]
        if not scen:
            return
        if m == 1:
            bad.extend((s, first, float(gs[s][first])) for s in scen)
            return
        h = m // 2
        descend(first, h, scen)
        descend(first + h, m - h, scen)

    total = ebn.pf_batch(job)[0]
    for first in range(0, n, window):
        descend(first, min(window, n - first), list(range(S)))
    # the windows add up to the one-call count (resume property), so nothing hides between them
    assert np.array_equal(total, np.array([int(want[s][n]) for s in range(S)]) +
                          np.array([sum((1 if not (g < 0) else -1) for (t, _, g) in bad if t == s) for s in range(S)]))
    return S * n, nan_rows / (S * n), bad


@pytest.mark.parametrize("fp32", [False, True])
@pytest.mark.parametrize("fixed_v", [False, True])
def test_sign_only_predicate_on_adversarial_marginals(ebn, fp32, fixed_v):
    """mc_kernel<VehicleSuspension<FastOps>, StaticPlan, 0> (the kernel bench.py times) against the oracle."""
    L = ebn._lib
    from ebn_b200 import workloads as W
    n = 2_500_001          # 4 scenarios: 1e7 samples per plan
    job = L.Job(L.LS_VEHICLE_SUSPENSION, W.vehicle_consts(), L.make_inputs(_adversarial_rows(L, fixed_v)), n,
                seed=0xADBE_0001 + 2 * fp32 + fixed_v, strict=False, fp32_sampling=fp32)
    ebn.pf_batch(dataclasses.replace(job, n=1000))
    assert ebn.last_stats()["specialised"] == 1, "the compile-time sampling plan must be the one under test"
    rows, nan_frac, bad = _explain_by_ties(ebn, job, ocpu.LS_VEHICLE, job.consts, both_signs=True)
    assert nan_frac > 0.10, nan_frac
    not_ties = [(s, r, g) for (s, r, g) in bad if not abs(g) < TIE]
    assert not not_ties, f"sign-only predicate disagrees off a tie: {not_ties[:5]}"
    print(f"adversarial fails(): {rows} samples, {100 * nan_frac:.1f}% NaN g, {len(bad)} tie flips")
    # strict mode of the same job: the fused kernel evaluates g in Julia order -> no flips at all
    sjob = dataclasses.replace(job, strict=True, n=500_001)
    _, _, sbad = _explain_by_ties(ebn, sjob, ocpu.LS_VEHICLE, job.consts)
    assert not sbad, sbad[:5]


def test_sign_only_predicate_extreme_rows(ebn):
    """Degenerate marginals: tiny and huge spreads, Ck exactly on the g4 threshold's side, K and C of
    either sign with |K| down to 1e-6 — counts still equal the oracle's up to ties."""
    L = ebn._lib
    from ebn_b200 import workloads as W
    k = W.vehicle_consts()
    rows = []
    for (cmu, csd), (kmu, ksd), (ckmu, cksd) in (
            ((391.0, 1e-3), (55.0, 1e-6), (k[4], 1e-9)),      # g2 threshold C = 8.6394/(4000 (Mg)^-1.5) ~ 391.2; Ck on the g4 threshold
            ((-5.0, 5.0), (1e-6, 1e-6), (1500.0, 1.0)),
            ((1e4, 1e4), (-300.0, 500.0), (-1442.0, 3000.0)),
            ((0.5, 0.5), (0.0, 1e3), (1442.642281188081, 1e-6))):
        rows.append([(L.IN_PARAM, 0.8, 0, 0, 0, -INF, INF), (L.IN_PARAM, 0.27, 0, 0, 0, -INF, INF),
                     (L.IN_UNIFORM, 7.0, 12.0, 0, 0, -INF, INF), (L.IN_NORMAL, cmu, csd, 0, 0, -INF, INF),
                     (L.IN_NORMAL, ckmu, cksd, 0, 0, -INF, INF), (L.IN_NORMAL, kmu, ksd, 0, 0, -INF, INF)])
    job = L.Job(L.LS_VEHICLE_SUSPENSION, k, L.make_inputs(rows), 400_001, seed=77, strict=False)
    ebn.pf_batch(dataclasses.replace(job, n=1000))
    assert ebn.last_stats()["specialised"] == 1
    _, _, bad = _explain_by_ties(ebn, job, ocpu.LS_VEHICLE, k, window=1 << 16)
    # near-degenerate rows sit on a threshold by construction: ties are expected, anything else is not
    not_ties = [(s, r, g) for (s, r, g) in bad if not abs(g) < TIE]
    assert not not_ties, not_ties[:5]


def test_compiled_formula_fast_mode_on_adversarial_marginals(ebn):
    """JitExpr<FastOps>: the vehicle closures as text, compiled at run time in fast mode (constants folded,
    divisions by constants turned into multiplications) against the oracle's strict stack machine."""
    L = ebn._lib
    from ebn_b200 import expr, workloads as W
    if not L.jit_available():
        pytest.skip("no NVRTC on this machine")
    k = W.vehicle_consts()
    prog = expr.compile_program(
        ["A", "b0", "V", "C", "Ck", "K"],
        [("g1", "1 .- (pi .* m .* df.V .* df.A) ./ (df.b0 .* df.K .* g .^ 2) .* ((df.Ck ./ (m .+ M) .- (df.C ./ M)) .^ 2"
                " .+ df.C .^ 2 ./ (m .* M) .+ df.Ck .* df.K .^ 2 ./ (m .* M .^ 2))"),
         ("g2", "4000 .* df.C .* Mg15 .- 8.6394"),
         ("g3", "2 .* sqrt(M .* g .* (df.K .^ 2 .* df.Ck ./ (df.C .* (m .+ M)) .+ df.C)) .- 1"),
         ("g4", "df.Ck .- gMm")],
        "min(g1, g2, g3, g4)", params={"M": k[0], "m": k[1], "g": k[2], "Mg15": k[3], "gMm": k[4]})
    n = 1_000_001
    job = L.Job(L.LS_EXPRESSION, prog.consts, L.make_inputs(_adversarial_rows(L)), n, seed=0xADBE_0009, strict=False)
    rows, nan_frac, bad = _explain_by_ties(ebn, job, ocpu.LS_EXPRESSION, prog.consts, both_signs=True)
    assert nan_frac > 0.10
    # fast mode re-associates: every mechanism is `1 - ratio` or `x - threshold`, so a flip needs |g| ~ 1e-15
    not_ties = [(s, r, g) for (s, r, g) in bad if not abs(g) < TIE]
    assert not not_ties, not_ties[:5]
    # the compiled formula and the hand-written functor see the same samples: strict counts are identical
    sj = dataclasses.replace(job, strict=True)
    hj = L.Job(L.LS_VEHICLE_SUSPENSION, k, job.inputs, n, seed=job.seed, strict=True)
    assert np.array_equal(ebn.pf_batch(sj)[0], ebn.pf_batch(hj)[0])
    print(f"adversarial JitExpr<FastOps>: {rows} samples, {100 * nan_frac:.1f}% NaN g, {len(bad)} tie flips")


@pytest.mark.parametrize("seed", [0, 7])
def test_sign_only_predicates_generated_for_formulas(ebn, seed):
    """JitExpr<FastOps>::fails, the predicate the generator derives for `g < 0` in fast mode (no division, no
    square root where only the sign matters): 40 random chains of mechanisms under min / max over inputs that
    change sign — so denominators change sign and square-root and log arguments go negative (NaN: safe) — each
    compared sample by sample with the oracle's strict stack machine on the replayed matrix. The two may
    differ on ties only. Seed 7 holds the chain on which NVRTC miscompiled the first, bool-valued form of the
    predicate inside the fused kernel (ebn_expr.cu, note above FailsGen)."""
    L = ebn._lib
    from ebn_b200 import expr
    if not L.jit_available():
        pytest.skip("no NVRTC on this machine")
    rng = np.random.default_rng(77001 + seed + FUZZ_SEED)
    rows = [[(L.IN_NORMAL, 0.3, 1.5, 0, 0, -INF, INF), (L.IN_NORMAL, -0.2, 1.0, 0, 0, -INF, INF),
             (L.IN_UNIFORM, -1.0, 2.0, 0, 0, -INF, INF), (L.IN_NORMAL, 1.0, 0.7, 0, 0, -INF, INF)]]
    names = ["x0", "x1", "x2", "x3"]
    total_nan, total_fail, total = 0.0, 0, 0
    for trial in range(40):
        models, perf, consts = formula_fuzz.chain(rng, names)
        prog = expr.compile_program(names, models, perf, consts)
        n = 150_001
        job = L.Job(L.LS_EXPRESSION, prog.consts, L.make_inputs(rows), n, seed=0x51C0 + trial, strict=False)
        _, nan_frac, bad = _explain_by_ties(ebn, job, ocpu.LS_EXPRESSION, prog.consts, window=1 << 15)
        not_ties = []
        for (s, r, g) in bad:   # only samples whose sign is a matter of rounding may differ (ties, cancellation)
            x = ebn.sample_matrix(job, s, job.sample_offset + r, 1)
            oracle_g = lambda Y: ocpu.eval_matrix(ocpu.LS_EXPRESSION, prog.consts, Y, want_g=True)[1]  # noqa: E731
            if not formula_fuzz.unstable_rows(oracle_g, x, np.array([g]))[0]:
                not_ties.append((s, r, g))
        assert not not_ties, (trial, models, perf, not_ties[:5])
        total_nan += nan_frac
        total_fail += int(ebn.pf_batch(job)[0][0])
        total += n
    print(f"generated predicates: {total} samples, {100 * total_nan / 40:.1f}% NaN g, pf over all formulas {total_fail / total:.3f}")
    assert total_nan / 40 > 0.05 and 0.05 < total_fail / total < 0.95
