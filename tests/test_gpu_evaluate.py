"""GPU tests of `evaluate!` through the host mirror of the reference API — the reference's own
hot-path tests restated (test/networks/ebn/evaluate/evaluate_node.jl, evaluate_net.jl,
test/inference/variableselimination.jl:133-206, demo/straub_example.jl, demo/vehicle_suspension.jl),
with the Monte Carlo running in libebncuda. Tolerances are the reference's (atol 0.02 … 0.2) or
tighter where the value is analytic."""
import math
import warnings

import numpy as np
import pytest
from scipy.special import ndtr

pytestmark = pytest.mark.gpu

INF = math.inf


@pytest.fixture(scope="module")
def E(ebn):
    return ebn


def _root(E, name, states, probs, params=None):
    cpt = E.DiscreteConditionalProbabilityTable(name)
    for s, p in zip(states, probs):
        cpt[(name, s)] = p
    return E.DiscreteNode(name, cpt, params or {})


def _cont_root(E, name, dist, disc=None, kind="precise"):
    cpt = E.ContinuousConditionalProbabilityTable(kind=kind)
    cpt[()] = dist
    return E.ContinuousNode(name, cpt, disc)


def _ab_roots(E):
    a = _root(E, "A", ["a1", "a2"], [0.5, 0.5], {"a1": [E.Parameter(1, "A")], "a2": [E.Parameter(2, "A")]})
    return a, _cont_root(E, "B", E.Normal())


def test_evaluate_discrete_functional_node(E):
    """evaluate_node.jl:94-157."""
    from ebn_b200 import evaluation as ev
    a, b = _ab_roots(E)
    sim = E.MonteCarlo(100_000)
    f = E.DiscreteFunctionalNode("C", [E.Model(E.Poly([(1, ["A"]), (1, ["B"])]), "C")], E.Poly([(2, []), (-1, ["C"])]), sim)
    net = E.EnhancedBayesianNetwork([a, b, f])
    E.add_child(net, a, f)
    E.add_child(net, b, f)
    E.order(net)
    out = ev._evaluate_node(net, f)
    assert out.name == "C" and isinstance(out, E.DiscreteNode)
    assert out.cpt.column("A") == ["a1", "a1", "a2", "a2"]
    assert out.cpt.column("C") == ["C_fail", "C_safe", "C_fail", "C_safe"]
    assert np.allclose(out.cpt.column("Π"), [0.16, 0.84, 0.5, 0.5], atol=0.02)
    assert out.parameters == f.parameters
    for key in (("a1",), ("a2",)):
        info = out.additional_info[key]
        assert info["samples"].shape == (sim.n, 3) and isinstance(info["cov"], float)
    rows = out.additional_info[("a1",)]["samples"].fetch(0, 1000)
    assert set(rows) == {"A", "B", "C", "g"} and np.all(rows["A"] == 1.0)
    assert np.allclose(rows["g"], 2 - (rows["A"] + rows["B"]))
    # pf and the replayed frame agree exactly: pf*n == #(g < 0) over the whole stream
    allrows = out.additional_info[("a1",)]["samples"].fetch(0, sim.n)
    assert int(np.sum(allrows["g"] < 0)) == round(out.cpt[{"A": "a1", "C": "C_fail"}] * sim.n)
    # no discrete ancestor (evaluate_node.jl:123-139)
    f2 = E.DiscreteFunctionalNode("C", [E.Model(E.Poly([(1, ["B"]), (1, [])]), "C")], E.Poly([(2, []), (-1, ["C"])]), sim)
    net = E.EnhancedBayesianNetwork([b, f2])
    E.add_child(net, b, f2)
    E.order(net)
    out = ev._evaluate_node(net, f2)
    assert out.isroot() and out.cpt.column("C") == ["C_fail", "C_safe"]
    assert np.allclose(out.cpt.column("Π"), [0.1587, 0.8413], atol=0.01)
    assert ev._evaluate_node(net, f2, False).additional_info == {}


def test_evaluate_continuous_functional_node(E):
    """evaluate_node.jl:1-62: the evaluated node is a ContinuousNode; KDE replaced by a histogram."""
    from ebn_b200 import evaluation as ev
    a, b = _ab_roots(E)
    sim = E.MonteCarlo(200_000)
    disc = E.ApproximatedDiscretization([-1, 0, 1], 2)
    cf = E.ContinuousFunctionalNode("C", [E.Model(E.Poly([(1, ["A"]), (1, ["B"])]), "C")], sim, disc)
    net = E.EnhancedBayesianNetwork([a, b, cf])
    E.add_child(net, a, cf)
    E.add_child(net, b, cf)
    E.order(net)
    out = ev._evaluate_node(net, cf)
    assert isinstance(out, E.ContinuousNode) and out.name == "C" and out.discretization == disc
    assert out.cpt.column("A") == ["a1", "a2"]
    d1, d2 = out.cpt.column("Π")
    assert isinstance(d1, E.EmpiricalDistribution)
    assert abs(d1.mean() - 1) < 0.01 and abs(d2.mean() - 2) < 0.01 and abs(d1.std() - 1) < 0.01
    for x in (-0.5, 0.0, 1.0, 2.5):
        assert abs(float(d1.cdf(x)) - ndtr(x - 1)) < 5e-3
    assert out.additional_info[("a1",)]["samples"].shape == (sim.n, 3)
    assert ev._evaluate_node(net, cf, False).additional_info == {}
    # root case → ExactDiscretization (evaluate_node.jl:12-14,60)
    cf2 = E.ContinuousFunctionalNode("C", [E.Model(E.Poly([(1, ["B"]), (1, [])]), "C")], sim, disc)
    net = E.EnhancedBayesianNetwork([b, cf2])
    E.add_child(net, b, cf2)
    E.order(net)
    out = ev._evaluate_node(net, cf2)
    assert out.discretization == E.ExactDiscretization(disc.intervals) and out.isroot()
    # imprecise parent → the reference's error (evaluate_node.jl:37, test :65-92)
    r2 = _cont_root(E, "Bi", (0.1, 0.3), kind="interval")
    cf3 = E.ContinuousFunctionalNode("C", [E.Model(E.Poly([(1, ["A"]), (1, ["Bi"])]), "C")], sim)
    net = E.EnhancedBayesianNetwork([a, r2, cf3])
    E.add_child(net, a, cf3)
    E.add_child(net, r2, cf3)
    E.order(net)
    with pytest.raises(ValueError, match="No method for extracting failure probability p-box"):
        ev._evaluate_node(net, cf3)


def test_form_branch(E):
    """test/networks/ebn/evaluate/evaluate_node.jl:141-158: FORM on C = 1 + B, performance 2 - C, B ~ N(0,1):
    Π ≈ [0.1587, 0.8413] (atol 0.1 in the reference; a linear limit state makes FORM exact: β = 1)."""
    b = _cont_root(E, "B", E.Normal())
    node = E.DiscreteFunctionalNode("C", [E.Model(E.Expression("1 .+ df.B"), "C")], E.Expression("2 .- df.C"), E.FORM())
    net = E.EnhancedBayesianNetwork([b, node])
    E.add_child(net, b, node)
    E.order(net)
    from ebn_b200.evaluation import _evaluate_node
    ev = _evaluate_node(net, node)
    assert isinstance(ev, E.DiscreteNode) and ev.isroot() and ev.name == "C"
    assert ev.cpt.column("C") == ["C_fail", "C_safe"]
    pi = ev.cpt.column("Π")
    assert abs(pi[0] - ndtr(-1.0)) < 1e-9 and pi[1] == 1 - pi[0]
    info = ev.additional_info[()]
    assert abs(info["β"] - 1.0) < 1e-9 and abs(info["design_point"]["B"] - 1.0) < 1e-6 and abs(info["α"]["B"] - 1.0) < 1e-9
    # curved limit state with a non-normal marginal: against a Monte Carlo estimate on the same functor
    a, b2 = _ab_roots(E)
    u = _cont_root(E, "U", E.Gumbel(1.0, 0.5))
    models = [E.Model(E.Expression("df.A .* df.B .+ df.U .^ 2 ./ 4"), "C")]
    perf = E.Expression("4.5 .- df.C")
    nf = E.DiscreteFunctionalNode("C", models, perf, E.FORM(n=30, tol=1e-8))
    nm = E.DiscreteFunctionalNode("C", models, perf, E.MonteCarlo(4_000_000))
    out = []
    for nd in (nf, nm):
        net = E.EnhancedBayesianNetwork([a, b2, u, nd])
        for p in (a, b2, u):
            E.add_child(net, p, nd)
        E.order(net)
        out.append(_evaluate_node(net, nd))
    pf_form = [r["Π"] for r in out[0].cpt.data if r["C"] == "C_fail"]
    pf_mc = [r["Π"] for r in out[1].cpt.data if r["C"] == "C_fail"]
    assert len(pf_form) == 2
    for f, m in zip(pf_form, pf_mc):     # first-order approximation: right order of magnitude, not exact
        assert 0.5 < f / m < 2.0, (f, m)
    assert set(out[0].additional_info) == {("a1",), ("a2",)} and set(out[0].additional_info[("a1",)]) == {"β", "design_point", "α"}


def _imprecise_net(E, sim):
    a = _root(E, "A", ["a1", "a2"], [0.5, 0.5], {"a1": [E.Parameter(1, "A")], "a2": [E.Parameter(2, "A")]})
    b = _cont_root(E, "B", (0.10, 1.30), kind="interval")
    p = _cont_root(E, "P", E.Uniform(-10, 10))
    f = E.DiscreteFunctionalNode("C", [E.Model(E.Poly([(1, ["A"]), (1, ["B"]), (1, ["P"])]), "C")],
                                 E.Poly([(2, []), (-1, ["C"])]), sim)
    net = E.EnhancedBayesianNetwork([a, b, p, f])
    for par in (a, b, p):
        E.add_child(net, par, f)
    E.order(net)
    return net, f


@pytest.mark.parametrize("method", ["optimise", "grid"])
def test_double_loop_interval_parent(E, method):
    """evaluate_node.jl:159-204: pf = (8 + A + B)/20 over B in [0.10, 1.30]."""
    from ebn_b200 import evaluation as ev
    net, f = _imprecise_net(E, E.DoubleLoop(E.MonteCarlo(100_000), method=method))
    out = ev._evaluate_node(net, f)
    assert not out.isprecise() and out.additional_info[("a1",)] == {} and out.additional_info[("a2",)] == {}
    assert out.cpt.column("A") == ["a1", "a1", "a2", "a2"]
    assert out.cpt.column("C") == ["C_fail", "C_safe", "C_fail", "C_safe"]
    want = [(0.45191, 0.51868), (0.48132, 0.54809), (0.50418, 0.56792), (0.43208, 0.49582)]
    for got, w in zip(out.cpt.column("Π"), want):
        assert np.allclose(got, w, atol=0.2)       # the reference's tolerance
    exact = [(0.455, 0.515), (0.485, 0.545), (0.505, 0.565), (0.435, 0.495)]
    for got, w in zip(out.cpt.column("Π"), exact):
        assert np.allclose(got, w, atol=0.006)     # analytic value, 3 SE at n = 1e5 is 0.0047


def test_random_slicing_and_child_intervals(E):
    """evaluate_node.jl:206-282."""
    from ebn_b200 import evaluation as ev
    sim = E.RandomSlicing(E.MonteCarlo(100_000))
    net, f = _imprecise_net(E, sim)
    out = ev._evaluate_node(net, f)
    for key in (("a1",), ("a2",)):
        lb, ub = out.additional_info[key]["lb"], out.additional_info[key]["ub"]
        assert isinstance(lb[0], float) and isinstance(ub[0], float)
    exact = [(0.455, 0.515), (0.485, 0.545), (0.505, 0.565), (0.435, 0.495)]
    for got, w in zip(out.cpt.column("Π"), exact):
        assert np.allclose(got, w, atol=0.006)
    # child with interval rows per parent state (evaluate_node.jl:233-282)
    a = _root(E, "A", ["a1", "a2"], [0.5, 0.5], {"a1": [E.Parameter(1, "A")], "a2": [E.Parameter(2, "A")]})
    b = _cont_root(E, "B", E.Normal())
    d = _root(E, "D", ["d1", "d2"], [0.5, 0.5], {"d1": [E.Parameter(1, "D")], "d2": [E.Parameter(2, "D")]})
    cptc = E.ContinuousConditionalProbabilityTable(["A"], kind="interval")
    cptc[("A", "a1")], cptc[("A", "a2")] = (0.1, 0.3), (0.7, 0.8)
    c1 = E.ContinuousNode("C1", cptc)
    fn = E.DiscreteFunctionalNode("F1", [E.Model(E.Poly([(1, ["D"]), (1, ["C1"]), (1, ["B"])]), "C2")],
                                  E.Poly([(2, []), (-1, ["C2"])]), sim)
    net = E.EnhancedBayesianNetwork([a, b, d, c1, fn])
    E.add_child(net, a, c1)
    for par in (b, d, c1):
        E.add_child(net, par, fn)
    E.order(net)
    out = ev._evaluate_node(net, fn)
    assert out.name == "F1"
    assert out.cpt.column("A") == ["a1", "a1", "a2", "a2"] * 2
    assert out.cpt.column("D") == ["d1"] * 4 + ["d2"] * 4
    assert out.cpt.column("F1") == ["F1_fail", "F1_safe"] * 4
    want = [(0.198, 0.2406), (0.759, 0.802), (0.371, 0.4248), (0.5752, 0.629), (0.537, 0.6224), (0.377, 0.463),
            (0.749, 0.7869), (0.213, 0.251)]
    for got, w in zip(out.cpt.column("Π"), want):
        assert np.allclose(got, w, atol=0.2)
    # analytic: pf = Phi(D + C1 - 2)
    ex = {("d1", "a1"): (ndtr(-0.9), ndtr(-0.7)), ("d1", "a2"): (ndtr(-0.3), ndtr(-0.2)),
          ("d2", "a1"): (ndtr(0.1), ndtr(0.3)), ("d2", "a2"): (ndtr(0.7), ndtr(0.8))}
    for (dd, aa), w in ex.items():
        assert np.allclose(out.cpt[{"D": dd, "A": aa, "F1": "F1_fail"}], w, atol=0.006)
    # the same node with its closures as text: monotonicity is then checked numerically on replayed
    # samples; a limit state that is not monotone in the interval input is refused
    fe = E.DiscreteFunctionalNode("F1", [E.Model(E.Expression("df.D .+ df.C1 .+ df.B"), "C2")], E.Expression("2 .- df.C2"), sim)
    net = E.EnhancedBayesianNetwork([a, b, d, c1, fe])
    E.add_child(net, a, c1)
    for par in (b, d, c1):
        E.add_child(net, par, fe)
    E.order(net)
    oute = ev._evaluate_node(net, fe)
    for (dd, aa), w in ex.items():
        assert np.allclose(oute.cpt[{"D": dd, "A": aa, "F1": "F1_fail"}], w, atol=0.006)
    # a limit state that is NOT monotone in the interval input (refused until round 2): the per-sample
    # extremes over the box are found inside the kernel; the bounds equal a brute-force oracle on the
    # replayed samples (fine grid over the box, NumPy) and contain every fixed-value probability
    fbad = E.DiscreteFunctionalNode("F1", [E.Model(E.Expression("df.D .+ (df.C1 .- 0.2) .^ 2 .+ df.B"), "C2")],
                                    E.Expression("2 .- df.C2"), sim)
    net = E.EnhancedBayesianNetwork([a, b, d, c1, fbad])
    E.add_child(net, a, c1)
    for par in (b, d, c1):
        E.add_child(net, par, fbad)
    E.order(net)
    outb = ev._evaluate_node(net, fbad)
    boxes = {"a1": (0.1, 0.3), "a2": (0.7, 0.8)}
    for dd, dval in (("d1", 1.0), ("d2", 2.0)):
        for aa, (lo, hi) in boxes.items():
            lb, ub = outb.cpt[{"D": dd, "A": aa, "F1": "F1_fail"}]
            grid = np.linspace(lo, hi, 201)
            gext = dval + (grid - 0.2) ** 2                     # 2 - (D + (C1-0.2)^2 + B) < 0  <=>  B > 2 - gext
            want_ub, want_lb = ndtr(-(2 - gext.max())), ndtr(-(2 - gext.min()))
            assert lb <= ub and abs(lb - want_lb) < 0.006 and abs(ub - want_ub) < 0.006, (dd, aa, lb, ub, want_lb, want_ub)


def test_random_slicing_general_pbox_two_parameters_against_brute_force(E):
    """RandomSlicing without the monotone / one-parameter restrictions (evaluate_node.jl:111-117):
    a Normal p-box with interval mean AND interval standard deviation next to an interval input, a model
    that is non-monotone in both. Oracle: the job's own replayed standard variates, per sample the p-box
    slice [min, max] over the parameter vertices, then min / max of g over a 41 x 41 grid of the box."""
    from ebn_b200 import uq
    sim = E.RandomSlicing(E.MonteCarlo(200_000, seed=77))
    models = [E.Model(E.Expression("(df.X .- 1.0) .^ 2 .+ df.Y .* df.Y .- df.Z"), "M")]
    perf = E.Expression("1.5 .- df.M")
    scen = []
    for ylo, yhi in ((-0.5, 0.25), (0.2, 0.9)):
        scen.append([E.ProbabilityBox(E.Normal, (E.Interval(0.8, 1.3, "mu"), E.Interval(0.5, 0.9, "sd")), "X"),
                     E.Interval(ylo, yhi, "Y"), E.RandomVariable(E.Normal(0.0, 0.3), "Z")])
    res = uq.random_slicing(models, perf, scen, sim)
    for (bounds, (sd_lb, smp_lb), (sd_ub, smp_ub)), ins in zip(res, scen):
        lb, ub = bounds
        rows = smp_lb.fetch(0, 20_000)
        s_x, z = rows["X__s"], rows["Z"]
        qs = [mu + sg * s_x for mu in (0.8, 1.3) for sg in (0.5, 0.9)]
        xlo, xhi = np.min(qs, axis=0), np.max(qs, axis=0)
        lam = np.linspace(0, 1, 41)
        gmin = np.full_like(z, np.inf)
        gmax = np.full_like(z, -np.inf)
        for lx in lam:
            x = xlo + lx * (xhi - xlo)
            for y in np.linspace(ins[1].lb, ins[1].ub, 41):
                g = 1.5 - ((x - 1.0) ** 2 + y * y - z)
                gmin, gmax = np.minimum(gmin, g), np.maximum(gmax, g)
        brute_ub, brute_lb = np.mean(gmin < 0), np.mean(gmax < 0)
        # the device probes ends and middle of every interval: an inner approximation of the brute-force box
        # (exact when the extreme sits at a probed point); compare on the SAME 20000 samples through g
        dev = smp_ub.fetch(0, 20_000)["g"]
        dev_l = smp_lb.fetch(0, 20_000)["g"]
        assert np.all(dev >= gmin - 1e-9) and np.all(dev_l <= gmax + 1e-9)
        assert abs(np.mean(dev < 0) - brute_ub) < 0.02 and abs(np.mean(dev_l < 0) - brute_lb) < 0.02
        assert lb <= ub and abs(ub - brute_ub) < 0.03 and abs(lb - brute_lb) < 0.03
        assert sd_lb >= 0 and sd_ub >= 0
    # Gumbel p-boxes slice through the uniform variate; a truncated p-box is refused with a clear message
    g = [E.ProbabilityBox(E.Gumbel, (E.Interval(0.0, 0.5, "mu"), E.Interval(1.0, 1.2, "beta")), "X"),
         E.Interval(0.0, 0.1, "Y"), E.RandomVariable(E.Normal(0.0, 0.3), "Z")]
    (lbg, ubg), _, _ = uq.random_slicing(models, perf, [g], sim)[0]
    assert 0.0 <= lbg <= ubg <= 1.0
    t = [E.ProbabilityBox(E.Normal, (E.Interval(0.8, 1.3, "mu"), E.Interval(0.5, 0.9, "sd")), "X", lb=0.0),
         E.Interval(0.0, 0.1, "Y"), E.RandomVariable(E.Normal(0.0, 0.3), "Z")]
    with pytest.raises(E.UnsupportedModel, match="untruncated"):
        uq.random_slicing(models, perf, [t], sim)


def test_evaluate_net_chain_and_error_messages(E):
    """evaluate_net.jl: a continuous functional node with children is fused, one with a discretisation
    is evaluated then discretised; wrong simulation type → the reference's two fixed messages."""
    a, b = _ab_roots(E)
    sim = E.MonteCarlo(20_000)
    # B -> C1 (fused away) -> C2 (discretised: evaluated to a histogram, then split into C2_d + C2) -> F,
    # the flow of test/inference/variableselimination.jl:207-284 (Straub with R4, R5 discretised)
    c1 = E.ContinuousFunctionalNode("C1", [E.Model(E.Poly([(1, ["B"]), (1, [])]), "C1")], sim)
    c2 = E.ContinuousFunctionalNode("C2", [E.Model(E.Poly([(1, ["C1"]), (0.5, ["B"])]), "C2")], sim,
                                    E.ApproximatedDiscretization([0.0, 1.0, 2.0, 3.0], 1.5))
    f = E.DiscreteFunctionalNode("F", [E.Model(E.Poly([(1, ["C2"]), (1, ["A"])]), "G")], E.Poly([(2.5, []), (-1, ["G"])]), sim)
    net = E.EnhancedBayesianNetwork([a, b, c1, c2, f])
    E.add_child(net, b, c1)
    E.add_child(net, c1, c2)
    E.add_child(net, b, c2)
    E.add_child(net, c2, f)
    E.add_child(net, a, f)
    E.order(net)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        E.evaluate(net, True, False)
    names = sorted(n.name for n in net.nodes)
    assert names == ["A", "B", "C2", "C2_d", "F"] and all(not isinstance(n, (E.DiscreteFunctionalNode, E.ContinuousFunctionalNode)) for n in net.nodes)
    c2d = net.node("C2_d")
    assert c2d.isroot() and abs(sum(c2d.cpt.column("Π")) - 1) < 1e-9
    # C2 = 1 + 1.5 B: P(C2 in [1, 2]) = Phi(2/3) - Phi(0)
    assert abs(c2d.cpt[("C2_d", "[1.0, 2.0]")] - (ndtr(2 / 3) - 0.5)) < 0.012
    fnode = net.node("F")
    assert sorted(fnode.cpt.names[:-1]) == ["A", "C2_d"] and len(fnode.cpt.data) == 2 * 2 * len(c2d.states())
    pf = {(r["A"], r["C2_d"]): r["Π"] for r in fnode.cpt.data if r["F"] == "F_fail"}
    # g = 2.5 - (C2 + A). The evaluated C2 is a ROOT continuous node, so its bins carry
    # truncated(EmpiricalDistribution, a, b) (discretize.jl:53-55), shipped to the device as inverse-CDF
    # tables: with C2 ~ N(1, 1.5), P(C2 > 1.5 | [1, 2]) and P(C2 > 0.5 | [0, 1]) are analytic.
    p_a1 = (ndtr(1 / 1.5) - ndtr(0.5 / 1.5)) / (ndtr(1 / 1.5) - 0.5)
    p_a2 = (0.5 - ndtr(-0.5 / 1.5)) / (0.5 - ndtr(-1 / 1.5))
    assert pf[("a1", "[2.0, 3.0]")] == 1.0 and pf[("a1", "[0.0, 1.0]")] == 0.0 and abs(pf[("a1", "[1.0, 2.0]")] - p_a1) < 0.02
    assert pf[("a2", "[1.0, 2.0]")] == 1.0 and abs(pf[("a2", "[0.0, 1.0]")] - p_a2) < 0.02
    assert all(not n.additional_info for n in net.nodes if isinstance(n, E.DiscreteNode))
    # precise simulation with an imprecise parent (evaluate_net.jl:268-281)
    net, f = _imprecise_net(E, E.MonteCarlo(100))
    with pytest.raises(ValueError, match=r"has .* as simulation technique, but have \['B'\] as imprecise parent/s. "
                                         "DoubleLoop or RandomSlicing technique must be employeed instead"):
        E.evaluate(net)
    # DoubleLoop without imprecise parents (evaluate_net.jl:218-266)
    a, b = _ab_roots(E)
    f = E.DiscreteFunctionalNode("C", [E.Model(E.Poly([(1, ["A"]), (1, ["B"])]), "C")], E.Poly([(2, []), (-1, ["C"])]),
                                 E.DoubleLoop(E.MonteCarlo(100)))
    net = E.EnhancedBayesianNetwork([a, b, f])
    E.add_child(net, a, f)
    E.add_child(net, b, f)
    E.order(net)
    with pytest.raises(ValueError, match="A prices simulation technique must be selected!"):
        E.evaluate(net)
    # a model with no device functor is an error, never a fallback
    f = E.DiscreteFunctionalNode("C", [E.Model(lambda df: df.A + df.B, "C")], "C", E.MonteCarlo(100))
    net = E.EnhancedBayesianNetwork([a, b, f])
    E.add_child(net, a, f)
    E.add_child(net, b, f)
    E.order(net)
    with pytest.raises(E.UnsupportedModel, match="no CPU fallback"):
        E.evaluate(net)


def test_no_ancestors_case(E):
    """evaluate_net.jl:218-244 ("No Ancestors case"): a functional node whose only parent is a continuous root has
    one (empty) scenario; the evaluated node is a root of the reduced net."""
    for kind in ("continuous", "discrete"):
        b = _cont_root(E, "B", E.Normal())
        model = E.Model(E.Expression("df.B .+ 1"), "C")
        sim = E.MonteCarlo(100_000)
        node = (E.ContinuousFunctionalNode("C", [model], sim) if kind == "continuous"
                else E.DiscreteFunctionalNode("C", [model], E.Expression("2 .- df.C"), sim))
        net = E.EnhancedBayesianNetwork([b, node])
        E.add_child(net, b, node)
        E.order(net)
        E.evaluate(net)
        last = net.nodes[-1]
        assert isinstance(last, E.ContinuousNode if kind == "continuous" else E.DiscreteNode) and last.isroot()
        if kind == "discrete":   # P(2 - (B + 1) < 0) = P(B > 1)
            assert last.cpt.column("C") == ["C_fail", "C_safe"]
            assert abs(last.cpt[("C", "C_fail")] - ndtr(-1.0)) < 0.005
        else:                    # C = B + 1 ~ N(1, 1)
            d = last.cpt[()]
            assert abs(d.mean() - 1.0) < 0.02 and abs(d.cdf(1.0) - 0.5) < 0.01


def test_imprecise_node_with_discretization_messages(E):
    """evaluate_net.jl:246-281: a p-box parent that is discretised stops being imprecise, so a DoubleLoop on its
    child is refused; an interval parent under a precise simulation is refused the other way round."""
    a = _root(E, "A", ["y", "n"], [0.5, 0.5])
    cpt = E.ContinuousConditionalProbabilityTable(["A"], kind="pbox")
    cpt[("A", "y")] = E.UnamedProbabilityBox(E.Normal, (E.Interval(-0.5, 0.5, "μ"), E.Interval(1, 2, "σ")))
    cpt[("A", "n")] = E.UnamedProbabilityBox(E.Normal, (E.Interval(1, 2, "μ"), E.Interval(1, 2, "σ")))
    b = E.ContinuousNode("B", cpt, E.ApproximatedDiscretization([-1, 0, 1], 2))
    model = E.Model(E.Expression("df.B .+ 1"), "C")
    f = E.DiscreteFunctionalNode("C", [model], E.Expression("2 .- df.C"), E.DoubleLoop(E.MonteCarlo(100_000)))
    net = E.EnhancedBayesianNetwork([a, b, f])
    E.add_child(net, a, b)
    E.add_child(net, b, f)
    E.order(net)
    with pytest.raises(ValueError, match="node C has as simulation .*DoubleLoop.*but its imprecise parents will be "
                                         "discretized and approximated with Uniform and Exponential assumption, therefore "
                                         "are no longer imprecise. A prices simulation technique must be selected!"):
        E.evaluate(net)
    b = _cont_root(E, "B", (-1, 1), kind="interval")
    f = E.DiscreteFunctionalNode("C", [model], E.Expression("2 .- df.C"), E.MonteCarlo(100_000))
    net = E.EnhancedBayesianNetwork([b, f])
    E.add_child(net, b, f)
    E.order(net)
    with pytest.raises(ValueError, match=r"node C has .* as simulation technique, but have \['B'\] as imprecise parent/s. "
                                         "DoubleLoop or RandomSlicing technique must be employeed instead"):
        E.evaluate(net)


def straub_net(E, n, as_text=False, strict=False):
    """demo/straub_example.jl in the current API (test/inference/variableselimination.jl:133-206).
    as_text: the model closures typed in as their own text (Expression) instead of catalogue entries."""
    lam, zeta = E.distribution_parameters(150, 30, "LogNormal")
    ur = _cont_root(E, "Uᵣ", E.Normal())
    v = _cont_root(E, "V", E.Gamma(*E.distribution_parameters(60, 12, "Gamma")))
    h = _cont_root(E, "H", E.Gumbel(*E.distribution_parameters(50, 20, "Gumbel")))
    sim = E.MonteCarlo(n, strict=strict)
    if as_text:
        cap = E.Expression("exp(rand(Normal(λ + ρ * ζ * df.Uᵣ, sqrt((1 - ρ^2) * ζ^2))))", λ=lam, ζ=zeta, ρ=0.5477)
        frm = E.Expression("minimum([df.r1 + df.r2 + df.r4 + df.r5 - 5 * df.H, df.r2 + 2 * df.r3 + df.r4 - 5 * df.V, "
                           "df.r1 + 2 * df.r3 + 2 * df.r4 + df.r5 - 5 * df.H - 5 * df.V])")
        perf = E.Expression("df.G")
    else:
        cap = E.Catalogue("straub_capacity", (lam, zeta, 0.5477), ("Uᵣ",))
        frm = E.Catalogue("straub_frame", (), ("r1", "r2", "r3", "r4", "r5", "V", "H"))
        perf = "G"
    rs = [E.ContinuousFunctionalNode(f"R{i}", [E.Model(cap, f"r{i}")], sim) for i in range(1, 6)]
    frame = E.DiscreteFunctionalNode("E", [E.Model(frm, "G")], perf, sim)
    net = E.EnhancedBayesianNetwork([ur, v, h] + rs + [frame])
    for r in rs:
        E.add_child(net, ur, r)
        E.add_child(net, r, frame)
    E.add_child(net, v, frame)
    E.add_child(net, h, frame)
    E.order(net)
    return net


def test_straub_end_to_end(E):
    """demo/straub_example.jl:71: Π ≈ [0.026129, 0.973871] atol 0.01 at n = 10^6."""
    net = straub_net(E, 10**6)
    E.evaluate(net)
    e = net.nodes[-1]
    assert e.name == "E" and sorted(n.name for n in net.nodes) == ["E", "H", "Uᵣ", "V"]
    assert np.allclose(e.cpt.column("Π"), [0.026129, 0.973871], atol=0.01)
    assert abs(e.cpt.column("Π")[0] - 0.0258) < 0.001
    bn = E.dispatch(net)
    assert isinstance(bn, E.BayesianNetwork) and [n.name for n in bn.nodes] == ["E"]


VEHICLE_TEXT = [  # demo/vehicle_suspension.jl:17-23, one model per line of the closure
    ("g1", "1 .- (pi .* m .* df.V .* df.A) ./ (df.b₀ .* df.K .* g .^ 2) .* ((df.Cₖ ./ (m .+ M) .- (df.C ./ M)) .^ 2"
           " .+ df.C .^ 2 ./ (m .* M) .+ df.Cₖ .* df.K .^ 2 ./ (m .* M .^ 2))"),
    ("g2", "4000 .* df.C .* Mg15 .- 8.6394"),
    ("g3", "2 .* sqrt.(M .* g .* (df.K .^ 2 .* df.Cₖ ./ (df.C .* (m .+ M)) .+ df.C)) .- 1"),
    ("g4", "df.Cₖ .- gMm"),
    ("y", "minimum([df.g1[1], df.g2, df.g3, df.g4[1]])"),
]


def vehicle_net(E, n, v_dist=None, sim=None, as_text=False):
    """demo/vehicle_suspension.jl restated in the current API (the demo file itself is stale)."""
    M, m, g = 3.2633, 0.8158, 981.0
    consts = (M, m, g, (M * g) ** -1.5, (g * (M + m)) ** 0.877)
    A = _root(E, "A", ["road", "offroad"], [0.7, 0.3],
              {"road": [E.Parameter(0.15915, "A")], "offroad": [E.Parameter(0.8, "A")]})
    b0 = _root(E, "b₀", ["normal_load", "over_load"], [0.7, 0.3],
               {"normal_load": [E.Parameter(0.27, "b₀")], "over_load": [E.Parameter(0.5, "b₀")]})
    if v_dist is None:
        V = _cont_root(E, "V", E.Uniform(7, 12), E.ExactDiscretization([8.5, 9.5, 10.5, 11.5]))
    else:
        V = _cont_root(E, "V", v_dist, E.ExactDiscretization([8.5, 9.5, 10.5, 11.5]), kind="interval")
    C = _cont_root(E, "C", E.Normal(431.7221, 10))
    Ck = _cont_root(E, "Cₖ", E.Normal(1475.5503, 10))
    K = _cont_root(E, "K", E.Normal(55.0406, 10))
    from ebn_b200 import uq
    if as_text:
        k = dict(M=M, m=m, g=g, Mg15=consts[3], gMm=consts[4])
        models = [E.Model(E.Expression(text, **k), name) for name, text in VEHICLE_TEXT]
    else:
        models = [E.Model(E.Catalogue("vehicle_suspension", consts, uq.VEHICLE_ARGS), "y")]
    node = E.DiscreteFunctionalNode("E", models, "y", sim or E.MonteCarlo(n))
    net = E.EnhancedBayesianNetwork([A, b0, V, C, Ck, K, node])
    for p in (A, b0, V, C, Ck, K):
        E.add_child(net, p, node)
    E.order(net)
    return net


def test_vehicle_suspension_network(E):
    """BASELINE config 2 through the node API: 20 scenarios, pf per BASELINE.md §2 / the oracle."""
    net = vehicle_net(E, 4_000_000)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        E.evaluate(net, True, False)
    e = net.node("E")
    assert e.cpt.names == ["A", "b₀", "V_d", "E"] and len(e.cpt.data) == 40
    pf = {(r["A"], r["b₀"], r["V_d"]): r["Π"] for r in e.cpt.data if r["E"] == "E_fail"}
    want = {"[7.0, 8.5]": 0.0023, "[8.5, 9.5]": 0.0431, "[9.5, 10.5]": 0.2151, "[10.5, 11.5]": 0.5164, "[11.5, 12.0]": 0.7373}
    for vbin, w in want.items():
        got = pf[("offroad", "normal_load", vbin)]
        assert abs(got - w) < 4 * math.sqrt(w * (1 - w) / 4e6) + 2e-4, (vbin, got, w)
    others = [v for k, v in pf.items() if k[:2] != ("offroad", "normal_load")]
    assert len(others) == 15 and all(abs(v - 5.2e-4) < 1e-4 for v in others)
    for r in e.cpt.data:  # safe = 1 - fail exactly as the reference writes it (evaluate_node.jl:91-92)
        if r["E"] == "E_safe":
            assert r["Π"] == 1 - pf[(r["A"], r["b₀"], r["V_d"])]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bn = E.dispatch(net)
    assert isinstance(bn, E.BayesianNetwork) and sorted(n.name for n in bn.nodes) == ["A", "E", "V_d", "b₀"]


def test_networks_with_models_given_as_text(E):
    """The two demos with their closures typed in as text (Expression → run-time compiled functor):
    the CPTs agree with those of the hand-written functors within Monte Carlo error (the input
    columns follow the parents' order, so the streams differ; bit-level equality on identical streams
    is in test_gpu_expression.py), and the reference anchors hold."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cpts = []
        for as_text in (False, True):
            net = vehicle_net(E, 500_000, sim=E.MonteCarlo(500_000, strict=True), as_text=as_text)
            E.evaluate(net, True, True)
            cpts.append(net.node("E"))
        for r0, r1 in zip(cpts[0].cpt.data, cpts[1].cpt.data):
            assert {k: v for k, v in r0.items() if k != "Π"} == {k: v for k, v in r1.items() if k != "Π"}
            p = r0["Π"]
            assert abs(p - r1["Π"]) < 5 * math.sqrt(2 * p * (1 - p) / 500_000) + 1e-12
        # the replayed DataFrame carries the model columns of the closure chain
        smp = next(iter(cpts[1].additional_info.values()))["samples"].fetch(0, 1000)
        assert sorted(list(smp)[:6]) == sorted(["A", "b₀", "V", "C", "Cₖ", "K"]) and list(smp)[6:] == ["g1", "g2", "g3", "g4", "y", "g"]
        assert np.array_equal(smp["y"], smp["g"]) and np.array_equal(smp["y"], np.minimum.reduce([smp[f"g{i}"] for i in (1, 2, 3, 4)]))
        pfs = []
        for as_text in (False, True):
            net = straub_net(E, 10**6, as_text=as_text, strict=True)
            E.evaluate(net)
            pfs.append(net.node("E").cpt.column("Π"))
        assert np.allclose(pfs[1], [0.026129, 0.973871], atol=0.01)     # demo/straub_example.jl:71
        assert np.allclose(pfs[0], pfs[1], atol=1e-3)


def test_examples_run(E):
    """examples/ mirror the reference's two demos end to end."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    mods = {}
    import sys
    sys.path.insert(0, os.path.join(root, "examples"))
    for name in ("vehicle_suspension", "straub_example", "vehicle_suspension_imprecise", "custom_model"):
        spec = importlib.util.spec_from_file_location(f"example_{name}", os.path.join(root, "examples", f"{name}.py"))
        mods[name] = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mods[name])
    e = mods["vehicle_suspension"].main(500_000)
    pf = {(r["A"], r["b₀"], r["V_d"]): r["Π"] for r in e.cpt.data if r["E"] == "E_fail"}
    assert len(pf) == 20 and abs(pf[("offroad", "normal_load", "[11.5, 12.0]")] - 0.7373) < 0.005
    pi = mods["straub_example"].main(10**6)
    assert np.allclose(pi, [0.026129, 0.973871], atol=0.01)
    cm = mods["custom_model"]
    pf, sd, smp = E.probability_of_failure([cm.stress], cm.performance, cm.inputs, E.MonteCarlo(2_000_000))
    rows = smp.fetch(0, 1000)
    assert 1e-4 < pf < 0.2 and list(rows) == ["fy", "P", "e", "A", "W", "sigma", "g"]
    assert np.allclose(rows["g"], rows["fy"] - (1000 * rows["P"] / 5400.0 + 1000 * rows["P"] * np.abs(rows["e"]) / 3.9e5), rtol=1e-12)
    ei = mods["vehicle_suspension_imprecise"].main(200_000)
    assert not ei.isprecise()
    bounds = {r["V_d"]: r["Π"] for r in ei.cpt.data
              if r["E"] == "E_fail" and (r["A"], r["b₀"]) == ("offroad", "normal_load")}
    assert len(bounds) == 5 and all(lb < ub for lb, ub in bounds.values())
    lbs = [bounds[k][0] for k in ("[7.0, 8.5]", "[8.5, 9.5]", "[9.5, 10.5]", "[10.5, 11.5]", "[11.5, 13.0]")]
    assert lbs == sorted(lbs)                      # pf grows with V: the lower bounds are the left bin ends


def test_vehicle_suspension_imprecise(E):
    """BASELINE config 4 (demo/vehicle_suspension_imprecise.jl): V an interval (7, 13), discretised into
    interval bins; DoubleLoop sweeps V over each bin. pf is monotone in V, so the bounds are the pf at
    the bin ends — checked against precise runs with V fixed there."""
    sim = E.DoubleLoop(E.MonteCarlo(400_000), method="grid", grid=9)
    net = vehicle_net(E, 0, v_dist=(7, 13), sim=sim)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        E.evaluate(net, True, False)
    e = net.node("E")
    assert not e.isprecise() and len(e.cpt.data) == 40
    got = {r["V_d"]: r["Π"] for r in e.cpt.data if r["E"] == "E_fail" and (r["A"], r["b₀"]) == ("offroad", "normal_load")}
    assert set(got) == {"[7.0, 8.5]", "[8.5, 9.5]", "[9.5, 10.5]", "[10.5, 11.5]", "[11.5, 13.0]"}
    from ebn_b200 import uq
    M, m, g = 3.2633, 0.8158, 981.0
    consts = (M, m, g, (M * g) ** -1.5, (g * (M + m)) ** 0.877)
    model = E.Model(E.Catalogue("vehicle_suspension", consts, uq.VEHICLE_ARGS), "y")

    def pf_at(v):
        inputs = [E.Parameter(0.8, "A"), E.Parameter(0.27, "b₀"), E.Parameter(v, "V"),
                  E.RandomVariable(E.Normal(431.7221, 10), "C"), E.RandomVariable(E.Normal(1475.5503, 10), "Cₖ"),
                  E.RandomVariable(E.Normal(55.0406, 10), "K")]
        return E.probability_of_failure(model, "y", inputs, E.MonteCarlo(400_000, seed=77))[0]

    for name, (lo, hi) in {"[8.5, 9.5]": (8.5, 9.5), "[11.5, 13.0]": (11.5, 13.0)}.items():
        lb, ub = got[name]
        assert lb < ub and abs(lb - pf_at(lo)) < 0.005 and abs(ub - pf_at(hi)) < 0.005
    bn = E.dispatch(net)
    assert isinstance(bn, E.CredalNetwork)


def test_evaluate_with_envelopes(E):
    """evaluate_net.jl:93-216: two independent components (the Straub frame with a discrete parent M of V,
    and a small second frame sharing M) are split into Markov envelopes and evaluated separately."""
    lam, zeta = E.distribution_parameters(150, 30, "LogNormal")
    a, th = E.distribution_parameters(60, 12, "Gamma")
    ur = _cont_root(E, "Uᵣ", E.Normal())
    m = _root(E, "M", ["new", "old"], [0.5, 0.5])
    cptv = E.ContinuousConditionalProbabilityTable(["M"])
    cptv[("M", "new")], cptv[("M", "old")] = E.Gamma(a, th), E.Gamma(a - 1, 2.4)
    v = E.ContinuousNode("V", cptv)
    h = _cont_root(E, "H", E.Gumbel(*E.distribution_parameters(50, 20, "Gumbel")))
    sim = E.MonteCarlo(200_000)
    rs = [E.ContinuousFunctionalNode(f"R{i}", [E.Model(E.Catalogue("straub_capacity", (lam, zeta, 0.5477), ("Uᵣ",)), f"r{i}")], sim)
          for i in range(1, 6)]
    frame = E.DiscreteFunctionalNode(
        "E", [E.Model(E.Catalogue("straub_frame", (), ("r1", "r2", "r3", "r4", "r5", "V", "H")), "G")], "G", sim)
    # second component: L | M and R9 feeding a second frame
    cptl = E.ContinuousConditionalProbabilityTable(["M"])
    cptl[("M", "new")], cptl[("M", "old")] = E.Normal(1, 1), E.Normal(2, 1)
    L = E.ContinuousNode("L", cptl)
    r9 = _cont_root(E, "R9", E.Normal(4, 1))
    frame2 = E.DiscreteFunctionalNode("E2", [E.Model(E.Poly([(1, ["R9"]), (-1, ["L"])]), "G2")], "G2", sim)
    net = E.EnhancedBayesianNetwork([ur, m, v, h] + rs + [frame, L, r9, frame2])
    E.add_child(net, m, v)
    for r in rs:
        E.add_child(net, ur, r)
        E.add_child(net, r, frame)
    E.add_child(net, v, frame)
    E.add_child(net, h, frame)
    E.add_child(net, m, L)
    E.add_child(net, L, frame2)
    E.add_child(net, r9, frame2)
    E.order(net)
    envs = E.markov_envelope(net)
    assert len(envs) == 2
    assert {n.name for n in envs[0]} == {"Uᵣ", "M", "V", "H", "R1", "R2", "R3", "R4", "R5", "E"}   # evaluate_net.jl:197
    assert {n.name for n in envs[1]} == {"L", "R9", "E2", "M"}                                      # :205
    subs = E.evaluate_with_envelopes(net)
    assert [sorted(n.name for n in s.nodes) for s in subs] == [["E", "H", "M", "Uᵣ", "V"], ["E2", "L", "M", "R9"]]
    assert not any(isinstance(n, (E.DiscreteFunctionalNode, E.ContinuousFunctionalNode)) for s in subs for n in s.nodes)
    e = subs[0].node("E")
    assert e.cpt.names == ["M", "E"] and abs(e.cpt[{"M": "new", "E": "E_fail"}] - 0.0258) < 0.003
    assert e.cpt[{"M": "old", "E": "E_fail"}] < e.cpt[{"M": "new", "E": "E_fail"}]      # lighter vertical load
    e2 = subs[1].node("E2")
    assert abs(e2.cpt[{"M": "new", "E2": "E2_fail"}] - ndtr(-3 / math.sqrt(2))) < 0.002
    assert abs(e2.cpt[{"M": "old", "E2": "E2_fail"}] - ndtr(-2 / math.sqrt(2))) < 0.004


def test_joint_distribution_gaussian_copula(E):
    """UQ.jl `JointDistribution(marginals, GaussianCopula(R))` as a probability_of_failure input
    (BASELINE config 3): five log-normal capacities with equicorrelated scores
    (rho^2 = 0.3) reproduce the common-factor dependence of demo/straub_example.jl:16-26, so the pf
    equals the Straub frame's."""
    lam, zeta = E.distribution_parameters(150, 30, "LogNormal")
    R = np.full((5, 5), 0.5477**2)
    np.fill_diagonal(R, 1.0)
    caps = E.JointDistribution([E.RandomVariable(E.LogNormal(lam, zeta), f"r{i}") for i in range(1, 6)],
                               E.GaussianCopula(R))
    v = E.RandomVariable(E.Gamma(*E.distribution_parameters(60, 12, "Gamma")), "V")
    h = E.RandomVariable(E.Gumbel(*E.distribution_parameters(50, 20, "Gumbel")), "H")
    g1 = E.Model(E.Poly([(1, ["r1"]), (1, ["r2"]), (1, ["r4"]), (1, ["r5"]), (-5, ["H"])]), "g1")
    g2 = E.Model(E.Poly([(1, ["r2"]), (2, ["r3"]), (1, ["r4"]), (-5, ["V"])]), "g2")
    g3 = E.Model(E.Poly([(1, ["r1"]), (2, ["r3"]), (2, ["r4"]), (1, ["r5"]), (-5, ["H"]), (-5, ["V"])]), "g3")
    n = 1_000_000
    pfs = [E.probability_of_failure([g], g.name, [caps, v, h], E.MonteCarlo(n, seed=5))[0] for g in (g1, g2, g3)]
    # each mechanism alone fails less often than their minimum; the frame pf (0.0258) is bounded by the union
    assert all(0 < p < 0.03 for p in pfs) and sum(pfs) > 0.0258 - 0.002
    # the min of the three through the min_affine catalogue entry, same inputs → Straub's pf
    k = [3.0, 0, 1, 1, 0, 1, 1, 0, -5, 0, 0, 1, 2, 1, 0, -5, 0, 0, 1, 0, 2, 2, 1, -5, -5.0]
    frame = E.Model(E.Catalogue("min_affine", k, ("r1", "r2", "r3", "r4", "r5", "V", "H")), "G")
    pf, sd, _ = E.probability_of_failure(frame, "G", [caps, v, h], E.MonteCarlo(n, seed=5))
    assert abs(pf - 0.0258) < 4 * sd + 5e-4
    # independent capacities (identity copula) give a different, smaller pf: the dependence matters
    ind = E.JointDistribution(caps.marginals, E.GaussianCopula(np.eye(5)))
    pf_ind = E.probability_of_failure(frame, "G", [ind, v, h], E.MonteCarlo(n, seed=5))[0]
    assert pf_ind < pf - 0.003
    with pytest.raises(ValueError, match="correlation matrix"):
        E.GaussianCopula([[1.0, 0.5], [0.4, 1.0]])


def test_straub_with_evidence_posterior(E):
    """test/inference/variableselimination.jl:207-284: R4 and R5 are continuous functional nodes with
    21 cut points each (22 bins after the support ends are added); they are evaluated to histograms,
    discretised, and the frame node is evaluated for all 22 x 22 scenarios with
    truncated(EmpiricalDistribution) capacities. The reference then asserts
    P(E | R4 in [140,150], R5 in [90.01,100.01]) ≈ [0.035, 0.965] (atol 0.05) — with both parents
    observed that posterior is the CPT row itself."""
    lam, zeta = E.distribution_parameters(150, 30, "LogNormal")
    ur = _cont_root(E, "Uᵣ", E.Normal())
    v = _cont_root(E, "V", E.Gamma(*E.distribution_parameters(60, 12, "Gamma")))
    h = _cont_root(E, "H", E.Gumbel(*E.distribution_parameters(50, 20, "Gumbel")))
    n2 = 2000
    sim = E.MonteCarlo(n2)
    cap = lambda i: E.Model(E.Catalogue("straub_capacity", (lam, zeta, 0.5477), ("Uᵣ",)), f"r{i}")
    rs = [E.ContinuousFunctionalNode(f"R{i}", [cap(i)], sim) for i in (1, 2, 3)]
    # collect(range(50, 250, 21)) / collect(range(50.01, 250.01, 21)): Julia's ranges give the exact decimals
    d1 = E.ApproximatedDiscretization([50.0 + 10 * i for i in range(21)], 1)
    d2 = E.ApproximatedDiscretization([round(50.01 + 10 * i, 2) for i in range(21)], 1)
    r4 = E.ContinuousFunctionalNode("R4", [cap(4)], sim, d1)
    r5 = E.ContinuousFunctionalNode("R5", [cap(5)], sim, d2)
    frame = E.DiscreteFunctionalNode(
        "E", [E.Model(E.Catalogue("straub_frame", (), ("r1", "r2", "r3", "R4", "R5", "V", "H")), "G")], "G", sim)
    net = E.EnhancedBayesianNetwork([ur, v, h] + rs + [r4, r5, frame])
    for r in rs + [r4, r5]:
        E.add_child(net, ur, r)
        E.add_child(net, r, frame)
    E.add_child(net, v, frame)
    E.add_child(net, h, frame)
    E.order(net)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        E.evaluate(net, True, False)
    assert sorted(n.name for n in net.nodes) == ["E", "H", "R4", "R4_d", "R5", "R5_d", "Uᵣ", "V"]
    e = net.node("E")
    assert sorted(e.cpt.names[:-1]) == ["R4_d", "R5_d"]
    n4, n5 = len(net.node("R4_d").states()), len(net.node("R5_d").states())
    assert n4 >= 20 and n5 >= 20 and len(e.cpt.data) == 2 * n4 * n5
    post = e.cpt[{"R4_d": "[140.0, 150.0]", "R5_d": "[90.01, 100.01]", "E": "E_fail"}]
    assert abs(post - 0.035) < 0.05 and abs((1 - post) - 0.965) < 0.05          # the reference's assertion
    # bin probabilities of R4_d follow LogNormal(lambda, zeta): P(140 < R4 < 150)
    from scipy.stats import lognorm
    p_bin = lognorm(zeta, scale=math.exp(lam)).cdf(150) - lognorm(zeta, scale=math.exp(lam)).cdf(140)
    assert abs(net.node("R4_d").cpt[("R4_d", "[140.0, 150.0]")] - p_bin) < 0.03   # n = 2000
    # weak capacities fail more often than strong ones
    lo = e.cpt[{"R4_d": "[60.0, 70.0]", "R5_d": "[60.01, 70.01]", "E": "E_fail"}]
    hi = e.cpt[{"R4_d": "[200.0, 210.0]", "R5_d": "[200.01, 210.01]", "E": "E_fail"}]
    assert lo > post > hi
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        E.reduce(net)
    assert sorted(n.name for n in net.nodes) == ["E", "R4_d", "R5_d"]          # adjacency :263-267
    assert net.adj_matrix.sum() == 2 and net.adj_matrix[:, net.topology_dict["E"]].sum() == 2
