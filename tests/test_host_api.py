"""CPU tests of the host-side mirror of the reference API (no GPU): CPTs, nodes, graph helpers,
discretisation, transmission, scenario enumeration and input marshalling — the parts of the hot path
that are NOT arithmetic. Each test names the reference test it restates."""
import math
import warnings

import numpy as np
import pytest

import ebn_b200 as E
from ebn_b200 import evaluation as ev
from ebn_b200 import network as nw
from ebn_b200 import uq
from ebn_b200.nodes import PI, interval_symbol, jl_float

INF = math.inf


def _root(name, states, probs, params=None):
    cpt = E.DiscreteConditionalProbabilityTable(name)
    for s, p in zip(states, probs):
        cpt[(name, s)] = p
    return E.DiscreteNode(name, cpt, params or {})


def _cont_root(name, dist, disc=None):
    cpt = E.ContinuousConditionalProbabilityTable()
    cpt[()] = dist
    return E.ContinuousNode(name, cpt, disc)


def test_cpt_setindex_getindex_and_errors():
    """test/nodes/cpts.jl: setindex!/getindex by evidence, wrong columns error."""
    cpt = E.DiscreteConditionalProbabilityTable(["x", "y"])
    cpt[{"x": "x1", "y": "y1"}] = 0.3
    cpt[{"x": "x1", "y": "y2"}] = 0.7
    cpt[{"y": "y1", "x": "x1"}] = 0.4
    assert cpt[{"x": "x1", "y": "y1"}] == 0.4 and len(cpt.data) == 2
    with pytest.raises(ValueError, match="Cannot set index"):
        cpt[{"x": "x1"}] = 0.1
    with pytest.raises(KeyError):
        cpt[{"x": "x9", "y": "y1"}]
    assert cpt.states("y") == ["y1", "y2"] and cpt.scenarios("y") == [{"x": "x1"}]
    assert not cpt.isroot() and E.DiscreteConditionalProbabilityTable("a").isroot()


def test_discrete_node_constructor_rules():
    """test/nodes/discretenodes.jl + test/utils/verify_discrete.jl: column order, sorting, Σ=1 with the
    0.05 auto-normalisation, parameter keys."""
    cpt = E.DiscreteConditionalProbabilityTable(["x", "p"])
    for p, x, v in [("b", "yes", 0.6), ("b", "no", 0.4), ("a", "yes", 0.2), ("a", "no", 0.8)]:
        cpt[{"p": p, "x": x}] = v
    n = E.DiscreteNode("x", cpt)
    assert n.cpt.names == ["p", "x"]
    assert [(r["p"], r["x"]) for r in n.cpt.data] == [("a", "no"), ("a", "yes"), ("b", "no"), ("b", "yes")]
    assert n.states() == ["no", "yes"] and n.scenarios() == [{"p": "a"}, {"p": "b"}]
    with pytest.warns(UserWarning, match="will be normalized"):
        m = _root("m", ["m1", "m2"], [0.5, 0.48])
    assert math.isclose(sum(r[PI] for r in m.cpt.data), 1.0)
    with pytest.raises(ValueError, match="not exhaustive"):
        _root("m", ["m1", "m2"], [0.5, 0.3])
    with pytest.raises(ValueError, match="non-negative"):
        _root("m", ["m1", "m2"], [-0.5, 1.5])
    with pytest.raises(ValueError, match="parameters keys"):
        _root("m", ["m1", "m2"], [0.5, 0.5], {"m1": [E.Parameter(1, "m")]})
    with pytest.raises(ValueError, match="does not contain a column"):
        E.DiscreteNode("zz", E.DiscreteConditionalProbabilityTable("m"))


def test_uq_inputs_marshalling():
    """test/nodes/continuousnode.jl:100-145, test/nodes/discretenodes.jl:279-301."""
    a = _root("A", ["a1", "a2"], [0.5, 0.5], {"a1": [E.Parameter(1, "A")], "a2": [E.Parameter(2, "A")]})
    assert a._uq_inputs({"A": "a2", "other": "o"}) == [E.Parameter(2, "A")]
    with pytest.raises(KeyError, match="does not contain the node A"):
        a._uq_inputs({"B": "b"})
    cpt = E.ContinuousConditionalProbabilityTable(["A"])
    cpt[("A", "a1")] = E.Normal(0, 1)
    cpt[("A", "a2")] = E.Normal(3, 1)
    c = E.ContinuousNode("C", cpt)
    assert c._uq_inputs({"A": "a2", "zzz": "q"}) == [E.RandomVariable(E.Normal(3, 1), "C")]
    iv = E.ContinuousConditionalProbabilityTable(kind="interval")
    iv[()] = (0.1, 0.3)
    assert E.ContinuousNode("B", iv)._uq_inputs() == [E.Interval(0.1, 0.3, "B")]
    pb = E.ContinuousConditionalProbabilityTable(kind="pbox")
    pb[()] = E.UnamedProbabilityBox(E.Normal, (E.Interval(0, 1, "μ"), E.Interval(1, 2, "σ")))
    got = E.ContinuousNode("P", pb)._uq_inputs()[0]
    assert isinstance(got, E.ProbabilityBox) and got.name == "P" and got.family is E.Normal
    with pytest.raises(ValueError, match="Root node must have ExactDiscretization"):
        _cont_root("z", E.Normal(), E.ApproximatedDiscretization([0, 1], 1))


def test_format_interval_approximate_and_bin_probabilities():
    """test/networks/ebn/discretization/discretize.jl:3-57."""
    tn = E.truncated(E.Normal(0, 1), -10, 10)
    assert ev._format_interval(_cont_root("z1", tn, E.ExactDiscretization([-10, 0, 10]))) == [[-10, 0.0], [0.0, 10]]
    for cuts, msg in [([-9, 10], "minimum intervals value -9.0 > support lower bound -10.0"),
                      ([-11, 10], "minimum intervals value -11.0 < support lower bound -10.0"),
                      ([-10, 9], "maximum intervals value 9.0 < support upper bound 10.0"),
                      ([-10, 11], "maximum intervals value 11.0 > support upper bound 10.0")]:
        with pytest.warns(UserWarning, match=msg):
            got = ev._format_interval(_cont_root("z1", tn, E.ExactDiscretization(cuts)))
        assert got[0][0] == -10 and got[-1][1] == 10
    intervals = [[-INF, -1.0], [-1.0, 0.0], [0.0, 1.0], [1.0, INF]]
    approx = [ev._approximate(iv, 2) for iv in intervals]
    assert approx == [-E.Exponential(2) - 1 if False else (-E.Exponential(2)) + (-1.0), E.Uniform(-1, 0), E.Uniform(0.0, 1.0),
                      E.Exponential(2) + 1]
    probs = [0.15865525393145702, 0.341344746068543, 0.34134474606854304, 0.15865525393145696]
    assert np.allclose(ev._bin_probabilities(E.Normal(), intervals), probs, rtol=0, atol=1e-15)
    # the shifted exponentials are proper distributions on their half lines
    lo = approx[0]
    assert lo.support() == (-INF, -1.0) and math.isclose(float(lo.cdf(-1.0)), 1.0) and float(lo.cdf(-3.0)) == pytest.approx(math.exp(-1))
    assert float(lo.quantile(0.5)) == pytest.approx(-1 - 2 * math.log(2))


def test_discretize_root_nodes():
    """discretize.jl:59-110 (precise root, then a Normal p-box root)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        d, c = ev._discretize(_cont_root("z1", E.Normal(), E.ExactDiscretization([0, INF])))
    assert d.name == "z1_d" and d.cpt.column("z1_d") == ["[-Inf, 0.0]", "[0.0, Inf]"] and d.cpt.column(PI) == [0.5, 0.5]
    assert c.name == "z1" and c.cpt.column("z1_d") == ["[-Inf, 0.0]", "[0.0, Inf]"]
    assert c.cpt.column(PI) == [E.truncated(E.Normal(), -INF, 0.0), E.truncated(E.Normal(), 0.0, INF)]
    cpt = E.ContinuousConditionalProbabilityTable(kind="pbox")
    box = E.UnamedProbabilityBox(E.Normal, (E.Interval(-0.5, 0.5, "μ"), E.Interval(1, 2, "σ")))
    cpt[()] = box
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        d, c = ev._discretize(E.ContinuousNode("z1", cpt, E.ExactDiscretization([-1, 0, 1])))
    assert d.cpt.column("z1_d") == ["[-1.0, 0.0]", "[-Inf, -1.0]", "[0.0, 1.0]", "[1.0, Inf]"]
    want = [(0.24173033745712885, 0.2901687869569368), (0.06680720126885804, 0.4012936743170763),
            (0.2417303374571288, 0.2901687869569368), (0.06680720126885809, 0.401293674317076)]
    for got, w in zip(d.cpt.column(PI), want):
        assert np.allclose(got, w, atol=1e-3)
    assert not d.isprecise() and c.cpt.kind == "pbox"
    assert c.cpt[("z1_d", "[-Inf, -1.0]")] == E.UnamedProbabilityBox(E.Normal, box.parameters, -INF, -1.0)


def test_jl_float_and_symbols():
    assert [jl_float(v) for v in (7.0, 8.5, -INF, INF, 1e-5, 1.5e20, 0.1)] == \
        ["7.0", "8.5", "-Inf", "Inf", "1.0e-5", "1.5e20", "0.1"]
    assert interval_symbol([9.5, 10.5]) == "[9.5, 10.5]"


def _toy_net(sim=None):
    """The A, B → C net of test/networks/ebn/evaluate/evaluate_node.jl:88-105."""
    a = _root("A", ["a1", "a2"], [0.5, 0.5], {"a1": [E.Parameter(1, "A")], "a2": [E.Parameter(2, "A")]})
    b = _cont_root("B", E.Normal())
    f = E.DiscreteFunctionalNode("C", [E.Model(E.Poly([(1, ["A"]), (1, ["B"])]), "C")], E.Poly([(2, []), (-1, ["C"])]),
                                 sim or E.MonteCarlo(100_000))
    net = E.EnhancedBayesianNetwork([a, b, f])
    E.add_child(net, a, f)
    E.add_child(net, b, f)
    E.order(net)
    return net, a, b, f


def test_network_graph_helpers_and_add_child_rules():
    """test/networks/networks_common.jl, test/networks/ebn/enhancedbn.jl."""
    net, a, b, f = _toy_net()
    assert nw.parents(net, f)[1] == ["A", "B"] and nw.children(net, a)[2] == [f]
    assert nw.discrete_ancestors(net, f) == [a]
    with pytest.raises(ValueError, match="Recursion"):
        E.add_child(net, a, a)
    with pytest.raises(ValueError, match="is a root node and cannot have parents"):
        E.add_child(net, b, a)
    with pytest.raises(ValueError, match="can have only functional children"):
        cptc = E.ContinuousConditionalProbabilityTable(["A"])
        cptc[("A", "a1")], cptc[("A", "a2")] = E.Normal(), E.Normal(1, 1)
        net2 = E.EnhancedBayesianNetwork([a, f, E.ContinuousNode("K", cptc)])
        E.add_child(net2, f, "K")
    with pytest.raises(ValueError, match="names must be unique"):
        E.EnhancedBayesianNetwork([a, a])
    with pytest.raises(ValueError, match="states must be unique"):
        E.EnhancedBayesianNetwork([a, _root("A2", ["a1", "zz"], [0.5, 0.5])])
    cyc = E.EnhancedBayesianNetwork([a, b, f])
    cyc.adj_matrix[0, 2] = cyc.adj_matrix[2, 0] = 1
    with pytest.raises(ValueError, match="cyclic"):
        E.order(cyc)
    lonely = E.EnhancedBayesianNetwork([a, b, f])
    with pytest.raises(ValueError, match="must have at least one parent"):
        E.order(lonely)


def test_is_eliminable_auxiliary_functions():
    """test/networks/ebn/evaluate/evaluate_net.jl:283-335 ("Auxiliary functions"): the eight-node net with three
    continuous functional nodes between the roots and two discrete functional nodes."""
    x = _root("x", ["x1", "x2"], [0.3, 0.7], {"x1": [E.Parameter(0.5, "x")], "x2": [E.Parameter(0.7, "x")]})
    y = _cont_root("y", E.Normal())
    z = _root("z", ["z1", "z2"], [0.3, 0.7], {"z1": [E.Parameter(0.5, "z")], "z2": [E.Parameter(0.7, "z")]})
    sim = E.MonteCarlo(300)
    cf1 = E.ContinuousFunctionalNode("cf1", [E.Model(E.Expression("df.x .^ 2 .- 0.7 .+ df.y"), "c1")], sim)
    cf2 = E.ContinuousFunctionalNode("cf2", [E.Model(E.Expression("df.z .^ 2 .- 0.7 .+ df.y"), "c2")], sim)
    fd1 = E.DiscreteFunctionalNode("fd1", [E.Model(E.Expression("df.c1 .* 0.5 .+ df.c2"), "final1")],
                                   E.Expression("df.final1 .- 0.5"), sim,
                                   {"fail_fd1": [E.Parameter(1, "fd1")], "safe_fd1": [E.Parameter(0, "fd1")]})
    c3 = E.ContinuousFunctionalNode("c3", [E.Model(E.Expression("df.c2 .* 0.5"), "c3")], sim)
    fd = E.DiscreteFunctionalNode("fd", [E.Model(E.Expression("0.5 .+ df.c3"), "tot")], E.Expression("0.5 .- df.tot"), sim)
    net = E.EnhancedBayesianNetwork([x, y, z, cf1, cf2, fd1, c3, fd])
    for par, ch in ((x, cf1), (y, cf1), (y, cf2), (z, cf2), (cf1, fd1), (cf2, fd1), (cf2, c3), (fd1, fd), (c3, fd)):
        E.add_child(net, par, ch)
    E.order(net)
    with pytest.raises(TypeError, match="node elimination algorithm is for continuous nodes and x is discrete"):
        ev._is_eliminable(net, x)
    assert ev._is_eliminable(net, cf2) is False
    assert ev._is_eliminable(net, y) is True
    assert ev._is_eliminable(net, net.topology_dict["y"]) is True      # by topology index (the reference's `2`)
    assert ev._is_eliminable(net, "y") is True


def test_scenario_enumeration_order_through_continuous_parents():
    """evaluate_node.jl:52-56 + networks_common.jl:37-45: direct discrete parents first, then the
    ancestors reached through continuous parents; combinations sorted lexicographically."""
    a = _root("A", ["a2", "a1"], [0.5, 0.5], {"a1": [E.Parameter(1, "A")], "a2": [E.Parameter(2, "A")]})
    d = _root("D", ["d1", "d2"], [0.5, 0.5], {"d1": [E.Parameter(1, "D")], "d2": [E.Parameter(2, "D")]})
    cptc = E.ContinuousConditionalProbabilityTable(["A"])
    cptc[("A", "a1")], cptc[("A", "a2")] = E.Normal(0, 1), E.Normal(5, 1)
    c1 = E.ContinuousNode("C1", cptc)
    f = E.DiscreteFunctionalNode("F1", [E.Model(E.Poly([(1, ["D"]), (1, ["C1"])]), "C2")], E.Poly([(2, []), (-1, ["C2"])]),
                                 E.MonteCarlo(10))
    net = E.EnhancedBayesianNetwork([a, d, c1, f])
    E.add_child(net, a, c1)
    E.add_child(net, d, f)
    E.add_child(net, c1, f)
    E.order(net)
    anc, combos, evs = ev._evidences(net, f)
    assert [x.name for x in anc] == ["D", "A"]
    assert combos == [("d1", "a1"), ("d1", "a2"), ("d2", "a1"), ("d2", "a2")]
    assert evs[1] == {"D": "d1", "A": "a2"}
    inputs = ev._scenario_inputs(net, f, evs[1])
    assert inputs == [E.Parameter(1, "D"), E.RandomVariable(E.Normal(5, 1), "C1")]


def test_discretize_net_rewires_and_multiplies_scenarios():
    """discretize.jl:1-31 / test discretize.jl:294-305 on the vehicle-suspension V node."""
    v = _cont_root("V", E.Uniform(7, 12), E.ExactDiscretization([8.5, 9.5, 10.5, 11.5]))
    a = _root("A", ["road", "offroad"], [0.7, 0.3], {"road": [E.Parameter(0.15915, "A")], "offroad": [E.Parameter(0.8, "A")]})
    f = E.DiscreteFunctionalNode("E", [E.Model(E.Poly([(1, ["A"]), (1, ["V"])]), "y")], "y", E.MonteCarlo(10))
    net = E.EnhancedBayesianNetwork([a, v, f])
    E.add_child(net, a, f)
    E.add_child(net, v, f)
    E.order(net)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ev._discretize_net(net)
    assert sorted(n.name for n in net.nodes) == ["A", "E", "V", "V_d"]
    vd, vc = net.node("V_d"), net.node("V")
    assert vd.states() == ["[10.5, 11.5]", "[11.5, 12.0]", "[7.0, 8.5]", "[8.5, 9.5]", "[9.5, 10.5]"]
    assert np.allclose(vd.cpt.column(PI), [0.2, 0.1, 0.3, 0.2, 0.2])
    assert nw.parents(net, vc)[1] == ["V_d"] and nw.parents(net, f)[1] == ["A", "V"]
    anc, combos, evs = ev._evidences(net, f)
    assert [x.name for x in anc] == ["A", "V_d"] and len(evs) == 10
    assert combos[0] == ("offroad", "[10.5, 11.5]")
    inp = ev._scenario_inputs(net, f, evs[2])
    assert inp == [E.Parameter(0.8, "A"), E.RandomVariable(E.truncated(E.Uniform(7, 12), 7.0, 8.5), "V")]
    assert inp[1].dist.abi() == (E._lib.IN_UNIFORM, 7.0, 12.0, 0.0, 0.0, 7.0, 8.5)


def test_transmission_fuses_model_chains():
    """test/networks/ebn/transmission/transmission.jl: an un-discretised continuous functional node
    prepends its models to its children and disappears; simulation types must agree."""
    b = _cont_root("B", E.Normal())
    sim = E.MonteCarlo(10)
    r = E.ContinuousFunctionalNode("R", [E.Model(E.Poly([(2, ["B"])]), "R")], sim)
    f = E.DiscreteFunctionalNode("F", [E.Model(E.Poly([(1, ["R"]), (1, ["B"])]), "G")], "G", sim)
    net = E.EnhancedBayesianNetwork([b, r, f])
    E.add_child(net, b, r)
    E.add_child(net, r, f)
    E.add_child(net, b, f)
    E.order(net)
    ev._transfer_continuous(net)
    assert [n.name for n in net.nodes] == ["B", "F"] and [m.name for m in f.models] == ["R", "G"]
    assert nw.parents(net, f)[1] == ["B"]
    low = uq.lower(f.models, f.performance, ["B"])
    assert low.limit_state == E._lib.LS_POLYNOMIAL and low.consts[0] == 2.0
    # incoherent simulation types
    r2 = E.ContinuousFunctionalNode("R", [E.Model(E.Poly([(2, ["B"])]), "R")], sim)
    f2 = E.DiscreteFunctionalNode("F", [E.Model(E.Poly([(1, ["R"])]), "G")], "G", E.DoubleLoop(sim))
    net = E.EnhancedBayesianNetwork([b, r2, f2])
    E.add_child(net, b, r2)
    E.add_child(net, r2, f2)
    E.order(net)
    with pytest.raises(ValueError, match="non coherent with children simulation types"):
        ev._transfer_continuous(net)


def test_lowering_model_chains():
    assert uq.lower([E.Model(E.Poly([(1, ["A"]), (1, ["B"])]), "C")], E.Poly([(2, []), (-1, ["C"])]), ["A", "B"]).consts.tolist() == \
        [2, 2, 1, 1, 0, 1, 1, 1, 2, 2, 0, -1, 1, 2]
    veh = E.Model(E.Catalogue("vehicle_suspension", (1, 2, 3, 4, 5), uq.VEHICLE_ARGS), "y")
    low = uq.lower([veh], "y", ["K", "A", "b₀", "V", "C", "Cₖ"])
    assert low.limit_state == E._lib.LS_VEHICLE_SUSPENSION and low.columns == list(uq.VEHICLE_ARGS)
    with pytest.raises(E.UnsupportedModel, match="needs inputs"):
        uq.lower([veh], "y", ["A", "V"])
    with pytest.raises(E.UnsupportedModel, match="no device functor"):
        uq.lower([E.Model(lambda df: df, "f")], "f", ["A"])
    with pytest.raises(E.UnsupportedModel, match="not among"):
        uq.lower([E.Model(E.Poly([(1, ["nope"])]), "C")], "C", ["A"])
    caps = [E.Model(E.Catalogue("straub_capacity", (4.99, 0.198, 0.5477), ("Uᵣ",)), f"r{i}") for i in range(1, 6)]
    frame = E.Model(E.Catalogue("straub_frame", (), ("r1", "r2", "r3", "r4", "r5", "V", "H")), "G")
    low = uq.lower(caps + [frame], "G", ["H", "Uᵣ", "V"])
    assert low.limit_state == E._lib.LS_STRAUB_FRAME and low.columns == ["Uᵣ", "V", "H"] and len(low.hidden) == 5
    frame45 = E.Model(E.Catalogue("straub_frame", (), ("r1", "r2", "r3", "R4", "R5", "V", "H")), "G")
    low = uq.lower(caps[:3] + [frame45], "G", ["H", "R4", "R5", "Uᵣ", "V"])
    assert low.limit_state == E._lib.LS_STRAUB_FRAME_R45 and low.columns == ["Uᵣ", "V", "H", "R4", "R5"] and len(low.hidden) == 3
    low = uq.lower(caps[3:4], None, ["Uᵣ"])
    assert low.limit_state == E._lib.LS_STRAUB_CAPACITY and low.columns == ["Uᵣ"] and len(low.hidden) == 1


def test_distributions_host_side():
    g = E.Gamma(*E.distribution_parameters(60, 12, E.Gamma))
    assert (g.alpha, g.theta) == (25.0, 2.4) and g.mean() == 60.0
    mu, beta = E.distribution_parameters(50, 20, "Gumbel")
    assert (mu, beta) == (40.99893584908611, 15.593936024673523)
    lam, zeta = E.distribution_parameters(150, 30, "LogNormal")
    assert math.isclose(lam, 4.991024937519615) and math.isclose(zeta, 0.1980422004353651)
    t = E.truncated(E.Normal(0, 1), -1, 2)
    assert t.support() == (-1, 2) and float(t.cdf(-1)) == 0 and float(t.cdf(2)) == 1
    assert math.isclose(float(t.quantile(t.cdf(0.3))), 0.3)
    edges = np.linspace(-4, 4, 81)
    x = np.random.default_rng(0).normal(0, 1, 200_000)
    counts = np.bincount(np.searchsorted(edges, x, side="right"), minlength=len(edges) + 1)
    emp = E.EmpiricalDistribution(edges, counts, len(x), x.sum(), (x * x).sum())
    assert abs(float(emp.cdf(0.0)) - 0.5) < 5e-3 and abs(emp.mean()) < 0.01 and abs(emp.std() - 1) < 0.01
    assert abs(float(emp.quantile(0.975)) - 1.96) < 0.03
    from ebn_b200.distributions import table_for
    q = table_for(g, 50.0, 70.0, 1025)
    assert q[0] == 50.0 and q[-1] == 70.0 and np.all(np.diff(q) >= 0)
