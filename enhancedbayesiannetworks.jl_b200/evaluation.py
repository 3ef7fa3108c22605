"""`evaluate!` and its pre/post steps, with the Monte Carlo routed through libebncuda.

  reference                                                       here
  src/networks/ebn/evaluation/evaluate_net.jl:1-32                evaluate
  src/networks/ebn/evaluation/evaluate_net.jl:42-66               _is_eliminable
  src/networks/ebn/evaluation/evaluate_node.jl:41-122             _evaluate_node (DiscreteFunctionalNode)
  src/networks/ebn/evaluation/evaluate_node.jl:1-39               _evaluate_node (ContinuousFunctionalNode)
  src/networks/ebn/discretization/discretize.jl                   _discretize_net, _discretize, _approximate, _format_interval
  src/networks/ebn/transmission/transmission.jl                   _transfer_continuous
  src/networks/ebn/reduction/reduction.jl, src/networks/dispatch.jl   reduce, dispatch

What changes against the reference: the scenario loop of `_evaluate_node` (one UQ.jl call per
scenario) becomes ONE batched C-ABI call for all scenarios; everything around it — scenario
enumeration order, input marshalling, CPT rows, additional_info keys — is kept.
"""
from __future__ import annotations

import itertools
import warnings
from typing import Dict, List

import numpy as np

from . import uq
from .distributions import EmpiricalDistribution, Exponential, ProbabilityBox, Uniform
from .network import (BayesianNetwork, CredalNetwork, EnhancedBayesianNetwork, _add_node, _is_cyclic_dfs,
                      _remove_node, add_child, children, discrete_ancestors, order, parents)
from .nodes import (FORM, PI, AbstractMonteCarlo, ContinuousConditionalProbabilityTable,
                    ContinuousFunctionalNode, ContinuousNode, CudaDoubleLoop, CudaRandomSlicing,
                    DiscreteConditionalProbabilityTable, DiscreteFunctionalNode, DiscreteNode, ExactDiscretization,
                    FunctionalNode, UnamedProbabilityBox, _truncate, interval_symbol, isprecise)


# ---------------------------------------------------------------------------------------------
# discretisation (pre-step)
# ---------------------------------------------------------------------------------------------
def _format_interval(node: ContinuousNode) -> List[List[float]]:
    """discretize.jl:93-118: clip/extend the cut points to the support."""
    iv = [float(v) for v in node.discretization.intervals]
    mn, mx = iv[0], iv[-1]
    lo, hi = node._distribution_bounds()
    if mn > lo:
        warnings.warn(f"node {node.name} has minimum intervals value {mn} > support lower bound {lo}. Lower bound "
                      "will be used as intervals start")
        iv.insert(0, lo)
    if mn < lo:
        warnings.warn(f"node {node.name} has minimum intervals value {mn} < support lower bound {lo}. Lower bound "
                      "will be used as intervals start")
        iv = [v for v in iv if v > lo]
        iv.insert(0, lo)
    if mx < hi:
        warnings.warn(f"node {node.name} has maximum intervals value {mx} < support upper bound {hi}. Upper bound "
                      "will be used as intervals end")
        iv.append(hi)
    if mx > hi:
        warnings.warn(f"node {node.name} has maximum intervals value {mx} > support upper bound {hi}. Upper bound "
                      "will be used as intervals end")
        iv = [v for v in iv if v < hi]
        iv.append(hi)
    return [[iv[i], iv[i + 1]] for i in range(len(iv) - 1)]


def _approximate(iv, lam):
    """discretize.jl:83-91."""
    a, b = iv
    if np.isfinite(a) and np.isfinite(b):
        return Uniform(a, b)
    if np.isfinite(b):
        return -Exponential(lam) + b
    return Exponential(lam) + a


def _bin_probabilities(dist, intervals):
    """discretize.jl:68-81."""
    if isinstance(dist, tuple):
        return [(0, 1)] * len(intervals)
    if isinstance(dist, UnamedProbabilityBox):
        pb = ProbabilityBox(dist.family, tuple(dist.parameters), "temp")
        out = []
        for a, b in intervals:
            (ll, lu), (rl, ru) = pb.cdf_bounds(a), pb.cdf_bounds(b)
            out.append((min(rl - ll, ru - lu), max(rl - ll, ru - lu)))
        return out
    return [float(dist.cdf(b)) - float(dist.cdf(a)) for a, b in intervals]


def _discretize(node: ContinuousNode):
    """discretize.jl:33-65 → (discrete node X_d, continuous node X | X_d)."""
    intervals = _format_interval(node)
    syms = [interval_symbol(iv) for iv in intervals]
    name_d = node.name + "_d"
    cpt_d = DiscreteConditionalProbabilityTable(node.cpt.names + [name_d], precise=node.cpt.precise)
    for row in node.cpt.data:
        probs = _bin_probabilities(row[PI], intervals)
        for s, p in zip(syms, probs):
            cpt_d.data.append({**{k: row[k] for k in node.cpt.names}, name_d: s, PI: p})
    disc = DiscreteNode(name_d, cpt_d)
    if node.isroot():
        dists = [_truncate(row[PI], iv) for row in node.cpt.data for iv in intervals]
        cpt_c = ContinuousConditionalProbabilityTable([name_d], kind=node.cpt.kind)
    else:
        dists = [_approximate(iv, node.discretization.sigma) for iv in intervals]
        cpt_c = ContinuousConditionalProbabilityTable([name_d], kind="precise")
    for s, d in zip(syms, dists):
        cpt_c.data.append({name_d: s, PI: d})
    return disc, ContinuousNode(node.name, cpt_c)


def _discretize_net(net):
    """discretize.jl:1-31."""
    todo = [n for n in net.nodes if isinstance(n, ContinuousNode) and n.discretization.intervals]
    plans = [(n, parents(net, n)[2], children(net, n)[2], _discretize(n)) for n in todo]
    for node, pars, chs, (disc, cont) in plans:
        _remove_node(net, node)
        _add_node(net, disc)
        _add_node(net, cont)
        add_child(net, disc, cont)
        for p in pars:
            try:
                add_child(net, p, disc)
            except ValueError:
                warnings.warn(f"node {disc.name} is a root node and will be added as a child of {p.name}. This is "
                              "allowed only for network evaluation.")
                net.adj_matrix[net.topology_dict[p.name], net.topology_dict[disc.name]] = 1
        for c in chs:
            add_child(net, cont, c)
        order(net)


# ---------------------------------------------------------------------------------------------
# transmission (pre-step)
# ---------------------------------------------------------------------------------------------
def _transfer_continuous(net):
    """transmission.jl:1-29: an un-discretised continuous functional node hands its models to its
    children and disappears."""
    for node in [n for n in net.nodes if isinstance(n, ContinuousFunctionalNode)]:
        chs = children(net, node)[2]
        if node.discretization.intervals or not chs:
            continue
        pars = parents(net, node)[2]
        for ch in chs:
            ch.models = list(node.models) + list(ch.models)
        bad = [ch for ch in chs if type(ch.simulation) is not type(node.simulation)]
        if bad:
            raise ValueError(f"node {node.name} cannot be transferred into his children => {[c.name for c in bad]}, "
                             f"because its simulation type: {type(node.simulation).__name__} is non coherent with "
                             f"children simulation types: {[type(c.simulation).__name__ for c in bad]}")
        _remove_node(net, node)
        for p in pars:
            for ch in chs:
                add_child(net, p, ch)
        order(net)


# ---------------------------------------------------------------------------------------------
# node evaluation (THE HOT PATH)
# ---------------------------------------------------------------------------------------------
def _evidences(net, node):
    """evaluate_node.jl:52-56: ancestors, their state product sorted lexicographically, and the
    state → owning node mapping (state names are unique network-wide, enhancedbn.jl:11-17)."""
    ancestors = discrete_ancestors(net, node)
    combos = sorted(itertools.product(*[a.states() for a in ancestors]))
    evs = []
    for combo in combos:
        ev = {}
        for st in combo:
            owner = next((a for a in ancestors if st in a.states()), None)
            ev[owner.name if owner is not None else node.name] = st
        evs.append(ev)
    return ancestors, combos, evs


def _scenario_inputs(net, node, evidence):
    inputs = []
    for p in parents(net, node)[2]:
        inputs += p._uq_inputs(evidence)
    return inputs


def _evaluate_discrete(net, node: DiscreteFunctionalNode, collect_samples=True) -> DiscreteNode:
    ancestors, combos, evidences = _evidences(net, node)
    imprecise = [p for p in parents(net, node)[2] if not isprecise(p)]
    sim_imprecise = [p for p in imprecise if isinstance(p, ContinuousNode) and not p.discretization.intervals]
    no_anc = len(combos[0]) == 0
    names = [node.name] if no_anc else [a.name for a in ancestors] + [node.name]
    cpt = DiscreteConditionalProbabilityTable(names, precise=not sim_imprecise)
    if no_anc:
        evidences = [{}]
    scen = [_scenario_inputs(net, node, ev) for ev in evidences]
    sim = node.simulation
    info: Dict[tuple, dict] = {}
    fail, safe = f"{node.name}_fail", f"{node.name}_safe"
    if isinstance(sim, AbstractMonteCarlo):
        if any(uq._is_imprecise(i) for inp in scen for i in inp):
            raise TypeError("imprecise inputs with a precise simulation")
        pf, sd, samples, _ = uq.pf_scenarios(node.models, node.performance, scen, sim)   # ONE device call
        for s, ev in enumerate(evidences):
            cpt[{**ev, node.name: fail}] = float(pf[s])
            cpt[{**ev, node.name: safe}] = 1 - float(pf[s])
            if collect_samples:
                info[tuple(ev.values())] = {"cov": float(sd[s]), "samples": samples[s]}
    elif isinstance(sim, FORM):     # evaluate_node.jl:97-102
        for inp, ev in zip(scen, evidences):
            pf, beta, dp, alpha = uq.form(node.models, node.performance, inp, sim)
            cpt[{**ev, node.name: fail}] = pf
            cpt[{**ev, node.name: safe}] = 1 - pf
            if collect_samples:
                info[tuple(ev.values())] = {"β": beta, "design_point": dp, "α": alpha}
    elif isinstance(sim, CudaDoubleLoop):
        res = uq.double_loop(node.models, node.performance, scen, sim)
        for (lb, ub), ev in zip(res, evidences):
            cpt[{**ev, node.name: fail}] = (lb, ub)
            cpt[{**ev, node.name: safe}] = (1 - ub, 1 - lb)
            if collect_samples:
                info[tuple(ev.values())] = {}
    elif isinstance(sim, CudaRandomSlicing):
        res = uq.random_slicing(node.models, node.performance, scen, sim)
        for ((lb, ub), ilb, iub), ev in zip(res, evidences):
            cpt[{**ev, node.name: fail}] = (lb, ub)
            cpt[{**ev, node.name: safe}] = (1 - ub, 1 - lb)
            if collect_samples:
                info[tuple(ev.values())] = {"lb": ilb, "ub": iub}
    else:
        raise TypeError(f"simulation {sim!r} has no device path")
    return DiscreteNode(node.name, cpt, node.parameters, info)


def _evaluate_continuous(net, node: ContinuousFunctionalNode, collect_samples=True) -> ContinuousNode:
    if not all(isprecise(p) for p in parents(net, node)[2]):
        raise ValueError(f"node {node.name} is a continuousfunctionalnode with at least one parent with Interval or "
                         "p-boxes in its distributions. No method for extracting failure probability p-box have been "
                         "implemented yet")
    ancestors, combos, evidences = _evidences(net, node)
    if len(combos[0]) == 0:
        cpt = ContinuousConditionalProbabilityTable([], kind="precise")
        disc = ExactDiscretization(node.discretization.intervals)
        evidences = [{}]
    else:
        cpt = ContinuousConditionalProbabilityTable([a.name for a in ancestors], kind="precise")
        disc = node.discretization
    scen = [_scenario_inputs(net, node, ev) for ev in evidences]
    edges, counts, s1, s2, job, low = uq.output_histograms(node.models, scen, node.simulation)  # device
    info = {}
    for s, ev in enumerate(evidences):
        cpt[ev] = EmpiricalDistribution(edges, counts[s], job.n, float(s1[s]), float(s2[s]))
        # outputs below the first / above the last edge (heavy tails) or NaN: the shared grid is mean +- 8 sd of a
        # pilot run, so this is normally 0; say so when it is not instead of folding it away silently
        outside = float(counts[s][0] + counts[s][-1]) / float(job.n)
        if outside > 1e-3:
            import warnings
            warnings.warn(f"node {node.name}, scenario {tuple(ev.values())}: {100 * outside:.2f} % of the outputs fall "
                          "outside the histogram grid (heavy tail or NaN); they are kept in the edge bins")
        if collect_samples:
            info[tuple(ev.values())] = {"samples": uq.Samples(job, s, low.out_columns), "mass_outside_grid": outside}
    return ContinuousNode(node.name, cpt, disc, info)


def _evaluate_node(net, node, collect_samples=True):
    if isinstance(node, DiscreteFunctionalNode):
        return _evaluate_discrete(net, node, collect_samples)
    return _evaluate_continuous(net, node, collect_samples)


def _is_eliminable(net, node) -> bool:
    """evaluate_net.jl:42-56: reversing the node's edges must not create a cycle."""
    if isinstance(node, int):   # a topology index (evaluate_net.jl:57-61; 0-based here), not a position in net.nodes
        node = next(name for name, i in net.topology_dict.items() if i == node)
    if isinstance(node, str):
        node = net.node(node)
    if not isinstance(node, (ContinuousNode, ContinuousFunctionalNode)):
        raise TypeError(f"node elimination algorithm is for continuous nodes and {node.name} is discrete")
    idx = net.topology_dict[node.name]
    adj = net.adj_matrix.copy()
    for p in parents(net, idx)[0]:
        adj[p, idx], adj[idx, p] = 0, 1
    for c in children(net, idx)[0]:
        adj[idx, c], adj[c, idx] = 0, 1
    return not _is_cyclic_dfs(adj)


def evaluate(net: EnhancedBayesianNetwork, check: bool = True, collect_samples: bool = True):
    """`evaluate!(net, check, collect_samples)` (evaluate_net.jl:1-32)."""
    _discretize_net(net)
    _transfer_continuous(net)
    functional = [n for n in net.nodes if isinstance(n, FunctionalNode)]
    while functional:
        first = functional[0]
        cont = [p for p in parents(net, first)[2] if isinstance(p, ContinuousNode)]
        bad = [c.name for c in cont if not _is_eliminable(net, c)]
        if bad and check:
            raise ValueError(f"nodes elimnation algorithm will lead to a cyclic network when elimnating node/s {bad}")
        try:
            evaluated = _evaluate_node(net, first, collect_samples)
        except AssertionError:
            raise ValueError(f"node {first.name} has as simulation {first.simulation}, but its imprecise parents will "
                             "be discretized and approximated with Uniform and Exponential assumption, therefore are no "
                             "longer imprecise. A prices simulation technique must be selected!") from None
        except (uq.L.EbnError, uq.UnsupportedModel):
            raise  # device / catalogue errors are never rewritten (INTEGRATION.md "errors")
        except TypeError as e:
            imp = [p.name for p in parents(net, first)[2] if not isprecise(p)]
            raise ValueError(f"node {first.name} has {first.simulation} as simulation technique, but have {imp} as "
                             "imprecise parent/s. DoubleLoop or RandomSlicing technique must be employeed instead") from e
        net.nodes[net.nodes.index(first)] = evaluated
        _discretize_net(net)
        _transfer_continuous(net)
        functional = [n for n in net.nodes if isinstance(n, FunctionalNode)]


# ---------------------------------------------------------------------------------------------
# reduction / dispatch (post-steps)
# ---------------------------------------------------------------------------------------------
def _eliminate_continuous_node(net, node):
    ip, ic = parents(net, node)[0], children(net, node)[0]
    for p in ip:
        for c in ic:
            net.adj_matrix[p, c] = 1
    _remove_node(net, node)


def reduce(net, check: bool = True, collect_samples: bool = True):
    """`reduce!` (reduction.jl:2-9)."""
    if any(isinstance(n, FunctionalNode) for n in net.nodes):
        evaluate(net, check, collect_samples)
    for n in [n for n in net.nodes if isinstance(n, ContinuousNode)]:
        _eliminate_continuous_node(net, n)


def dispatch(net):
    """dispatch.jl:1-15."""
    if any(isinstance(n, ContinuousNode) for n in net.nodes):
        reduce(net)
    if not any(isinstance(n, ContinuousNode) for n in net.nodes):
        if all(isprecise(n) for n in net.nodes):
            return BayesianNetwork(net.nodes, net.topology_dict, net.adj_matrix)
        return CredalNetwork(net.nodes, net.topology_dict, net.adj_matrix)
    return net


# ---------------------------------------------------------------------------------------------
# evaluate_with_envelopes (evaluate_net.jl:34-40, :68-95)
# ---------------------------------------------------------------------------------------------
def _add_root2envelope(net, envelope):
    envelope = list(envelope)
    while True:
        missing = []
        for n in envelope:
            for p in parents(net, n)[2]:
                if not any(p is e for e in envelope) and not any(p is m for m in missing):
                    missing.append(p)
        if not missing:
            return envelope
        envelope += missing


def _build_envelope_edges(net, envelope):
    sub = EnhancedBayesianNetwork(envelope)
    for n in envelope:
        for p in parents(net, n)[2]:
            if any(p is e for e in envelope):
                add_child(sub, p, n)
    order(sub)
    return sub


def evaluate_with_envelopes(net):
    """One sub-network per Markov envelope, each evaluated on its own."""
    from .network import markov_envelope
    subs = [_build_envelope_edges(net, _add_root2envelope(net, env)) for env in markov_envelope(net)]
    for s in subs:
        evaluate(s)
    return subs
