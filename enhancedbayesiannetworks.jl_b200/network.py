"""EnhancedBayesianNetwork container and the graph helpers the hot path uses per scenario.

  reference                                              here
  src/networks/ebn/enhancedbn.jl:1-61                    EnhancedBayesianNetwork, _is_cyclic_dfs
  src/networks/networks_common.jl:1-56                   parents, children, discrete_ancestors
  src/networks/networks_common.jl:58-121                 add_child, order
  src/networks/networks_common.jl:123-160                _remove_node, _add_node
  src/networks/networks_common.jl:162-218                _verify_net and friends
  src/util/verification_add_child.jl                     the add_child checks
Julia's `f!(net)` mutators keep their name without the bang; indices are 0-based.
"""
from __future__ import annotations

import itertools
import warnings
from typing import Dict, List, Sequence, Tuple

import numpy as np

from .nodes import AbstractContinuousNode, AbstractDiscreteNode, DiscreteNode, FunctionalNode


def _is_cyclic_dfs(adj: np.ndarray) -> bool:
    n = adj.shape[0]
    state = [0] * n  # 0 unvisited, 1 on stack, 2 done

    def dfs(v):
        state[v] = 1
        for w in np.flatnonzero(adj[v]):
            if state[w] == 1 or (state[w] == 0 and dfs(int(w))):
                return True
        state[v] = 2
        return False

    return any(state[v] == 0 and dfs(v) for v in range(n))


class EnhancedBayesianNetwork:
    def __init__(self, nodes: Sequence, topology_dict: Dict[str, int] = None, adj_matrix: np.ndarray = None):
        nodes = list(nodes)
        names = [n.name for n in nodes]
        if len(set(names)) != len(names):
            raise ValueError("network nodes names must be unique")
        states = [s for n in nodes if isinstance(n, DiscreteNode) for s in n.states()]
        if len(set(states)) != len(states):
            raise ValueError("network nodes states must be unique")
        self.nodes = nodes
        self.topology_dict = dict(topology_dict) if topology_dict is not None else {n: i for i, n in enumerate(names)}
        self.adj_matrix = adj_matrix if adj_matrix is not None else np.zeros((len(nodes), len(nodes)), dtype=np.int8)

    def node(self, name: str):
        for n in self.nodes:
            if n.name == name:
                return n
        raise KeyError(name)

    def __repr__(self):
        return f"EnhancedBayesianNetwork({[n.name for n in self.nodes]})"


def _index(net, x) -> int:
    if isinstance(x, (int, np.integer)):
        return int(x)
    return net.topology_dict[x if isinstance(x, str) else x.name]


def parents(net, x) -> Tuple[List[int], List[str], List]:
    """(indices, names, nodes): nodes in `net.nodes` order (networks_common.jl:1-7)."""
    idx = _index(net, x)
    rev = {v: k for k, v in net.topology_dict.items()}
    indices = [int(i) for i in np.flatnonzero(net.adj_matrix[:, idx])]
    names = [rev[i] for i in indices]
    return indices, names, [n for n in net.nodes if n.name in names]


def children(net, x) -> Tuple[List[int], List[str], List]:
    idx = _index(net, x)
    rev = {v: k for k, v in net.topology_dict.items()}
    indices = [int(i) for i in np.flatnonzero(net.adj_matrix[idx, :])]
    names = [rev[i] for i in indices]
    return indices, names, [n for n in net.nodes if n.name in names]


def discrete_ancestors(net, node) -> List:
    """Discrete parents, then the discrete ancestors reached through continuous parents
    (networks_common.jl:37-45)."""
    pars = parents(net, node)[2]
    disc = [p for p in pars if isinstance(p, AbstractDiscreteNode)]
    cont = [p for p in pars if isinstance(p, AbstractContinuousNode)]
    out = list(disc)
    for c in cont:
        for a in discrete_ancestors(net, c):
            if not any(a is o for o in out):
                out.append(a)
    return out


def _theoretical_scenarios(net, node) -> List[Dict[str, str]]:
    pars = [p for p in discrete_ancestors(net, node) if isinstance(p, AbstractDiscreteNode)]
    combos = itertools.product(*[[(p.name, s) for s in p.states()] for p in pars])
    return [dict(c) for c in combos]


def add_child(net, par, ch):
    """networks_common.jl:58-83 with the checks of src/util/verification_add_child.jl."""
    ip, ic = _index(net, par), _index(net, ch)
    rev = {v: k for k, v in net.topology_dict.items()}
    p, c = net.node(rev[ip]), net.node(rev[ic])
    if p is c:
        raise ValueError(f"Recursion on the same node '{p.name}' is not allowed in EnhancedBayesianNetworks")
    if c.isroot():
        raise ValueError(f"node '{c.name}' is a root node and cannot have parents")
    if not isinstance(c, FunctionalNode) and not isinstance(p, FunctionalNode):
        if p.name not in c.cpt.names:
            raise ValueError(f"trying to set node '{c.name}' as child of node '{p.name}', but '{c.name}' has a cpt "
                             f"that does not contains '{p.name}' in the scenarios: {c.cpt}")
        seen = set(c.cpt.column(p.name))
        if any(s not in seen for s in p.states()):
            raise ValueError(f"child node '{c.name}' has scenarios {sorted(seen)}, that is not coherent with its "
                             f"parent node '{p.name}' with states {p.states()}")
    if isinstance(p, FunctionalNode) and not isinstance(c, FunctionalNode):
        raise ValueError(f"functional node '{p.name}' can have only functional children. '{c.name}' is not a "
                         "functional node")
    net.adj_matrix[ip, ic] = 1


def _verify_net(net):
    for n in net.nodes:
        if isinstance(n, FunctionalNode):
            pars = parents(net, n)[2]
            if not pars:
                raise ValueError(f"functional node '{n.name}' must have at least one parent")
            if not any(isinstance(p, AbstractContinuousNode) for p in pars):
                warnings.warn(f"functional node '{n.name}' have no continuous parents. All the simulations will "
                              "return the same output")
            for dp in (p for p in pars if isinstance(p, DiscreteNode)):
                if not dp.parameters:
                    raise ValueError(f"node '{dp.name}' is a discrete parent of a functional node and cannot have an "
                                     "empty parameters vector")
        elif not n.isroot():
            cols = [c for c in n.cpt.names if c != n.name]
            if set(cols) != set(parents(net, n)[1]):
                raise ValueError(f"node '{n.name}''s cpt requires exctly the nodes '{cols}' to be its parents, but "
                                 f"provided parents are '{parents(net, n)[1]}'")
            th = _theoretical_scenarios(net, n)
            got = n.scenarios()
            if sorted(map(lambda d: sorted(d.items()), th)) != sorted(map(lambda d: sorted(d.items()), got)):
                raise ValueError(f"node '{n.name}' has defined cpt scenarios {n.cpt} not coherent with the "
                                 f"theoretical one {th}")


def order(net):
    """Topological re-ordering of nodes / adjacency / topology_dict (networks_common.jl:85-121)."""
    if _is_cyclic_dfs(net.adj_matrix):
        raise ValueError("network is cyclic!")
    n = net.adj_matrix.shape[0]
    rev = {v: k for k, v in net.topology_dict.items()}
    roots = [j for j in range(n) if not net.adj_matrix[:, j].any()]
    todo = [j for j in range(n) if j not in roots]
    while todo:
        new = [j for j in todo if all(p in roots for p in np.flatnonzero(net.adj_matrix[:, j]))]
        roots += new
        todo = [j for j in todo if j not in new]
    conv = {old: new for new, old in enumerate(roots)}
    adj = np.zeros_like(net.adj_matrix)
    for i, j in zip(*np.nonzero(net.adj_matrix)):
        adj[conv[int(i)], conv[int(j)]] = 1
    net.nodes = [net.node(rev[old]) for old in roots]
    net.topology_dict = {rev[old]: new for new, old in enumerate(roots)}
    net.adj_matrix = adj
    _verify_net(net)


def _remove_node(net, x):
    idx = _index(net, x)
    rev = {v: k for k, v in net.topology_dict.items()}
    name = rev[idx]
    keep = [i for i in range(net.adj_matrix.shape[0]) if i != idx]
    net.adj_matrix = net.adj_matrix[np.ix_(keep, keep)]
    net.nodes = [n for n in net.nodes if n.name != name]
    net.topology_dict = {k: (v - 1 if v > idx else v) for k, v in net.topology_dict.items() if k != name}


def _add_node(net, node):
    net.nodes.append(node)
    n = net.adj_matrix.shape[0]
    net.topology_dict[node.name] = n
    adj = np.zeros((n + 1, n + 1), dtype=net.adj_matrix.dtype)
    adj[:n, :n] = net.adj_matrix
    net.adj_matrix = adj


class BayesianNetwork:
    """What `dispatch` returns when every node is discrete and precise (src/networks/bn/bayesnet.jl).
    Container only: discrete inference is outside the hot path."""

    def __init__(self, nodes, topology_dict, adj_matrix):
        self.nodes, self.topology_dict, self.adj_matrix = list(nodes), dict(topology_dict), adj_matrix


class CredalNetwork(BayesianNetwork):
    """All-discrete network with at least one imprecise node (src/networks/cn/credalnet.jl)."""


# ---- Markov blankets / envelopes (src/networks/ebn/enhancedbn.jl:63-118) -----------------------
def markov_blanket(net, x) -> Tuple[List[int], List[str], List]:
    """Parents, children and the children's other parents of a node (enhancedbn.jl:63-74)."""
    idx = _index(net, x)
    rev = {v: k for k, v in net.topology_dict.items()}
    blanket: List[int] = []
    for ch in children(net, idx)[0]:
        blanket += parents(net, ch)[0]
        blanket.append(ch)
    blanket += parents(net, idx)[0]
    indices = []
    for i in blanket:
        if i != idx and i not in indices:
            indices.append(i)
    names = [rev[i] for i in indices]
    return indices, names, [n for n in net.nodes if n.name in names]


def _get_markov_group(net, node) -> List:
    """Closure of a continuous node under "continuous members of the Markov blanket"
    (enhancedbn.jl:86-96)."""
    def grow(lst):
        out = list(lst)
        for n in lst:
            for m in markov_blanket(net, node)[2]:
                if isinstance(m, AbstractContinuousNode) and not any(m is o for o in out):
                    out.append(m)
        return out
    cur = [node]
    new = grow(cur)
    while {n.name for n in new} != {n.name for n in cur}:
        cur, new = new, grow(new)
    return new


def markov_envelope(net) -> List[List]:
    """Groups of nodes that have to be evaluated together (enhancedbn.jl:98-118)."""
    cont = [n for n in net.nodes if isinstance(n, AbstractContinuousNode)]
    envelopes = []
    for c in cont:
        env = []
        for x in _get_markov_group(net, c):
            for m in markov_blanket(net, x)[2] + [x]:
                if not any(m is o for o in env):
                    env.append(m)
        envelopes.append(env)
    envelopes.sort(key=len, reverse=True)
    final = []
    while envelopes:
        head, rest = envelopes[0], envelopes[1:]
        names = {n.name for n in head}
        envelopes = [e for e in rest if any(n.name not in names for n in e)]
        final.append(head)
    return final
