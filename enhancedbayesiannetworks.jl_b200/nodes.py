"""Node and CPT data model of the hot path — a host-side mirror of the reference's Julia types so
that the parity tests read like the reference's own tests.

  reference file                         here
  src/nodes/cpts.jl                      DiscreteConditionalProbabilityTable, Continuous...
  src/nodes/nodes.jl                     ExactDiscretization, ApproximatedDiscretization, UnamedProbabilityBox
  src/nodes/discretenodes.jl:1-49        DiscreteNode (+ _uq_inputs)
  src/nodes/continuousnodes.jl:1-81      ContinuousNode (+ _uq_inputs, _truncate, _distribution_bounds)
  src/nodes/functionalnodes.jl           DiscreteFunctionalNode, ContinuousFunctionalNode
  src/util/verify_discrete.jl            _verify_probabilities (sum to one, 0.05 auto-normalise)

Julia Symbols are Python strings; a DataFrame is a list of row dicts with an ordered column list
(the `Π` column is called "Π" here as well).
"""
from __future__ import annotations

import math
import warnings
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

from .distributions import (Interval, Parameter, ProbabilityBox, RandomVariable, Truncated)

PI = "Π"
Evidence = Dict[str, str]


def jl_float(x: float) -> str:
    """Julia's `string(::Float64)` for the values that appear in state names."""
    if math.isinf(x):
        return "Inf" if x > 0 else "-Inf"
    r = repr(float(x))
    if "e" in r:
        m, e = r.split("e")
        if "." not in m:
            m += ".0"
        return f"{m}e{int(e)}"
    return r


def interval_symbol(iv: Sequence[float]) -> str:
    """`Symbol([a, b])` → "[a, b]" (discretize.jl:35)."""
    return "[" + ", ".join(jl_float(v) for v in iv) + "]"


# ---------------------------------------------------------------------------------------------
# CPTs (src/nodes/cpts.jl)
# ---------------------------------------------------------------------------------------------
class _CPT:
    def __init__(self, names: Union[str, Sequence[str], None] = None):
        if names is None:
            names = []
        self.names: List[str] = [names] if isinstance(names, str) else list(names)
        self.data: List[Dict[str, Any]] = []

    # cpt[evidence] = value  (cpts.jl:81-102). The reference subsets the DataFrame per call
    # (O(S) each, O(S^2) per node — SURVEY.md §8f row 4); here a hash index over the key columns
    # makes the write-back O(1) per row.
    def _key(self, ev) -> tuple:
        return tuple(ev[n] for n in self.names)

    def _reindex(self):
        self._index = {tuple(r[n] for n in self.names): r for r in self.data}
        self._index_rows, self._index_names = len(self.data), tuple(self.names)

    def __setitem__(self, key, value):
        ev = self._as_evidence(key)
        if set(ev) != set(self.names):
            raise ValueError(f"Cannot set index with {list(ev)} into a CPT initialized with {self.names}")
        if getattr(self, "_index_rows", -1) != len(self.data) or getattr(self, "_index_names", None) != tuple(self.names):
            self._reindex()
        k = self._key(ev)
        row = self._index.get(k)
        if row is not None:
            row[PI] = value
            return
        row = {**{n: ev[n] for n in self.names}, PI: value}
        self.data.append(row)
        self._index[k] = row
        self._index_rows = len(self.data)

    def __getitem__(self, key):
        ev = self._as_evidence(key)
        if set(ev) == set(self.names):
            if getattr(self, "_index_rows", -1) != len(self.data) or getattr(self, "_index_names", None) != tuple(self.names):
                self._reindex()
            row = self._index.get(self._key(ev))
            if row is None:
                raise KeyError(f"index not find in the CPT {self}")
            return row[PI]
        hits = [r for r in self.data if all(r.get(k) == v for k, v in ev.items())]
        if not hits:
            raise KeyError(f"index not find in the CPT {self}")
        assert len(hits) == 1
        return hits[0][PI]

    @staticmethod
    def _as_evidence(key) -> Evidence:
        if key is None or key == ():
            return {}
        if isinstance(key, dict):
            return dict(key)
        if isinstance(key, tuple) and len(key) == 2 and isinstance(key[0], str):
            return {key[0]: key[1]}
        return {k: v for k, v in key}

    def column(self, name: str) -> List[Any]:
        return [r[name] for r in self.data]

    def sort(self):
        self.data.sort(key=lambda r: tuple(str(r[n]) for n in self.names))

    def subset(self, ev: Evidence) -> List[Dict[str, Any]]:
        return [r for r in self.data if all(r[k] == v for k, v in ev.items())]

    def __eq__(self, other):
        return type(self) is type(other) and self.names == other.names and \
            [{k: r[k] for k in self.names + [PI]} for r in self.data] == \
            [{k: r[k] for k in other.names + [PI]} for r in other.data] \
            and getattr(self, "precise", None) == getattr(other, "precise", None)

    def __repr__(self):
        return f"{type(self).__name__}({self.names}, {len(self.data)} rows)"


class DiscreteConditionalProbabilityTable(_CPT):
    """`DiscreteConditionalProbabilityTable{P}(names)`; precise=True ⇔ P = PreciseDiscreteProbability."""

    def __init__(self, names, precise: bool = True):
        super().__init__(names)
        self.precise = precise

    def isroot(self) -> bool:
        return len(self.names) == 1

    def states(self, name: str) -> List[str]:
        out = []
        for r in self.data:
            if r[name] not in out:
                out.append(r[name])
        return out

    def scenarios(self, name: str) -> List[Evidence]:
        out = []
        for r in self.data:
            s = {k: r[k] for k in self.names if k != name}
            if s not in out:
                out.append(s)
        return out


class ContinuousConditionalProbabilityTable(_CPT):
    """`ContinuousConditionalProbabilityTable{P}(names)`; kind in {"precise", "interval", "pbox"}."""

    def __init__(self, names=None, kind: str = "precise"):
        super().__init__(names)
        self.kind = kind

    @property
    def precise(self) -> bool:
        return self.kind == "precise"

    def isroot(self) -> bool:
        return len(self.names) == 0

    def distributions(self):
        return self.column(PI)

    def scenarios(self) -> List[Evidence]:
        out = []
        for r in self.data:
            s = {k: r[k] for k in self.names}
            if s not in out:
                out.append(s)
        return [s for s in out if s]


# ---------------------------------------------------------------------------------------------
# discretisations and p-boxes (src/nodes/nodes.jl)
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class UnamedProbabilityBox:
    family: type
    parameters: Tuple[Interval, ...]
    lb: float = -math.inf
    ub: float = math.inf


class ExactDiscretization:
    def __init__(self, intervals: Sequence[float] = ()):
        intervals = list(intervals)
        if intervals != sorted(intervals):
            raise ValueError(f"interval values {intervals} are not sorted")
        self.intervals = intervals

    def __eq__(self, o):
        return type(o) is ExactDiscretization and self.intervals == o.intervals

    def __repr__(self):
        return f"ExactDiscretization({self.intervals})"


class ApproximatedDiscretization:
    def __init__(self, intervals: Sequence[float] = (), sigma: float = 0):
        intervals = list(intervals)
        if intervals != sorted(intervals):
            raise ValueError(f"interval values {intervals} are not sorted")
        if sigma < 0:
            raise ValueError("variance must be positive")
        if sigma > 2:
            warnings.warn(f"Selected variance values {sigma} can be too big, and the approximation not realistic")
        self.intervals, self.sigma = intervals, sigma

    def __eq__(self, o):
        return type(o) is ApproximatedDiscretization and self.intervals == o.intervals and self.sigma == o.sigma

    def __repr__(self):
        return f"ApproximatedDiscretization({self.intervals}, {self.sigma})"


# ---------------------------------------------------------------------------------------------
# plain nodes
# ---------------------------------------------------------------------------------------------
def _verify_probabilities(cpt: DiscreteConditionalProbabilityTable, name: str):
    """src/util/verify_discrete.jl:1-45."""
    for r in cpt.data:
        for p in ([r[PI]] if cpt.precise else list(r[PI])):
            if p < 0:
                raise ValueError(f"probabilities must be non-negative: {p}")
            if p > 1:
                raise ValueError(f"probabilities must be lower or equal than 1: {p}")
        if not cpt.precise and r[PI][0] > r[PI][1]:
            raise ValueError(f"interval probabilities must have a lower bound smaller than upper bound. {cpt.data}")
    groups: Dict[tuple, list] = {}
    others = [n for n in cpt.names if n != name]
    for r in cpt.data:
        groups.setdefault(tuple(r[n] for n in others), []).append(r)
    for rows in groups.values():
        if cpt.precise:
            tot = sum(r[PI] for r in rows)
            if tot != 1:
                if abs(tot - 1) <= 0.05:
                    if abs(tot - 1) > 1e-12:  # pf + (1 - pf) can miss 1 by an ulp: normalise silently
                        warnings.warn(f"total probability should be one, but the evaluated value is {tot}, and will be normalized")
                    for r in rows:
                        r[PI] = r[PI] / tot
                else:
                    raise ValueError(f"states are not exhaustive and mutually exclusives for the following cpt: {rows}")
        else:
            if sum(r[PI][0] for r in rows) > 1:
                raise ValueError(f"sum of intervals lower bounds is bigger than 1: {rows}")
            if sum(r[PI][1] for r in rows) < 1:
                raise ValueError(f"sum of intervals upper bounds is smaller than 1: {rows}")


class AbstractNode:
    name: str


class DiscreteNode(AbstractNode):
    """src/nodes/discretenodes.jl:1-22."""

    def __init__(self, name: str, cpt: DiscreteConditionalProbabilityTable,
                 parameters: Optional[Dict[str, List[Parameter]]] = None, additional_info: Optional[dict] = None):
        if name not in cpt.names:
            raise ValueError(f"defined cpt does not contain a column refered to node name {name}: {cpt}")
        _verify_probabilities(cpt, name)
        parameters = parameters or {}
        if parameters and set(cpt.states(name)) != set(parameters):
            raise ValueError(f"parameters keys {list(parameters)} must be coherent with states {cpt.states(name)}")
        cpt.names = [n for n in cpt.names if n != name] + [name]  # node column last before Π
        cpt.sort()
        self.name, self.cpt, self.parameters = name, cpt, parameters
        self.additional_info = additional_info if additional_info is not None else {}

    def states(self):
        return self.cpt.states(self.name)

    def scenarios(self):
        return self.cpt.scenarios(self.name)

    def isprecise(self):
        return self.cpt.precise

    def isroot(self):
        return self.cpt.isroot()

    def _uq_inputs(self, evidence: Evidence) -> List[Parameter]:
        """discretenodes.jl:43-49."""
        if self.name not in evidence:
            raise KeyError(f"evidence {evidence} does not contain the node {self.name}")
        return list(self.parameters[evidence[self.name]])

    def __eq__(self, o):
        return isinstance(o, DiscreteNode) and (self.name, self.cpt, self.parameters) == (o.name, o.cpt, o.parameters)

    def __hash__(self):
        return hash(("DiscreteNode", self.name))

    def __repr__(self):
        return f"DiscreteNode({self.name})"


class ContinuousNode(AbstractNode):
    """src/nodes/continuousnodes.jl:1-29."""

    def __init__(self, name: str, cpt: ContinuousConditionalProbabilityTable, discretization=None,
                 additional_info: Optional[dict] = None):
        if discretization is None:
            discretization = ExactDiscretization() if cpt.isroot() else ApproximatedDiscretization()
        if cpt.isroot() and isinstance(discretization, ApproximatedDiscretization):
            raise ValueError(f"Root node must have ExactDiscretization as discretization structure, provided "
                             f"discretization is {discretization} and node cpt is {cpt}")
        if not cpt.isroot() and isinstance(discretization, ExactDiscretization):
            raise ValueError(f"Child node must have ApproximatedDiscretization as discretization structure, "
                             f"provided discretization is {discretization} and node cpt is {cpt}")
        cpt.sort()
        self.name, self.cpt, self.discretization = name, cpt, discretization
        self.additional_info = additional_info if additional_info is not None else {}

    def distributions(self):
        return self.cpt.distributions()

    def scenarios(self):
        return self.cpt.scenarios()

    def isprecise(self):
        return self.cpt.precise

    def isroot(self):
        return self.cpt.isroot()

    def _uq_inputs(self, evidence: Optional[Evidence] = None):
        """continuousnodes.jl:39-50: the scenario's row(s) as RandomVariable / Interval / ProbabilityBox."""
        evidence = evidence or {}
        ev = {k: v for k, v in evidence.items() if k in self.cpt.names}
        rows = self.cpt.subset(ev)
        out = []
        for r in rows:
            d = r[PI]
            if self.cpt.kind == "precise":
                out.append(RandomVariable(d, self.name))
            elif isinstance(d, tuple):
                out.append(Interval(d[0], d[1], self.name))
            else:
                out.append(ProbabilityBox(d.family, tuple(d.parameters), self.name, d.lb, d.ub))
        return out

    def _distribution_bounds(self) -> Tuple[float, float]:
        """continuousnodes.jl:54-69."""
        los, his = [], []
        for d in self.cpt.distributions():
            if isinstance(d, tuple):
                lo, hi = d
            elif isinstance(d, UnamedProbabilityBox):
                lo = min([p.lb for p in d.parameters] + [d.lb])
                hi = max([p.ub for p in d.parameters] + [d.ub])
            else:
                lo, hi = d.support()
            los.append(lo)
            his.append(hi)
        return min(los), max(his)

    def __eq__(self, o):
        return isinstance(o, ContinuousNode) and (self.name, self.cpt, self.discretization) == \
            (o.name, o.cpt, o.discretization)

    def __hash__(self):
        return hash(("ContinuousNode", self.name))

    def __repr__(self):
        return f"ContinuousNode({self.name})"


def _truncate(dist, iv: Sequence[float]):
    """continuousnodes.jl:71-81."""
    if isinstance(dist, tuple):
        return (iv[0], iv[1])
    if isinstance(dist, UnamedProbabilityBox):
        return UnamedProbabilityBox(dist.family, dist.parameters, iv[0], iv[1])
    return Truncated(dist, iv[0], iv[1])


# ---------------------------------------------------------------------------------------------
# models, simulations, functional nodes (src/nodes/functionalnodes.jl)
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Poly:
    """A polynomial of DataFrame columns: sum of coef * prod(columns), evaluated left to right.
    `Poly([(1, ["A"]), (1, ["B"])])` is `df.A .+ df.B`; `Poly([(2, []), (-1, ["C"])])` is `2 .- df.C`."""
    terms: Tuple[Tuple[float, Tuple[str, ...]], ...]

    def __init__(self, terms):
        object.__setattr__(self, "terms", tuple((float(c), tuple(cols)) for c, cols in terms))


@dataclass(frozen=True)
class Catalogue:
    """A limit state compiled into the kernels (include/ebncuda.h EBN_LS_*). `args` names the
    DataFrame columns in the functor's argument order; `consts` its constants."""
    limit_state: str
    consts: Tuple[float, ...]
    args: Tuple[str, ...]

    def __init__(self, limit_state: str, consts=(), args=()):
        object.__setattr__(self, "limit_state", limit_state)
        object.__setattr__(self, "consts", tuple(float(c) for c in consts))
        object.__setattr__(self, "args", tuple(args))


@dataclass(frozen=True)
class Expression:
    """The body of a model closure (or of `performance`) as text, Julia or Python syntax:
    `Model(Expression("df.A .+ df.B"), "C")` stands for `Model(df -> df.A .+ df.B, :C)`.
    Keyword arguments are named constants of the formula; they travel as run-time constants, so the
    same formula with other values reuses the compiled kernel. Compiled for the GPU at run time
    (expr.py → EBN_LS_EXPRESSION → NVRTC)."""
    text: str
    constants: Tuple[Tuple[str, float], ...]

    def __init__(self, text: str, **constants):
        object.__setattr__(self, "text", text)
        object.__setattr__(self, "constants", tuple((k, float(v)) for k, v in constants.items()))


@dataclass(frozen=True)
class Model:
    """UQ.jl `Model(func, name)`. A Python closure cannot run on the GPU: `func` is an `Expression`
    (the closure's formula as text), a `Poly`, or a `Catalogue` entry; anything else raises when the
    node is evaluated (no CPU fallback)."""
    func: Any
    name: str


@dataclass(frozen=True)
class CudaMonteCarlo:
    """The `simulation` of a functional node that routes it to libebncuda
    (`struct CudaMonteCarlo <: AbstractMonteCarlo` in julia/src/cuda/EBNCuda.jl)."""
    n: int
    seed: int = 0x5EED0001
    strict: bool = False
    fp32_sampling: bool = False   # EBN_JOB_FP32_SAMPLING: untruncated normals through the fp32 Box-Muller
    qmc: bool = False             # EBN_JOB_QMC_SOBOL (set by CudaSobolSampling)


MonteCarlo = CudaMonteCarlo


@dataclass(frozen=True)
class CudaSobolSampling(CudaMonteCarlo):
    """UQ.jl `SobolSampling(n)`, a quasi-Monte Carlo `AbstractMonteCarlo` (accepted by
    `ContinuousFunctionalNode.simulation::AbstractMonteCarlo`, src/nodes/functionalnodes.jl:4): sample i of a
    scenario is point i of an Owen-scrambled Sobol sequence, every marginal through its inverse CDF.
    A power of two for `n` keeps the point set a net."""
    qmc: bool = True


SobolSampling = CudaSobolSampling


@dataclass(frozen=True)
class CudaDoubleLoop:
    """UQ.jl `DoubleLoop(sim)`: outer search over interval / p-box parameters, inner Monte Carlo.
    `method`: "optimise" (bounded derivative-free search per bound, like the reference's BOBYQA) or
    "grid" (sweep). Inner estimates use common random numbers, so the objective is deterministic."""
    lb: CudaMonteCarlo
    ub: Optional[CudaMonteCarlo] = None
    method: str = "optimise"
    grid: int = 33


DoubleLoop = CudaDoubleLoop


@dataclass(frozen=True)
class FORM:
    """UQ.jl `FORM(n, tol)`: first-order reliability method (Hasofer-Lind / Rackwitz-Fiessler
    iteration in standard-normal space). Not a sampling method: each iteration evaluates the limit
    state on 2d+1 points, which go to the device in one `ebn_eval_matrix` call (strict arithmetic).
    The result lands in the reference's FORM branch (evaluate_node.jl:97-102): pf plus β, design
    point and α in `additional_info`."""
    n: int = 10
    tol: float = 1e-3
    h: float = 1e-4      # central finite-difference step in standard-normal space


@dataclass(frozen=True)
class CudaRandomSlicing:
    """UQ.jl `RandomSlicing(sim)`: per sample, the p-box/interval inputs become boxes and the model
    is bounded over them; supported on the device for models that are monotone in each imprecise
    input (polynomial programs are checked term by term)."""
    lb: CudaMonteCarlo
    ub: Optional[CudaMonteCarlo] = None


RandomSlicing = CudaRandomSlicing
AbstractMonteCarlo = (CudaMonteCarlo,)


class DiscreteFunctionalNode(AbstractNode):
    def __init__(self, name: str, models, performance, simulation, parameters: Optional[dict] = None):
        self.name = name
        self.models: List[Model] = [models] if isinstance(models, Model) else list(models)
        self.performance = performance
        self.simulation = simulation
        self.parameters = parameters or {}

    def isroot(self):
        return False

    def __hash__(self):
        return hash(("DiscreteFunctionalNode", self.name))

    def __repr__(self):
        return f"DiscreteFunctionalNode({self.name})"


class ContinuousFunctionalNode(AbstractNode):
    def __init__(self, name: str, models, simulation, discretization: Optional[ApproximatedDiscretization] = None):
        if not isinstance(simulation, AbstractMonteCarlo):
            raise TypeError("ContinuousFunctionalNode.simulation must be an AbstractMonteCarlo")
        self.name = name
        self.models: List[Model] = [models] if isinstance(models, Model) else list(models)
        self.simulation = simulation
        self.discretization = discretization or ApproximatedDiscretization()

    def isroot(self):
        return False

    def __hash__(self):
        return hash(("ContinuousFunctionalNode", self.name))

    def __repr__(self):
        return f"ContinuousFunctionalNode({self.name})"


FunctionalNode = (DiscreteFunctionalNode, ContinuousFunctionalNode)
AbstractDiscreteNode = (DiscreteNode, DiscreteFunctionalNode)
AbstractContinuousNode = (ContinuousNode, ContinuousFunctionalNode)


def isprecise(node) -> bool:
    return node.isprecise() if hasattr(node, "isprecise") else True
