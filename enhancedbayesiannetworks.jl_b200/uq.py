"""`probability_of_failure` and `sample`/`evaluate!` for functional nodes, lowered to ONE C-ABI call
per node for all its scenarios.

Replaces the three UQ.jl call sites of the reference hot path:
  src/networks/ebn/evaluation/evaluate_node.jl:83      probability_of_failure(models, performance, inputs, sim)
  src/networks/ebn/evaluation/evaluate_node.jl:24-26   sample / evaluate! / EmpiricalDistribution
and the DoubleLoop / RandomSlicing branches at :104-117.
There is no CPU fallback: a model that is not an `Expression`, a `Poly` or a `Catalogue` entry raises.
"""
from __future__ import annotations

import dataclasses
import itertools
import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L
from .distributions import (INF, Gamma, Interval, JointDistribution, NeedsTable, Normal, Parameter,
                            ProbabilityBox, RandomVariable, table_for)
from . import expr as _expr
from .nodes import FORM, Catalogue, CudaDoubleLoop, CudaMonteCarlo, CudaRandomSlicing, Expression, Model, Poly

VEHICLE_ARGS = ("A", "b₀", "V", "C", "Cₖ", "K")
HIST_BINS = 2048  # fine grid a ContinuousFunctionalNode's output is histogrammed on


class UnsupportedModel(TypeError):
    """The model chain has no device functor (include/ebncuda.h EBN_LS_*)."""


@dataclass
class Lowered:
    limit_state: int
    consts: np.ndarray
    columns: List[str]          # input names in functor argument order
    hidden: List[tuple]         # extra ABI rows appended to every scenario (in-model draws)
    out_columns: List[str]      # names of all columns (inputs + model outputs) for `samples`


def _poly_program(stages: List[Poly], columns: List[str], stage_names: List[str]) -> np.ndarray:
    cols = list(columns)
    out = [float(len(stages))]
    for st, name in zip(stages, stage_names):
        out.append(float(len(st.terms)))
        for coef, names in st.terms:
            out.append(coef)
            out.append(float(len(names)))
            for nm in names:
                if nm not in cols:
                    raise UnsupportedModel(f"model column '{nm}' is not among the inputs/outputs {cols}")
                out.append(float(cols.index(nm)))
        cols.append(name)
    return np.array(out)


def _poly_text(p: Poly) -> str:
    """A Poly as expression text with the same evaluation order: ((coef*a)*b) + ((coef2*c)) + ..."""
    terms = []
    for coef, names in p.terms:
        terms.append("(" + "*".join([repr(float(coef))] + [f"df.{n}" for n in names]) + ")")
    return " + ".join(terms)


def _lower_expression(models: List[Model], performance, input_names: List[str]) -> Lowered:
    consts: Dict[str, float] = {}
    chain = []
    for m in models:
        f = m.func
        if isinstance(f, Expression):
            for k, v in f.constants:
                if consts.setdefault(k, v) != v:
                    raise UnsupportedModel(f"constant '{k}' has two different values in one model chain")
            chain.append((m.name, f.text))
        elif isinstance(f, Poly):
            chain.append((m.name, _poly_text(f)))
        else:
            raise UnsupportedModel(f"model '{m.name}': an Expression chain takes Expression and Poly models only")
    if isinstance(performance, Expression):
        for k, v in performance.constants:
            if consts.setdefault(k, v) != v:
                raise UnsupportedModel(f"constant '{k}' has two different values in one model chain")
        perf = performance.text
    elif isinstance(performance, Poly):
        perf = _poly_text(performance)
    elif isinstance(performance, str):
        perf = f"df.{performance}"
    elif performance is None and chain:
        perf = f"df.{chain[-1][0]}"
    else:
        raise UnsupportedModel("performance must be an Expression, a Poly or a column name")
    try:
        prog = _expr.compile_program(input_names, chain, perf, consts)
        L.expression_check(prog.consts, prog.d)
    except (_expr.ExpressionError, L.EbnError) as e:
        raise UnsupportedModel(str(e)) from None
    hidden = [Normal(0.0, 1.0).abi()] * prog.n_hidden
    return Lowered(L.LS_EXPRESSION, prog.consts, list(input_names), hidden, prog.columns)


# A polynomial chain runs through the interpreted PolynomialProgram functor and the runtime-dispatched
# sampling plan (7.6e10 samples/s on A + B); the same chain as a formula compiles into a kernel with a
# compile-time plan (4.2e11, profiles/r01_t_workloads.txt). Compiling costs ~0.4 s once: worth it above:
JIT_POLY_MIN_SAMPLES = 5e10


def lower(models: Sequence[Model], performance, input_names: Sequence[str], prefer_expression: bool = False) -> Lowered:
    """Maps a model chain (+ performance) onto a device limit state. `prefer_expression`: send a
    polynomial chain through the run-time compiled path too (large jobs)."""
    models = list(models)
    seen, uniq = set(), []
    for m in models:  # transmission may hand the same model twice (reference quirk, transmission.jl:13)
        if m.name not in seen:
            uniq.append(m)
            seen.add(m.name)
    models = uniq
    input_names = list(input_names)
    if any(isinstance(m.func, Expression) for m in models) or isinstance(performance, Expression):
        return _lower_expression(models, performance, input_names)
    if all(isinstance(m.func, Poly) for m in models) and (performance is None or isinstance(performance, (Poly, str))):
        if prefer_expression and models:
            return _lower_expression(models, performance, input_names)
        stages = [m.func for m in models]
        names = [m.name for m in models]
        if isinstance(performance, Poly):
            stages.append(performance)
            names.append("__performance")
        elif isinstance(performance, str) and (not models or performance != models[-1].name):
            stages.append(Poly([(1.0, [performance])]))
            names.append("__performance")
        if not stages:
            raise UnsupportedModel("empty model chain")
        if len(input_names) + len(stages) > L.EBN_MAX_INPUTS + 1:
            raise UnsupportedModel("too many columns for the polynomial functor")
        prog = _poly_program(stages, input_names, names)
        return Lowered(L.LS_POLYNOMIAL, prog, input_names, [], input_names + [m.name for m in models])
    cats = [m for m in models if isinstance(m.func, Catalogue)]
    if len(cats) != len(models):
        bad = [m.name for m in models if not isinstance(m.func, (Poly, Catalogue))]
        raise UnsupportedModel(f"models {bad or [m.name for m in models]} have no device functor: pass the closure's "
                               "formula as Expression(text), or use Poly / Catalogue (a Python closure cannot run "
                               "on the GPU and there is no CPU fallback)")
    last = cats[-1]
    if isinstance(performance, str) and performance != last.name:
        raise UnsupportedModel("a Catalogue chain needs `performance = df -> df.<last model>`")
    if isinstance(performance, Poly):
        raise UnsupportedModel("a Catalogue chain cannot be followed by a polynomial performance")
    ls = last.func.limit_state
    if ls == "straub_frame":
        # demo/straub_example.jl after transmission: the capacity models r1..r5 (each with its own
        # in-model randn → one hidden standard-normal input) followed by the frame model. With R4, R5
        # discretised (test/inference/variableselimination.jl:207-217) only r1..r3 are fused and the
        # frame reads the R4, R5 columns.
        caps = cats[:-1]
        if any(c.func.limit_state != "straub_capacity" for c in caps) or len(last.func.args) != 7:
            raise UnsupportedModel("straub_frame expects straub_capacity models before it and 7 arguments")
        consts = np.array(caps[0].func.consts) if caps else np.array(last.func.consts)
        made = [c.name for c in caps]
        r_args, (v_arg, h_arg) = list(last.func.args[:5]), last.func.args[5:]
        ur = caps[0].func.args[0]
        if all(a in made for a in r_args):
            ls, cols = "straub_frame", [ur, v_arg, h_arg]
        elif all(a in made for a in r_args[:3]) and not any(a in made for a in r_args[3:]):
            ls, cols = "straub_frame_r45", [ur, v_arg, h_arg, r_args[3], r_args[4]]
        else:
            raise UnsupportedModel("straub_frame supports r1..r5 (or r1..r3) produced by straub_capacity models")
        hidden = [Normal(0.0, 1.0).abi()] * sum(a in made for a in r_args)
    elif ls == "straub_capacity":
        if len(cats) != 1:
            raise UnsupportedModel("straub_capacity evaluates one capacity; chains end in straub_frame")
        consts = np.array(last.func.consts)
        cols = [last.func.args[0]]
        hidden = [Normal(0.0, 1.0).abi()]
    else:
        consts = np.array(last.func.consts)
        if len(cats) != 1:
            raise UnsupportedModel(f"limit state '{ls}' is a single-model functor")
        cols = list(last.func.args) or list(input_names)
        hidden = []
    missing = [c for c in cols if c not in input_names]
    if missing:
        raise UnsupportedModel(f"limit state '{ls}' needs inputs {missing} which the node's parents do not provide")
    return Lowered(L.limit_state_id(ls), consts, cols, hidden, cols + [m.name for m in models])


class _TablePool:
    def __init__(self):
        self.chunks, self.size = [], 0

    def add(self, q: np.ndarray) -> Tuple[int, int]:
        off = self.size
        self.chunks.append(np.asarray(q, dtype=float))
        self.size += len(q)
        return off, len(q)

    def array(self):
        return np.concatenate(self.chunks) if self.chunks else None


def _abi_row(inp, pool: _TablePool) -> tuple:
    if isinstance(inp, Parameter):
        return (L.IN_PARAM, float(inp.value), 0.0, 0.0, 0.0, -INF, INF)
    if isinstance(inp, RandomVariable):
        try:
            return inp.dist.abi()
        except NeedsTable as nt:
            off, n = pool.add(table_for(nt.dist, nt.lo, nt.hi))
            return (L.IN_TABULATED, float(off), float(n), 0.0, 0.0, -INF, INF)
    raise TypeError(f"input {inp} is imprecise: use DoubleLoop or RandomSlicing")


def _flatten(inputs):
    """Expands JointDistribution inputs; returns (flat inputs, {name: (group id, position)}, copulas)."""
    flat, where, copulas = [], {}, []
    for i in inputs:
        if isinstance(i, JointDistribution):
            g = len(copulas)
            copulas.append(np.asarray(i.copula.correlation))
            for k, m in enumerate(i.marginals):
                flat.append(m)
                where[m.name] = (g, k)
        else:
            flat.append(i)
    return flat, where, copulas


def input_names(inputs) -> List[str]:
    return [i.name for i in _flatten(inputs)[0]]


def build_job(low: Lowered, scenario_inputs: Sequence[Sequence], sim: CudaMonteCarlo, stream_id=None) -> L.Job:
    pool = _TablePool()
    rows, chols = [], []
    d = len(low.columns) + len(low.hidden)
    any_copula = False
    for inputs in scenario_inputs:
        flat, where, copulas = _flatten(inputs)
        by_name = {}
        for i in flat:
            if i.name in by_name:
                raise ValueError(f"two inputs named '{i.name}' in one scenario")
            by_name[i.name] = i
        row = []
        for c in low.columns:
            inp = by_name[c]
            if (copulas or getattr(sim, "qmc", False)) and isinstance(inp, RandomVariable) and isinstance(inp.dist, Gamma):
                # no closed-form quantile on the device: Gamma under a copula or a Sobol sequence is shipped tabulated
                off, n = pool.add(table_for(inp.dist, -INF, INF, 8193))
                row.append((L.IN_TABULATED, float(off), float(n), 0.0, 0.0, -INF, INF))
            else:
                row.append(_abi_row(inp, pool))
        rows.append(row + list(low.hidden))
        if copulas:
            any_copula = True
            R = np.eye(d)
            for a, ca in enumerate(low.columns):
                for b, cb in enumerate(low.columns):
                    if ca in where and cb in where and where[ca][0] == where[cb][0]:
                        R[a, b] = copulas[where[ca][0]][where[ca][1], where[cb][1]]
            chols.append(np.linalg.cholesky(R))
        else:
            chols.append(np.eye(d))
    return L.Job(low.limit_state, low.consts, L.make_inputs(rows), int(sim.n), int(sim.seed), bool(sim.strict),
                 stream_id=stream_id, tables=pool.array(), corr_chol=np.stack(chols) if any_copula else None,
                 fp32_sampling=bool(getattr(sim, "fp32_sampling", False)), qmc=bool(getattr(sim, "qmc", False)))


class Samples:
    """Stands where the reference stores the n-row DataFrame of a scenario
    (`additional_info[...][:samples]`, evaluate_node.jl:93-95). At 1e8+ samples per scenario the
    frame cannot be materialised; it is replayed from the counter-based streams on demand."""

    def __init__(self, job: L.Job, scenario: int, columns: List[str]):
        self.job, self.scenario, self.columns = job, scenario, columns
        self.shape = (job.n, len(columns))

    def fetch(self, first: int = 0, n: int = 1000) -> Dict[str, np.ndarray]:
        n = min(n, self.job.n - first)
        ncols = (min(len(self.columns), L.EBN_MAX_INPUTS)
                 if self.job.limit_state in (L.LS_POLYNOMIAL, L.LS_EXPRESSION) else self.job.d)
        X, g = L.sample_matrix(self.job, self.scenario, self.job.sample_offset + first, n, want_g=True, ncols=ncols)
        out = {c: X[:, j] for j, c in enumerate(self.columns[:X.shape[1]])}
        out["g"] = g
        return out

    def __len__(self):
        return self.job.n


def _is_imprecise(i) -> bool:
    return isinstance(i, (Interval, ProbabilityBox))


def pf_scenarios(models, performance, scenario_inputs, sim: CudaMonteCarlo):
    """All scenarios of a node in one ebn_pf_batch call → (pf[S], std[S], [Samples])."""
    big = len(scenario_inputs) * float(sim.n) >= JIT_POLY_MIN_SAMPLES and L.jit_available()
    low = lower(models, performance, input_names(scenario_inputs[0]), prefer_expression=big)
    job = build_job(low, scenario_inputs, sim)
    cnt, pf, var = L.pf_batch(job)
    return pf, np.sqrt(var), [Samples(job, s, low.out_columns) for s in range(job.S)], cnt


def probability_of_failure(models, performance, inputs, sim):
    """UQ.jl signature, one scenario: (pf, std, samples) for Monte Carlo; `Interval`-like (lb, ub)
    for DoubleLoop; ((lb, ub), info_lb, info_ub) for RandomSlicing."""
    models = [models] if isinstance(models, Model) else list(models)
    if isinstance(sim, CudaMonteCarlo):
        if any(_is_imprecise(i) for i in inputs):
            raise AssertionError("imprecise inputs need DoubleLoop or RandomSlicing")
        pf, sd, smp, _ = pf_scenarios(models, performance, [inputs], sim)
        return float(pf[0]), float(sd[0]), smp[0]
    if isinstance(sim, FORM):
        return form(models, performance, inputs, sim)
    if isinstance(sim, CudaDoubleLoop):
        return double_loop(models, performance, [inputs], sim)[0]
    if isinstance(sim, CudaRandomSlicing):
        return random_slicing(models, performance, [inputs], sim)[0]
    raise TypeError(f"simulation {sim!r} has no device path (MonteCarlo, DoubleLoop and RandomSlicing do)")


# ---- FORM ------------------------------------------------------------------------------------
def form(models, performance, inputs, sim: FORM):
    """UQ.jl `probability_of_failure(models, performance, inputs, ::FORM)` → (pf, β, design point, α).
    HL-RF: y ← (∇G·y − G(y)) ∇G / |∇G|² with G(y) = g(x(y)), x_j = Q_j(Φ(y_j)); central differences.
    All evaluations of g run on the device (ebn_eval_matrix, strict); the iteration itself is host
    logic on d numbers."""
    from scipy.special import ndtr
    flat, _, copulas = _flatten(inputs)
    if copulas:
        raise UnsupportedModel("FORM under a copula is not implemented")
    if any(_is_imprecise(i) for i in flat):
        raise AssertionError("imprecise inputs need DoubleLoop or RandomSlicing")
    low = lower(models, performance, [i.name for i in flat])
    by_name = {i.name: i for i in flat}
    cols = [by_name[c] for c in low.columns]
    rv = [j for j, c in enumerate(cols) if isinstance(c, RandomVariable)]
    n_hidden = len(low.hidden)
    d = len(rv) + n_hidden

    def to_x(Y):
        X = np.empty((Y.shape[0], len(cols) + n_hidden))
        for j, c in enumerate(cols):
            if isinstance(c, Parameter):
                X[:, j] = c.value
        for k, j in enumerate(rv):
            u = np.clip(ndtr(Y[:, k]), 1e-300, 1.0 - 1e-16)
            X[:, j] = [cols[j].dist.quantile(float(v)) for v in u]
        X[:, len(cols):] = Y[:, len(rv):]          # in-model standard normals
        return X

    def G(Y):
        return L.eval_matrix(low.limit_state, low.consts, to_x(Y), strict=True, want_g=True)[1]

    y = np.zeros(d)
    alpha = np.zeros(d)
    beta = 0.0
    for _ in range(max(1, sim.n)):
        pts = np.vstack([y] + [y + sgn * sim.h * np.eye(d)[k] for k in range(d) for sgn in (1.0, -1.0)])
        g = G(pts)
        grad = (g[1::2] - g[2::2]) / (2.0 * sim.h)
        nrm2 = float(grad @ grad)
        if not (nrm2 > 0.0) or not np.all(np.isfinite(g)):
            raise ValueError("FORM: the limit state has no usable gradient at the current point")
        y_new = (float(grad @ y) - float(g[0])) * grad / nrm2
        alpha = -grad / math.sqrt(nrm2)
        beta_new = float(alpha @ y_new)
        done = abs(beta_new - beta) < sim.tol and np.linalg.norm(y_new - y) < sim.tol * max(1.0, np.linalg.norm(y_new))
        y, beta = y_new, beta_new
        if done:
            break
    pf = float(ndtr(-beta))
    dp = dict(zip([low.columns[j] for j in rv] + [f"__randn{i + 1}" for i in range(n_hidden)], to_x(y[None, :])[0][rv + list(range(len(cols), len(cols) + n_hidden))]))
    return pf, beta, dp, dict(zip(dp.keys(), alpha))


# ---- imprecise inputs ------------------------------------------------------------------------
def _imprecise_dims(inputs) -> List[Tuple[int, int, float, float]]:
    """(input position, parameter position or -1, lb, ub) per free coordinate of the outer loop."""
    dims = []
    for pos, i in enumerate(inputs):
        if isinstance(i, Interval):
            dims.append((pos, -1, float(i.lb), float(i.ub)))
        elif isinstance(i, ProbabilityBox):
            for k, p in enumerate(i.parameters):
                dims.append((pos, k, float(p.lb), float(p.ub)))
    return dims


def _realise(inputs, dims, x) -> list:
    """`map_to_precise`: an Interval becomes Parameter(x), a p-box RandomVariable(T(x...))."""
    out = list(inputs)
    pbox_vals: Dict[int, Dict[int, float]] = {}
    for (pos, k, _, _), v in zip(dims, x):
        if k < 0:
            out[pos] = Parameter(float(v), inputs[pos].name)
        else:
            pbox_vals.setdefault(pos, {})[k] = float(v)
    for pos, vals in pbox_vals.items():
        pb = inputs[pos]
        out[pos] = RandomVariable(pb.member([vals[k] for k in range(len(pb.parameters))]), pb.name)
    return out


def _eval_candidates(models, performance, scenario_inputs, dims_per_s, cands_per_s, sim: CudaMonteCarlo):
    """pf of every (scenario, candidate) in ONE batch; candidates of a scenario share its Philox
    stream (common random numbers), so pf is a deterministic function of the candidate."""
    flat, stream, owner = [], [], []
    for s, (inputs, dims, cands) in enumerate(zip(scenario_inputs, dims_per_s, cands_per_s)):
        for x in cands:
            flat.append(_realise(inputs, dims, x))
            stream.append(s)
            owner.append(s)
    low = lower(models, performance, input_names(flat[0]))
    job = build_job(low, flat, sim, stream_id=np.array(stream, dtype=np.uint32))
    _, pf, _ = L.pf_batch(job)
    out, k = [], 0
    for cands in cands_per_s:
        out.append(pf[k:k + len(cands)])
        k += len(cands)
    return out


def double_loop(models, performance, scenario_inputs, sim: CudaDoubleLoop) -> List[Tuple[float, float]]:
    """UQ.jl DoubleLoop (A5): per scenario (pf_lb, pf_ub) over the imprecise box, all scenarios and
    all outer candidates batched on the device."""
    inner = sim.lb
    dims_per_s = [_imprecise_dims(inp) for inp in scenario_inputs]
    if any(not d for d in dims_per_s):
        raise AssertionError("DoubleLoop needs at least one Interval / ProbabilityBox input")
    S = len(scenario_inputs)
    los = [np.array([d[2] for d in dims]) for dims in dims_per_s]
    his = [np.array([d[3] for d in dims]) for dims in dims_per_s]
    if sim.method == "grid":
        cands = []
        for lo, hi in zip(los, his):
            axes = [np.linspace(a, b, sim.grid) for a, b in zip(lo, hi)]
            cands.append([np.array(p) for p in itertools.product(*axes)])
        pf = _eval_candidates(models, performance, scenario_inputs, dims_per_s, cands, inner)
        return [(float(p.min()), float(p.max())) for p in pf]
    # "optimise": compass search from the box midpoint (where the reference starts BOBYQA), both
    # bounds and all scenarios advancing in lock-step; the vertices are always probed first.
    best_lo = [None] * S
    best_hi = [None] * S
    x_lo = [0.5 * (a + b) for a, b in zip(los, his)]
    x_hi = [0.5 * (a + b) for a, b in zip(los, his)]
    step = [0.25 * (b - a) for a, b in zip(los, his)]
    first = True
    for _ in range(14):
        cands = []
        for s in range(S):
            c = [x_lo[s], x_hi[s]]
            for x0 in (x_lo[s], x_hi[s]):
                for k in range(len(x0)):
                    for sg in (-1.0, 1.0):
                        y = x0.copy()
                        y[k] = min(max(y[k] + sg * step[s][k], los[s][k]), his[s][k])
                        c.append(y)
            if first:
                c += [np.array(v) for v in itertools.product(*zip(los[s], his[s]))]
            cands.append(c)
        pf = _eval_candidates(models, performance, scenario_inputs, dims_per_s, cands, inner)
        moved = False
        for s in range(S):
            i_min, i_max = int(np.argmin(pf[s])), int(np.argmax(pf[s]))
            if best_lo[s] is None or pf[s][i_min] < best_lo[s]:
                moved |= best_lo[s] is not None
                best_lo[s], x_lo[s] = float(pf[s][i_min]), cands[s][i_min].copy()
            if best_hi[s] is None or pf[s][i_max] > best_hi[s]:
                moved |= best_hi[s] is not None
                best_hi[s], x_hi[s] = float(pf[s][i_max]), cands[s][i_max].copy()
        if not moved and not first:
            step = [0.5 * st for st in step]
            if all(np.all(st <= 1e-3 * (b - a) + 1e-300) for st, a, b in zip(step, los, his)):
                break
        first = False
    return list(zip(best_lo, best_hi))


def _poly_monotone(models, performance, name: str) -> bool:
    """True if every polynomial stage is affine in `name` with sign-definite coefficient, i.e. the
    limit state is monotone in that column (sufficient condition, checked syntactically)."""
    derived = {name}
    for st in [m.func for m in models] + ([performance] if isinstance(performance, Poly) else []):
        if not isinstance(st, Poly):
            return False
        for _, cols in st.terms:
            hits = [c for c in cols if c in derived]
            if len(hits) > 1 or (hits and len(cols) > 1):
                return False
    for m in models:
        if any(c in derived for _, cols in m.func.terms for c in cols):
            derived.add(m.name)
    return True


def _monotone_on_replay(models, performance, inputs, dims, k: int, seed: int, rows: int = 2048, grid: int = 9) -> bool:
    """Numerical check for limit states that are not polynomial programs: is g monotone, in ONE
    direction for all samples, in imprecise dimension k of this scenario? The other imprecise
    dimensions sit at their midpoints; the scenario is replayed on the device at `grid` values of
    dimension k with common random numbers (same stream, same seed), `rows` samples each.
    A sufficient condition checked on a sample, not a proof: stated in `random_slicing`."""
    mid = np.array([0.5 * (d[2] + d[3]) for d in dims])
    G = []
    for v in np.linspace(dims[k][2], dims[k][3], grid):
        x = mid.copy()
        x[k] = v
        real = _realise(inputs, dims, x)
        low = lower(models, performance, input_names(real))
        job = build_job(low, [real], CudaMonteCarlo(rows, seed, True))
        G.append(L.sample_matrix(job, 0, 0, rows, want_g=True)[1])
    D = np.diff(np.array(G), axis=0)
    D = D[:, np.all(np.isfinite(D), axis=0)]
    return bool(np.all(D >= 0.0) or np.all(D <= 0.0))


# Quantile of a p-box family as formula text over the family's parameters p0, p1 and ONE standard
# variate s per sample (N(0,1) for the normal family, U(0,1) otherwise): every member of the family
# sees the same s, which is what slices the p-box into an interval per sample.
_PBOX_QUANTILE = {
    "Normal": ("normal", "({p0} + {p1} * {s})"),
    "LogNormal": ("normal", "exp({p0} + {p1} * {s})"),
    "Uniform": ("uniform", "({p0} + ({p1} - {p0}) * {s})"),
    "Gumbel": ("uniform", "({p0} - {p1} * log(-log({s})))"),
}
SLICING_LEVELS = (0.0, 0.5, 1.0)   # where each per-sample interval is probed: both ends and the middle


def _slicing_texts(models, performance):
    """(model name, text) chain + performance text of a chain that can be written as formulas."""
    chain = []
    consts: Dict[str, float] = {}
    for m in models:
        f = m.func
        if isinstance(f, Expression):
            consts.update(dict(f.constants))
            chain.append((m.name, f.text))
        elif isinstance(f, Poly):
            chain.append((m.name, _poly_text(f)))
        else:
            return None
    if isinstance(performance, Expression):
        consts.update(dict(performance.constants))
        perf = performance.text
    elif isinstance(performance, Poly):
        perf = _poly_text(performance)
    elif isinstance(performance, str):
        perf = f"df.{performance}"
    elif performance is None and chain:
        perf = f"df.{chain[-1][0]}"
    else:
        return None
    return chain, perf, consts


def random_slicing_general(models, performance, scenario_inputs, sim: CudaRandomSlicing, levels=SLICING_LEVELS):
    """UQ.jl RandomSlicing (A6) for model chains that are formulas, without a monotonicity assumption.

    Per sample an Interval input is its box [lb, ub]; a p-box input is sliced by ONE standard variate
    into [min, max] of the family's quantile over the vertices of its parameter box (any number of
    interval parameters). The model chain is then evaluated, inside the kernel, at every combination
    of `levels` (ends and middle) of the per-sample intervals, and the failure indicator of the
    sample's smallest / largest g is counted: pf_ub = P(min g < 0), pf_lb = P(max g < 0). Both sides
    are scenarios of ONE ebn_pf_batch call on the same Philox streams.
    The whole construction is a rewrite of the formulas (the candidates become one expression under
    min / max), so it runs through EBN_LS_EXPRESSION with no extra kernel.
    Exact when the per-sample extreme lies at a probed point (always for chains that are monotone or
    convex/concave with the extreme at an end or the middle); otherwise an inner approximation that
    tightens with more levels."""
    import ast
    import copy
    texts = _slicing_texts(list(models), performance)
    if texts is None:
        raise UnsupportedModel("general RandomSlicing needs Expression / Poly models")
    chain, perf_text, consts = texts
    base = list(scenario_inputs[0])
    norm = _expr._norm
    # ---- columns of the rewritten job ---------------------------------------------------------
    columns: List[str] = []             # input names of the rewritten job, in order
    makers = []                         # per column: function(scenario inputs) -> UQ input
    interval_of: Dict[str, Tuple[str, str]] = {}   # imprecise input -> (text of per-sample lower end, upper end)
    for pos, inp in enumerate(base):
        if isinstance(inp, Interval):
            lo, hi = f"{inp.name}__lo", f"{inp.name}__hi"
            columns += [lo, hi]
            makers += [lambda ins, pos=pos, lo=lo: Parameter(float(ins[pos].lb), lo),
                       lambda ins, pos=pos, hi=hi: Parameter(float(ins[pos].ub), hi)]
            interval_of[norm(inp.name)] = (norm(lo), norm(hi))
        elif isinstance(inp, ProbabilityBox):
            fam = inp.family.__name__
            if fam not in _PBOX_QUANTILE or len(inp.parameters) != 2:
                raise UnsupportedModel(f"RandomSlicing on the device has no quantile formula for a {fam} p-box")
            if inp.lb > -INF or inp.ub < INF:
                raise UnsupportedModel("RandomSlicing on the device takes untruncated p-boxes")
            kind, qtext = _PBOX_QUANTILE[fam]
            sname = f"{inp.name}__s"
            columns.append(sname)
            makers.append(lambda ins, sname=sname, kind=kind:
                          RandomVariable(Normal(0.0, 1.0), sname) if kind == "normal"
                          else RandomVariable(_uniform01(), sname))
            pnames = []
            for k in range(2):
                for side in ("lo", "hi"):
                    nm = f"{inp.name}__p{k}{side}"
                    columns.append(nm)
                    makers.append(lambda ins, pos=pos, k=k, side=side, nm=nm:
                                  Parameter(float(getattr(ins[pos].parameters[k], "lb" if side == "lo" else "ub")), nm))
                pnames.append((norm(f"{inp.name}__p{k}lo"), norm(f"{inp.name}__p{k}hi")))
            verts = [qtext.format(p0=a, p1=b, s=norm(sname)) for a in pnames[0] for b in pnames[1]]
            interval_of[norm(inp.name)] = ("min(" + ", ".join(verts) + ")", "max(" + ", ".join(verts) + ")")
        else:
            columns.append(inp.name)
            makers.append(lambda ins, pos=pos: ins[pos])
    if not interval_of:
        raise AssertionError("RandomSlicing needs at least one Interval / ProbabilityBox input")
    columns.append("__side")
    if len(columns) > L.EBN_MAX_INPUTS:
        raise UnsupportedModel("too many columns after slicing the imprecise inputs")

    # ---- the chain as ONE expression per candidate: model names and imprecise inputs substituted ----
    class Subst(ast.NodeTransformer):
        def __init__(self, table):
            self.table = table

        def visit_Name(self, node):
            return copy.deepcopy(self.table[node.id]) if node.id in self.table else node

    def tree(text):
        t = _expr._parse(text)
        if _expr._count_hidden(t):
            raise UnsupportedModel("RandomSlicing on the device: in-model randomness cannot be sliced")
        return t

    model_trees = [(norm(n), tree(t)) for n, t in chain]
    perf_tree = tree(perf_text)
    ends = {name: (tree(lo), tree(hi)) for name, (lo, hi) in interval_of.items()}
    names = list(ends)
    cands = []
    for combo in itertools.product(levels, repeat=len(names)):
        table = {}
        for name, lam in zip(names, combo):
            lo, hi = ends[name]
            if lam == 0.0:
                table[name] = lo
            elif lam == 1.0:
                table[name] = hi
            else:       # lo + lam * (hi - lo)
                table[name] = ast.BinOp(copy.deepcopy(lo), ast.Add(),
                                        ast.BinOp(ast.Constant(float(lam)), ast.Mult(),
                                                  ast.BinOp(copy.deepcopy(hi), ast.Sub(), copy.deepcopy(lo))))
        for mname, mt in model_trees:      # a later model sees the earlier ones already substituted
            table[mname] = Subst(table).visit(copy.deepcopy(mt))
        cands.append(ast.unparse(Subst(table).visit(copy.deepcopy(perf_tree))))
    gmin = "min(" + ", ".join(cands) + ")" if len(cands) > 1 else cands[0]
    gmax = "max(" + ", ".join(cands) + ")" if len(cands) > 1 else cands[0]
    final = f"_ifelse(__side < 0.5, {gmax}, {gmin})"   # side 0: lower bound of pf, side 1: upper bound
    try:
        prog = _expr.compile_program(columns, [], final, consts)
        L.expression_check(prog.consts, prog.d)
    except (_expr.ExpressionError, L.EbnError) as e:
        raise UnsupportedModel(f"RandomSlicing on the device: {e}") from None
    low = Lowered(L.LS_EXPRESSION, prog.consts, columns, [], prog.columns)
    # ---- one batch: every scenario twice (side 0 / 1) on the scenario's stream -------------------
    rows, stream = [], []
    for s, ins in enumerate(scenario_inputs):
        if [type(i) for i in ins] != [type(i) for i in base]:
            raise UnsupportedModel("RandomSlicing on the device: scenarios differ in the kind of their inputs")
        real = [mk(list(ins)) for mk in makers]
        for side in (0.0, 1.0):
            rows.append(real + [Parameter(side, "__side")])
            stream.append(s)
    inner = sim.lb
    job = build_job(low, rows, inner, stream_id=np.array(stream, dtype=np.uint32))
    _, pf, _ = L.pf_batch(job)
    out, n = [], inner.n
    for s in range(len(scenario_inputs)):
        lb, ub = float(pf[2 * s]), float(pf[2 * s + 1])
        out.append(((lb, ub), (math.sqrt(max(lb - lb * lb, 0) / n), Samples(job, 2 * s, low.out_columns)),
                    (math.sqrt(max(ub - ub * ub, 0) / n), Samples(job, 2 * s + 1, low.out_columns))))
    return out


def _uniform01():
    from .distributions import Uniform
    return Uniform(0.0, 1.0)


def random_slicing(models, performance, scenario_inputs, sim: CudaRandomSlicing):
    """UQ.jl RandomSlicing (A6) for limit states monotone in each imprecise input: the per-sample
    extremes over the input boxes are then attained at the box vertices, so
    pf_lb / pf_ub are the extreme vertex probabilities (evaluated with common random numbers).
    Returns [((lb, ub), (std_lb, Samples), (std_ub, Samples))] per scenario.
    Chains that are formulas (Expression / Poly) take `random_slicing_general` instead: per-sample
    extremes over the sliced boxes inside the kernel, no monotonicity needed, p-boxes with two interval
    parameters. What is left here are the catalogue functors (compiled-in limit states)."""
    if _slicing_texts(list(models), performance) is not None:
        return random_slicing_general(models, performance, scenario_inputs, sim)
    inner = sim.lb
    dims_per_s = [_imprecise_dims(inp) for inp in scenario_inputs]
    for inputs, dims in zip(scenario_inputs, dims_per_s):
        for i in inputs:
            if isinstance(i, ProbabilityBox) and len(i.parameters) > 1:
                raise UnsupportedModel("RandomSlicing on the device supports p-boxes with one interval parameter")
        for k, d in enumerate(dims):
            name = inputs[d[0]].name
            # polynomial programs are checked syntactically; anything else (formulas, catalogue
            # functors) numerically on replayed samples (2048 rows x 9 values per dimension)
            if not (_poly_monotone(models, performance, name)
                    or _monotone_on_replay(models, performance, inputs, dims, k, inner.seed)):
                raise UnsupportedModel(f"RandomSlicing on the device needs a limit state monotone in '{name}'")
    cands = [[np.array(v) for v in itertools.product(*[(d[2], d[3]) for d in dims])] for dims in dims_per_s]
    pf = _eval_candidates(models, performance, scenario_inputs, dims_per_s, cands, inner)
    out = []
    n = inner.n
    for p in pf:
        lb, ub = float(p.min()), float(p.max())
        out.append(((lb, ub), (math.sqrt(max(lb - lb * lb, 0) / n), None), (math.sqrt(max(ub - ub * ub, 0) / n), None)))
    return out


# ---- ContinuousFunctionalNode ------------------------------------------------------------------
def output_histograms(models, scenario_inputs, sim: CudaMonteCarlo, edges: Optional[np.ndarray] = None):
    """`sample` + `evaluate!` for a ContinuousFunctionalNode: per-scenario histogram and moments of
    the last model's output. Without `edges`, a pilot call finds the output range."""
    big = len(scenario_inputs) * float(sim.n) >= JIT_POLY_MIN_SAMPLES and L.jit_available()
    low = lower(models, None, input_names(scenario_inputs[0]), prefer_expression=big)
    job = build_job(low, scenario_inputs, sim)
    if edges is None:
        # every field of the job carries over to the pilot (copula factors, stream ids, fp32 sampling, tables)
        pilot = dataclasses.replace(job, n=min(job.n, 1 << 16))
        _, s1, s2 = L.hist_batch(pilot, np.array([0.0]))
        m = s1 / pilot.n
        sd = np.sqrt(np.maximum(s2 / pilot.n - m * m, 0.0))
        lo, hi = float(np.min(m - 8 * sd)), float(np.max(m + 8 * sd))
        if not hi > lo:
            lo, hi = lo - 1.0, hi + 1.0
        edges = np.linspace(lo, hi, min(HIST_BINS, L.EBN_MAX_EDGES))
    counts, s1, s2 = L.hist_batch(job, edges)
    return edges, counts, s1, s2, job, low
