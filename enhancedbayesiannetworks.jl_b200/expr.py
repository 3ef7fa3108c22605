"""Expression limit states: user-written models as text → the stack program of EBN_LS_EXPRESSION.

The reference's functional nodes carry arbitrary Julia closures, ``Model(df -> ..., :name)`` and a
``performance`` function (src/nodes/functionalnodes.jl:27-52). A closure cannot cross the C ABI, so
this path takes the *expression* each closure evaluates — Julia or Python syntax, broadcasting dots
and ``df.`` prefixes allowed::

    Expression("g1", "1 .- (pi*m .* df.V .* df.A) ./ (df.b0 .* df.K .* g^2) .* (...)")

and compiles it, in Julia's evaluation order, into the program include/ebncuda.h documents
("expression program"). The library turns the program into a CUDA functor and compiles it for the
GPU at run time (NVRTC); there is no interpreter and no CPU fallback.

Julia semantics kept: ``a*b*c`` and ``a+b+c`` fold left to right; ``x^2`` is ``x*x`` and ``x^3`` is
``(x*x)*x`` (literal_pow); ``min``/``max``/``minimum([...])``/``maximum([...])`` propagate NaN;
integer literals take part as Float64 (exact). ``randn()`` and ``rand(Normal(mu, sigma))`` inside a
model (demo/straub_example.jl:26) become hidden standard-normal input columns, one per textual
occurrence, drawn by the device sampler: ``rand(Normal(mu, sigma))`` is ``mu + sigma * z`` as in
Distributions.jl. Not supported: anything that is not a scalar formula of the row (loops, array
indexing other than ``x[1]`` on a scalar, other in-model distributions).
"""
from __future__ import annotations

import ast
import math
import re
import unicodedata
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import _lib as L


class ExpressionError(ValueError):
    pass


_UNARY = {"sqrt": L.OP_SQRT, "exp": L.OP_EXP, "log": L.OP_LOG, "abs": L.OP_ABS, "sin": L.OP_SIN, "cos": L.OP_COS}
_NAMED_CONSTS = {"pi": math.pi, "π": math.pi, "_euler": math.e, "Inf": math.inf}


def _to_python(src: str) -> str:
    """Julia surface syntax → something Python's parser accepts (same tree shape)."""
    # Julia identifiers may carry subscripts Python's parser rejects (b₀) or silently normalises
    # (Uᵣ): NFKC the whole text (b₀ → b0, Cₖ → Ck); column and constant names get the same treatment
    s = unicodedata.normalize("NFKC", src.strip().replace("ℯ", "_euler"))
    s = re.sub(r"\.(\^|\*|/|\+|-|<=|>=|<|>)", r"\1", s)      # broadcasting operators: .* ./ .^ .+ .-
    s = re.sub(r"([^\W\d]|[)\]])\.\(", r"\1(", s)               # broadcast calls: sqrt.(x)
    s = s.replace("^", "**")
    s = re.sub(r"\bdf\.([^\W\d]\w*)", r"\1", s)              # df.x → x
    s = re.sub(r"\bdf\[\s*[:!]\s*,\s*:([^\W\d]\w*)\s*\]", r"\1", s)  # df[:, :x] / df[!, :x]
    s = s.replace("π", "pi")
    s = re.sub(r"\bifelse\(", "_ifelse(", s)
    return s


@dataclass
class _Emitter:
    columns: Dict[str, int]            # name → column index (inputs, then stored models)
    params: Dict[str, int]             # name → runtime-constant index
    hidden_base: int = 0               # first hidden standard-normal column
    n_hidden: int = 0                  # hidden columns handed out so far
    code: List[Tuple[int, float]] = field(default_factory=list)

    def _hidden(self):
        self.emit(L.OP_INPUT, self.hidden_base + self.n_hidden)
        self.n_hidden += 1

    def emit(self, op: int, arg: float = 0.0):
        self.code.append((op, float(arg)))

    # -- expression tree, post-order ------------------------------------------------------------
    def expr(self, n: ast.AST):
        if isinstance(n, ast.Constant) and isinstance(n.value, (int, float)) and not isinstance(n.value, bool):
            self.emit(L.OP_CONST, float(n.value))
        elif isinstance(n, ast.Name):
            if n.id in self.columns:
                self.emit(L.OP_INPUT, self.columns[n.id])
            elif n.id in self.params:
                self.emit(L.OP_PARAM, self.params[n.id])
            elif n.id in _NAMED_CONSTS:
                self.emit(L.OP_CONST, _NAMED_CONSTS[n.id])
            else:
                raise ExpressionError(f"unknown name '{n.id}': not an input column, an earlier model, or a constant")
        elif isinstance(n, ast.UnaryOp) and isinstance(n.op, (ast.USub, ast.UAdd)):
            if isinstance(n.op, ast.UAdd):
                self.expr(n.operand)
            elif isinstance(n.operand, ast.Constant) and isinstance(n.operand.value, (int, float)):
                self.emit(L.OP_CONST, -float(n.operand.value))     # a negative literal
            else:
                self.expr(n.operand)
                self.emit(L.OP_NEG)
        elif isinstance(n, ast.BinOp):
            if isinstance(n.op, ast.Pow):
                k = _literal_int(n.right)
                self.expr(n.left)
                if k is not None:
                    self.emit(L.OP_POWI, k)
                else:
                    self.expr(n.right)
                    self.emit(L.OP_POW)
                return
            op = {ast.Add: L.OP_ADD, ast.Sub: L.OP_SUB, ast.Mult: L.OP_MUL, ast.Div: L.OP_DIV}.get(type(n.op))
            if op is None:
                raise ExpressionError(f"operator {type(n.op).__name__} is not supported")
            self.expr(n.left)
            self.expr(n.right)
            self.emit(op)
        elif isinstance(n, ast.Compare) and len(n.ops) == 1:
            a, b = n.left, n.comparators[0]
            if isinstance(n.ops[0], ast.Lt):
                self.expr(a); self.expr(b)
            elif isinstance(n.ops[0], ast.Gt):
                self.expr(b); self.expr(a)
            else:
                raise ExpressionError("only < and > comparisons are supported")
            self.emit(L.OP_LT)
        elif isinstance(n, ast.IfExp):                              # a if c else b
            self.expr(n.test); self.expr(n.body); self.expr(n.orelse)
            self.emit(L.OP_SELECT)
        elif isinstance(n, ast.Subscript):                          # g1[1] on a scalar row value
            self.expr(n.value)
        elif isinstance(n, ast.Call) and isinstance(n.func, ast.Name):
            f, args = n.func.id, n.args
            if f in _UNARY and len(args) == 1:
                self.expr(args[0])
                self.emit(_UNARY[f])
            elif f in ("min", "max") and len(args) >= 2:
                self._fold(args, L.OP_MIN if f == "min" else L.OP_MAX)
            elif f in ("minimum", "maximum") and len(args) == 1 and isinstance(args[0], (ast.List, ast.Tuple)) \
                    and len(args[0].elts) >= 1:
                self._fold(args[0].elts, L.OP_MIN if f == "minimum" else L.OP_MAX)
            elif f == "_ifelse" and len(args) == 3:
                self.expr(args[0]); self.expr(args[1]); self.expr(args[2])
                self.emit(L.OP_SELECT)
            elif f == "float" and len(args) == 1:
                self.expr(args[0])
            elif f == "randn" and not args:
                self._hidden()
            elif f == "rand" and len(args) == 1 and _is_normal_call(args[0]):
                pa = args[0].args
                if len(pa) == 0:
                    self._hidden()
                elif len(pa) == 2:                                  # mu + sigma * randn()
                    self.expr(pa[0]); self.expr(pa[1]); self._hidden()
                    self.emit(L.OP_MUL); self.emit(L.OP_ADD)
                else:
                    raise ExpressionError("rand(Normal(...)) takes Normal() or Normal(mu, sigma)")
            else:
                raise ExpressionError(f"function '{f}' with {len(args)} argument(s) is not supported")
        else:
            raise ExpressionError(f"unsupported syntax: {ast.dump(n)[:80]}")

    def _fold(self, items: Sequence[ast.AST], op: int):
        self.expr(items[0])
        for it in items[1:]:
            self.expr(it)
            self.emit(op)


def _norm(name: str) -> str:
    return unicodedata.normalize("NFKC", name)


def _is_normal_call(n: ast.AST) -> bool:
    return isinstance(n, ast.Call) and isinstance(n.func, ast.Name) and n.func.id == "Normal"


def _count_hidden(n: ast.AST) -> int:
    c = 0
    for sub in ast.walk(n):
        if isinstance(sub, ast.Call) and isinstance(sub.func, ast.Name):
            if (sub.func.id == "randn" and not sub.args) or \
                    (sub.func.id == "rand" and len(sub.args) == 1 and _is_normal_call(sub.args[0])):
                c += 1
    return c


def _literal_int(n: ast.AST):
    if isinstance(n, ast.Constant) and isinstance(n.value, int) and not isinstance(n.value, bool):
        return n.value
    if isinstance(n, ast.UnaryOp) and isinstance(n.op, ast.USub) and isinstance(n.operand, ast.Constant) \
            and isinstance(n.operand.value, int):
        return -n.operand.value
    return None


def _parse(src: str) -> ast.AST:
    try:
        return ast.parse(_to_python(src), mode="eval").body
    except SyntaxError as e:
        raise ExpressionError(f"cannot parse '{src}': {e.msg}") from None


@dataclass
class Program:
    """A compiled chain of models + performance: what ebn_job_t.consts carries."""
    consts: np.ndarray           # the expression program
    inputs: List[str]            # the caller's column names
    stored: List[str]            # model output names, the last columns
    params: List[str]            # runtime constants, in k[] order
    n_hidden: int = 0            # hidden N(0,1) columns between the inputs and the stored columns

    @property
    def d(self) -> int:
        """n_inputs of the job: the caller's inputs plus the hidden standard normals."""
        return len(self.inputs) + self.n_hidden

    @property
    def ncols(self) -> int:
        return self.d + len(self.stored)

    @property
    def columns(self) -> List[str]:
        return self.inputs + [f"__randn{i + 1}" for i in range(self.n_hidden)] + self.stored

    def with_params(self, values: Dict[str, float]) -> np.ndarray:
        """Same code, new runtime constants (no recompilation on the device)."""
        c = self.consts.copy()
        for i, name in enumerate(self.params):
            if name in values:
                c[2 + i] = float(values[name])
        return c


def compile_program(inputs: Sequence[str], models: Sequence[Tuple[str, str]], performance: str,
                    params: Dict[str, float] | None = None) -> Program:
    """inputs: the job's column names in order; models: (output name, expression) in evaluation order
    (evaluate!(models, df) appends one column per model); performance: expression whose value is g.
    params: named constants kept as *runtime* constants (a new value reuses the compiled kernel)."""
    params = dict(params or {})
    inputs = list(inputs)
    if len(set(inputs)) != len(inputs):
        raise ExpressionError("duplicate input names")
    trees = [(name, _parse(src)) for name, src in models]
    perf = _parse(performance)
    n_hidden = sum(_count_hidden(t) for _, t in trees) + _count_hidden(perf)
    # Python's parser NFKC-normalises identifiers (Uᵣ → Ur, Cₖ → Ck): look names up the same way
    em = _Emitter({_norm(n): i for i, n in enumerate(inputs)}, {_norm(n): i for i, n in enumerate(params)},
                  hidden_base=len(inputs))
    if len(em.columns) != len(inputs) or len(em.params) != len(params) or set(em.columns) & set(em.params):
        raise ExpressionError("input / constant names collide after Unicode normalisation")
    stored = []
    for name, tree in trees:
        em.expr(tree)
        em.emit(L.OP_STORE)
        em.columns[_norm(name)] = len(inputs) + n_hidden + len(stored)     # a later model may shadow an earlier name
        stored.append(name)
    em.expr(perf)
    assert em.n_hidden == n_hidden
    flat = [float(len(params)), float(len(em.code))] + [float(v) for v in params.values()]
    for op, arg in em.code:
        flat += [float(op), arg]
    return Program(np.asarray(flat, dtype=np.float64), inputs, stored, list(params), n_hidden)
