"""Random formula chains in the shapes the sign-only predicate generator rewrites (ebn_expr.cu: FailsGen), shared
by the GPU test (fused count against the oracle by bisection) and the CPU test (generated functor compiled with
g++ behind a shim)."""


def rational(rng, atoms, depth):
    if depth == 0 or rng.random() < 0.25:
        return str(rng.choice(atoms)) if rng.random() < 0.8 else repr(round(float(rng.normal(0, 2)), 2))
    a, b = rational(rng, atoms, depth - 1), rational(rng, atoms, depth - 1)
    return f"({a} {rng.choice(['+', '-', '*', '/', '/'])} {b})"


def mechanism(rng, atoms):
    """Scaled square roots against constants on either side, differences of fractions, squares of fractions,
    and things the generator must leave alone."""
    R = lambda: rational(rng, atoms, int(rng.integers(1, 4)))  # noqa: E731
    c = repr(round(float(rng.uniform(0.2, 3.0)), 2))
    k = str(rng.choice(["k0", "k1", repr(round(float(rng.normal(0.5, 1.0)), 2))]))
    return str(rng.choice([
        f"{c} * sqrt({R()}) - {k}", f"sqrt({R()}) * {c} - {k}", f"{k} - {c} * sqrt({R()})", f"sqrt({R()}) - {k}",
        f"{k} - {R()}", f"{R()} - {k}", f"1 - ({R()} / {R()}) * {R()}", f"({R()})^2 - {k}", f"-({R()})",
        f"sqrt({R()}) - sqrt(abs({R()}))", f"log({R()}) - {k}", f"abs({R()}) - {c}", f"{R()} + {k}",
        f"min({R()}, {R()}) - {k}", f"exp(min(max({R()}, -40), 40) / 8) - {c}"]))


def unstable_rows(eval_g, X, g):
    """Rows on which the SIGN of g is not a property of the formula but of its rounding: g changes sign (or turns
    NaN, or stops being NaN) when every input moves by a relative 1e-12 — catastrophic cancellation between huge
    intermediates, a square-root argument that is a rounding error away from zero. A predicate evaluated in
    another order of operations may legitimately disagree there, exactly as it may on a tie.
    eval_g(X) -> g is the reference evaluator."""
    import numpy as np
    bad = np.zeros(len(X), dtype=bool)
    for eps in (1e-12, -1e-12):
        gp = eval_g(X * (1.0 + eps))
        bad |= (np.isnan(gp) != np.isnan(g)) | ((gp < 0) != (g < 0))
    return bad


def chain(rng, names):
    """(models, performance, constants) of one random chain over the input columns `names`."""
    consts = {"k0": float(rng.normal(0.5, 1.0)), "k1": float(rng.uniform(0.1, 2.0))}
    models, avail = [], list(names)
    for t in range(int(rng.integers(1, 4))):
        models.append((f"m{t}", mechanism(rng, avail)))
        avail.append(f"m{t}")
    stored = [f"m{t}" for t in range(len(models))]
    extra = [mechanism(rng, avail) for _ in range(int(rng.integers(0, 3)))]
    shape = int(rng.integers(0, 4))
    terms = stored + extra
    if shape == 0 or len(terms) == 1:
        perf = "minimum([" + ", ".join(terms) + "])"
    elif shape == 1:
        perf = "max(" + ", ".join(terms[:2]) + ")"
    elif shape == 2:
        perf = "min(" + terms[0] + ", max(" + ", ".join(terms[1:3] if len(terms) > 2 else terms[:2]) + "))"
    else:
        perf = terms[-1]
    return models, perf, consts
