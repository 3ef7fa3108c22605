"""demo/vehicle_suspension.jl of the reference, in the current node API, on the GPU.

The model closure of the demo is passed as text (one Expression per line of `composite_model`); the
library generates a CUDA functor from it and compiles it at run time. 20 scenarios (A x b0 x the five
bins of V) are evaluated by ONE ebn_pf_batch call.

    python examples/vehicle_suspension.py [samples per scenario, default 2e6 like the demo]
"""
import os
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ebn_b200 as E  # noqa: E402


def root(name, states, params):
    cpt = E.DiscreteConditionalProbabilityTable(name)
    for s, p in states.items():
        cpt[(name, s)] = p
    return E.DiscreteNode(name, cpt, params)


def continuous_root(name, dist, discretization=None):
    cpt = E.ContinuousConditionalProbabilityTable()
    cpt[()] = dist
    return E.ContinuousNode(name, cpt, discretization)


def build(n):
    M, m, g = 3.2633, 0.8158, 981
    A = root("A", {"road": 0.7, "offroad": 0.3}, {"road": [E.Parameter(0.15915, "A")], "offroad": [E.Parameter(0.8, "A")]})
    b0 = root("b₀", {"normal_load": 0.7, "over_load": 0.3},
              {"normal_load": [E.Parameter(0.27, "b₀")], "over_load": [E.Parameter(0.5, "b₀")]})
    V = continuous_root("V", E.Uniform(7, 12), E.ExactDiscretization([8.5, 9.5, 10.5, 11.5]))
    C = continuous_root("C", E.Normal(431.7221, 10))
    Ck = continuous_root("Cₖ", E.Normal(1475.5503, 10))
    K = continuous_root("K", E.Normal(55.0406, 10))
    # the body of composite_model, line by line; the two non-integer powers of constants are computed
    # here, by the caller, so that they carry the caller's pow() rounding
    k = dict(M=M, m=m, g=g, Mg15=(M * g) ** -1.5, gMm=(g * (M + m)) ** 0.877)
    models = [
        E.Model(E.Expression("1 .- (π .* m .* df.V .* df.A) ./ (df.b₀ .* df.K .* g .^ 2) .* ((df.Cₖ ./ (m .+ M) .- (df.C ./ M)) .^ 2"
                             " .+ df.C .^ 2 ./ (m .* M) .+ df.Cₖ .* df.K .^ 2 ./ (m .* M .^ 2))", **k), "g1"),
        E.Model(E.Expression("4000 .* df.C .* Mg15 .- 8.6394", **k), "g2"),
        E.Model(E.Expression("2 .* sqrt.(M .* g .* (df.K .^ 2 .* df.Cₖ ./ (df.C .* (m .+ M)) .+ df.C)) .- 1", **k), "g3"),
        E.Model(E.Expression("df.Cₖ .- gMm", **k), "g4"),
        E.Model(E.Expression("minimum([df.g1, df.g2, df.g3, df.g4])"), "y"),
    ]
    node = E.DiscreteFunctionalNode("E", models, E.Expression("df.y"), E.MonteCarlo(n))
    net = E.EnhancedBayesianNetwork([A, b0, V, C, Ck, K, node])
    for p in (A, b0, V, C, Ck, K):
        E.add_child(net, p, node)
    E.order(net)
    return net


def main(n=2 * 10**6):
    net = build(n)
    t = time.time()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")      # pf + (1 - pf) rounding notes of the CPT check
        E.evaluate(net, True, False)
    dt = time.time() - t
    e = net.node("E")
    st = E.last_stats()
    print(f"evaluate!: {dt:.2f} s wall for {len(e.cpt.data) // 2} scenarios x {n:.0e} samples "
          f"(kernel {st['kernel_ms']:.2f} ms, {st['jit_compiles']} kernel compiled at run time)")
    for r in e.cpt.data:
        if r["E"] == "E_fail":
            print(f"  A={r['A']:8s} b₀={r['b₀']:12s} V_d={r['V_d']:13s} pf = {r['Π']:.6f}")
    return e


if __name__ == "__main__":
    main(int(float(sys.argv[1])) if len(sys.argv) > 1 else 2 * 10**6)
