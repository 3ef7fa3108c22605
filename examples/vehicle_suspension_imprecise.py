"""demo/vehicle_suspension_imprecise.jl of the reference on the GPU (BASELINE config 4): the speed V is
only known as the interval (7, 13). Discretisation turns it into interval bins; a DoubleLoop searches
each bin for the smallest and largest failure probability. All outer-loop candidates of all scenarios
run in one batch per search step, with common random numbers (one Philox stream per scenario), so the
inner estimate is a deterministic function of V and the search sees a smooth objective.

    python examples/vehicle_suspension_imprecise.py [inner samples, default 1e6]
"""
import os
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ebn_b200 as E  # noqa: E402
from vehicle_suspension import continuous_root, root  # noqa: E402  (same directory)


def build(n, method="optimise"):
    M, m, g = 3.2633, 0.8158, 981
    A = root("A", {"road": 0.7, "offroad": 0.3}, {"road": [E.Parameter(0.15915, "A")], "offroad": [E.Parameter(0.8, "A")]})
    b0 = root("b₀", {"normal_load": 0.7, "over_load": 0.3},
              {"normal_load": [E.Parameter(0.27, "b₀")], "over_load": [E.Parameter(0.5, "b₀")]})
    cpt = E.ContinuousConditionalProbabilityTable(kind="interval")
    cpt[()] = (7, 13)                                   # demo lines 28-29
    V = E.ContinuousNode("V", cpt, E.ExactDiscretization([8.5, 9.5, 10.5, 11.5]))
    C = continuous_root("C", E.Normal(430, 10))        # demo line 38
    Ck = continuous_root("Cₖ", E.Normal(1475.5503, 10))
    K = continuous_root("K", E.Normal(55.0406, 10))
    k = dict(M=M, m=m, g=g, Mg15=(M * g) ** -1.5, gMm=(g * (M + m)) ** 0.877)
    models = [
        E.Model(E.Expression("1 .- (π .* m .* df.V .* df.A) ./ (df.b₀ .* df.K .* g .^ 2) .* ((df.Cₖ ./ (m .+ M) .- (df.C ./ M)) .^ 2"
                             " .+ df.C .^ 2 ./ (m .* M) .+ df.Cₖ .* df.K .^ 2 ./ (m .* M .^ 2))", **k), "g1"),
        E.Model(E.Expression("4000 .* df.C .* Mg15 .- 8.6394", **k), "g2"),
        E.Model(E.Expression("2 .* sqrt.(M .* g .* (df.K .^ 2 .* df.Cₖ ./ (df.C .* (m .+ M)) .+ df.C)) .- 1", **k), "g3"),
        E.Model(E.Expression("df.Cₖ .- gMm", **k), "g4"),
        E.Model(E.Expression("minimum([df.g1, df.g2, df.g3, df.g4])"), "y"),
    ]
    sim = E.DoubleLoop(E.MonteCarlo(n), method=method)
    node = E.DiscreteFunctionalNode("E", models, E.Expression("df.y"), sim)
    net = E.EnhancedBayesianNetwork([A, b0, V, C, Ck, K, node])
    for p in (A, b0, V, C, Ck, K):
        E.add_child(net, p, node)
    E.order(net)
    return net


def main(n=10**6, method="optimise"):
    net = build(n, method)
    t = time.time()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        E.evaluate(net, True, False)
    dt = time.time() - t
    e = net.node("E")
    print(f"evaluate!: {dt:.2f} s wall, inner n = {n:.0e}, outer search '{method}'")
    for r in e.cpt.data:
        if r["E"] == "E_fail" and r["A"] == "offroad" and r["b₀"] == "normal_load":
            lb, ub = r["Π"]
            print(f"  A=offroad b₀=normal_load V_d={r['V_d']:13s} pf in [{lb:.5f}, {ub:.5f}]")
    return e


if __name__ == "__main__":
    main(int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**6)
