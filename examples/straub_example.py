"""demo/straub_example.jl of the reference on the GPU: five plastic-moment capacities, each with its
own in-model `rand(Normal(...))`, fused into the frame node by `evaluate!` and sampled by one kernel.
Expected (demo line 71): Π ≈ [0.026129, 0.973871] (atol 0.01).

    python examples/straub_example.py [samples, default 1e6 like the demo]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ebn_b200 as E  # noqa: E402


def continuous_root(name, dist):
    cpt = E.ContinuousConditionalProbabilityTable()
    cpt[()] = dist
    return E.ContinuousNode(name, cpt)


def build(n):
    ur = continuous_root("Uᵣ", E.Normal())
    v = continuous_root("V", E.Gamma(*E.distribution_parameters(60, 60 * 0.2, "Gamma")))
    h = continuous_root("H", E.Gumbel(*E.distribution_parameters(50, 0.4 * 50, "Gumbel")))
    lam, zeta = E.distribution_parameters(150, 150 * 0.2, "LogNormal")
    # plastic_moment_capacities(uᵣ), demo lines 16-27
    capacity = E.Expression("exp(rand(Normal(λ + ρ * ζ * df.Uᵣ, sqrt((1 - ρ^2) * ζ^2))))", λ=lam, ζ=zeta, ρ=0.5477)
    sim = E.MonteCarlo(n)
    rs = [E.ContinuousFunctionalNode(f"R{i}", [E.Model(capacity, f"r{i}")], sim) for i in range(1, 6)]
    # frame_model, demo lines 35-40
    frame_model = E.Expression("minimum([df.r1 + df.r2 + df.r4 + df.r5 - 5 * df.H, df.r2 + 2 * df.r3 + df.r4 - 5 * df.V, "
                               "df.r1 + 2 * df.r3 + 2 * df.r4 + df.r5 - 5 * df.H - 5 * df.V])")
    frame = E.DiscreteFunctionalNode("E", [E.Model(frame_model, "G")], E.Expression("df.G"), sim)
    net = E.EnhancedBayesianNetwork([ur, v, h] + rs + [frame])
    for r in rs:
        E.add_child(net, ur, r)
        E.add_child(net, r, frame)
    E.add_child(net, v, frame)
    E.add_child(net, h, frame)
    E.order(net)
    return net


def main(n=10**6):
    net = build(n)
    t = time.time()
    E.evaluate(net)
    dt = time.time() - t
    pi = net.node("E").cpt.column("Π")
    print(f"evaluate!: {dt:.2f} s wall, n = {n:.0e}; Π = {pi}  (reference: [0.026129, 0.973871] atol 0.01)")
    return pi


if __name__ == "__main__":
    main(int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**6)
