"""A user's own limit state on the GPU in ten lines: UQ.jl-style `probability_of_failure` with the
model closure given as text (the last block of demo/vehicle_suspension.jl, "MonteCarlo Checking",
is this call with the vehicle model).

    python examples/custom_model.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ebn_b200 as E  # noqa: E402

# a short column under axial load P with an eccentricity e: capacity minus demand, strength fy lognormal
inputs = [
    E.RandomVariable(E.LogNormal(*E.distribution_parameters(355.0, 25.0, "LogNormal")), "fy"),   # MPa
    E.RandomVariable(E.Gumbel(*E.distribution_parameters(900.0, 120.0, "Gumbel")), "P"),          # kN
    E.RandomVariable(E.Normal(12.0, 4.0), "e"),                                                   # mm
    E.Parameter(5400.0, "A"),                                                                     # mm^2
    E.Parameter(3.9e5, "W"),                                                                      # mm^3
]
stress = E.Model(E.Expression("1000 * df.P / df.A + 1000 * df.P * abs(df.e) / df.W"), "sigma")     # MPa
performance = E.Expression("df.fy - df.sigma")

if __name__ == "__main__":
    pf, std, samples = E.probability_of_failure([stress], performance, inputs, E.MonteCarlo(10**9))
    st = E.last_stats()
    print(f"pf = {pf:.6e} +- {std:.1e}   (1e9 samples, kernel {st['kernel_ms']:.1f} ms, "
          f"{st['jit_compiles']} kernel compiled at run time)")
    rows = samples.fetch(0, 5)          # the first rows of the DataFrame UQ.jl would have built, replayed
    print({k: [round(float(x), 3) for x in v] for k, v in rows.items()})
