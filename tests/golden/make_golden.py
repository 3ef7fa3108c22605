"""Generates the golden fixtures under tests/golden/ (committed together with this script).

The reference holds no sample matrices or bit-level vectors for this path and
cannot be executed here (Julia), so the fixtures are produced by the ORACLE:
  * limit-state fixtures: seeded NumPy sample matrices + the C oracle's g and fail count
    (cross-checked against the independent NumPy restatement before writing);
  * stream fixture: the first samples of the all-marginals scenario as the NumPy twin of the
    stream layout (oracle/replay.py) draws them — this pins Philox + layout + transforms.
Anyone with Julia can regenerate the limit-state half from the real reference with
baseline/julia/export_fixtures.jl and compare.

    python tests/golden/make_golden.py
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import cpu as ocpu  # noqa: E402
from oracle import replay  # noqa: E402

INF = math.inf
VEH_K = np.array([3.2633, 0.8158, 981.0, (3.2633 * 981.0) ** (-1.5), (981.0 * (3.2633 + 0.8158)) ** 0.877])
STRAUB_K = np.array([4.991024937519615, 0.1980422004353651, 0.5477])

ALL_KINDS_ROW = [
    (0, 3.25, 0, 0, 0, -INF, INF), (1, 1.0, 2.0, 0, 0, -INF, INF), (1, 1.0, 2.0, 0, 0, 0.5, 4.0),
    (2, 0.3, 0.25, 0, 0, -INF, INF), (2, 0.3, 0.25, 0, 0, 1.0, 2.0), (3, 7.0, 12.0, 0, 0, 8.5, 9.5),
    (4, 41.0, 15.6, 0, 0, -INF, INF), (4, 41.0, 15.6, 0, 0, 30.0, 80.0), (5, 25.0, 2.4, 0, 0, -INF, INF),
    (5, 0.6, 1.5, 0, 0, -INF, INF), (6, 2.0, 1.0, 3.0, 0, -INF, INF), (6, 2.0, -1.0, 3.0, 0, 0.0, INF),
    (7, 0.0, 5.0, 0, 0, -INF, INF),
]
TABLE = np.array([-1.0, 0.0, 0.5, 2.0, 10.0])


VEHICLE_TEXT = [  # demo/vehicle_suspension.jl:17-23
    ("g1", "1 .- (π .* m .* df.V .* df.A) ./ (df.b₀ .* df.K .* g .^ 2) .* ((df.Cₖ ./ (m .+ M) .- (df.C ./ M)) .^ 2"
           " .+ df.C .^ 2 ./ (m .* M) .+ df.Cₖ .* df.K .^ 2 ./ (m .* M .^ 2))"),
    ("g2", "4000 .* df.C .* Mg15 .- 8.6394"),
    ("g3", "2 .* sqrt.(M .* g .* (df.K .^ 2 .* df.Cₖ ./ (df.C .* (m .+ M)) .+ df.C)) .- 1"),
    ("g4", "df.Cₖ .- gMm"),
    ("y", "minimum([g1[1], g2, g3, g4[1]])"),
]
CAPACITY_TEXT = "exp(rand(Normal(λ + ρ * ζ * df.Uᵣ, sqrt((1 - ρ^2) * ζ^2))))"      # demo/straub_example.jl:16-27
FRAME_TEXT = ("minimum([r1 + r2 + r4 + r5 - 5 * df.H, r2 + 2 * r3 + r4 - 5 * df.V, "
              "r1 + 2 * r3 + 2 * r4 + r5 - 5 * df.H - 5 * df.V])")                  # :35-40


def main():
    rng = np.random.default_rng(20261019)
    n = 2048
    Xv = np.column_stack([rng.choice([0.15915, 0.8], n), rng.choice([0.27, 0.5], n), rng.uniform(7, 12, n),
                          rng.normal(431.7221, 10, n), rng.normal(1475.5503, 10, n), rng.normal(55.0406, 10, n)])
    cv, gv = ocpu.eval_matrix(ocpu.LS_VEHICLE, VEH_K, Xv, want_g=True)
    assert np.array_equal(gv, replay.g_vehicle(Xv, VEH_K))
    Xs = np.column_stack([rng.normal(0, 1, n), rng.gamma(25.0, 2.4, n),
                          rng.gumbel(40.99893584908611, 15.593936024673523, n), rng.normal(0, 1, (n, 5))])
    cs, gs = ocpu.eval_matrix(ocpu.LS_STRAUB, STRAUB_K, Xs, want_g=True)
    assert np.allclose(gs, replay.g_straub(Xs, STRAUB_K), rtol=0, atol=1e-10)
    np.savez_compressed(os.path.join(HERE, "limit_states.npz"), vehicle_X=Xv, vehicle_k=VEH_K, vehicle_g=gv,
                        vehicle_count=cv, straub_X=Xs, straub_k=STRAUB_K, straub_g=gs, straub_count=cs)
    inputs = ocpu.make_inputs([ALL_KINDS_ROW])
    seed = 0xABCDEF0123456789
    blocks = {}
    for first in (0, 7, 2**33 + 5):
        blocks[f"first_{first}"] = replay.sample_matrix(inputs[0], 3, seed, first, 64, tables=TABLE)
    np.savez_compressed(os.path.join(HERE, "stream_v4.npz"), inputs=inputs, table=TABLE, seed=np.uint64(seed),
                        stream=np.uint32(3), **blocks)
    # expression programs: the two demos' closures as text, compiled by the product's text compiler
    # (pins the compiler's instruction order) and evaluated by the oracle's stack machine on the same
    # matrices; g must equal the hand-written restatements above bit for bit
    from ebn_b200 import expr
    veh = expr.compile_program(["A", "b₀", "V", "C", "Cₖ", "K"], VEHICLE_TEXT, "df.y",
                               params={"M": VEH_K[0], "m": VEH_K[1], "g": VEH_K[2], "Mg15": VEH_K[3], "gMm": VEH_K[4]})
    stb = expr.compile_program(["Uᵣ", "V", "H"], [(f"r{i}", CAPACITY_TEXT) for i in range(1, 6)] + [("G", FRAME_TEXT)],
                               "df.G", params={"λ": STRAUB_K[0], "ζ": STRAUB_K[1], "ρ": STRAUB_K[2]})
    cve, gve = ocpu.eval_matrix(ocpu.LS_EXPRESSION, veh.consts, Xv, want_g=True)
    cse, gse = ocpu.eval_matrix(ocpu.LS_EXPRESSION, stb.consts, Xs, want_g=True)
    assert cve == cv and np.array_equal(gve, gv) and cse == cs and np.array_equal(gse, gs)
    np.savez_compressed(os.path.join(HERE, "expression.npz"), vehicle_program=veh.consts, straub_program=stb.consts)
    print("vehicle count", cv, "straub count", cs, "program sizes", veh.consts.size, stb.consts.size)


if __name__ == "__main__":
    main()
