"""CPU oracle — TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package. The product
(``enhancedbayesiannetworks.jl_b200``) never does.

Two independent restatements live here:

* ``oracle.cpu``    — ctypes front end of ``ebn_oracle.c`` (plain C, Julia
  evaluation order): limit states over a sample matrix and the structural
  Monte Carlo of UQ.jl (``sample`` → ``evaluate!`` → ``sum(g .< 0)/n``).
* ``oracle.replay`` — NumPy twin: Philox4x32-10 and the sampler's stream layout
  restated from DESIGN.md, so the exact uniforms the GPU consumed can be
  regenerated on the CPU, transformed with SciPy's special functions and pushed
  through the NumPy limit states.

Parity status: bit-level parity with the reference is UNPINNED (the reference
holds no seeded vectors for this path and Julia cannot run here — SURVEY.md
§8c); the oracle is pinned to the reference's statistical and analytic anchors
in ``tests/test_oracle.py``.
"""
