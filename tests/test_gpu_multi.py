"""Multi-GPU tests (need >= 2 devices; skipped on a 1-GPU box): single process driving several
devices — the mode a Julia caller uses — with the int64 count vectors combined by ncclAllReduce
inside the library. Counts must be bit-identical to the 1-GPU result."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_single_process_multi_device_counts_match_one_gpu():
    import ebn_b200
    from ebn_b200 import workloads as W
    ndev = ebn_b200.device_count()
    if ndev < 2:
        pytest.skip("needs at least 2 GPUs")
    ebn_b200.init([0])
    job = W.w1_vehicle(n=3_000_001)
    base = ebn_b200.pf_batch(job)[0]
    prog = np.array([1, 2, 1.0, 1, 0, 1.0, 1, 1])
    rows = [[(0, a, 0, 0, 0, -np.inf, np.inf), (1, 0.0, 1.0, 0, 0, -np.inf, np.inf)] for a in (1.0, 2.0, 3.0)]
    hjob = ebn_b200.Job(ebn_b200._lib.LS_POLYNOMIAL, prog, ebn_b200.make_inputs(rows), n=1_000_001, seed=3)
    edges = np.linspace(-3, 7, 41)
    hbase = ebn_b200.hist_batch(hjob, edges)
    # run-time compiled kernels: one module per device (an expression limit state and a plan compiled on demand)
    from ebn_b200 import expr
    ep = expr.compile_program(["A", "B"], [("C", "df.A .+ df.B")], "2 .- df.C")
    ejob = ebn_b200.Job(ebn_b200._lib.LS_EXPRESSION, ep.consts, hjob.inputs, n=1_000_001, seed=3)
    ebase = ebn_b200.pf_batch(ejob)[0]
    tjob = W.w1_vehicle(n=1_000_001)
    tjob.inputs[:, 3]["lo"] = 400.0
    tjob.inputs[:, 5]["kind"], tjob.inputs[:, 5]["p"] = ebn_b200._lib.IN_LOGNORMAL, (4.0, 0.18, 0, 0)
    # C truncated, K log-normal: no compiled-in plan, compiled at run time
    tjob.jit_plan = True
    tbase = ebn_b200.pf_batch(tjob)[0]
    for devs in ([0, 1], list(range(ndev))):
        ebn_b200.init(devs)
        assert np.array_equal(ebn_b200.pf_batch(ejob)[0], ebase), devs
        assert np.array_equal(ebn_b200.pf_batch(tjob)[0], tbase), devs
        got = ebn_b200.pf_batch(job)[0]
        st = ebn_b200.last_stats()
        assert np.array_equal(got, base), devs
        assert st["devices"] == len(devs) and st["launches"] == len(devs)
        h = ebn_b200.hist_batch(hjob, edges)
        assert np.array_equal(h[0], hbase[0]) and np.allclose(h[1], hbase[1], rtol=1e-12)
    ebn_b200.init([0])
