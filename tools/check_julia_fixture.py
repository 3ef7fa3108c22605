"""Checks a fixture written by baseline/julia/export_fixtures.jl (real reference output) against the
C oracle and — on a GPU box — against ebn_eval_matrix(strict=1), bit for bit.

    python tools/check_julia_fixture.py vehicle_fixture.bin vehicle_fixture.json
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(bin_path, json_path):
    from oracle import cpu as ocpu
    meta = json.load(open(json_path))
    n, cols = meta["n"], meta["columns"]
    raw = np.fromfile(bin_path, dtype="<f8").reshape(len(cols), n).T
    X, y = raw[:, :6], raw[:, 6]
    k = np.array(meta["consts"])
    cnt, g = ocpu.eval_matrix(ocpu.LS_VEHICLE, k, X, want_g=True)
    ties = int(np.sum(np.abs(y) < 1e-12))
    print(f"oracle: count {cnt} vs reference {meta['count']}; max |g - y| = {np.max(np.abs(g - y)):.3e}; ties {ties}")
    ok = cnt == meta["count"] and np.array_equal(g, y)
    try:
        import ebn_b200
        ebn_b200.init()
        c2, g2 = ebn_b200.eval_matrix(ebn_b200._lib.LS_VEHICLE_SUSPENSION, k, X, strict=True, want_g=True)
        print(f"libebncuda strict: count {c2}; bit-identical g: {np.array_equal(g2, y)}")
        ok = ok and c2 == meta["count"]
    except Exception as e:  # no GPU here
        print("libebncuda not checked:", e)
    print("PASS" if ok else "MISMATCH (if only the last ulp of g differs, compare the consts: Julia's `^` vs libm pow)")


if __name__ == "__main__":
    main(*sys.argv[1:3])
