"""CPU tests of the boundary: the shared library loads, exports every symbol the header declares,
struct layouts match, and — without a GPU — every compute entry point fails loudly (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import ebn_b200
from ebn_b200 import _lib as L
from ebn_b200 import workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAS_GPU = L.load().ebn_device_count() > 0


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "ebncuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ebn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = _declared_functions()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ebncuda.h but not exported"
    assert sorted(L.EXPORTS) == names
    assert lib.ebn_abi_version() == 2


def test_struct_layouts():
    assert L.INPUT_DTYPE.itemsize == 56 and C.sizeof(L.EbnJob) == 104 and C.sizeof(L.EbnStats) == 48
    assert L.EbnJob.consts.offset == 8 and L.EbnJob.inputs.offset == 24 and L.EbnJob.n_per_scenario.offset == 32
    assert L.EbnJob.corr_chol.offset == 56 and L.EbnJob.shard_rank.offset == 96


def test_limit_state_catalogue():
    assert L.limit_state_id("vehicle_suspension") == L.LS_VEHICLE_SUSPENSION
    assert L.limit_state_id("straub_frame") == L.LS_STRAUB_FRAME
    assert L.limit_state_name(L.LS_POLYNOMIAL) == "polynomial"
    assert L.limit_state_arity(L.LS_VEHICLE_SUSPENSION) == (6, 5)
    assert L.limit_state_arity(L.LS_MIN_AFFINE) == (-1, -1)
    with pytest.raises(L.EbnError, match="EBN_ELIMITSTATE"):
        L.limit_state_id("tower")  # demo/Diegos_tower is a discrete credal net: no limit state


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU behaviour")
def test_no_gpu_means_error_not_fallback():
    with pytest.raises(L.EbnError, match="EBN_ENODEV.*no CPU fallback"):
        ebn_b200.init()
    lib = L.load()
    job = W.w1_vehicle(n=1000)
    cj = job.c_struct()
    cnt = np.zeros(job.S, dtype=np.int64)
    rc = lib.ebn_pf_batch(C.byref(cj), cnt.ctypes.data_as(C.POINTER(C.c_int64)), None, None)
    assert rc == -6 and b"ebn_init" in lib.ebn_last_error()
    assert (cnt == 0).all()


def test_partition_is_host_logic_and_covers_every_pair_once():
    for n, off in ((1, 0), (999, 0), (10**6 + 1, 7), (10**8, 0), (10**10, 2**40 + 1)):
        job = W.w1_vehicle(n=n)
        job.sample_offset = off
        for world in (1, 2, 3, 8):
            job.shard_world = world
            tot, grids = 0, set()
            for r in range(world):
                job.shard_rank = r
                p = L.partition(job)
                tot += p["local_items"]
                grids.add((p["pairs_per_item"], p["items_per_scenario"], p["pair_base"]))
            assert len(grids) == 1  # every rank derives the same grid
            ppi, ips, base = next(iter(grids))
            assert tot == ips * job.S
            pairs = ((off + n + 1) >> 1) - (off >> 1)
            assert (ips - 1) * ppi < pairs <= ips * ppi and base == off >> 1
            assert ppi % 128 == 0   # a multiple of the block size (EBN_THREADS)


def test_partition_validates_jobs_without_a_gpu():
    job = W.w1_vehicle(n=1000)
    job.inputs[2, 4]["p"][1] = -1.0  # sigma < 0
    with pytest.raises(L.EbnError, match="EBN_EMARGINAL.*sigma"):
        L.partition(job)
    job = W.w1_vehicle(n=1000)
    job.shard_rank, job.shard_world = 5, 4
    with pytest.raises(L.EbnError, match="EBN_EINVAL.*shard_rank"):
        L.partition(job)
    job = W.w2_straub(n=1000)
    job.inputs[0, 1]["lo"] = 10.0  # truncated Gamma must come tabulated
    with pytest.raises(L.EbnError, match="EBN_EMARGINAL.*TABULATED"):
        L.partition(job)


def test_frozen_profile_belongs_to_this_build():
    """bench.py quotes instruction counts from profiles/w1_kernel_profile.json and refuses to run when the file
    was frozen for another build of the kernels: keep the two in step (python tools/freeze_profile.py)."""
    import json
    prof = json.load(open(os.path.join(ROOT, "profiles", "w1_kernel_profile.json")))
    assert prof["build_id"] == L.build_id(), "kernel sources changed: run `python tools/freeze_profile.py`"
    assert 40 < prof["sass"]["fp64_instructions_per_sample"] < 80
    assert len(L.build_id()) == 16
