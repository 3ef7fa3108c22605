"""N > 1 path on CPU (gloo, world_size 2): each rank takes the work items `ebn_partition` deals
it, evaluates them with the ORACLE's replay of the device streams, and the int64 count vectors are
all-reduced. The sum must equal the single-rank count — the invariant the NCCL path relies on."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _item_counts(job, rank, world):
    """Fail counts of the items of `rank`, via the NumPy twin of the sampler (oracle)."""
    from ebn_b200 import _lib as L
    from oracle import replay
    from oracle.cpu import INPUT_DTYPE
    job.shard_rank, job.shard_world = rank, world
    p = L.partition(job)
    ppi, ips, base = p["pairs_per_item"], p["items_per_scenario"], p["pair_base"]
    first, last = job.sample_offset, job.sample_offset + job.n
    counts = np.zeros(job.S, dtype=np.int64)
    inputs = job.inputs.view(INPUT_DTYPE)
    n_items = 0
    for w in range(rank, ips * job.S, world):
        s, c = divmod(w, ips)
        lo = max(2 * (base + c * ppi), first)
        hi = min(2 * (base + (c + 1) * ppi), last)
        if hi <= lo:
            continue
        X = replay.sample_matrix(inputs[s], s, job.seed, lo, hi - lo)
        counts[s] += int(np.sum(replay.limit_state(job.limit_state, job.consts, X) < 0))
        n_items += 1
    assert n_items <= p["local_items"]
    return counts


def _worker(rank, world, port, n, out):
    sys.path.insert(0, ROOT)
    import ebn_b200  # noqa: F401
    from ebn_b200 import workloads as W
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    job = W.w1_vehicle(n=n)
    job.inputs = job.inputs[[2, 3, 4, 7]].copy()  # 4 scenarios with pf 0.2 .. 0.7 and 5e-4
    job.sample_offset = 3
    t = torch.from_numpy(_item_counts(job, rank, world))
    dist.all_reduce(t)  # the int64 count all-reduce (ncclAllReduce(ncclInt64, ncclSum) on the GPU path)
    if rank == 0:
        np.save(out, t.numpy())
    dist.destroy_process_group()


def test_two_rank_sharding_reproduces_single_rank_counts(tmp_path):
    sys.path.insert(0, ROOT)
    import ebn_b200  # noqa: F401
    from ebn_b200 import workloads as W
    n = 60_001
    out = str(tmp_path / "counts.npy")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, n, out), nprocs=2, join=True)
    got = np.load(out)
    job = W.w1_vehicle(n=n)
    job.inputs = job.inputs[[2, 3, 4, 7]].copy()
    job.sample_offset = 3
    want = _item_counts(job, 0, 1)
    assert np.array_equal(got, want) and want.sum() > 0
    # and the partition-free count over the whole range agrees as well
    from oracle import replay
    from oracle.cpu import INPUT_DTYPE
    inputs = job.inputs.view(INPUT_DTYPE)
    for s in range(job.S):
        X = replay.sample_matrix(inputs[s], s, job.seed, 3, n)
        assert want[s] == int(np.sum(replay.limit_state(job.limit_state, job.consts, X) < 0))
