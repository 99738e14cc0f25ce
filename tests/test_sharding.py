"""Multi-GPU path on CPU: two `gloo` ranks shard a batch of dimer searches by replica (openfly_b200/shard.py),
each runs its slice (here with the CPU oracle standing in for the device call -- tests may), the records are
merged on the host, and the result equals the single-process batch bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

from tests.common import ROOT, WEN, bcc_vacancy


def test_replica_ids_partition():
    from openfly_b200 import shard

    for total in (1, 7, 64, 2048):
        for world in (1, 2, 4, 8):
            got = np.concatenate([shard.replica_ids(total, world, r) for r in range(world)])
            assert sorted(got.tolist()) == list(range(total))
            sizes = [len(shard.replica_ids(total, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.replica_ids(8, 2, 2)


def test_merge_without_process_group():
    from openfly_b200 import shard

    ids = np.array([2, 0, 1])
    out = shard.merge_records(ids, {"code": np.array([20, 0, 10])}, 3)
    assert out["code"].tolist() == [0, 10, 20]
    with pytest.raises(ValueError):
        shard.merge_records(np.array([0, 1]), {"code": np.array([0, 1])}, 3)


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    from openfly_b200 import shard
    from oracle import oracle as O

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        base, ex, ids = bcc_vacancy(3, a=2.855, remove=(1,), hydrogen=True)
        fz = np.zeros(len(ids), np.uint8)
        tab = O.Tables(WEN)
        osys = O.System(tab, ex)
        r0, _, failed, _ = osys.minimise(base, ids, fz, threads=1)
        starts = [osys.perturb(r0, fz, len(ids) - 1, (1, 3, 3, 4), rep) for rep in range(total)]
        P, A = np.stack([s[0] for s in starts]), np.stack([s[1] for s in starts])
        mine, (p, a) = shard.shard((P, A), total, world, rank)
        iopt = O.DIMER_IOPT.copy()
        iopt[1] = 15  # short searches: this test is about the plumbing
        o_r, o_ax, code, cnt = osys.find_sp_batch(p, a, ids, fz, r0, threads=1, iopt=iopt)
        merged = shard.merge_records(mine, {"pos": o_r, "code": code, "evals": cnt[:, 0]}, total)
        if rank == 0:
            full = osys.find_sp_batch(P, A, ids, fz, r0, threads=2, iopt=iopt)
            q.put((np.array_equal(merged["pos"], full[0]), merged["code"].tolist() == full[2].tolist(),
                   merged["evals"].tolist() == full[3][:, 0].tolist()))
    finally:
        dist.destroy_process_group()


def test_two_gloo_ranks_equal_one_process(oracle_lib):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 6, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == (True, True, True)


def test_bench_start_states_do_not_depend_on_the_sharding():
    """bench.py draws the start state of global search g from the stream (seed, g): a rank's shard of the global queue is the
    same set of states whatever the number of ranks, so the N-GPU runs of the scaling bench work on ONE global queue."""
    import bench
    from openfly_b200 import inputs, shard

    pos, ex, ids = inputs.iron_hydrogen_cell(3, 2.855, remove=(1,), hydrogen=True)
    total = 12
    P, A = bench.start_states(pos, ex, len(ids) - 1, np.arange(total), seed=5)
    for world in (2, 4):
        for rank in range(world):
            gids = shard.replica_ids(total, world, rank)
            p, a = bench.start_states(pos, ex, len(ids) - 1, gids, seed=5)
            assert np.array_equal(p, P[gids]) and np.array_equal(a, A[gids])
    cfg = bench.config_dict(1334)
    assert cfg["queued_searches_per_gpu_per_step"] == bench.QUEUE and cfg["resident_searches_per_gpu"] == bench.SLOTS
