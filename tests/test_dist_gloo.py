"""world_size-2 gloo tests (CPU) of the multi-rank host logic: unique-id hand-out, sub-file split, max-over-ranks timing,
and the property the multi-GPU path relies on: sharded int64 fixed-point planes sum to the unsharded plane exactly,
whatever the split (the reference's float MPI_Reduce does not have this property)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys_path_root = os.path.dirname(os.path.abspath(__file__))
import sys

if sys_path_root not in sys.path:
    sys.path.insert(0, sys_path_root)
if os.path.dirname(sys_path_root) not in sys.path:
    sys.path.insert(0, os.path.dirname(sys_path_root))
import dist_helpers as sdist  # noqa: E402  (test-side helpers: rank plumbing on any torch.distributed backend)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = sdist.broadcast_unique_id(lambda: bytes(range(128)), rank)
        t = sdist.max_over_ranks(1.0 + rank)
        # every rank deposits its shard of the accepted particles with the oracle's fixed-point accumulator
        import sys

        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from oracle.oracle_bindings import Oracle

        orc = Oracle()
        rng = np.random.default_rng(5)
        n, nn = 20000, 32
        xs, ys = rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32)
        ms = (rng.random(n) * 3).astype(np.float32)
        mine = sdist.balanced_subfiles(n, world, rank)
        part = orc.gridist_w_fixed(xs[mine], ys[mine], ms[mine], nn, 40)
        total = sdist.sum_int64_planes(part)
        # the bench's hardware check of ncclReduce(int64): linear checksums of the per-rank planes add up to the reduced plane's
        import bench

        mine_cs = torch.tensor([v - (1 << 64) if v >= (1 << 63) else v for v in bench.linear_checksums(part)], dtype=torch.int64)
        allcs = [torch.zeros_like(mine_cs) for _ in range(world)]
        dist.all_gather(allcs, mine_cs)
        lo, hi = bench.shard_range(1000003, world, rank)
        if rank == 0:
            full = orc.gridist_w_fixed(xs, ys, ms, nn, 40)
            got = bench.linear_checksums(total.numpy())
            want = tuple(sum(int(c[k]) for c in allcs) & ((1 << 64) - 1) for k in range(2))
            q.put(dict(uid_ok=uid == bytes(range(128)), tmax=t, exact=bool(np.array_equal(total.numpy(), full)), checks=got == want,
                       shard=(lo, hi)))
        else:
            q.put(dict(uid_ok=uid == bytes(range(128)), tmax=t, shard=(lo, hi)))
    finally:
        dist.destroy_process_group()


def test_two_rank_plumbing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r["uid_ok"] for r in res) and all(r["tmax"] == 2.0 for r in res)
    assert any(r.get("exact") for r in res)
    assert any(r.get("checks") for r in res)
    shards = sorted(r["shard"] for r in res)
    assert shards[0][0] == 0 and shards[0][1] == shards[1][0] and shards[1][1] == 1000003 and shards[0][1] % 4 == 0


def test_subfile_split_matches_reference():
    # slicer-v2.cpp:162-175
    assert [sdist.subfile_range(10, 4, r) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 10)]
    assert [sdist.subfile_range(2, 4, r) for r in range(4)] == [(0, 0), (0, 0), (0, 0), (0, 2)]  # only the last rank works
    assert sorted(sum((sdist.balanced_subfiles(10, 4, r) for r in range(4)), [])) == list(range(10))
