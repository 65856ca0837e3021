"""N > 1 host logic on CPU: contiguous sharding and the statistics gather over a 2-rank gloo group."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from dgsqp_b200.sharding import shard_bounds, shard_stats, combine_stats, gather_stats


def test_shard_bounds_cover_and_balance():
    for total in (0, 1, 7, 10000, 10001):
        for ws in (1, 2, 4, 8):
            b = [shard_bounds(total, ws, r) for r in range(ws)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _fake_results(n, seed):
    rng = np.random.default_rng(seed)
    status = rng.choice([0, 0, 0, 1, 2, 4], size=n)
    iters = np.where(status == 2, 50, rng.integers(5, 30, size=n))
    qp = iters + rng.integers(0, 10, size=n)
    cond = np.abs(rng.normal(size=(n, 3))) * 1e-4
    return status, iters, qp, cond


def _worker(rank, world, port, total, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    status, iters, qp, cond = _fake_results(total, 0)
    lo, hi = shard_bounds(total, world, rank)
    res = gather_stats(shard_stats(status[lo:hi], iters[lo:hi], qp[lo:hi], cond[lo:hi]))
    if rank == 0:
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_stats_two_ranks_equals_single():
    total = 1001
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    status, iters, qp, cond = _fake_results(total, 0)
    single = combine_stats([shard_stats(status, iters, qp, cond)])
    assert res["count"] == total
    for k in single:
        assert np.isclose(res[k], single[k]), k
    assert res["converged"] == int((status <= 1).sum())
    assert np.isclose(res["mean_iters"], iters.mean()) and np.isclose(res["std_iters"], iters.std())
