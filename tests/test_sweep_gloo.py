"""world_size-2 gloo test of the sweep sharding host logic (the N>1 path of bench.py / scuff-em_b200/sweep.py)."""
import importlib.util
import os
import socket
import sys

import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load():
    spec = importlib.util.spec_from_file_location("sweep", os.path.join(ROOT, "scuff-em_b200", "sweep.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sweep = _load()
    omegas = [0.05 * 1.5 ** j for j in range(11)]           # the deterministic xi_j sweep of SURVEY.md §8d
    mine = sweep.points_for_rank(omegas, rank, world)
    res = sweep.run_sweep(omegas, lambda w: {"omega": w, "rank": rank, "val": w * w}, rank, world)
    q.put((rank, [j for j, _ in mine], res))
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_sharding_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict()
    for _ in range(2):
        rank, idx, res = q.get(timeout=120)
        got[rank] = (idx, res)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][0] == [0, 2, 4, 6, 8, 10] and got[1][0] == [1, 3, 5, 7, 9]
    assert got[1][1] is None
    res = got[0][1]
    assert [r["rank"] for r in res] == [j % 2 for j in range(11)]
    assert all(abs(r["val"] - (0.05 * 1.5 ** j) ** 2) < 1e-15 for j, r in enumerate(res))


def test_single_rank_and_errors():
    sweep = _load()
    assert sweep.run_sweep([1.0, 2.0, 3.0], lambda w: 2 * w) == [2.0, 4.0, 6.0]
    assert sweep.points_for_rank(list("abcde"), 1, 4) == [(1, "b")]
    assert sweep.points_for_rank([], 0, 2) == []
    import pytest
    with pytest.raises(RuntimeError):
        sweep.gather_results([(0, 1.0)], 2, 0, 1)
    with pytest.raises(RuntimeError):
        sweep.gather_results([(0, 1.0), (0, 2.0)], 1, 0, 1)


def _gather_worker(rank, world, port, q):
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sweep = _load()
    n = 37
    ranges = sweep.column_ranges(n, world)
    # column-major n x n complex matrix as 2*n*n doubles; rank r fills its own columns with (r+1)*1000 + flat index
    flat = torch.zeros(2 * n * n, dtype=torch.float64)
    c0, c1 = ranges[rank]
    flat[2 * n * c0:2 * n * c1] = (rank + 1) * 1000.0 + torch.arange(2 * n * c0, 2 * n * c1, dtype=torch.float64)
    sweep.gather_columns(flat, n, 2, ranges, rank, world, dst=0)
    q.put((rank, ranges, flat.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_column_block_gather_world2():
    import torch
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        rank, ranges, flat = q.get(timeout=120)
        got[rank] = (ranges, flat)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ranges, flat = got[0]
    n = 37
    assert ranges[0][0] == 0 and ranges[-1][1] == n and ranges[0][1] == ranges[1][0]
    for r, (c0, c1) in enumerate(ranges):
        want = (r + 1) * 1000.0 + torch.arange(2 * n * c0, 2 * n * c1, dtype=torch.float64)
        assert torch.equal(flat[2 * n * c0:2 * n * c1], want)


def test_column_ranges_balance_upper_triangle():
    sweep = _load()
    for n, w in ((20280, 8), (501, 4), (10, 3), (5, 8)):
        rs = sweep.column_ranges(n, w)
        assert rs[0][0] == 0 and rs[-1][1] == n and all(a[1] == b[0] for a, b in zip(rs, rs[1:])) and all(c1 >= c0 for c0, c1 in rs)
    rs = sweep.column_ranges(20280, 8)
    areas = [c1 * c1 - c0 * c0 for c0, c1 in rs]
    assert max(areas) / min(areas) < 1.01


class _FakePS:
    """stands in for pyscuff in the CPU test of share_fippi_table: a 'table' of nrec records of which this rank fills its share"""
    def __init__(self, nrec):
        import numpy as np
        self.np = np; self.nrec = nrec; self.part = (0, 1); self.rec = None; self.committed = None
    def fippi_set_partition(self, rank, world): self.part = (rank, world)
    def fippi_host_records(self): return self.rec
    def fippi_commit(self): self.committed = self.rec.copy()
    def cache_records(self): return self.nrec


class _FakeG:
    def __init__(self, ps): self.ps = ps
    def PrepareNearPlans(self):
        np = self.ps.np; r, w = self.ps.part
        full = np.arange(self.ps.nrec * 36, dtype=np.float64).reshape(self.ps.nrec, 36) + 1.0
        self.ps.rec = np.where((np.arange(self.ps.nrec) % w == r)[:, None], full, 0.0)


def _fippi_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sweep = _load()
    ps = _FakePS(1001)
    n = sweep.share_fippi_table(ps, _FakeG(ps), rank, world)
    mismatch = None
    try:
        bad = _FakePS(1001 + rank)                       # ranks that do not hold the same geometry are detected
        sweep.share_fippi_table(bad, _FakeG(bad), rank, world)
    except RuntimeError as e:
        mismatch = str(e)
    q.put((rank, n, ps.part, float(ps.committed.sum()), bool((ps.committed > 0).all()), mismatch))
    dist.barrier()
    dist.destroy_process_group()


def test_fippi_table_built_once_by_all_ranks_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_fippi_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = float(sum(range(1, 1001 * 36 + 1)))
    for rank, n, part, total, complete, mismatch in got:
        assert n == 1001 and part == (0, 1)              # partition reset after the build
        assert total == want and complete                # every rank ends with the union of the shares
        assert mismatch and "differ in size" in mismatch
