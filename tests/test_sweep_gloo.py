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
