"""N > 1 host logic on CPU (gloo, world_size 2): band sharding and the max-over-ranks timing reduction that bench.py
uses under torchrun. The path shards by independent bands: no data-path collective exists to test."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bench.bands_for_rank(64, rank, world)
    ms_local = 10.0 + 5.0 * rank  # rank 1 is the slow one
    ms_max = bench.reduce_step_time(ms_local)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        out.put((ms_max, gathered, bench.aggregate_msps(1000, 4, world, ms_max)))
    dist.destroy_process_group()


def test_band_sharding_and_time_reduction_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ms_max, gathered, msps = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ms_max == 15.0  # max over ranks
    assert sorted(gathered[0] + gathered[1]) == list(range(64)) and not set(gathered[0]) & set(gathered[1])
    assert gathered[0] == list(range(0, 64, 2)) and gathered[1] == list(range(1, 64, 2))  # band b -> GPU b mod G
    assert abs(msps - 2 * 1000 * 4 / 0.015 / 1e6) < 1e-9


def test_single_rank_is_identity():
    assert bench.reduce_step_time(3.5) == 3.5
    assert bench.bands_for_rank(8, 0, 1) == list(range(8))
    assert bench.aggregate_msps(67108864, 10, 1, 1000.0) == pytest.approx(671.08864)
