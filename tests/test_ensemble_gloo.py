"""The N > 1 host logic (scene-ensemble sharding and the whole-job reduction bench.py uses) on CPU with the gloo
backend, world_size 2."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ensemble


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = ensemble.scene_shard(7, rank, world)
    # pretend every scene took (rank+1) ms per frame and ran 10 ADMM iterations
    ms_local = 5.0 * (rank + 1)
    units_local = 10.0 * len(mine)
    ms_job, units_job = ensemble.reduce_job(ms_local, units_local)
    dist.barrier()
    q.put((rank, mine, ms_job, units_job))
    dist.destroy_process_group()


def test_shards_cover_every_scene_once():
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 64):
            shards = [ensemble.scene_shard(n, r, world) for r in range(world)]
            flat = sorted(i for s in shards for i in s)
            assert flat == list(range(n))
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1


def test_world_size_two_gloo_reduction():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5]
    for _, _, ms_job, units_job in res:
        assert ms_job == 10.0           # MAX over ranks
        assert units_job == 70.0        # SUM over ranks
    assert ensemble.throughput(10.0, 70.0) == 7000.0


def test_single_process_passthrough():
    assert ensemble.reduce_job(3.0, 12.0) == (3.0, 12.0)
