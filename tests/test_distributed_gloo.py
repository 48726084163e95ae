"""world_size-2 gloo test of the clip sharding + clip-boundary all-gather (CPU)."""
import os
import socket

import torch
import torch.multiprocessing as mp

from evoworld_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")  # the container hostname may not resolve
    r, _, w = D.init_from_env("gloo")
    clips = list(D.shard_range(5, r, w, start_idx=10))
    lat = torch.full((1, 3, 4, 2, 4), float(r + 1))
    g = D.gather_latents(lat)
    q.put((r, clips, tuple(g.shape), [float(g[i].mean()) for i in range(w)]))
    import torch.distributed as dist

    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [10, 11, 12] and res[1][1] == [13, 14]
    for r in res:
        assert r[2] == (2, 1, 3, 4, 2, 4) and r[3] == [1.0, 2.0]


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 24):
        for w in (1, 2, 4, 8):
            got = [i for r in range(w) for i in D.shard_range(n, r, w, 3)]
            assert got == list(range(3, 3 + n))
    assert D.gather_latents(torch.zeros(2, 3)).shape == (1, 2, 3)
