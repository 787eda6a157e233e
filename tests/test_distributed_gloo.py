"""The N > 1 host logic on CPU: two gloo ranks check placement, the
max-over-ranks timing reduction and the fixed-size blob exchange that carries
CUDA IPC handles on the GPU box."""

import os
import socket

import pytest

from covalent_ssh_plugin_b200 import distributed as D


def test_round_robin_placement_is_a_partition():
    n, world = 10000, 8
    seen = []
    for r in range(world):
        ids = list(D.shard(n, r, world))
        assert all(D.owner_of(i, world) == r for i in ids)
        seen += ids
    assert sorted(seen) == list(range(n))
    assert D.shard_sizes(n, world) == [1250] * 8 and sum(D.shard_sizes(10, 4)) == 10
    # BASELINE config 4: fan-out -> reduce; 7/8 of the edges into node 999 cross GPUs
    edges = [(i, 999) for i in range(1, 999)]
    crossing = D.crossing_edges(edges, 8)
    assert abs(len(crossing) / len(edges) - 7 / 8) < 0.01


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        slow = D.max_over_ranks(1.0 + rank)                    # every rank must see the slowest
        blobs = D.all_gather_bytes(bytes([rank]) * 72)         # same size as an IPC record
        mine = list(D.shard(11, rank, world))
        # shared dependency: rank 0's array lands in every rank's arena (host path under gloo)
        import numpy as np

        from covalent_ssh_plugin_b200 import _b2lib, engine as engine_mod
        from tests.fake_b2lib import FakeLib

        _b2lib._lib = FakeLib(n_devices=2)
        eng = engine_mod.get_engine(rank, arena_bytes=48 << 20, staging_bytes=8 << 20, n_segments=2,
                                    max_batch=16, queue_cap=64)
        src = np.arange(600, dtype=np.float32).reshape(20, 30) if rank == 0 else None
        da = D.broadcast_shared(eng, src, root=0)
        shared_ok = bool(np.array_equal(da.numpy(), np.arange(600, dtype=np.float32).reshape(20, 30)))
        q.put((rank, slow, [b[0] for b in blobs], [len(b) for b in blobs], mine, shared_ok))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_gloo_ranks():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=150) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [g[1] for g in got] == [2.0, 2.0]
    assert all(g[2] == [0, 1] and g[3] == [72, 72] for g in got)
    assert sorted(got[0][4] + got[1][4]) == list(range(11))
    assert all(g[5] for g in got)
