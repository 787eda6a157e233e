"""Run under torchrun on >= 2 GPUs (not collected by pytest): every rank maps its
peers' arenas through CUDA IPC handles and pushes payloads into the next rank's
arena with the span mover (K4); each rank then verifies what landed.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/mp_peer_check.py
"""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ.setdefault("HOME", "/root")


def payload(sender: int, n: int, sz: int) -> np.ndarray:
    return np.random.default_rng([7, sender]).integers(0, 256, n * sz, dtype=np.uint8)


def main():
    import torch
    import torch.distributed as dist

    from covalent_ssh_plugin_b200 import _b2lib as L, distributed as D

    rank, world, dev = D.env_rank(), D.env_world(), D.env_local_rank()
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    arena = 1 << 30
    L.call("b2_init", dev, arena, 16 << 20, 16 << 20, 1 << 12)
    peers = D.exchange_peer_arenas(dev, arena)
    assert sorted(peers) == [r for r in range(world) if r != rank], peers

    n, sz = 2000, 65536
    mine = payload(rank, n, sz)
    L.call("b2_write", dev, 0, mine.ctypes.data, mine.size)
    dst_rank = (rank + 1) % world
    perm = np.random.default_rng([8, rank]).permutation(n)
    spans = D.spans_array([(int(perm[i]) * sz, (512 << 20) + i * sz, sz) for i in range(n)])

    ev = [C.c_uint32(), C.c_uint32()]
    for e in ev:
        L.call("b2_event_create", dev, C.byref(e))
    L.call("b2_peer_pack", dev, peers[dst_rank], spans.ctypes.data, n)      # warm-up + correctness
    L.call("b2_sync", dev)
    dist.barrier()
    L.call("b2_event_record", dev, ev[0].value)
    reps = 10
    for _ in range(reps):
        L.call("b2_peer_pack", dev, peers[dst_rank], spans.ctypes.data, n)
    L.call("b2_event_record", dev, ev[1].value)
    ms = C.c_float()
    L.call("b2_event_elapsed_ms", dev, ev[0].value, ev[1].value, C.byref(ms))
    dist.barrier()

    src_rank = (rank - 1) % world
    got = np.empty(n * sz, dtype=np.uint8)
    L.call("b2_read", dev, 512 << 20, got.ctypes.data, got.size)
    sent = payload(src_rank, n, sz).reshape(n, sz)
    sperm = np.random.default_rng([8, src_rank]).permutation(n)
    ok = np.array_equal(got.reshape(n, sz), sent[sperm])
    gbs = n * sz * reps / (ms.value / 1e3) / 1e9
    worst = D.max_over_ranks(ms.value / reps)
    print(f"rank {rank}: peer_pack -> rank {dst_rank}: {'OK' if ok else 'MISMATCH'}; "
          f"{gbs:.1f} GB/s per direction (kernel remote stores over NVLink), slowest rank {worst:.3f} ms", flush=True)
    dist.barrier()
    L.call("b2_shutdown", dev)

    # shared dependency: uploaded on rank 0, ncclBroadcast arena -> arena, then used as an
    # ordinary electron operand on every rank
    import asyncio
    import tempfile

    from covalent_ssh_plugin_b200 import SSHExecutor, electrons
    from covalent_ssh_plugin_b200.engine import get_engine, shutdown_all

    eng = get_engine(dev, arena_bytes=2 << 30, staging_bytes=256 << 20, n_segments=4, max_batch=256)
    shared_host = np.random.default_rng(99).standard_normal(1 << 20, dtype=np.float32)      # 4 MiB
    da = D.broadcast_shared(eng, shared_host if rank == 0 else None, root=0)
    tmp = tempfile.mkdtemp()
    key = Path(tmp) / "k"
    key.touch()
    ex = SSHExecutor(username="u", hostname=f"cuda:{dev}", ssh_key_file=str(key), cache_dir=tmp)
    x = np.random.default_rng([5, rank]).standard_normal(1 << 20, dtype=np.float32)
    y = asyncio.run(ex.run(electrons.vec_add, [x, da], {}, {"dispatch_id": "shared", "node_id": rank}))
    ok2 = np.array_equal(y, x + shared_host)
    print(f"rank {rank}: shared operand via ncclBroadcast + vec_add on cuda:{dev}: {'OK' if ok2 else 'MISMATCH'}", flush=True)
    dist.barrier()
    shutdown_all()
    dist.destroy_process_group()
    if not (ok and ok2):
        sys.exit(1)


if __name__ == "__main__":
    main()
