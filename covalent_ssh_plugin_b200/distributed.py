"""One process per GPU: placement, peer-arena exchange and timing reductions.

The hot path shards by electron: electrons handed to ``run``
are independent, so rank r of W simply owns electrons ``i % W == r`` (the
"round-robin" of BASELINE.json config 3) and no collective touches the data
path.  ``torch.distributed`` is plumbing only: rendezvous, the barrier around
timed regions, the max-over-ranks of a timing, and the one-off all-gather of
CUDA IPC handles that lets a rank's kernels address its peers' arenas over
NVLink (K4, ``b2_peer_pack`` / ``b2_peer_copy``).  Works on ``gloo`` (CPU tests)
and ``nccl``.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Sequence

import numpy as np


def env_rank() -> int:
    return int(os.environ.get("RANK", "0"))


def env_world() -> int:
    return int(os.environ.get("WORLD_SIZE", "1"))


def env_local_rank() -> int:
    return int(os.environ.get("LOCAL_RANK", "0"))


def owner_of(node_id: int, world: int) -> int:
    """Rank that runs electron ``node_id`` under round-robin placement."""
    return node_id % world


def shard(n_electrons: int, rank: int, world: int) -> range:
    """Electron ids owned by ``rank``: i with i % world == rank."""
    return range(rank, n_electrons, world)


def shard_sizes(n_electrons: int, world: int) -> List[int]:
    return [len(shard(n_electrons, r, world)) for r in range(world)]


def crossing_edges(edges: Sequence[tuple], world: int) -> List[tuple]:
    """Dependency edges (producer, consumer) whose endpoints live on different
    ranks: these are the payloads K4 has to route over NVLink."""
    return [(p, c) for p, c in edges if owner_of(p, world) != owner_of(c, world)]


def _tensor_device(dist):
    import torch

    if dist.get_backend() == "nccl":
        return torch.device("cuda", env_local_rank())
    return torch.device("cpu")


def max_over_ranks(value: float) -> float:
    """Max of a host scalar over all ranks (the timing rule for multi-GPU runs)."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_tensor_device(dist))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_gather_bytes(blob: bytes) -> List[bytes]:
    """Every rank's fixed-size blob, in rank order."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [blob]
    dev = _tensor_device(dist)
    mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    out = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return [bytes(t.cpu().numpy().tobytes()) for t in out]


def exchange_peer_arenas(ordinal: int, arena_bytes: int) -> Dict[int, int]:
    """Map every other rank's arena into this rank's engine.

    Each rank exports its arena as a CUDA IPC handle (``b2_ipc_export``), the 72
    byte records (handle + arena size) are all-gathered, and every remote one is
    opened with ``b2_ipc_open``.  Returns {peer rank: arena id} for use in
    ``b2_task.flags`` / ``b2_peer_pack``.
    """
    import torch.distributed as dist

    from . import _b2lib as L

    handle = (C.c_uint8 * 64)()
    L.call("b2_ipc_export", ordinal, handle)
    record = bytes(handle) + int(arena_bytes).to_bytes(8, "little")
    records = all_gather_bytes(record)
    me = dist.get_rank() if dist.is_initialized() else 0
    arena_ids: Dict[int, int] = {}
    for r, rec in enumerate(records):
        if r == me:
            continue
        buf = (C.c_uint8 * 64).from_buffer_copy(rec[:64])
        aid = C.c_uint32()
        L.call("b2_ipc_open", ordinal, buf, int.from_bytes(rec[64:72], "little"), C.byref(aid))
        arena_ids[r] = aid.value
    return arena_ids


class _ArenaWindow:
    """Exposes a byte range of an arena through __cuda_array_interface__ so that
    torch can wrap it without copying."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                         "version": 2}


def broadcast_shared(engine, array, root: int = 0, stats: dict = None):
    """A shared dependency — the same operand for electrons on every rank — is
    uploaded once on ``root`` and broadcast into every rank's arena (the only
    collective on this path).  Returns a DeviceArray that can be
    passed as an electron argument.  ``array`` is only read on ``root``.  ``stats`` (optional
    dict) receives ``collective_seconds``, the ncclBroadcast alone.

    NCCL backend: ncclBroadcast directly between arenas (device to device over
    NVLink).  Other backends (gloo in the CPU tests): broadcast of the host bytes,
    then a local upload.
    """
    import torch
    import torch.distributed as dist

    single = not dist.is_initialized() or dist.get_world_size() == 1
    me = 0 if single else dist.get_rank()
    if me == root:
        a = np.ascontiguousarray(array, dtype=np.float32)
        meta = [tuple(a.shape), int(a.nbytes)]
    else:
        a, meta = None, [None, None]
    if not single:
        dist.broadcast_object_list(meta, src=root)
    shape, nbytes = tuple(meta[0]), int(meta[1])
    if single:
        return engine.put(a)
    if dist.get_backend() == "nccl":
        from . import _b2lib as L

        da = engine.put(a) if me == root else engine.alloc(nbytes, shape)
        ptr, size = C.c_void_p(), C.c_uint64()
        L.call("b2_arena_base", engine.ordinal, 0, C.byref(ptr), C.byref(size))
        L.call("b2_sync", engine.ordinal)
        window = _ArenaWindow(ptr.value + da.offset, nbytes)
        t = torch.as_tensor(window, device=torch.device("cuda", engine.ordinal))
        import time

        t0 = time.perf_counter()
        dist.broadcast(t, src=root)
        torch.cuda.synchronize(engine.ordinal)
        if stats is not None:
            stats["collective_seconds"] = time.perf_counter() - t0
        return da
    buf = torch.from_numpy(a.reshape(-1).view(np.uint8).copy()) if me == root else torch.empty(nbytes, dtype=torch.uint8)
    dist.broadcast(buf, src=root)
    return engine.put(buf.numpy().view(np.float32).reshape(shape))


def spans_array(rows: Sequence[tuple]) -> np.ndarray:
    """(src_off, dst_off, bytes) rows as the uint64 array the C ABI takes."""
    return np.ascontiguousarray(np.array(rows, dtype=np.uint64).reshape(-1, 3))
