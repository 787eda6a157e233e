"""TEST INFRASTRUCTURE — numpy model of the CUDA kernels' byte-level contracts.

The kernels in covalent_ssh_plugin_b200/csrc have no counterpart in the
reference (which moves bytes with SCP, ssh.py:360-361/451, and computes with
whatever the electron's Python body does under numpy, exec.py:37).  Their
contracts are therefore restated here in numpy so that the GPU parity tests can
compare byte for byte on seeded inputs:

* ``move_spans`` / ``pack`` / ``unpack`` — what K1/K2/K4 (b2_spans.cuh) must
  leave in the arena: the transport is value-transparent, exactly like SCP.
* ``run_callback`` — what each built-in device callback (b2_worker.cuh) must
  write, using the same fp32 operation order as the registered electron's
  Python body (electrons.py), which is the code the reference would execute on
  the remote host.
* ``run_tasks`` — the worker's batch semantics: every electron runs once, after
  all of its predecessors.

Never imported by the product; tests/fake_b2lib.py builds a CPU stand-in of the
C ABI on top of it for the host-logic tests.
"""

from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

K_NONE, K_COPY, K_SCALE_ADD, K_ADD, K_MUL, K_AXPY, K_SUM, K_STACK_SUM, K_IADD, K_IMUL, K_ISCALE_ADD = range(11)
TS_PENDING, TS_OK, TS_OVERFLOW, TS_BAD_FN, TS_BAD_ARG, TS_CANCELLED = 0, 1, 2, 3, 4, 5
F_IMM0, F_IMM1, F_CANCELLED = 1, 2, 4
INT64_MIN, INT64_MAX = -(1 << 63), (1 << 63) - 1


def move_spans(arena: np.ndarray, spans: Sequence[Tuple[int, int, int]],
               dst_arena: np.ndarray = None) -> None:
    """arena[dst:dst+n] = arena[src:src+n] for every (src, dst, n)."""
    dst_arena = arena if dst_arena is None else dst_arena
    staged = [(d, arena[s:s + n].copy()) for s, d, n in spans]   # reads see the pre-move state
    for d, buf in staged:
        dst_arena[d:d + buf.size] = buf


def pack_layout(sizes: Sequence[int], base: int, align: int) -> Tuple[List[int], int]:
    """Contiguous-side offsets b2_pack/b2_unpack assign, and the run length."""
    offs, cur = [], base
    for n in sizes:
        offs.append(cur)
        cur += (int(n) + align - 1) & ~(align - 1)
    return offs, cur - base


def pack(arena: np.ndarray, src_offs, sizes, dst_off: int, align: int = 16) -> List[int]:
    offs, _ = pack_layout(sizes, dst_off, align)
    move_spans(arena, [(int(s), o, int(n)) for s, o, n in zip(src_offs, offs, sizes)])
    return offs


def unpack(arena: np.ndarray, src_off: int, dst_offs, sizes, align: int = 16) -> List[int]:
    offs, _ = pack_layout(sizes, src_off, align)
    move_spans(arena, [(o, int(d), int(n)) for o, d, n in zip(offs, dst_offs, sizes)])
    return offs


def _f32(arena: np.ndarray, off: int, count: int) -> np.ndarray:
    return arena[off:off + 4 * count].view(np.float32)


def _s64(v: int) -> int:
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >> 63 else v


def run_callback(arena: np.ndarray, fn_entry: dict, task: dict, peers: dict = None) -> int:
    """Execute one task record against ``arena`` (uint8); returns its status.
    ``peers`` maps arena ids (b2_task.flags bits 8..19) to other devices' arenas."""
    kind = fn_entry.get("kind", K_NONE)
    flags = int(task.get("flags", 0))
    ids = ((flags >> 8) & 15, (flags >> 12) & 15, (flags >> 16) & 15)
    if any(ids):
        peers = peers or {}
        if any(i and i not in peers for i in ids):
            return TS_BAD_ARG
        src0 = peers[ids[0]] if ids[0] else arena
        src1 = peers[ids[1]] if ids[1] else arena
        dst = peers[ids[2]] if ids[2] else arena
        # evaluate on a scratch arena laid out like the local one, then write through
        n_ = int(task["count"])
        scratch = arena
        t2 = dict(task)
        t2["flags"] = flags & 0xFF
        if ids[0] or ids[1]:
            esz = 1 if kind == K_COPY else 4
            span = n_ * esz + (int(task["nrows"]) - 1) * int(task["arg1"]) if kind == K_STACK_SUM else n_ * esz
            saved = []
            for which, srcx, off in ((0, src0, int(task["arg0"])), (1, src1, int(task["arg1"]))):
                if ids[which] and (which == 0 or kind in (K_ADD, K_MUL, K_AXPY)):
                    saved.append((off, arena[off:off + span].copy()))
                    arena[off:off + span] = srcx[off:off + span]
            st = run_callback(arena, fn_entry, t2 if not ids[2] else t2, None)
            out_bytes = 4 if kind == K_SUM else n_ * (1 if kind == K_COPY else 4)
            result = arena[int(task["out"]):int(task["out"]) + out_bytes].copy()
            for off, buf in reversed(saved):
                arena[off:off + buf.size] = buf
            dst[int(task["out"]):int(task["out"]) + out_bytes] = result
            return st
        st = run_callback(arena, fn_entry, t2, None)
        return st
    n = int(task["count"])
    out = int(task["out"])
    a0, a1 = int(task["arg0"]), int(task["arg1"])
    if kind == K_NONE:
        return TS_OK if task["fn_id"] == 0 else TS_BAD_FN
    if kind in (K_IADD, K_IMUL, K_ISCALE_ADD):
        x, y = _s64(a0), _s64(a1)
        in_range = lambda v: INT64_MIN <= v <= INT64_MAX  # noqa: E731
        if kind == K_IADD:
            r, ok = x + y, True
        elif kind == K_IMUL:
            r, ok = x * y, True
        else:
            p = int(fn_entry["ialpha"]) * x
            r, ok = p + int(fn_entry["ibeta"]), in_range(p)
        ok = ok and in_range(r)
        arena[out:out + 8].view(np.int64)[0] = np.int64(_s64(r))   # wrapped value; ignored on overflow
        return TS_OK if ok else TS_OVERFLOW
    if (a0 | out) & 15:
        return TS_BAD_ARG
    if kind == K_COPY:
        arena[out:out + n] = arena[a0:a0 + n].copy()
        return TS_OK
    alpha = np.float32(fn_entry.get("alpha", 0.0))
    beta = np.float32(fn_entry.get("beta", 0.0))
    x = _f32(arena, a0, n).copy()
    if kind == K_SCALE_ADD:
        _f32(arena, out, n)[:] = alpha * x + beta          # two fp32 roundings, as numpy does
        return TS_OK
    if kind == K_SUM:
        # the device accumulates in double and rounds once; compare with tolerance
        arena[out:out + 4].view(np.float32)[0] = np.float32(np.sum(x, dtype=np.float64))
        return TS_OK
    if kind == K_STACK_SUM:
        stride = a1
        if stride & 15 or int(task["nrows"]) == 0:
            return TS_BAD_ARG
        acc = x
        for r in range(1, int(task["nrows"])):
            acc = acc + _f32(arena, a0 + r * stride, n)     # row order, fp32
        _f32(arena, out, n)[:] = acc
        return TS_OK
    if a1 & 15:
        return TS_BAD_ARG
    y = _f32(arena, a1, n).copy()
    if kind == K_ADD:
        _f32(arena, out, n)[:] = x + y
    elif kind == K_MUL:
        _f32(arena, out, n)[:] = x * y
    elif kind == K_AXPY:
        _f32(arena, out, n)[:] = alpha * x + y
    else:
        return TS_BAD_FN
    return TS_OK


def run_tasks(arena: np.ndarray, fn_table: dict, tasks: Sequence[dict], succ: Sequence[int],
              peers: dict = None) -> List[int]:
    """Batch semantics of the worker: Kahn order over the successor lists."""
    n = len(tasks)
    deps = [int(t["deps"]) for t in tasks]
    status = [TS_PENDING] * n
    ready = [i for i in range(n) if deps[i] == 0]
    while ready:
        i = ready.pop(0)
        t = tasks[i]
        entry = fn_table.get(int(t["fn_id"]), {"kind": K_NONE})
        if int(t["flags"]) & F_CANCELLED:      # tombstone: nothing runs, successors are still released
            status[i] = TS_CANCELLED
        else:
            status[i] = run_callback(arena, entry, t, peers)
        for e in range(int(t["succ_count"])):
            j = int(succ[int(t["succ_begin"]) + e])
            deps[j] -= 1
            if deps[j] == 0:
                ready.append(j)
    return status
