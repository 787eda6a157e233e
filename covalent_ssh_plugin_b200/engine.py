"""Per-GPU engine: batching, staging and completion on top of the C ABI.

One ``DeviceEngine`` exists per (process, CUDA device) and lives in a module
registry — never on an executor instance, whose attributes must stay plain
data.  It plays the role of the reference's SSH connection
(/root/reference/covalent_ssh_plugin/ssh.py:237-282) plus the remote cache
directory (ssh.py:528-532): electrons are staged into pinned host memory,
shipped in batches (one H2D, one worker launch, one D2H per batch instead of
three SCP sessions and a remote interpreter per electron), and completed
through a future the submitting coroutine awaits.

Arena layout on the device (all offsets in bytes):

    [0, S_in)             ingress mirror: pinned ingress offset X <-> arena offset X
    [S_in, S_in + S_out)  egress mirror:  pinned egress offset Y  <-> arena offset S_in + Y
    [S_in + S_out, A)     heap: device-resident slots (lattice intermediates, pack scratch)

Both staging buffers are cut into ``n_segments`` equal segments; a batch owns
one ingress and one egress segment from first staging until its last result
has been read, so a batch is always one contiguous copy each way.
"""

from __future__ import annotations

import asyncio
import ctypes as C
import os
import queue
import struct
import threading
from typing import Any, Dict, List, Optional

import numpy as np

from . import _b2lib as L
from .electrons import DeviceElectron, EncodedTask, decode, registry_size, registry_snapshot

_TASK = struct.Struct("<IIQQQQIIIIQ")
assert _TASK.size == 64

ALIGN = 256                      # payload alignment inside a segment
MiB = 1 << 20

DEFAULTS = {
    "arena_bytes": int(os.environ.get("B2_ARENA_BYTES", 4096 * MiB)),
    "staging_bytes": int(os.environ.get("B2_STAGING_BYTES", 1024 * MiB)),
    "n_segments": int(os.environ.get("B2_SEGMENTS", 8)),
    "max_batch": int(os.environ.get("B2_MAX_BATCH", 1024)),
    "queue_cap": int(os.environ.get("B2_QUEUE_CAP", 1 << 16)),
}


def _round_up(n: int, a: int = ALIGN) -> int:
    return (n + a - 1) & ~(a - 1)


class DeviceTaskError(RuntimeError):
    """The worker kernel reported a non-OK status for an electron."""

    def __init__(self, status: int, what: str):
        names = {L.TS_OVERFLOW: "int64 overflow", L.TS_BAD_FN: "no callback registered",
                 L.TS_BAD_ARG: "operand outside arena or misaligned", L.TS_PENDING: "never ran"}
        super().__init__(f"{what}: device status {status} ({names.get(status, 'unknown')})")
        self.status = status


class Batch:
    """Electrons staged together: one segment pair, one launch, one future."""

    __slots__ = ("engine", "seg", "n", "in_used", "out_used", "table", "future", "loop",
                 "ticket", "unread", "done", "flushed", "error")

    def __init__(self, engine: "DeviceEngine", seg: int, loop):
        self.engine = engine
        self.seg = seg
        self.n = 0
        self.in_used = 0
        self.out_used = 0
        self.table = engine._tables[seg]
        self.loop = loop
        self.future = loop.create_future() if loop is not None else None
        self.ticket = 0
        self.unread = 0
        self.done = False
        self.flushed = False
        self.error: Optional[BaseException] = None


class Staged:
    """Handle of one staged electron (what the executor keeps between steps)."""

    __slots__ = ("batch", "index", "out_off", "out_bytes", "meta", "spec")

    def __init__(self, batch: Batch, index: int, out_off: int, out_bytes: int, meta: tuple,
                 spec: DeviceElectron):
        self.batch = batch
        self.index = index
        self.out_off = out_off
        self.out_bytes = out_bytes
        self.meta = meta
        self.spec = spec


class DeviceEngine:
    def __init__(self, ordinal: int, **cfg):
        c = dict(DEFAULTS)
        c.update({k: v for k, v in cfg.items() if v is not None})
        self.ordinal = ordinal
        self.n_segments = int(c["n_segments"])
        self.max_batch = int(c["max_batch"])
        self.queue_cap = int(c["queue_cap"])
        if self.max_batch > self.queue_cap:
            self.max_batch = self.queue_cap
        staging = int(c["staging_bytes"])
        # each egress segment ends with max_batch status words
        self.seg_bytes = (staging // self.n_segments) & ~(ALIGN - 1)
        self.status_bytes = _round_up(self.max_batch * 4)
        if self.seg_bytes <= self.status_bytes + ALIGN:
            raise ValueError("staging_bytes too small for n_segments/max_batch")
        self.staging_in = self.seg_bytes * self.n_segments
        self.staging_out = self.seg_bytes * self.n_segments
        self.arena_bytes = max(int(c["arena_bytes"]), self.staging_in + self.staging_out + 64 * MiB)
        self.arena_bytes = _round_up(self.arena_bytes)
        self.egress_base = self.staging_in                      # arena offset of egress mirror
        self.heap_base = self.staging_in + self.staging_out
        self.heap_bytes = self.arena_bytes - self.heap_base

        L.call("b2_init", ordinal, self.arena_bytes, self.staging_in, self.staging_out, self.queue_cap)
        self.info = L.DeviceInfo()
        L.call("b2_get_info", ordinal, C.byref(self.info))
        self._h_in = self._host_view(0, self.staging_in)
        self._h_out = self._host_view(1, self.staging_out)
        self._tables = [bytearray(64 * self.max_batch) for _ in range(self.n_segments)]
        self._free_segments = list(range(self.n_segments - 1, -1, -1))
        self._segment_waiters: List[asyncio.Future] = []
        self._open: Optional[Batch] = None
        self._inflight: List[Batch] = []
        self._registered = 0
        self._lock = threading.RLock()
        self._waitq: "queue.SimpleQueue[Optional[Batch]]" = queue.SimpleQueue()
        self._waiter = threading.Thread(target=self._wait_loop, name=f"b2-wait-{ordinal}", daemon=True)
        self._waiter.start()
        self.counters = {"electrons": 0, "batches": 0, "bytes_in": 0, "bytes_out": 0}
        self.closed = False
        self.sync_registry()

    # ------------------------------------------------------------- plumbing --
    def _host_view(self, which: int, nbytes: int) -> np.ndarray:
        ptr = C.c_void_p()
        size = C.c_uint64()
        L.call("b2_host_buffer", self.ordinal, which, C.byref(ptr), C.byref(size))
        if size.value < nbytes or not ptr.value:
            raise RuntimeError("pinned staging buffer smaller than requested")
        buf = (C.c_uint8 * nbytes).from_address(ptr.value)
        return np.frombuffer(buf, dtype=np.uint8)

    def sync_registry(self) -> None:
        """Push callback-table entries registered since the last call (b2_register)."""
        if self._registered == registry_size():
            return
        with self._lock:
            specs = registry_snapshot()
            for spec in specs[self._registered:]:
                L.call("b2_register", self.ordinal, spec.fn_id, spec.kind, spec.alpha, spec.beta,
                       spec.ialpha, spec.ibeta)
            self._registered = len(specs)

    def stats(self) -> L.Stats:
        s = L.Stats()
        L.call("b2_get_stats", self.ordinal, C.byref(s))
        return s

    # --------------------------------------------------------------- staging --
    def fits(self, enc: EncodedTask) -> bool:
        need_in = sum(_round_up(p.size) for p in enc.payloads)
        need_out = _round_up(max(enc.out_bytes, 16))
        return need_in <= self.seg_bytes and need_out <= self.seg_bytes - self.status_bytes

    async def stage(self, enc: EncodedTask) -> Staged:
        """Copy one electron's operands into pinned ingress memory and append
        its task record to the open batch (ssh.py:147-150 + ssh.py:360-361)."""
        need_in = sum(_round_up(p.size) for p in enc.payloads)
        need_out = _round_up(max(enc.out_bytes, 16))
        if not self.fits(enc):
            raise ValueError("electron payload larger than one staging segment")
        loop = asyncio.get_running_loop()
        while True:
            b = self._open
            if b is not None and (b.loop is not loop or b.n >= self.max_batch
                                  or b.in_used + need_in > self.seg_bytes
                                  or b.out_used + need_out > self.seg_bytes - self.status_bytes):
                self.flush()
                b = None
            if b is None:
                if not self._free_segments:
                    self._reap()
                if not self._free_segments:
                    fut = loop.create_future()
                    self._segment_waiters.append(fut)
                    await fut
                    continue
                b = Batch(self, self._free_segments.pop(), loop)
                self._open = b
                loop.call_soon(self._idle_flush, b)
            break
        return self._append(b, enc)

    def _append(self, b: Batch, enc: EncodedTask) -> Staged:
        spec = enc.spec
        if spec.fn_id > self._registered:
            self.sync_registry()
        seg_base = b.seg * self.seg_bytes
        offs = [0, 0]
        flags = 0
        if enc.payloads:
            if spec.kind == L.K_STACK_SUM:
                # rows are staged back to back at a fixed stride
                stride = _round_up(enc.payloads[0].size if len(enc.payloads) > 1
                                   else enc.payloads[0].size // enc.nrows, 16)
                base = seg_base + b.in_used
                if len(enc.payloads) == 1:
                    if stride * enc.nrows != enc.payloads[0].size:
                        rows = enc.payloads[0].reshape(enc.nrows, -1)
                        for r in range(enc.nrows):
                            o = base + r * stride
                            self._h_in[o:o + rows.shape[1]] = rows[r]
                    else:
                        self._h_in[base:base + enc.payloads[0].size] = enc.payloads[0]
                else:
                    for r, p in enumerate(enc.payloads):
                        o = base + r * stride
                        self._h_in[o:o + p.size] = p
                b.in_used += _round_up(stride * enc.nrows)
                offs = [base, stride]
            else:
                for i, p in enumerate(enc.payloads):
                    o = seg_base + b.in_used
                    self._h_in[o:o + p.size] = p
                    offs[i] = o
                    b.in_used += _round_up(p.size)
        else:
            offs = [enc.imm[0] & 0xFFFFFFFFFFFFFFFF, enc.imm[1] & 0xFFFFFFFFFFFFFFFF]
            flags = L.F_IMM0 | L.F_IMM1
        out_off = seg_base + b.out_used                      # egress-relative
        b.out_used += _round_up(max(enc.out_bytes, 16))
        idx = b.n
        _TASK.pack_into(b.table, idx * 64, spec.fn_id, flags, offs[0], offs[1],
                        self.egress_base + out_off, enc.count, enc.nrows, 0, 0, 0, 0)
        b.n = idx + 1
        b.unread += 1
        return Staged(b, idx, out_off, enc.out_bytes, enc.result_meta, spec)

    def _idle_flush(self, b: Batch) -> None:
        if self._open is b and not b.flushed:
            self.flush()

    def flush(self) -> Optional[Batch]:
        """Ship the open batch: H2D, worker launch, D2H, completion ticket."""
        b = self._open
        if b is None:
            return None
        self._open = None
        if b.n == 0:
            self._release_segment(b.seg)
            return None
        b.flushed = True
        dev = self.ordinal
        seg_base = b.seg * self.seg_bytes
        status_off = seg_base + self.seg_bytes - self.status_bytes
        try:
            if b.in_used:
                L.call("b2_h2d", dev, seg_base, seg_base, b.in_used)
            tab = (C.c_char * (64 * b.n)).from_buffer(b.table)
            L.call("b2_enqueue", dev, C.addressof(tab), b.n, None, 0, None)
            t = C.c_uint64()
            L.call("b2_launch", dev, status_off, C.byref(t))
            L.call("b2_d2h", dev, self.egress_base + seg_base, seg_base, b.out_used)
            L.call("b2_mark", dev, C.byref(t))
            b.ticket = t.value
        except BaseException as err:  # surface on the awaiting coroutines
            b.error = err
            self._complete(b)
            return b
        self.counters["batches"] += 1
        self.counters["electrons"] += b.n
        self.counters["bytes_in"] += b.in_used
        self.counters["bytes_out"] += b.out_used
        self._inflight.append(b)
        self._waitq.put(b)
        self._reap()
        return b

    # ------------------------------------------------------------ completion --
    def _reap(self) -> None:
        """Complete finished batches without leaving the event-loop thread."""
        while self._inflight:
            b = self._inflight[0]
            if not b.done:
                rc = L.load().b2_poll(self.ordinal, b.ticket)
                if rc == L.B2_E_NOTREADY:
                    return
                if rc < 0:
                    b.error = L.B2Error(rc, L.last_error(), "b2_poll")
                self._complete(b)
            self._inflight.pop(0)

    def _wait_loop(self) -> None:
        while True:
            b = self._waitq.get()
            if b is None:
                return
            if b.done:
                continue
            try:
                L.call("b2_wait", self.ordinal, b.ticket)
            except BaseException as err:
                b.error = err
            try:
                b.loop.call_soon_threadsafe(self._complete, b)
            except RuntimeError:      # loop already closed
                self._complete(b)

    def _complete(self, b: Batch) -> None:
        if b.done:
            return
        b.done = True
        fut = b.future
        if fut is not None and not fut.done():
            if b.error is not None:
                fut.set_exception(b.error)
            else:
                fut.set_result(b)

    def status_of(self, s: Staged) -> int:
        b = s.batch
        off = b.seg * self.seg_bytes + self.seg_bytes - self.status_bytes + 4 * s.index
        return int(self._h_out[off:off + 4].view(np.uint32)[0])

    def result_of(self, s: Staged) -> Any:
        """Decode one result out of pinned egress memory (ssh.py:451-456)."""
        return decode(s.meta, self._h_out[s.out_off:s.out_off + max(s.out_bytes, 16)])

    def release(self, s: Staged) -> None:
        """The caller is done with this electron's staging (ssh.py:284-315)."""
        b = s.batch
        b.unread -= 1
        if b.unread == 0 and b.done:
            self._release_segment(b.seg)

    def _release_segment(self, seg: int) -> None:
        self._free_segments.append(seg)
        while self._segment_waiters:
            fut = self._segment_waiters.pop(0)
            if not fut.done():
                fut.set_result(None)
                break

    # ----------------------------------------------------------------- close --
    def close(self) -> None:
        if self.closed:
            return
        self.closed = True
        self._waitq.put(None)
        try:
            L.call("b2_sync", self.ordinal)
        finally:
            L.call("b2_shutdown", self.ordinal)


# ---------------------------------------------------------------------------
_engines: Dict[int, DeviceEngine] = {}
_engines_lock = threading.Lock()


def get_engine(ordinal: int, **cfg) -> DeviceEngine:
    """The process-wide engine of one GPU (created on first use)."""
    with _engines_lock:
        eng = _engines.get(ordinal)
        if eng is None or eng.closed:
            eng = DeviceEngine(ordinal, **cfg)
            _engines[ordinal] = eng
        return eng


def existing_engine(ordinal: int) -> Optional[DeviceEngine]:
    eng = _engines.get(ordinal)
    return None if eng is None or eng.closed else eng


def shutdown_all() -> None:
    with _engines_lock:
        for eng in list(_engines.values()):
            try:
                eng.close()
            except Exception:
                pass
        _engines.clear()


def device_count() -> int:
    n = C.c_int()
    L.call("b2_device_count", C.byref(n))
    return n.value
