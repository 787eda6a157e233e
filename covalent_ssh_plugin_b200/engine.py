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
import collections
import ctypes as C
import logging
import os
import queue
import struct
import threading
import weakref
from typing import Any, Dict, List, Optional

import numpy as np

from . import _b2lib as L
from .electrons import DeviceElectron, DeviceRef, EncodedTask, decode, registry_size, registry_snapshot

app_log = logging.getLogger(__name__)

BACKPRESSURE = False               # stage_fast: eligible call, but no staging segment pair is free

_b2host = None                     # compiled staging helper (hostext/b2host.c), loaded with the first engine


def _load_b2host():
    """Import the CPython staging helper.  Deferred to engine creation so that importing the package —
    which `python -m covalent_ssh_plugin_b200.build` itself has to do — works in a tree that has not been
    built yet.  Never a silent fallback: without the helper there is no engine."""
    global _b2host
    if _b2host is None:
        try:
            from . import _b2host as helper
        except ImportError as err:
            raise ImportError(
                "covalent_ssh_plugin_b200._b2host is not built: run `python -m covalent_ssh_plugin_b200.build` "
                f"({err})") from err
        if helper.IDX_BITS != IDX_BITS or helper.TASK_BYTES != 64 or helper.JOB_BYTES != 24:
            raise ImportError("covalent_ssh_plugin_b200._b2host does not match this source tree: rebuild it")
        _b2host = helper
    return _b2host

_TASK = struct.Struct("<IIQQQQIIIIQ")
assert _TASK.size == 64

ALIGN = 256                      # payload alignment inside a segment
SPLIT_BYTES = 1 << 20            # large elementwise electrons are sliced into tasks of this size
_ELEMENTWISE = (L.K_COPY, L.K_SCALE_ADD, L.K_ADD, L.K_MUL, L.K_AXPY)
_FAST1 = (L.K_COPY, L.K_SCALE_ADD, L.K_SUM)        # kinds the C stager takes in one call
_FAST2 = (L.K_ADD, L.K_MUL, L.K_AXPY)
_FASTI = (L.K_IADD, L.K_IMUL)
IDX_BITS = 20                     # = B2H_IDX_BITS in hostext/b2host.c (checked when the helper loads)
IDX_MASK = (1 << IDX_BITS) - 1
MiB = 1 << 20

DEFAULTS = {
    "arena_bytes": int(os.environ.get("B2_ARENA_BYTES", 4096 * MiB)),
    "staging_bytes": int(os.environ.get("B2_STAGING_BYTES", 1024 * MiB)),
    "n_segments": int(os.environ.get("B2_SEGMENTS", 32)),      # 32 MiB per batch each way by default
    "n_out_segments": int(os.environ.get("B2_OUT_SEGMENTS", 0)) or None,   # default: same as n_segments
    "max_batch": int(os.environ.get("B2_MAX_BATCH", 4096)),
    "queue_cap": int(os.environ.get("B2_QUEUE_CAP", 1 << 16)),
    "zero_copy": os.environ.get("B2_ZERO_COPY", "1") != "0",
    # resident worker (b2_resident_enable): the worker kernel outlives its batches; batches are published
    # through a doorbell instead of a launch.  Off by default: at the batch sizes run() produces it is on a
    # par with launching (profiles/r02_notes.md), so the simpler mode stays the default.
    "resident": os.environ.get("B2_RESIDENT", "0") == "1",
    "gc_tune": os.environ.get("B2_GC_TUNE", "1") != "0",
    # general path only: operands of at least this many bytes become copy jobs for the
    # library's copy threads (one host thread moves ~9-11 GB/s, 8-16 threads 25-40 GB/s,
    # tools/staging_probe.py on the B200 host); smaller ones are copied on the spot.  The
    # fast path (stage_fast) always queues jobs.  Thread count: B2_STAGE_THREADS (library).
    "parallel_stage_bytes": int(os.environ.get("B2_PARALLEL_STAGE_BYTES", 64 * 1024)),
}

_gc_tuned = False


def tune_gc() -> None:
    """One-time cyclic-GC settings for a process that pushes tens of thousands of
    electrons per second through asyncio.

    Every ``run()`` allocates a task, a coroutine frame, a future callback and a
    handful of small records; with CPython's default thresholds that is a
    generation-0 collection every 700 allocations, and the periodic full
    collections walk every tracked object in the process — ~150 000 more of
    them once ``torch`` has been imported for the multi-GPU plumbing.  Measured
    on the B200 host: 277 -> 352 ms per 10 000 electrons with torch imported,
    all of it GC.  ``gc.freeze()`` parks what exists now in the permanent
    generation (it is never walked again) and the young-generation threshold is
    raised so collections happen per ~100 000 allocations.  Reference counting
    is untouched, so ordinary objects are freed exactly as before; only
    reference *cycles* wait a little longer.  Set B2_GC_TUNE=0 to opt out."""
    global _gc_tuned
    if _gc_tuned:
        return
    _gc_tuned = True
    import gc

    gc.collect()
    gc.freeze()
    g0, g1, g2 = gc.get_threshold()
    gc.set_threshold(max(g0, 100_000), max(g1, 50), max(g2, 100))


def _round_up(n: int, a: int = ALIGN) -> int:
    return (n + a - 1) & ~(a - 1)


class Batch:
    """Electrons staged together: one ingress segment, one egress segment, one
    b2_submit (copies, H2D, worker launch, D2H), one completion ticket."""

    __slots__ = ("engine", "seg", "oseg", "n", "n_jobs", "in_used", "out_used", "table", "table_mem", "loop",
                 "ticket", "unread", "done", "flushed", "error", "lease", "zero_copy", "out_freed",
                 "waiters", "keep", "all_ok", "f32", "in_base", "out_base")

    def __init__(self, engine: "DeviceEngine", seg: int, oseg: int, loop):
        self.engine = engine
        self.seg = seg
        self.oseg = oseg
        self.n = 0                  # task records (while the batch is open the stager holds the truth)
        self.n_jobs = 0
        self.in_used = 0
        self.out_used = 0
        self.table = engine._tables[seg]
        self.table_mem = engine._table_mem[seg]
        self.in_base = seg * engine.seg_bytes
        self.out_base = oseg * engine.seg_bytes
        self.loop = loop
        self.ticket = 0
        self.unread = 0
        self.done = False
        self.flushed = False
        self.error: Optional[BaseException] = None
        self.lease = None           # ctypes view of the egress segment that result arrays alias
        self.zero_copy = False
        self.out_freed = False
        self.waiters: list = []     # one future per electron awaiting this batch
        self.keep: list = []        # operand objects the queued copy jobs still read from
        self.all_ok = False         # every status word of the batch is B2_TS_OK (set on completion)
        self.f32 = None             # float32 view of the egress segment (zero-copy results are slices of it)


class Staged:
    """Handle of one staged electron (what the executor keeps between steps)."""

    __slots__ = ("batch", "index", "out_off", "out_bytes", "meta", "spec", "parts", "cancelled")

    def __init__(self, batch: Batch, index: int, out_off: int, out_bytes: int, meta: tuple,
                 spec: DeviceElectron):
        self.batch = batch
        self.index = index
        self.parts = 1                  # task records this electron occupies (large ones are sliced)
        self.out_off = out_off          # byte offset inside the batch's egress segment
        self.out_bytes = out_bytes
        self.meta = meta
        self.spec = spec
        self.cancelled = False


class Heap:
    """First-fit allocator over the arena heap [base, base + size), 256 B granules."""

    def __init__(self, base: int, size: int):
        self.base, self.size = base, size
        self.free = [(base, size)]            # sorted, coalesced (offset, bytes)
        self.used: Dict[int, int] = {}
        self.lock = threading.Lock()

    def alloc(self, nbytes: int) -> int:
        need = _round_up(max(nbytes, 1))
        with self.lock:
            for i, (off, sz) in enumerate(self.free):
                if sz >= need:
                    if sz == need:
                        self.free.pop(i)
                    else:
                        self.free[i] = (off + need, sz - need)
                    self.used[off] = need
                    return off
        raise MemoryError(f"arena heap exhausted: need {need} bytes (raise B2_ARENA_BYTES)")

    def release(self, off: int) -> None:
        with self.lock:
            sz = self.used.pop(off, None)
            if sz is None:
                return
            import bisect

            i = bisect.bisect_left(self.free, (off, 0))
            self.free.insert(i, (off, sz))
            if i + 1 < len(self.free) and self.free[i][0] + self.free[i][1] == self.free[i + 1][0]:
                self.free[i] = (self.free[i][0], self.free[i][1] + self.free.pop(i + 1)[1])
            if i > 0 and self.free[i - 1][0] + self.free[i - 1][1] == self.free[i][0]:
                self.free[i - 1] = (self.free[i - 1][0], self.free[i - 1][1] + self.free.pop(i)[1])

    def bytes_in_use(self) -> int:
        with self.lock:
            return sum(self.used.values())


class DeviceArray:
    """A float32 array that lives in a GPU's arena (the device-resident form of
    an electron result).  ``view(offset_bytes, shape)`` makes non-owning windows,
    e.g. the rows of a fan-in block."""

    __slots__ = ("engine", "offset", "nbytes", "shape", "order", "owner", "alive", "__weakref__")
    __b200_device_array__ = True

    def __init__(self, engine: "DeviceEngine", offset: int, nbytes: int, shape: tuple, order: str = "C",
                 owner: Optional["DeviceArray"] = None):
        self.engine, self.offset, self.nbytes = engine, offset, nbytes
        self.shape, self.order, self.owner = shape, order, owner
        self.alive = True

    @property
    def ordinal(self) -> int:
        return self.engine.ordinal

    def view(self, byte_offset: int, nbytes: int, shape: tuple) -> "DeviceArray":
        return DeviceArray(self.engine, self.offset + byte_offset, nbytes, shape, "C", owner=self.owner or self)

    def numpy(self) -> np.ndarray:
        return self.engine.get(self)

    def free(self) -> None:
        if self.alive and self.owner is None and not self.engine.closed:
            self.engine.heap.release(self.offset)
        self.alive = False

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceEngine:
    def __init__(self, ordinal: int, **cfg):
        c = dict(DEFAULTS)
        c.update({k: v for k, v in cfg.items() if v is not None})
        self.ordinal = ordinal
        self.n_segments = int(c["n_segments"])
        self.n_out_segments = int(c["n_out_segments"] or self.n_segments)
        self.max_batch = int(c["max_batch"])
        self.queue_cap = int(c["queue_cap"])
        self.zero_copy = bool(c["zero_copy"])
        self.parallel_stage_bytes = int(c["parallel_stage_bytes"])
        if self.max_batch > self.queue_cap:
            self.max_batch = self.queue_cap
        staging = int(c["staging_bytes"])
        self.seg_bytes = (staging // self.n_segments) & ~(ALIGN - 1)
        # every egress segment ends with max_batch status words
        self.status_bytes = _round_up(self.max_batch * 4)
        self.out_data_bytes = self.seg_bytes - self.status_bytes
        if self.out_data_bytes <= ALIGN:
            raise ValueError("staging_bytes too small for n_segments/max_batch")
        self.staging_in = self.seg_bytes * self.n_segments
        self.ctrl_off = self.seg_bytes * self.n_out_segments       # status words of DAG launches
        self.staging_out = self.ctrl_off + _round_up(self.queue_cap * 4)
        self.arena_bytes = max(int(c["arena_bytes"]), self.staging_in + self.staging_out + 64 * MiB)
        self.arena_bytes = _round_up(self.arena_bytes)
        self.egress_base = self.staging_in                      # arena offset of egress mirror
        self.heap_base = self.staging_in + self.staging_out
        self.heap_bytes = self.arena_bytes - self.heap_base

        L.call("b2_init", ordinal, self.arena_bytes, self.staging_in, self.staging_out, self.queue_cap)
        self.resident = bool(c.get("resident"))
        if self.resident:
            L.call("b2_resident_enable", ordinal, 0)
        self.info = L.DeviceInfo()
        L.call("b2_get_info", ordinal, C.byref(self.info))
        self._h_in, _ = self._host_view(0, self.staging_in)
        self._h_out, self._h_out_addr = self._host_view(1, self.staging_out)
        self._h_out_f32 = self._h_out[:self.staging_out & ~3].view(np.float32)
        # per ingress segment: the batch's task table, copy-job list and b2_batch descriptor.
        # They belong to the batch that owns the segment and must stay valid until its ticket
        # completes — which is exactly when the segment returns to the pool.
        self._tables = [np.zeros(self.max_batch, dtype=L.TASK_DTYPE) for _ in range(self.n_segments)]
        self._table_mem = [memoryview(t).cast("B") for t in self._tables]
        self._tables_addr = [t.ctypes.data for t in self._tables]
        self._cap_jobs = 2 * self.max_batch + 64
        self._jobs = [np.zeros(self._cap_jobs, dtype=L.JOB_DTYPE) for _ in range(self.n_segments)]
        self._jobs_addr = [j.ctypes.data for j in self._jobs]
        self._descs = [L.Batch() for _ in range(self.n_segments)]
        self._stager = _load_b2host().Stager(np.ndarray)
        self._free_segments = list(range(self.n_segments - 1, -1, -1))
        self._free_out = list(range(self.n_out_segments - 1, -1, -1))
        self._seg_lock = threading.Lock()          # lease finalizers may run on any thread
        self._segment_waiters: collections.deque = collections.deque()     # (loop, future) of parked stagers
        self._leases_out = 0                       # egress segments currently aliased by results
        self._open: Optional[Batch] = None
        self._inflight: List[Batch] = []
        self._registered = 0
        self._lock = threading.RLock()
        self._waitq: "queue.SimpleQueue[Optional[Batch]]" = queue.SimpleQueue()
        self._waiter = threading.Thread(target=self._wait_loop, name=f"b2-wait-{ordinal}", daemon=True)
        self._waiter.start()
        self.counters = {"electrons": 0, "batches": 0, "bytes_in": 0, "bytes_out": 0,
                         "zero_copy_batches": 0}
        self.closed = False
        if c.get("gc_tune"):
            tune_gc()
        self.heap = Heap(self.heap_base, self.heap_bytes)
        self._peer_ids: Dict[int, int] = {}
        self.sync_registry()

    # ------------------------------------------------- device-resident data --
    def alloc(self, nbytes: int, shape=None, order: str = "C") -> "DeviceArray":
        """A float32 slot in the arena heap (lattice intermediates, shared operands)."""
        off = self.heap.alloc(max(int(nbytes), 16))
        return DeviceArray(self, off, int(nbytes), tuple(shape) if shape is not None else (int(nbytes) // 4,), order)

    def put(self, array: np.ndarray) -> "DeviceArray":
        """Host ndarray -> arena heap (synchronous; used for lattice inputs and
        shared operands, not on the per-electron path)."""
        a = np.asarray(array)
        if a.dtype != np.float32:
            raise TypeError("device arrays are float32")
        order = "F" if (a.flags.f_contiguous and not a.flags.c_contiguous) else "C"
        flat = np.ascontiguousarray(a.reshape(-1, order="A" if order == "F" else "C"))
        da = self.alloc(flat.nbytes, a.shape, order)
        if flat.nbytes:
            L.call("b2_write", self.ordinal, da.offset, flat.ctypes.data, flat.nbytes)
        return da

    def get(self, da: "DeviceArray") -> np.ndarray:
        out = np.empty(da.shape, dtype=np.float32, order=da.order)
        if da.nbytes:
            flat = out.reshape(-1, order="A" if da.order == "F" else "C")
            L.call("b2_read", self.ordinal, da.offset, flat.ctypes.data, da.nbytes)
        return out

    # Bulk transfers of device-resident arrays: many arrays cross PCIe as ONE contiguous
    # copy per staging segment and are scattered to / gathered from their heap slots by the
    # span mover (K2 unpack / K1 pack) — instead of one synchronous copy per array.
    def _take_one(self, pool: list) -> int:
        """A free staging segment from ``pool`` (ingress or egress), waiting for in-flight
        batches if none is free (synchronous: used by the lattice compiler, not by run())."""
        for _ in range(3):
            with self._seg_lock:
                if pool:
                    return pool.pop()
            self.flush()
            L.call("b2_sync", self.ordinal)
            self._reap()
        raise MemoryError("no free staging segment (results of earlier electrons still hold all of them)")

    def _copy_into_pinned(self, jobs) -> None:
        """(ingress offset, flat uint8 array) pairs -> pinned ingress memory, on the library's
        copy threads (b2_stage_copy; synchronous)."""
        jobs = [j for j in jobs if j[1].size]
        if not jobs:
            return
        tab = np.zeros(len(jobs), dtype=L.JOB_DTYPE)
        tab["src"] = [j[1].ctypes.data for j in jobs]
        tab["dst_off"] = [j[0] for j in jobs]
        tab["bytes"] = [j[1].size for j in jobs]
        L.call("b2_stage_copy", self.ordinal, tab.ctypes.data, len(jobs))

    def put_many(self, arrays) -> List["DeviceArray"]:
        """Host ndarrays -> heap slots: per staging segment one H2D + one unpack launch."""
        out: List[DeviceArray] = []
        flats = []
        for a in arrays:
            a = np.asarray(a)
            if a.dtype != np.float32:
                raise TypeError("device arrays are float32")
            order = "F" if (a.flags.f_contiguous and not a.flags.c_contiguous) else "C"
            flat = np.ascontiguousarray(a.reshape(-1, order="A" if order == "F" else "C")).view(np.uint8)
            out.append(self.alloc(flat.size, a.shape, order))
            flats.append(flat)
        i, n = 0, len(flats)
        inflight: List[tuple] = []            # (segment, ticket): at most two, so that filling the
                                              # next segment overlaps the previous H2D + unpack
        def retire(limit: int) -> None:
            while len(inflight) > limit:
                seg_, t_ = inflight.pop(0)
                L.call("b2_wait", self.ordinal, t_)
                self._release_in(seg_)

        try:
            while i < n:
                if _round_up(flats[i].size) > self.seg_bytes:          # larger than a segment: direct copy
                    if flats[i].size:
                        L.call("b2_write", self.ordinal, out[i].offset, flats[i].ctypes.data, flats[i].size)
                    i += 1
                    continue
                retire(1)
                seg = self._take_one(self._free_segments)
                base, used, jobs, dsts, sizes = seg * self.seg_bytes, 0, [], [], []
                while i < n and used + _round_up(flats[i].size) <= self.seg_bytes:
                    if flats[i].size:
                        jobs.append((base + used, flats[i]))
                        dsts.append(out[i].offset)
                        sizes.append(flats[i].size)
                        used += _round_up(flats[i].size)
                    i += 1
                t = C.c_uint64()
                try:
                    if jobs:
                        self._copy_into_pinned(jobs)
                        L.call("b2_h2d", self.ordinal, base, base, used)
                        d = np.array(dsts, dtype=np.uint64)
                        z = np.array(sizes, dtype=np.uint64)
                        L.call("b2_unpack", self.ordinal, base, d.ctypes.data, z.ctypes.data, len(d), ALIGN,
                               None, None)
                        self.counters["bytes_in"] += used
                    L.call("b2_mark", self.ordinal, C.byref(t))
                except BaseException:
                    self._release_in(seg)
                    raise
                inflight.append((seg, t.value))
        finally:
            retire(0)
        return out

    def get_many(self, das) -> List[np.ndarray]:
        """Heap slots -> host ndarrays: per staging segment one pack launch + one D2H."""
        results: List[Optional[np.ndarray]] = [None] * len(das)
        i, n = 0, len(das)
        cap = self.out_data_bytes
        while i < n:
            if _round_up(das[i].nbytes) > cap:
                results[i] = self.get(das[i])
                i += 1
                continue
            oseg = self._take_one(self._free_out)
            base, used, idx, srcs, sizes, offs = oseg * self.seg_bytes, 0, [], [], [], []
            while i < n and used + _round_up(das[i].nbytes) <= cap:
                idx.append(i)
                if das[i].nbytes:
                    srcs.append(das[i].offset)
                    sizes.append(das[i].nbytes)
                offs.append(used)
                used += _round_up(das[i].nbytes)
                i += 1
            try:
                if srcs:
                    s_ = np.array(srcs, dtype=np.uint64)
                    z = np.array(sizes, dtype=np.uint64)
                    L.call("b2_pack", self.ordinal, s_.ctypes.data, z.ctypes.data, len(s_),
                           self.egress_base + base, ALIGN, None, None)
                    L.call("b2_d2h", self.ordinal, self.egress_base + base, base, used)
                    L.call("b2_sync", self.ordinal)
                    self.counters["bytes_out"] += used
                for k, j in enumerate(idx):
                    da = das[j]
                    arr = np.empty(da.shape, dtype=np.float32, order=da.order)
                    if da.nbytes:
                        flat = arr.reshape(-1, order="A" if da.order == "F" else "C")
                        flat.view(np.uint8)[:] = self._h_out[base + offs[k]:base + offs[k] + da.nbytes]
                    results[j] = arr
            finally:
                with self._seg_lock:
                    self._free_out.append(oseg)
                self._wake_waiters()
        return results

    def peer_arena_id(self, other_ordinal: int) -> int:
        """Arena id under which this GPU's kernels address ``other_ordinal``'s arena
        (same-process peers; maps it on first use, b2_peer_enable)."""
        if other_ordinal == self.ordinal:
            return 0
        aid = self._peer_ids.get(other_ordinal)
        if aid is None:
            get_engine(other_ordinal)                     # the peer must be up
            out = C.c_uint32()
            L.call("b2_peer_enable", self.ordinal, other_ordinal, C.byref(out))
            aid = self._peer_ids[other_ordinal] = out.value
        return aid

    def run_dag(self, table: np.ndarray, succ: Optional[np.ndarray] = None) -> np.ndarray:
        """One worker launch over a whole dependency graph of electrons whose
        operands already live in the arena: dependents are released on the
        device through the MPMC queue, no host round trip per edge.  Returns the
        per-electron status words.  ``table`` has dtype _b2lib.TASK_DTYPE."""
        return self.finish_dag(self.submit_dag(table, succ))

    def submit_dag(self, table: np.ndarray, succ: Optional[np.ndarray] = None):
        """Asynchronous half of ``run_dag``: returns a handle for ``finish_dag`` (so that
        several GPUs can work on their part of one lattice level at the same time).  Only
        one DAG per engine may be outstanding: its status words share a control region."""
        n = len(table)
        if n == 0:
            return (0, 0)
        if n > self.queue_cap:
            raise ValueError(f"DAG of {n} electrons exceeds queue capacity {self.queue_cap}")
        self.flush()                       # keep launch order = submission order
        self.sync_registry()
        table = np.ascontiguousarray(table)
        ns = 0 if succ is None else len(succ)
        sp = None
        if ns:
            succ = np.ascontiguousarray(succ, dtype=np.uint32)
            sp = succ.ctypes.data
        L.call("b2_enqueue", self.ordinal, table.ctypes.data, n, sp, ns, None)
        t = C.c_uint64()
        L.call("b2_launch", self.ordinal, self.ctrl_off, C.byref(t))
        self.counters["electrons"] += n
        self.counters["batches"] += 1
        return (t.value, n)

    def finish_dag(self, handle) -> np.ndarray:
        ticket, n = handle
        if n == 0:
            return np.zeros(0, dtype=np.uint32)
        L.call("b2_wait", self.ordinal, ticket)
        return self._h_out[self.ctrl_off:self.ctrl_off + 4 * n].view(np.uint32).copy()

    # ------------------------------------------------------------- plumbing --
    def _host_view(self, which: int, nbytes: int):
        ptr = C.c_void_p()
        size = C.c_uint64()
        L.call("b2_host_buffer", self.ordinal, which, C.byref(ptr), C.byref(size))
        if size.value < nbytes or not ptr.value:
            raise RuntimeError("pinned staging buffer smaller than requested")
        buf = (C.c_uint8 * nbytes).from_address(ptr.value)
        return np.frombuffer(buf, dtype=np.uint8), ptr.value

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
    def _need_in(self, enc: EncodedTask) -> int:
        """Ingress bytes ``_append`` will consume for this electron (the one formula
        ``fits``, ``stage`` and ``_append`` share)."""
        if enc.nrows and enc.spec.kind == L.K_STACK_SUM:
            return _round_up(_round_up(enc.count * 4, 16) * enc.nrows)     # rows at a 16 B aligned stride
        need = 0
        for p in enc.payloads:
            if type(p) is not DeviceRef:
                need += _round_up(p.size)
        return need

    def fits(self, enc: EncodedTask) -> bool:
        return (self._need_in(enc) <= self.seg_bytes
                and _round_up(max(enc.out_bytes, 16)) <= self.out_data_bytes)

    def _take_segments(self):
        with self._seg_lock:
            if self._free_segments and self._free_out:
                return self._free_segments.pop(), self._free_out.pop()
        return None

    def _open_batch(self, loop) -> Optional[Batch]:
        """A fresh open batch bound to ``loop`` (None when no staging segment pair is free)."""
        pair = self._take_segments()
        if pair is None:
            self._reap()
            pair = self._take_segments()
            if pair is None:
                return None
        b = Batch(self, pair[0], pair[1], loop)
        self._open = b
        self._stager.begin(self._jobs_addr[b.seg], self._cap_jobs, self._tables_addr[b.seg], self.max_batch,
                           b.in_base, self.seg_bytes, self.egress_base + b.out_base, self.out_data_bytes)
        if loop is not None:
            loop.call_soon(self._idle_flush, b, -1)
        return b

    def _pull(self, b: Batch) -> None:
        """Counters of the open batch: stager -> Batch (before Python-side staging / shipping)."""
        b.n, b.n_jobs, b.in_used, b.out_used = self._stager.state()

    def stage_fast(self, loop, spec: DeviceElectron, args):
        """The common case in one C call: a registered one-/two-array or int64 electron whose
        operands are plain C-contiguous float32 ndarrays (or Python ints).  Writes the task record
        and the copy job(s) into the open batch and returns ``(batch, packed)`` with
        ``packed = out_off << IDX_BITS | index``; returns None when the general path
        (``encode`` + ``stage``) has to look at this call, and ``BACKPRESSURE`` (False) when the
        call is fine but no staging segment pair is free right now (``segment_wait`` and retry)."""
        kind = spec.kind
        nargs = len(args)
        st = self._stager
        b = self._open
        if b is None or b.loop is not loop:
            if b is not None:
                self.flush()
            b = self._open_batch(loop)
            if b is None:
                return BACKPRESSURE
        if spec.fn_id > self._registered:
            self.sync_registry()
        for _ in range(2):
            if kind in _FAST1 and nargs == 1:
                r = st.stage1(spec.fn_id, kind, args[0])
            elif kind in _FAST2 and nargs == 2:
                r = st.stage2(spec.fn_id, args[0], args[1])
            elif kind in _FASTI and nargs == 2:
                r = st.stage_int(spec.fn_id, args[0], args[1])
            elif kind == L.K_ISCALE_ADD and nargs == 1:
                r = st.stage_int(spec.fn_id, args[0], 0)
            else:
                return None
            if r >= 0:
                b.keep.append(args)
                b.unread += 1
                return b, r
            if r != _b2host.FULL:
                return None
            self.flush()                       # the open batch is full: ship it, retry in a new one
            b = self._open_batch(loop)
            if b is None:
                return BACKPRESSURE
        return None

    def pass_wakeup(self) -> None:
        """A parked stager is leaving without staging (cancelled): if it had been woken for a free
        segment pair, somebody else has to be."""
        self._wake_waiters()

    def segment_wait(self, loop):
        """A future of ``loop`` that completes when a staging segment pair may be free again."""
        fut = loop.create_future()
        self._segment_waiters.append((loop, fut))
        return fut

    def waiter(self, b: Batch, loop):
        """A future of ``loop`` that completes when batch ``b`` has (one per electron: cancelling
        one ``run()`` must never cancel its batch-mates)."""
        fut = loop.create_future()
        b.waiters.append(fut)
        return fut

    async def stage(self, enc: EncodedTask) -> Staged:
        """General path: copy (or queue the copy of) one electron's operands into pinned
        ingress memory and append its task record to the open batch (ssh.py:147-150 +
        ssh.py:360-361).  Waits for a staging segment pair when none is free."""
        need_in = self._need_in(enc)
        need_out = _round_up(max(enc.out_bytes, 16))
        if need_in > self.seg_bytes or need_out > self.out_data_bytes:
            raise ValueError("electron payload larger than one staging segment")
        loop = asyncio.get_running_loop()
        while True:
            b = self._open
            if b is not None:
                self._pull(b)
                if (b.n >= self.max_batch or b.in_used + need_in > self.seg_bytes
                        or b.out_used + need_out > self.out_data_bytes or b.loop is not loop
                        or b.n_jobs + len(enc.payloads) > self._cap_jobs):
                    self.flush()
                    b = None
            if b is None:
                b = self._open_batch(loop)
                if b is None:
                    try:
                        await self.segment_wait(loop)
                    except BaseException:
                        self.pass_wakeup()
                        raise
                    continue
            break
        return self._append(b, enc)

    def _copy_in(self, b: Batch, off: int, p: np.ndarray) -> None:
        """One operand into pinned ingress memory at ``off``: small ones right here, large ones
        as a copy job for the library's copy threads (they run when the batch ships)."""
        if p.size < self.parallel_stage_bytes:
            self._h_in[off:off + p.size] = p
        else:
            if self._stager.add_job(p.ctypes.data, off, p.size) < 0:
                self._h_in[off:off + p.size] = p
            else:
                b.keep.append(p)

    def _append(self, b: Batch, enc: EncodedTask) -> Staged:
        """Reserve staging space, copy (or queue the copy of) the operands, write the
        task record(s).  ``b`` is the open batch; its counters live in the stager."""
        spec = enc.spec
        if spec.fn_id > self._registered:
            self.sync_registry()
        self._pull(b)
        seg_base = b.in_base
        h_in = self._h_in
        flags = 0
        payloads = enc.payloads
        if len(payloads) == 1 and not enc.nrows and type(payloads[0]) is not DeviceRef:
            p = payloads[0]                               # the common case, kept short
            a0 = seg_base + b.in_used
            self._copy_in(b, a0, p)
            b.in_used += _round_up(p.size)
            a1 = 0
        elif not payloads:
            a0, a1 = enc.imm[0] & 0xFFFFFFFFFFFFFFFF, enc.imm[1] & 0xFFFFFFFFFFFFFFFF
            flags = L.F_IMM0 | L.F_IMM1
        elif spec.kind == L.K_STACK_SUM:
            # rows are staged back to back at a fixed, 16 B aligned stride
            row_bytes = enc.count * 4
            stride = _round_up(row_bytes, 16)
            a0, a1 = seg_base + b.in_used, stride
            if len(payloads) == 1:
                rows = payloads[0].reshape(enc.nrows, row_bytes)
                if stride == row_bytes:
                    self._copy_in(b, a0, payloads[0])
                else:
                    for r in range(enc.nrows):
                        h_in[a0 + r * stride:a0 + r * stride + row_bytes] = rows[r]
            else:
                for r, p in enumerate(payloads):
                    self._copy_in(b, a0 + r * stride, p)
            b.in_used += _round_up(stride * enc.nrows)
        else:
            offs = []
            for i, p in enumerate(payloads):
                if type(p) is DeviceRef:                  # already in an arena: point at it
                    offs.append(p.offset)
                    if p.ordinal != self.ordinal:
                        flags |= self.peer_arena_id(p.ordinal) << (L.F_ARENA0_SHIFT, L.F_ARENA1_SHIFT)[i]
                    continue
                o = seg_base + b.in_used
                self._copy_in(b, o, p)
                offs.append(o)
                b.in_used += _round_up(p.size)
            a0, a1 = offs[0], (offs[1] if len(offs) > 1 else 0)
        assert b.in_used <= self.seg_bytes, "staging budget and staging layout disagree"
        out_off = b.out_used                                 # inside the egress segment
        b.out_used = out_off + _round_up(max(enc.out_bytes, 16))
        idx = b.n
        out_dev = self.egress_base + b.out_base + out_off
        parts = 1
        mem = b.table_mem
        if enc.out_bytes >= 2 * SPLIT_BYTES and spec.kind in _ELEMENTWISE and not (flags & L.F_IMM0):
            # One electron = one CTA on the device.  A batch of a few very large electrons would
            # leave most of the 592 resident CTAs idle, so a large elementwise electron goes out
            # as 1 MiB slices (disjoint ranges of the same operands); it has succeeded when every
            # slice has.
            esz = 1 if spec.kind == L.K_COPY else 4
            total = enc.count
            per = SPLIT_BYTES // esz
            parts = min(-(-total // per), self.max_batch - idx)
            if parts > 1:
                per = -(-total // parts)
                per = (per + 63) & ~63                       # keep slice starts 256 B aligned
                two = spec.kind in (L.K_ADD, L.K_MUL, L.K_AXPY)
                done, k = 0, 0
                while done < total:
                    n_el = min(per, total - done)
                    o = done * esz
                    _TASK.pack_into(mem, (idx + k) * 64, spec.fn_id, flags, a0 + o,
                                    (a1 + o) if two else a1, out_dev + o, n_el, 0, 0, 0, 0, 0)
                    done += n_el
                    k += 1
                parts = k
        if parts <= 1:
            parts = 1
            _TASK.pack_into(mem, idx * 64, spec.fn_id, flags, a0, a1, out_dev, enc.count, enc.nrows,
                            0, 0, 0, 0)
        b.n = idx + parts
        b.unread += 1
        _, n_jobs, _, _ = self._stager.state()                # add_job may have advanced it
        self._stager.set_state(b.n, n_jobs, b.in_used, b.out_used)
        st = Staged(b, idx, out_off, enc.out_bytes, enc.result_meta, spec)
        st.parts = parts
        return st

    def _idle_flush(self, b: Batch, seen: int) -> None:
        """Ships the open batch once a whole event-loop iteration has passed without it
        growing.  (Shipping on the first idle callback after opening cut batches short
        whenever the electrons that were about to join were still queued behind it — e.g.
        stagers woken by a segment release, whose wake-ups are scheduled after this callback.)"""
        if self._open is not b or b.flushed:
            return
        n = self._stager.state()[0]
        if n != seen:
            b.loop.call_soon(self._idle_flush, b, n)
            return
        self.flush()

    def flush(self) -> Optional[Batch]:
        """Ship the open batch: one b2_submit (copies, H2D, worker launch, D2H, ticket)."""
        b = self._open
        if b is None:
            return None
        self._open = None
        self._pull(b)
        self._stager.end()
        if b.n == 0:
            self._release_in(b.seg)
            self._release_out(b)
            return None
        return self._ship(b)

    def _ship(self, b: Batch) -> Batch:
        b.flushed = True
        status_off = b.out_base + self.out_data_bytes
        # results alias pinned memory only while the egress pool is comfortable
        with self._seg_lock:
            b.zero_copy = self.zero_copy and len(self._free_out) * 4 >= self.n_out_segments
        d = self._descs[b.seg]
        d.jobs = self._jobs_addr[b.seg]
        d.tasks = self._tables_addr[b.seg]
        d.n_jobs = b.n_jobs
        d.n_tasks = b.n
        d.flags = L.BATCH_INDEPENDENT              # run() electrons carry no dependency edges
        d.in_off = b.in_base
        d.in_bytes = b.in_used
        d.status_host_off = status_off
        d.out_dev_off = self.egress_base + b.out_base
        d.out_host_off = b.out_base
        d.out_bytes = b.out_used
        d.ticket = 0
        try:
            L.call("b2_submit", self.ordinal, C.byref(d))
            b.ticket = d.ticket
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
            except (RuntimeError, AttributeError):      # loop already closed / no loop
                try:
                    self._complete(b)
                except Exception:                       # this thread serves every later batch: it stays up
                    app_log.exception("completing a batch whose event loop is gone")

    def _complete(self, b: Batch) -> None:
        if b.done:
            return
        b.done = True
        b.keep = []                   # the copy jobs have run: operands may go
        self._release_in(b.seg)       # the copies out of pinned ingress have finished
        if b.error is None:
            words = self._h_out[b.out_base + self.out_data_bytes:
                                b.out_base + self.out_data_bytes + 4 * b.n].view(np.uint32)
            b.all_ok = bool((words == L.TS_OK).all())
        if b.error is not None or b.unread == 0:
            self._release_out(b)
        waiters, b.waiters = b.waiters, []
        for fut in waiters:
            if not fut.done():
                try:
                    fut.set_result(None)
                except RuntimeError:          # its loop was closed under it: nobody is left to wake
                    pass

    def status_of(self, s: Staged) -> int:
        b = s.batch
        if b.all_ok:
            return L.TS_OK
        off = b.out_base + self.out_data_bytes + 4 * s.index
        if s.parts == 1:
            return int(self._h_out[off:off + 4].view(np.uint32)[0])
        words = self._h_out[off:off + 4 * s.parts].view(np.uint32)
        bad = words[words != L.TS_OK]
        return int(bad[0]) if len(bad) else L.TS_OK

    def status_at(self, b: Batch, index: int) -> int:
        off = b.out_base + self.out_data_bytes + 4 * index
        return int(self._h_out[off:off + 4].view(np.uint32)[0])

    def _lease_of(self, b: Batch):
        lease = b.lease
        if lease is None:
            lease = (C.c_uint8 * self.out_data_bytes).from_address(self._h_out_addr + b.out_base)
            fin = weakref.finalize(lease, self._lease_dropped, b.oseg)
            fin.atexit = False
            b.lease = lease
            b.f32 = np.frombuffer(lease, dtype=np.float32)
            with self._seg_lock:
                self._leases_out += 1
            self.counters["zero_copy_batches"] += 1
        return lease

    def array_result(self, b: Batch, out_off: int, count: int) -> np.ndarray:
        """1-D float32 result of ``count`` elements at ``out_off`` of the batch's egress segment:
        a zero-copy slice of pinned memory when the pool has room, else a copy."""
        o4 = out_off >> 2
        if b.zero_copy and count:
            if b.lease is None:
                self._lease_of(b)
            return b.f32[o4:o4 + count]
        base4 = b.out_base >> 2
        return self._h_out_f32[base4 + o4:base4 + o4 + count].copy()

    def int_result(self, b: Batch, out_off: int) -> int:
        o = b.out_base + out_off
        return int(self._h_out[o:o + 8].view(np.int64)[0])

    def f32_result(self, b: Batch, out_off: int):
        o = b.out_base + out_off
        return np.float32(self._h_out[o:o + 4].view(np.float32)[0])

    def result_of(self, s: Staged) -> Any:
        """Materialise one result (ssh.py:451-456).  Array results alias the
        batch's pinned egress segment when the pool has room (no copy, no page
        faults); the segment returns to the pool when the last such array dies.
        Otherwise they are copied out."""
        b = s.batch
        meta = s.meta
        if b.zero_copy and meta[0] == "array" and s.out_bytes:
            arr = self.array_result(b, s.out_off, s.out_bytes // 4)
            return arr.reshape(meta[2], order=meta[3])
        off = b.out_base + s.out_off
        return decode(meta, self._h_out[off:off + max(s.out_bytes, 16)])

    def release(self, s: Staged) -> None:
        """The caller is done with this electron's staging (ssh.py:284-315)."""
        self.release_one(s.batch)

    def release_one(self, b: Batch) -> None:
        b.unread -= 1
        if b.unread == 0 and b.done:
            self._release_out(b)

    def cancel_slots(self, b: Batch, index: int, parts: int = 1) -> None:
        """Tombstone task records [index, index+parts) of batch ``b`` (the
        reference cannot cancel, ssh.py:460-464): directly in the table while the batch is
        still open, through ``b2_cancel`` once it has been submitted."""
        if b.done:
            return
        if not b.flushed:
            b.table["flags"][index:index + parts] |= L.F_CANCELLED
        else:
            L.call("b2_cancel", self.ordinal, b.ticket, index, parts)

    # -------------------------------------------------------- segment pools ---
    def _release_in(self, seg: int) -> None:
        with self._seg_lock:
            self._free_segments.append(seg)
        self._wake_waiters()

    def _release_out(self, b: Batch) -> None:
        """Egress segment back to the pool — unless result arrays alias it, in
        which case the lease finalizer does it when they are gone."""
        if b.out_freed:
            return
        b.out_freed = True
        if b.lease is not None:
            b.lease = None            # drop the batch's own references; arrays keep theirs
            b.f32 = None
            return
        with self._seg_lock:
            self._free_out.append(b.oseg)
        self._wake_waiters()

    def _lease_dropped(self, oseg: int) -> None:
        if self.closed:
            return
        with self._seg_lock:
            self._free_out.append(oseg)
            self._leases_out -= 1
        self._wake_waiters()

    def _wake_waiters(self) -> None:
        """Segments came back: wake the stagers parked on the pool — one batch worth per free
        segment pair.  (Waking one at a time shipped one-electron batches: the first opens a batch,
        the others must be there to join it.  Waking all of them on every release costs a wake-up
        and a re-park per parked electron per batch.)  Nothing is woken while no PAIR is free: a
        stager needs an ingress and an egress segment."""
        waiters = self._segment_waiters
        if not waiters:
            return
        with self._seg_lock:
            pairs = min(len(self._free_segments), len(self._free_out))
        if pairs == 0:
            return
        try:
            here = asyncio.get_running_loop()
        except RuntimeError:
            here = None
        budget = pairs * self.max_batch
        while budget > 0:
            try:
                loop, fut = waiters.popleft()
            except IndexError:
                return
            if fut.done():
                continue
            budget -= 1
            if loop is here:
                fut.set_result(None)
                continue
            try:
                loop.call_soon_threadsafe(_resolve, fut)
            except RuntimeError:
                continue

    # ----------------------------------------------------------------- close --
    def close(self) -> None:
        if self.closed:
            return
        self._waitq.put(None)
        if self._waiter is not threading.current_thread():
            self._waiter.join(timeout=10)      # it may be inside b2_wait on an event we are about to destroy
        try:
            L.call("b2_sync", self.ordinal)
        finally:
            with self._seg_lock:
                leaked = self._leases_out
            self.closed = True
            if leaked:
                # live result arrays still point into the pinned egress buffer:
                # hand it over to them instead of freeing it under their feet
                L.call("b2_detach_host_buffer", self.ordinal, 1)
            L.call("b2_shutdown", self.ordinal)


def _resolve(fut) -> None:
    if not fut.done():
        fut.set_result(None)


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


import atexit  # noqa: E402

atexit.register(shutdown_all)


def device_count() -> int:
    n = C.c_int()
    L.call("b2_device_count", C.byref(n))
    return n.value
