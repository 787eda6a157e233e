"""TEST DOUBLE of libb2exec.so for the CPU-only host-logic tests.

Implements the subset of the C ABI that engine.py / ssh.py call, in Python, on
top of the numpy contracts in oracle/device_model.py.  It exists so that the
batching, staging, completion and error-path logic of the Python host can be
tested where there is no GPU.  It is installed by the ``fake_lib`` fixture via
``_b2lib._lib = FakeLib()``; the product never imports it, and GPU tests never
use it.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from covalent_ssh_plugin_b200 import _b2lib as L
from oracle import device_model as M


def _deref(ref):
    """The ctypes object behind C.byref(obj) (or obj itself)."""
    return getattr(ref, "_obj", ref)


def _addr(p) -> int:
    if p is None:
        return 0
    if isinstance(p, int):
        return p
    return C.cast(p, C.c_void_p).value or 0


class _Dev:
    def __init__(self, arena, s_in, s_out, qcap):
        self.arena = np.zeros(arena, dtype=np.uint8)
        self.h_in = np.zeros(max(s_in, 1), dtype=np.uint8)
        self.h_out = np.zeros(max(s_out, 1), dtype=np.uint8)
        self.qcap = qcap
        self.fn = {0: {"kind": M.K_NONE}}
        self.pending, self.pending_succ = [], []
        self.peers = {}                   # arena id -> peer device ordinal
        self.ticket = 0
        self.stats = dict(launches_pack=0, launches_worker=0, launches_other=1, electrons=0,
                          bytes_packed=0, bytes_h2d=0, bytes_d2h=0)


class FakeLib:
    """Duck-typed stand-in for the ctypes CDLL object."""

    def __init__(self, n_devices: int = 1, delay_polls: int = 0, fail_init: int = 0):
        self.n_devices = n_devices
        self.devs = {}
        self.err = b""
        self.delay_polls = delay_polls     # b2_poll answers NOTREADY this many times per ticket
        self.fail_init = fail_init         # first N b2_init calls fail (retry-path tests)
        self.fail_launch = 0               # next N b2_launch calls fail (device lost mid-run)
        self._polls = {}
        self.calls = []

    # -- helpers ---------------------------------------------------------------
    def _fail(self, code, msg):
        self.err = msg.encode()
        return code

    def _dev(self, dev):
        return self.devs.get(dev)

    # -- lifecycle ---------------------------------------------------------------
    def b2_last_error(self):
        return self.err

    def b2_abi_version(self):
        return L.ABI_VERSION

    def b2_source_hash(self):
        return b"fake-numpy-model"

    def b2_device_count(self, ref):
        _deref(ref).value = self.n_devices
        return 0 if self.n_devices else self._fail(L.B2_E_CUDA, "no CUDA-capable device is detected")

    def b2_init(self, dev, arena, s_in, s_out, qcap):
        self.calls.append(("init", dev))
        if self.fail_init > 0:
            self.fail_init -= 1
            return self._fail(L.B2_E_CUDA, "device busy (injected)")
        if dev >= self.n_devices:
            return self._fail(L.B2_E_ARG, f"no CUDA device {dev} (have {self.n_devices})")
        if dev in self.devs:
            return self._fail(L.B2_E_STATE, "already initialised")
        self.devs[dev] = _Dev(arena, s_in, s_out, qcap)
        return 0

    def b2_shutdown(self, dev):
        self.devs.pop(dev, None)
        return 0

    def b2_is_initialized(self, dev):
        return int(dev in self.devs)

    def b2_get_info(self, dev, ref):
        info = _deref(ref)
        d = self._dev(dev)
        info.cc_major, info.cc_minor, info.sm_count = 10, 0, 148
        info.arena_bytes = d.arena.size
        info.staging_in_bytes, info.staging_out_bytes = d.h_in.size, d.h_out.size
        info.queue_cap, info.abi_version = d.qcap, L.ABI_VERSION
        info.name = b"FAKE B200 (numpy model)"
        return 0

    def b2_get_stats(self, dev, ref):
        s = _deref(ref)
        for k, v in self._dev(dev).stats.items():
            setattr(s, k, v)
        return 0

    # -- memory ------------------------------------------------------------------
    def b2_host_buffer(self, dev, which, pref, sref):
        d = self._dev(dev)
        buf = d.h_in if which == 0 else d.h_out
        _deref(pref).value = buf.ctypes.data
        _deref(sref).value = buf.size
        return 0

    def b2_arena_base(self, dev, arena_id, pref, sref):
        d = self._dev(dev)
        _deref(pref).value = d.arena.ctypes.data
        _deref(sref).value = d.arena.size
        return 0

    def b2_detach_host_buffer(self, dev, which):
        self.detached = getattr(self, "detached", []) + [(dev, which, self._dev(dev).h_out)]
        return 0

    def b2_h2d(self, dev, host_off, dev_off, n):
        d = self._dev(dev)
        if host_off + n > d.h_in.size or dev_off + n > d.arena.size:
            return self._fail(L.B2_E_ARG, "h2d range")
        d.arena[dev_off:dev_off + n] = d.h_in[host_off:host_off + n]
        d.stats["bytes_h2d"] += n
        return 0

    def b2_d2h(self, dev, dev_off, host_off, n):
        d = self._dev(dev)
        if host_off + n > d.h_out.size or dev_off + n > d.arena.size:
            return self._fail(L.B2_E_ARG, "d2h range")
        d.h_out[host_off:host_off + n] = d.arena[dev_off:dev_off + n]
        d.stats["bytes_d2h"] += n
        return 0

    def b2_write(self, dev, dev_off, src, n):
        d = self._dev(dev)
        if dev_off + n > d.arena.size:
            return self._fail(L.B2_E_ARG, "write range")
        d.arena[dev_off:dev_off + n] = np.frombuffer((C.c_uint8 * n).from_address(_addr(src)), dtype=np.uint8)
        return 0

    def b2_read(self, dev, dev_off, dst, n):
        d = self._dev(dev)
        if dev_off + n > d.arena.size:
            return self._fail(L.B2_E_ARG, "read range")
        np.frombuffer((C.c_uint8 * n).from_address(_addr(dst)), dtype=np.uint8)[:] = d.arena[dev_off:dev_off + n]
        return 0

    def _pack_like(self, dev, gather, scattered, sizes, n, run_off, align, out_offs, out_total):
        d = self._dev(dev)
        sc = np.frombuffer((C.c_uint64 * n).from_address(_addr(scattered)), dtype=np.uint64)
        sz = np.frombuffer((C.c_uint64 * n).from_address(_addr(sizes)), dtype=np.uint64)
        offs, total = M.pack_layout([int(x) for x in sz], int(run_off), int(align))
        if any(int(o) + int(z) > d.arena.size for o, z in zip(sc, sz)) or run_off + total > d.arena.size:
            return self._fail(L.B2_E_ARG, "span outside arena")
        spans = [(int(s), o, int(z)) if gather else (o, int(s), int(z)) for s, o, z in zip(sc, offs, sz)]
        M.move_spans(d.arena, spans)
        d.stats["launches_pack"] += 1
        d.stats["bytes_packed"] += int(sz.sum())
        return 0

    def b2_pack(self, dev, src_offs, sizes, n, dst_off, align, out_offs, out_total):
        return self._pack_like(dev, True, src_offs, sizes, n, dst_off, align, out_offs, out_total)

    def b2_unpack(self, dev, src_off, dst_offs, sizes, n, align, out_offs, out_total):
        return self._pack_like(dev, False, dst_offs, sizes, n, src_off, align, out_offs, out_total)

    # -- callback table + queue ------------------------------------------------------
    def b2_register(self, dev, fn_id, kind, alpha, beta, ialpha, ibeta):
        self._dev(dev).fn[fn_id] = dict(kind=kind, alpha=alpha, beta=beta, ialpha=ialpha, ibeta=ibeta)
        return 0

    def b2_enqueue(self, dev, tasks, n, succ, n_succ, first_ref):
        d = self._dev(dev)
        if len(d.pending) + n > d.qcap:
            return self._fail(L.B2_E_NOMEM, "batch exceeds queue capacity")
        raw = (C.c_uint8 * (64 * n)).from_address(_addr(tasks))
        arr = np.frombuffer(raw, dtype=L.TASK_DTYPE).copy()
        base_succ = len(d.pending_succ)
        if n_succ:
            sraw = (C.c_uint32 * n_succ).from_address(_addr(succ))
            d.pending_succ += list(sraw)
        for row in arr:
            t = {k: int(row[k]) for k in row.dtype.names}
            if t["succ_count"]:
                t["succ_begin"] += base_succ
            d.pending.append(t)
        return 0

    def b2_launch(self, dev, status_off, tref):
        if self.fail_launch > 0:
            self.fail_launch -= 1
            d = self._dev(dev)
            d.pending, d.pending_succ = [], []
            return self._fail(L.B2_E_CUDA, "cudaLaunchKernel -> unspecified launch failure (injected)")
        d = self._dev(dev)
        tasks, succ = d.pending, d.pending_succ
        d.pending, d.pending_succ = [], []
        if tasks:
            peers = {aid: self.devs[o].arena for aid, o in d.peers.items() if o in self.devs}
            status = M.run_tasks(d.arena, d.fn, tasks, succ, peers)
            d.stats["launches_worker"] += 1
            d.stats["electrons"] += len(tasks)
            if status_off != L.NO_STATUS_COPY:
                d.h_out[status_off:status_off + 4 * len(tasks)] = np.array(status, dtype=np.uint32).view(np.uint8)
        return self.b2_mark(dev, tref)

    # -- whole-batch submission (the real library does this on its own thread) ----------
    def _run_jobs(self, d, jobs, n):
        if not n:
            return 0
        rows = np.frombuffer((C.c_uint8 * (24 * n)).from_address(_addr(jobs)), dtype=L.JOB_DTYPE)
        self.jobs_run = getattr(self, "jobs_run", 0) + n
        for r in rows:
            src, off, nb = int(r["src"]), int(r["dst_off"]), int(r["bytes"])
            if off + nb > d.h_in.size:
                return self._fail(L.B2_E_ARG, "copy job outside ingress staging")
            if nb:
                d.h_in[off:off + nb] = np.frombuffer((C.c_uint8 * nb).from_address(src), dtype=np.uint8)
        return 0

    def b2_stage_copy(self, dev, jobs, n):
        return self._run_jobs(self._dev(dev), jobs, n)

    def b2_submit(self, dev, bref):
        b = _deref(bref)
        d = self._dev(dev)
        self.submits = getattr(self, "submits", 0) + 1
        rc = self._run_jobs(d, b.jobs, b.n_jobs)
        if rc == 0 and b.in_bytes:
            rc = self.b2_h2d(dev, b.in_off, b.in_off, b.in_bytes)
        if rc == 0 and b.n_tasks:
            rc = self.b2_enqueue(dev, b.tasks, b.n_tasks, None, 0, None)
        t = C.c_uint64()
        if rc == 0:
            rc = self.b2_launch(dev, b.status_host_off, t)
        if rc == 0 and b.out_bytes:
            rc = self.b2_d2h(dev, b.out_dev_off, b.out_host_off, b.out_bytes)
        if rc != 0:
            # like the real library: the submission is accepted, the failure surfaces on the ticket
            d.pending, d.pending_succ = [], []
            self.b2_mark(dev, t)
            self._ticket_err = getattr(self, "_ticket_err", {})
            self._ticket_err[(dev, t.value)] = (rc, self.err)
        b.ticket = t.value
        return 0

    def b2_table_create(self, dev, tasks, n, succ, n_succ, idref):
        d = self._dev(dev)
        d.tables = getattr(d, "tables", [])
        raw = np.frombuffer((C.c_uint8 * (64 * n)).from_address(_addr(tasks)), dtype=L.TASK_DTYPE).copy()
        sl = list((C.c_uint32 * n_succ).from_address(_addr(succ))) if n_succ else []
        d.tables.append((raw, sl))
        _deref(idref).value = len(d.tables) - 1
        return 0

    def b2_table_launch(self, dev, table_id, status_off, tref):
        d = self._dev(dev)
        tabs = getattr(d, "tables", [])
        if table_id >= len(tabs) or tabs[table_id] is None:
            return self._fail(L.B2_E_ARG, "unknown table")
        raw, sl = tabs[table_id]
        succ = (C.c_uint32 * max(1, len(sl)))(*sl)
        rc = self.b2_enqueue(dev, raw.ctypes.data, len(raw), succ if sl else None, len(sl), None)
        return rc or self.b2_launch(dev, status_off, tref)

    def b2_table_destroy(self, dev, table_id):
        d = self._dev(dev)
        tabs = getattr(d, "tables", [])
        if table_id >= len(tabs) or tabs[table_id] is None:
            return self._fail(L.B2_E_ARG, "unknown table")
        tabs[table_id] = None
        return 0

    def b2_resident_enable(self, dev, idle_us):
        return 0          # the numpy model has no launch gaps to hide

    def b2_resident_disable(self, dev):
        return 0

    def b2_cancel(self, dev, ticket, first, count):
        self.cancels = getattr(self, "cancels", []) + [(dev, ticket, first, count)]
        return 0          # the fake runs a batch inside b2_submit: nothing is left to tombstone

    def b2_peer_enable(self, dev, peer, aref):
        d = self._dev(dev)
        if peer not in self.devs:
            return self._fail(L.B2_E_STATE, "peer device is not initialised")
        for aid, o in d.peers.items():
            if o == peer:
                _deref(aref).value = aid
                return 0
        aid = len(d.peers) + 1
        d.peers[aid] = peer
        _deref(aref).value = aid
        return 0

    def b2_mark(self, dev, tref):
        d = self._dev(dev)
        d.ticket += 1
        if tref is not None:
            _deref(tref).value = d.ticket
        self._polls[(dev, d.ticket)] = self.delay_polls
        return 0

    def _ticket_failure(self, dev, ticket):
        err = getattr(self, "_ticket_err", {}).get((dev, ticket))
        if err is None:
            return 0
        self.err = err[1]
        return err[0]

    def b2_poll(self, dev, ticket):
        rc = self._ticket_failure(dev, ticket)
        if rc:
            return rc
        left = self._polls.get((dev, ticket), 0)
        if left > 0:
            self._polls[(dev, ticket)] = left - 1
            return L.B2_E_NOTREADY
        return 0

    def b2_wait(self, dev, ticket):
        rc = self._ticket_failure(dev, ticket)
        if rc:
            return rc
        self._polls[(dev, ticket)] = 0
        return 0

    def b2_sync(self, dev):
        return 0
