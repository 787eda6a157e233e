#!/usr/bin/env python
"""Host-path probe (no GPU): how many microseconds of Python does one SSHExecutor.run cost?

Installs a NULL device (tests/fake_b2lib.FakeLib with the numpy device model switched off:
every electron is reported OK, no payload byte is computed) and pushes N electrons through
``SSHExecutor.run`` exactly as bench.py's e2e leg does.  What is timed is therefore ONLY the host
side: asyncio, encode, staging into the (fake) pinned buffer, completion, result materialisation.

    python tools/host_probe.py [--n 10000] [--payload 65536] [--profile]
"""

from __future__ import annotations

import argparse
import asyncio
import os
import sys
import tempfile
import time
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
os.environ.setdefault("HOME", "/root")

import numpy as np  # noqa: E402

from covalent_ssh_plugin_b200 import _b2lib as L  # noqa: E402
from tests.fake_b2lib import FakeLib, _addr, _deref  # noqa: E402


class NullLib(FakeLib):
    """Completes every electron with status OK without touching payload bytes."""

    def b2_h2d(self, dev, host_off, dev_off, n):
        return 0

    def b2_d2h(self, dev, dev_off, host_off, n):
        return 0

    def b2_enqueue(self, dev, tasks, n, succ, n_succ, first_ref):
        self._dev(dev).pending.append(n)
        return 0

    def b2_launch(self, dev, status_off, tref):
        d = self._dev(dev)
        n = sum(d.pending)
        d.pending = []
        if status_off != L.NO_STATUS_COPY and n:
            d.h_out[status_off:status_off + 4 * n].view(np.uint32)[:] = L.TS_OK
        d.stats["launches_worker"] += 1
        d.stats["electrons"] += n
        return self.b2_mark(dev, tref)

    def b2_submit(self, dev, sref):
        import ctypes as C

        s = _deref(sref)
        d = self._dev(dev)
        n = s.n_tasks
        if n:
            d.h_out[s.status_host_off:s.status_host_off + 4 * n].view(np.uint32)[:] = L.TS_OK
        d.stats["launches_worker"] += 1
        d.stats["electrons"] += n
        d.ticket += 1
        s.ticket = d.ticket
        self._polls[(dev, d.ticket)] = 0
        return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10000)
    ap.add_argument("--payload", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--ints", action="store_true")
    ap.add_argument("--tight", action="store_true",
                    help="back-pressure shape: 1 GiB staging in 32 segments, 256-electron batches — the results "
                         "of one step need more egress segments than there are")
    args = ap.parse_args()

    L._lib = NullLib(n_devices=1)
    from covalent_ssh_plugin_b200 import SSHExecutor, electrons
    from covalent_ssh_plugin_b200.engine import get_engine, shutdown_all

    n, payload = args.n, args.payload
    count = payload // 4
    per_seg = 4096 * (payload + 256) + (1 << 20)
    nb = -(-n // 4096)
    if args.tight:
        eng = get_engine(0, arena_bytes=3 << 30, staging_bytes=1 << 30, n_segments=32, max_batch=256)
    else:
        eng = get_engine(0, arena_bytes=(nb + 1 + 2 * nb + 4) * per_seg + (64 << 20), staging_bytes=(nb + 1) * per_seg,
                         n_segments=nb + 1, n_out_segments=2 * nb + 4, max_batch=4096, queue_cap=1 << 16)
    tmp = tempfile.mkdtemp(prefix="b2-probe-")
    key = Path(tmp) / "id_stub"
    key.touch()
    ex = SSHExecutor(username="probe", hostname="cuda:0", ssh_key_file=str(key), cache_dir=tmp)
    xs = [np.random.default_rng([1, i]).standard_normal(count, dtype=np.float32) for i in range(n)]

    if args.ints:
        async def step(tag):
            return await asyncio.gather(*[
                ex.run(electrons.add_int, [i, 2 * i + 1], {}, {"dispatch_id": tag, "node_id": i}) for i in range(n)])
    else:
        async def step(tag):
            return await asyncio.gather(*[
                ex.run(electrons.scale_add, [xs[i]], {}, {"dispatch_id": tag, "node_id": i}) for i in range(n)])

    asyncio.run(step("warm"))
    if args.profile:
        import cProfile
        import pstats

        pr = cProfile.Profile()
        pr.enable()
        asyncio.run(step("prof"))
        pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(28)
        return
    best = 1e9
    for s in range(args.steps):
        t0 = time.perf_counter()
        out = asyncio.run(step(f"s{s}"))
        dt = time.perf_counter() - t0
        best = min(best, dt)
        del out
    print(f"{n} electrons x {payload} B: best {1e3 * best:.1f} ms/step = {1e6 * best / n:.2f} us/electron "
          f"= {n / best / 1e3:.1f} k electrons/s (host only, null device); batches {eng.counters['batches']}, "
          f"zero-copy batches {eng.counters['zero_copy_batches']}")
    shutdown_all()


if __name__ == "__main__":
    main()
