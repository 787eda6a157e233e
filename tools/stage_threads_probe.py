#!/usr/bin/env python
"""How does the size of the library's host copy pool (B2_STAGE_THREADS) move the host-bound legs?

One process per setting (the pool is sized once per process):

    for t in 4 8 16 32; do B2_STAGE_THREADS=$t python tools/stage_threads_probe.py; done

Prints, for the current setting: b2_stage_copy alone (pageable -> pinned ingress, 64 KiB and 1 MiB
pieces), `put_many` of 998 x 1 MiB (the upload of BASELINE config 4's diamond) and 1 MiB electrons
through SSHExecutor.run.
"""
import asyncio
import os
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ.setdefault("HOME", "/root")

import numpy as np  # noqa: E402

from covalent_ssh_plugin_b200 import SSHExecutor, electrons as E, _b2lib as L  # noqa: E402
from covalent_ssh_plugin_b200.engine import get_engine  # noqa: E402


def best_of(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def main():
    threads = os.environ.get("B2_STAGE_THREADS", "default")
    eng = get_engine(0, arena_bytes=6 << 30, staging_bytes=1 << 30, n_segments=8, max_batch=1024)
    out = [f"B2_STAGE_THREADS={threads}"]

    # b2_stage_copy alone: 512 MiB of distinct pageable source bytes into pinned ingress memory
    total = 512 << 20
    src = np.random.default_rng(0).integers(0, 255, total, dtype=np.uint8)
    for piece in (64 << 10, 1 << 20):
        k = total // piece
        tab = np.zeros(k, dtype=L.JOB_DTYPE)
        tab["src"] = src.ctypes.data + np.arange(k, dtype=np.uint64) * piece
        tab["dst_off"] = np.arange(k, dtype=np.uint64) * piece
        tab["bytes"] = piece
        s = best_of(lambda: L.call("b2_stage_copy", 0, tab.ctypes.data, k))
        out.append(f"stage_copy {piece >> 10} KiB pieces {total / s / 1e9:.1f} GB/s")

    # C4's upload: 998 x 1 MiB through put_many (copy pool -> H2D -> unpack, two segments in flight)
    ws = [np.random.default_rng(i).random(262144, dtype=np.float32) for i in range(998)]

    def upload():
        das = eng.put_many(ws)
        L.call("b2_sync", 0)
        for d in das:
            d.free()

    s = best_of(upload, 4)
    out.append(f"put_many 998 x 1 MiB {1e3 * s:.1f} ms = {998 * (1 << 20) / s / 1e9:.1f} GB/s")

    # 1 MiB electrons through run()
    tmp = tempfile.mkdtemp(prefix="b2-probe-")
    key = Path(tmp) / "id"
    key.touch()
    ex = SSHExecutor(username="probe", hostname="cuda:0", ssh_key_file=str(key), cache_dir=tmp)
    xs = ws[:512]

    async def step(tag):
        return await asyncio.gather(*[ex.run(E.scale_add, [x], {}, {"dispatch_id": tag, "node_id": i})
                                      for i, x in enumerate(xs)])

    loop = asyncio.new_event_loop()
    r = loop.run_until_complete(step("c"))
    assert np.array_equal(r[3], np.float32(2) * xs[3] + np.float32(1))
    del r
    s = best_of(lambda: loop.run_until_complete(step("t")), 5)
    out.append(f"run() 512 x 1 MiB {len(xs) / s / 1e3:.1f} k electrons/s = {2 * len(xs) * (1 << 20) / s / 1e9:.1f} GB/s in+out")
    loop.close()
    print("; ".join(out), flush=True)


if __name__ == "__main__":
    main()
