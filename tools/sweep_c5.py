"""BASELINE config 5: payload sweep 1 KiB - 16 MiB x {1k, 10k, 100k} electrons on one GPU.

For every cell: the worker kernel (scale_add electrons, operands resident in the arena heap),
the span mover (pack of the same payloads) and — where the cell's bytes fit the staging pool —
the end-to-end rate through SSHExecutor.run.  Cells whose bytes exceed the wave budget run in
waves; the wave size is reported.  Prints a markdown table (committed under profiles/).

    python tools/sweep_c5.py [--max-n 100000] [--wave-gib 8] [--e2e-gib 1]
"""
import argparse
import asyncio
import ctypes as C
import os
import sys
import tempfile
import time
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
os.environ.setdefault("HOME", "/root")

import numpy as np  # noqa: E402

from covalent_ssh_plugin_b200 import SSHExecutor, _b2lib as L, electrons  # noqa: E402
from covalent_ssh_plugin_b200.engine import get_engine  # noqa: E402

KiB, MiB, GiB = 1 << 10, 1 << 20, 1 << 30
PAYLOADS = [1 * KiB, 4 * KiB, 16 * KiB, 64 * KiB, 256 * KiB, 1 * MiB, 4 * MiB, 16 * MiB]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-n", type=int, default=100000)
    ap.add_argument("--wave-gib", type=float, default=8.0)
    ap.add_argument("--e2e-gib", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--payloads", default="", help="comma-separated payload bytes (default: the full list)")
    ap.add_argument("--ns", default="1000,10000,100000")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    dev = 0
    wave_bytes = int(args.wave_gib * GiB)
    qcap = 1 << 17
    eng = get_engine(dev, arena_bytes=2 * wave_bytes + 7 * GiB, staging_bytes=int(1.25 * GiB), n_segments=8,
                     n_out_segments=24, max_batch=4096, queue_cap=qcap)
    spec = electrons.lookup(electrons.scale_add)
    peak = 6551.0
    ev = [C.c_uint32(), C.c_uint32()]
    for e in ev:
        L.call("b2_event_create", dev, C.byref(e))

    def timed(fn, reps):
        fn()
        L.call("b2_sync", dev)
        L.call("b2_event_record", dev, ev[0].value)
        for _ in range(reps):
            fn()
        L.call("b2_event_record", dev, ev[1].value)
        ms = C.c_float()
        L.call("b2_event_elapsed_ms", dev, ev[0].value, ev[1].value, C.byref(ms))
        return ms.value / reps

    # fill the input half of the heap once: 64 MiB of noise, replicated on the device
    src0 = eng.heap_base
    dst0 = eng.heap_base + wave_bytes
    noise = np.random.default_rng(0).standard_normal(16 * MiB, dtype=np.float32)
    L.call("b2_write", dev, src0, noise.ctypes.data, noise.nbytes)
    reps_fill = wave_bytes // noise.nbytes
    spans = np.zeros((reps_fill - 1, 3), dtype=np.uint64)
    spans[:, 0] = src0
    spans[:, 1] = src0 + np.arange(1, reps_fill, dtype=np.uint64) * noise.nbytes
    spans[:, 2] = noise.nbytes
    L.call("b2_move_spans", dev, spans.ctypes.data, len(spans))
    L.call("b2_sync", dev)

    tmp = tempfile.mkdtemp(prefix="b2-c5-")
    key = Path(tmp) / "k"
    key.touch()
    ex = SSHExecutor(username="u", hostname="cuda:0", ssh_key_file=str(key), cache_dir=tmp)

    print("| payload | electrons | wave | worker M el/s | worker GB/s | of peak | pack GB/s | of peak | e2e k el/s | e2e GB/s (in+out) |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    ticket = C.c_uint64()
    payloads = [int(x) for x in args.payloads.split(",")] if args.payloads else PAYLOADS
    for P in payloads:
        for N in [int(x) for x in args.ns.split(",")]:
            if N > args.max_n:
                continue
            slot = (P + 255) & ~255
            wave = max(1, min(N, wave_bytes // slot))
            n_waves = -(-N // wave)
            tabs = []
            for lo in range(0, wave, qcap):
                k = min(qcap, wave - lo)
                tab = np.zeros(k, dtype=L.TASK_DTYPE)
                tab["fn_id"] = spec.fn_id
                tab["arg0"] = src0 + (lo + np.arange(k, dtype=np.uint64)) * slot
                tab["out"] = dst0 + (lo + np.arange(k, dtype=np.uint64)) * slot
                tab["count"] = P // 4
                tabs.append(tab)

            def one_wave():
                for tab in tabs:
                    L.call("b2_enqueue", dev, tab.ctypes.data, len(tab), None, 0, None)
                    L.call("b2_launch", dev, L.NO_STATUS_COPY, C.byref(ticket))

            ms_wave = timed(one_wave, args.reps)
            # spot check the last electron of the wave
            got = np.empty(P // 4, dtype=np.float32)
            L.call("b2_read", dev, dst0 + (wave - 1) * slot, got.ctypes.data, P)
            off = ((wave - 1) * slot) % noise.nbytes
            want = np.float32(2.0) * np.resize(noise.view(np.uint8)[off:].view(np.float32), P // 4) + np.float32(1.0) \
                if off + P <= noise.nbytes else None
            assert want is None or np.array_equal(got, want), (P, N)
            t_all = ms_wave * (N / wave)                       # waves are identical; extrapolate from one
            w_els = N / (t_all / 1e3)
            w_gbs = 2.0 * N * P / 1e9 / (t_all / 1e3)

            scat = src0 + np.random.default_rng(1).permutation(wave).astype(np.uint64) * slot
            sizes = np.full(wave, P, dtype=np.uint64)

            def one_pack():
                L.call("b2_pack", dev, scat.ctypes.data, sizes.ctypes.data, wave, dst0, 16, None, None)

            ms_pack = timed(one_pack, args.reps)
            p_gbs = 2.0 * wave * P / 1e9 / (ms_pack / 1e3)

            e2e_cell = "-"
            e2e_gbs = "-"
            if not args.no_e2e and N * P <= args.e2e_gib * GiB and P <= 16 * MiB and N <= 10000:
                xs = [np.resize(noise[i::7], P // 4).copy() for i in range(min(N, 64))]

                async def go():
                    return await asyncio.gather(*[
                        ex.run(electrons.scale_add, [xs[i % len(xs)]], {}, {"dispatch_id": f"c5_{P}", "node_id": i})
                        for i in range(N)])

                asyncio.run(go())
                t0 = time.perf_counter()
                out = asyncio.run(go())
                L.call("b2_sync", dev)
                dt = time.perf_counter() - t0
                assert np.array_equal(out[-1], np.float32(2.0) * xs[(N - 1) % len(xs)] + np.float32(1.0))
                del out
                e2e_cell = f"{N / dt / 1e3:.1f}"
                e2e_gbs = f"{2.0 * N * P / dt / 1e9:.1f}"
            label = f"{P // KiB} KiB" if P < MiB else f"{P // MiB} MiB"
            print(f"| {label} | {N} | {wave} x{n_waves} | {w_els / 1e6:.2f} | {w_gbs:.0f} | {w_gbs / peak:.2f} | "
                  f"{p_gbs:.0f} | {p_gbs / peak:.2f} | {e2e_cell} | {e2e_gbs} |", flush=True)


if __name__ == "__main__":
    main()
