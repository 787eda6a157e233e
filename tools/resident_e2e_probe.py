#!/usr/bin/env python
"""The C3 lattice through SSHExecutor.run with and without the resident worker (one process per mode):

    python tools/resident_e2e_probe.py --resident 0
    python tools/resident_e2e_probe.py --resident 1

Same engine shape as bench.py's e2e leg (256-electron batches), best and median of the timed steps.
"""
import argparse
import asyncio
import statistics
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--resident", type=int, default=0)
    ap.add_argument("--electrons", type=int, default=10000)
    ap.add_argument("--max-batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--shape", default="probe", choices=["probe", "bench", "default"],
                    help="probe: 1 GiB / 32 segments; bench: bench.py's engine_shape (a whole step in flight); "
                         "default: the engine's own defaults (max_batch ignored)")
    ap.add_argument("--bind", type=int, default=0, help="1: pin to the GPU's NUMA node as bench.py does")
    args = ap.parse_args()

    from covalent_ssh_plugin_b200 import SSHExecutor, electrons as E, _b2lib as L
    from covalent_ssh_plugin_b200.engine import get_engine

    if args.bind:
        import bench

        bench.bind_to_gpu_numa(0)
    if args.shape == "bench":
        import bench

        per_seg, n_seg, n_out = bench.engine_shape(args.electrons, 65536, args.max_batch)
        eng = get_engine(0, resident=bool(args.resident), arena_bytes=(per_seg * (n_seg + n_out)) + (3 << 30),
                         staging_bytes=per_seg * n_seg, n_segments=n_seg, n_out_segments=n_out,
                         max_batch=args.max_batch)
    elif args.shape == "default":
        eng = get_engine(0, resident=bool(args.resident))
    else:
        eng = get_engine(0, resident=bool(args.resident), arena_bytes=4 << 30, staging_bytes=1 << 30, n_segments=32,
                         max_batch=args.max_batch)
    tmp = tempfile.mkdtemp(prefix="b2-probe-")
    key = Path(tmp) / "id"
    key.touch()
    ex = SSHExecutor(username="probe", hostname="cuda:0", ssh_key_file=str(key), cache_dir=tmp)
    n = args.electrons
    xs = [np.random.default_rng([1, i]).standard_normal(16384, dtype=np.float32) for i in range(n)]

    async def step(tag):
        return await asyncio.gather(*[ex.run(E.scale_add, [x], {}, {"dispatch_id": tag, "node_id": i})
                                      for i, x in enumerate(xs)])

    loop = asyncio.new_event_loop()
    out = loop.run_until_complete(step("c"))
    ok = all(np.array_equal(out[i], np.float32(2) * xs[i] + np.float32(1)) for i in range(0, n, 97))
    del out
    for w in range(3):
        loop.run_until_complete(step(f"w{w}"))
    L.call("b2_sync", 0)
    times = []
    s0 = eng.stats()
    for s in range(args.steps):
        t0 = time.perf_counter()
        out = loop.run_until_complete(step(f"s{s}"))
        times.append(time.perf_counter() - t0)
        del out
    s1 = eng.stats()
    loop.close()
    print(f"resident={args.resident} shape={args.shape} bind={args.bind} max_batch={eng.max_batch} "
          f"segments={eng.n_segments}/{eng.n_out_segments} x {eng.seg_bytes >> 20} MiB, batches/step "
          f"{(eng.counters['batches']) // (args.steps + 4)}: best {n / min(times) / 1e3:.1f} k, median "
          f"{n / statistics.median(times) / 1e3:.1f} k electrons/s; exact={ok}; worker launches "
          f"{s1.launches_worker - s0.launches_worker}, resident batches {s1.resident_batches - s0.resident_batches}, "
          f"resident kernel instances {s1.resident_instances - s0.resident_instances}", flush=True)


if __name__ == "__main__":
    main()
