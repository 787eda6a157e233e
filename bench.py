#!/usr/bin/env python
"""bench.py — electrons/s on a 10 000-task fan-out lattice (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

Workload (config.workload = "C3"): 10 000 independent electrons per GPU, each
taking one 64 KiB fp32 ndarray and returning one (scale_add: 2x+1).  Weak
scaling: every rank runs its own 10 000 electrons on its own GPU; there is no
data-path collective.

A "step" is one pass of the hot path over the 10 000-electron batch.
  value  electrons/s with the operands already resident in HBM (device arena
         ingress region): per step, the same sequence of task-table uploads and
         worker-kernel launches the executor issues, timed with CUDA events on
         the library's stream, max over ranks.
  e2e    electrons/s through SSHExecutor.run with host numpy arrays: staging
         into pinned memory, H2D, worker, D2H and result materialisation all
         inside the timed region (wall clock bracketed by device syncs).
  roofline   the worker kernel (dominant kernel of the step): algorithmic bytes
         = (64 KiB in + 64 KiB out) per electron, against MEASURED_PEAKS.json.
  kernels    pack / unpack / worker achieved GB/s on the same 10 000 x 64 KiB set.
  cpu_baseline   the oracle port of the reference path, timed on a bounded sample.
"""

from __future__ import annotations

import argparse
import asyncio
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))
os.environ.setdefault("HOME", "/root")

import numpy as np  # noqa: E402

N_ELECTRONS = 10000
PAYLOAD = 65536
MiB = 1 << 20


def env_int(name, default):
    return int(os.environ.get(name, default))


# ------------------------------------------------------------------ helpers ---
class Clocks:
    """Samples nvidia-smi while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="b2clocks", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def measured_peak_gbs():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class Dist:
    """torch.distributed plumbing: barrier and max-over-ranks (NCCL when N > 1)."""

    def __init__(self, n_gpus: int):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.torch = None
        if self.world > 1 or os.environ.get("B2_BENCH_FORCE_TORCH") == "1":
            import torch
            import torch.distributed as dist

            self.torch, self.dist = torch, dist
            torch.cuda.set_device(self.local_rank)
            if self.world == 1:          # experiment switch: same plumbing as a multi-rank run
                os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
                os.environ.setdefault("MASTER_PORT", "29499")
                os.environ.setdefault("RANK", "0")
                os.environ.setdefault("WORLD_SIZE", "1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        if n_gpus != self.world:
            if self.world == 1 and n_gpus > 1:
                raise SystemExit(f"--gpus {n_gpus} needs torchrun --nproc-per-node {n_gpus}")

    def barrier(self):
        if self.torch is not None:
            self.dist.barrier()

    def max(self, v: float) -> float:
        if self.torch is None:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=f"cuda:{self.local_rank}")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.torch is not None:
            self.dist.destroy_process_group()


# ------------------------------------------------------------- CPU baseline ---
def cpu_baseline_sample(n_sample: int, payload: int, steps: int = 1, warmup: int = 0, budget_s: float = 150.0):
    """Times the oracle port of the reference path (oracle/reference_port.py) on
    n_sample electrons of the same workload, in waves of one coroutine per core.  The whole
    call is held to about ``budget_s`` seconds: if the first pass shows that warmup + steps
    passes of n_sample electrons would take longer, later passes use a smaller sample (the
    reference's cost per electron does not depend on the sample size)."""
    from covalent_ssh_plugin_b200 import electrons
    from oracle import reference_port as P

    cores = len(os.sched_getaffinity(0))
    calls = [(electrons.scale_add,
              [np.random.default_rng([0, i]).standard_normal(payload // 4, dtype=np.float32)], {})
             for i in range(n_sample)]
    times = []
    with tempfile.TemporaryDirectory(prefix="b2-cpu-") as tmp:
        s = P.PortSettings(tmp, pythonpath=str(REPO))
        # One BLAS/OpenMP thread per remote interpreter: with `cores` interpreters in flight the
        # default (one thread pool of `cores` threads EACH) oversubscribes the host ~cores-fold and
        # makes the reference path ~10x slower (15 vs 150 electrons/s on the 128-core B200 host).
        # The baseline is timed in its best configuration.
        for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            s.env[var] = "1"
        done = 0
        n_now = n_sample
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            out = asyncio.run(P.run_many(s, calls[:n_now], cores))
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
                done += n_now
            left = warmup + steps - it - 1
            if left and dt * left > budget_s:
                n_now = max(min(cores, n_sample), int(n_now * budget_s / (dt * left)))
        assert np.array_equal(out[0], np.float32(2.0) * calls[0][1][0] + np.float32(1.0))
    dt = sum(times) / len(times)
    n_sample = done / len(times)            # average sample per timed pass
    return {"value": n_sample / dt, "unit": "electrons/s", "cores": cores, "kind": "port",
            "sample": f"{n_sample:.0f} electrons x {payload} B fp32 scale_add through oracle/reference_port.py "
                      f"(reference ssh.py+exec.py path, local transport), waves of {cores} coroutines, "
                      f"OMP/BLAS threads pinned to 1 per remote interpreter",
            "seconds": dt}, times


def run_reference_arm(args, dist: Dist):
    if dist.rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    n_sample = max(32, min(256, 4 * cores))
    base, times = cpu_baseline_sample(n_sample, PAYLOAD, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "electrons/sec on 10k-task lattice", "value": base["value"],
        "unit": "electrons/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * base["seconds"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C3", "electrons_per_step": n_sample, "payload_bytes": PAYLOAD,
                   "electron": "scale_add (2x+1) fp32", "note": "bounded sample of the 10 000-electron lattice; "
                   "the reference path costs the same per electron regardless of lattice size"},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "electrons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa(index: int):
    """Pin this rank's host threads to the CPUs NVML reports as local to its GPU:
    the staging memcpy into pinned memory is the host's largest data mover."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


# --------------------------------------------------------------- CUDA arm -----
def run_cuda_arm(args, dist: Dist):
    from covalent_ssh_plugin_b200 import SSHExecutor, _b2lib as L, electrons
    from covalent_ssh_plugin_b200.engine import get_engine, shutdown_all

    dev = dist.local_rank
    if (dist.world > 1 and os.environ.get("B2_BENCH_BIND", "1") != "0") or os.environ.get("B2_BENCH_BIND") == "1":
        bind_to_gpu_numa(dev)
    n, payload = args.electrons, args.payload
    count = payload // 4
    total = n * payload
    max_batch = env_int("B2_MAX_BATCH", 4096)
    # one whole step must fit in flight each way; the egress pool also has to hold
    # every result of a step while the caller still references them (zero-copy)
    per_seg = max_batch * (payload + 256) + (1 << 20)
    n_batches = -(-n // max_batch)
    n_seg = max(env_int("B2_SEGMENTS", 8), n_batches + 1)
    n_out = max(env_int("B2_OUT_SEGMENTS", 0), 2 * n_batches + 4)
    staging = n_seg * per_seg
    heap = 2 * total + 512 * MiB
    arena = staging + n_out * per_seg + heap
    eng = get_engine(dev, arena_bytes=arena, staging_bytes=staging, n_segments=n_seg, n_out_segments=n_out,
                     max_batch=max_batch, queue_cap=1 << 16)
    spec = electrons.lookup(electrons.scale_add)
    peak, peak_src = measured_peak_gbs()

    ev = [C.c_uint32() for _ in range(2)]
    for e in ev:
        L.call("b2_event_create", dev, C.byref(e))

    def timed(fn, reps):
        L.call("b2_sync", dev)
        L.call("b2_event_record", dev, ev[0].value)
        for _ in range(reps):
            fn()
        L.call("b2_event_record", dev, ev[1].value)
        ms = C.c_float()
        L.call("b2_event_elapsed_ms", dev, ev[0].value, ev[1].value, C.byref(ms))
        return ms.value / reps

    # ---- device-resident inputs: fill the ingress mirror once -------------------
    rng = np.random.default_rng([0, dist.rank])
    seg_bytes = eng.seg_bytes
    slot = (payload + 255) & ~255
    batches = []
    host_in = eng._h_in
    for b0 in range(0, n, max_batch):
        k = min(max_batch, n - b0)
        seg = len(batches)
        base = seg * seg_bytes
        blk = rng.standard_normal(k * count, dtype=np.float32)
        tab = np.zeros(k, dtype=L.TASK_DTYPE)
        tab["fn_id"] = spec.fn_id
        tab["arg0"] = base + np.arange(k, dtype=np.uint64) * slot
        tab["out"] = eng.egress_base + base + np.arange(k, dtype=np.uint64) * slot   # egress segment = same index
        tab["count"] = count
        for i in range(k):
            o = base + i * slot
            host_in[o:o + payload] = blk[i * count:(i + 1) * count].view(np.uint8)
        L.call("b2_h2d", dev, base, base, k * slot)
        batches.append((tab, k, base, blk))
    L.call("b2_sync", dev)

    ticket = C.c_uint64()

    def device_step():
        for tab, k, base, _ in batches:
            L.call("b2_enqueue", dev, tab.ctypes.data, k, None, 0, None)
            L.call("b2_launch", dev, L.NO_STATUS_COPY, C.byref(ticket))

    # correctness of the device step (outside the timed region)
    device_step()
    tab, k, base, blk = batches[-1]
    L.call("b2_d2h", dev, eng.egress_base + base, base, k * slot)
    L.call("b2_sync", dev)
    got = eng._h_out[base + (k - 1) * slot: base + (k - 1) * slot + payload].view(np.float32)
    want = np.float32(2.0) * blk[(k - 1) * count:k * count] + np.float32(1.0)
    assert np.array_equal(got, want), "device step result differs from numpy"

    clocks = Clocks(dev)       # samples every 100 ms from here to the end of the e2e leg
    clocks.start()
    for _ in range(args.warmup):
        device_step()
    dist.barrier()
    s0 = eng.stats()
    ms_step = timed(device_step, args.steps)
    s1 = eng.stats()
    dist.barrier()
    ms_step = dist.max(ms_step)
    launches = int((s1.launches_worker - s0.launches_worker) + (s1.launches_pack - s0.launches_pack))
    traffic = None
    try:   # DRAM bytes per launch from the committed ncu capture (bench.py cannot run under ncu)
        tj = json.loads((REPO / "profiles" / "traffic.json").read_text())["worker_kernel"]
        traffic = tj["dram_bytes_per_electron"] * n / len(batches)
    except Exception:
        pass
    n_launch_per_step = len(batches)
    algo_bytes_per_launch = 2.0 * total / n_launch_per_step
    worker_gbs = (2.0 * total / 1e9) / (ms_step / 1e3)

    # ---- per-kernel numbers on the same payload set (rank-local) ----------------
    kernels = {"worker": {"achieved_gbs": worker_gbs, "frac": worker_gbs / peak,
                          "algorithmic_bytes_per_launch": algo_bytes_per_launch,
                          "launches_per_step": n_launch_per_step}}
    heap0 = eng.heap_base
    perm = np.random.default_rng(1).permutation(n).astype(np.uint64)
    scattered = heap0 + perm * payload
    sizes = np.full(n, payload, dtype=np.uint64)
    run_off = heap0 + total

    def do_pack():
        L.call("b2_pack", dev, scattered.ctypes.data, sizes.ctypes.data, n, run_off, 16, None, None)

    def do_unpack():
        L.call("b2_unpack", dev, run_off, scattered.ctypes.data, sizes.ctypes.data, n, 16, None, None)

    for name, fn in (("pack", do_pack), ("unpack", do_unpack)):
        for _ in range(5):          # > staging-ring depth, so no slot is sized inside the timed region
            fn()
        ms = timed(fn, max(5, args.steps))
        gbs = (2.0 * total / 1e9) / (ms / 1e3)
        kernels[name] = {"achieved_gbs": gbs, "frac": gbs / peak, "ms": ms,
                         "algorithmic_bytes_per_launch": 2.0 * total}

    # CPU context for the pack figure: the same gather done by numpy on one host core (what the
    # reference's transport would have to beat: it moves every payload byte through a file copy)
    if dist.rank == 0:
        n_host = max(1, min(2048, (128 * MiB) // payload))
        host_src = np.frombuffer(np.random.default_rng(3).bytes(n_host * payload), dtype=np.uint8)
        host_dst = np.empty(n_host * payload, dtype=np.uint8)
        hp = np.random.default_rng(4).permutation(n_host)
        host_dst[:] = 0                                    # touch the pages before timing
        t0 = time.perf_counter()
        for i, j in enumerate(hp):
            host_dst[i * payload:(i + 1) * payload] = host_src[j * payload:(j + 1) * payload]
        dt = time.perf_counter() - t0
        kernels["pack"]["cpu_one_core_gbs"] = 2.0 * n_host * payload / 1e9 / dt
        del host_src, host_dst

    # ---- end to end through the executor ------------------------------------------
    tmp = tempfile.mkdtemp(prefix="b2-bench-")
    key = Path(tmp) / "id_stub"
    key.touch()
    ex = SSHExecutor(username="bench", hostname=f"cuda:{dev}", ssh_key_file=str(key), cache_dir=tmp)
    xs = [np.random.default_rng([1, dist.rank, i]).standard_normal(count, dtype=np.float32) for i in range(n)]

    async def e2e_step(tag):
        return await asyncio.gather(*[
            ex.run(electrons.scale_add, [xs[i]], {}, {"dispatch_id": tag, "node_id": i}) for i in range(n)])

    e2e_steps = max(1, min(args.steps, 5))
    if args.skip_e2e:
        e2e_s, h2d, d2h = float("nan"), 0, 0
    else:
        out = asyncio.run(e2e_step("check"))
        for i in (0, n // 2, n - 1):
            assert np.array_equal(out[i], np.float32(2.0) * xs[i] + np.float32(1.0)), "e2e result differs from numpy"
        del out
        for w in range(max(1, min(args.warmup, 2))):
            asyncio.run(e2e_step(f"w{w}"))
        L.call("b2_sync", dev)
        dist.barrier()
        c0 = dict(eng.counters)
        t0 = time.perf_counter()
        for s in range(e2e_steps):
            asyncio.run(e2e_step(f"s{s}"))
        L.call("b2_sync", dev)
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        c1 = dict(eng.counters)
        dist.barrier()
        e2e_s = dist.max(e2e_s)
        h2d = (c1["bytes_in"] - c0["bytes_in"]) // e2e_steps + 64 * n
        d2h = (c1["bytes_out"] - c0["bytes_out"]) // e2e_steps + 4 * n

    clk = clocks.stop()

    # ---- NVLink peer path (N > 1): K4 pack into the next rank's arena ---------------------
    peer = None
    if dist.world > 1:
        from covalent_ssh_plugin_b200 import distributed as D

        ids = D.exchange_peer_arenas(dev, eng.arena_bytes)
        nxt = (dist.rank + 1) % dist.world
        m = 2000
        rows = np.zeros((m, 3), dtype=np.uint64)
        rows[:, 0] = heap0 + np.random.default_rng(2).permutation(m).astype(np.uint64) * payload
        rows[:, 1] = run_off + np.arange(m, dtype=np.uint64) * payload      # contiguous in the peer's heap
        rows[:, 2] = payload

        def do_peer():
            L.call("b2_peer_pack", dev, ids[nxt], rows.ctypes.data, m)

        for _ in range(3):
            do_peer()
        L.call("b2_sync", dev)
        dist.barrier()
        ms_peer = dist.max(timed(do_peer, 10))
        dist.barrier()
        gbs = m * payload / 1e9 / (ms_peer / 1e3)
        peer = {"kernel": "span mover writing into the next rank's arena (CUDA IPC mapping)",
                "gbs_per_direction": gbs, "all_ranks_sending": True, "of_measured_peer_copy_770": gbs / 770.0,
                "of_nominal_900": gbs / 900.0, "payloads": m, "payload_bytes": payload}

    # ---- CPU baseline (rank 0, N = 1 only) -------------------------------------------
    cpu = None
    if dist.world == 1 and not args.no_cpu_baseline:
        cores = len(os.sched_getaffinity(0))
        cpu, _ = cpu_baseline_sample(max(32, min(128, 2 * cores)), payload)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if dist.rank == 0:
        line = {
            "metric": "electrons/sec on 10k-task lattice", "value": dist.world * n / (ms_step / 1e3),
            "unit": "electrons/s", "n_gpus": dist.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C3", "electrons_per_gpu": n, "payload_bytes": payload,
                       "electron": "scale_add (2x+1) fp32, 64 KiB in / 64 KiB out", "max_batch": max_batch,
                       "l2": "inputs larger than L2 (1.31 GB touched per step)", "parallelism": f"fan-out x{dist.world}"},
            "e2e": {"value": dist.world * n / e2e_s, "unit": "electrons/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s, "steps": e2e_steps},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "b2::worker_kernel (scale_add callback)",
                         "achieved": worker_gbs, "peak": peak, "unit": "GB/s", "frac": worker_gbs / peak,
                         "traffic": traffic, "traffic_source": "profiles/traffic.json (ncu --set full capture)",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": algo_bytes_per_launch},
            "kernels": kernels,
            "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"],
                       "samples": clk["samples"]},
            "cpu_baseline": cpu,
            "peer": peer,
            "device": eng.info.name.decode(),
        }
        print(json.dumps(line), flush=True)
    shutdown_all()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--electrons", type=int, default=N_ELECTRONS)
    ap.add_argument("--payload", type=int, default=PAYLOAD)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: leave the executor leg out")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup
    if args.impl == "reference":
        # no GPU work on this arm: rank 0 alone measures, everybody else leaves
        rank = int(os.environ.get("RANK", "0"))
        if rank != 0:
            return
        d = Dist.__new__(Dist)
        d.world, d.rank, d.local_rank = 1, 0, 0
        run_reference_arm(args, d)
        return
    dist = Dist(args.gpus)
    try:
        run_cuda_arm(args, dist)
    finally:
        dist.close()


if __name__ == "__main__":
    main()
