#!/usr/bin/env python
"""bench.py — electrons/s on ONE 10 000-task fan-out lattice (BASELINE.json metric, config C3).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path

Workload C3 (BASELINE.json configs[2]): 10 000 independent electrons, each
taking one 64 KiB fp32 ndarray and returning one (scale_add: 2x+1); electron i runs on GPU
i mod G.  A "step" is one pass of the whole lattice through ``SSHExecutor.run``.

  value / e2e   electrons/s of the lattice through the executor's public call with HOST numpy
                arrays — pinned staging, H2D, worker kernel, D2H and result materialisation all
                inside the timed region (wall clock bracketed by device syncs and barriers, max over
                ranks).  STRONG scaling: the 10 000 electrons are sharded over the N ranks
                (rank r owns i % N == r, ``distributed.shard``), one dispatcher process per GPU.
                Every result of every timed step is checked against a checksum of the expected
                values.
  single_dispatcher   (N > 1) the same lattice driven by ONE process / one event loop over N
                executors cuda:0..N-1 round-robin — how a single Covalent dispatcher would use it.
  kernels.worker_step   the device-resident loop (operands already in HBM; per step the task-table
                uploads and worker launches the executor would issue), CUDA events on the library's
                stream, >= 0.5 s long, per-rank min / median / max.  ``roofline`` is this kernel.
  kernels.pack / unpack   span mover on the same 10 000 x 64 KiB set.
  c2, c4, c5    BASELINE configs 2 (10 000 int electrons), 4 (1 000-node diamond, 1 MiB operands,
                across the N GPUs) and 5 (payload sweep) — each with its own correctness check.
  peer, bcast   (N > 1) K4 NVLink peer pack (bytes verified) and the shared-operand ncclBroadcast.
  cpu_baseline  the reference's CPU path on this box's host cores, bounded sample, median of 3
                (``oracle/_ref`` = the unmodified reference when it was installed at build time,
                else the oracle port), plus the in-process floor (cloudpickle round trips only) and
                ``other_configs``: the reference beside C2 / C4 / C5 on bounded samples.
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
KiB, MiB, GiB = 1 << 10, 1 << 20, 1 << 30


def env_int(name, default):
    return int(os.environ.get(name, default))


# ------------------------------------------------------------------ helpers ---
class Clocks:
    """Samples nvidia-smi (one GPU, from ONE process) while the timed regions run."""

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
                 "-lms", str(max(50, env_int("B2_BENCH_CLOCKS_MS", 200)))], stdout=open(self.path, "w"),
                stderr=subprocess.DEVNULL)
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
    """torch.distributed plumbing: barriers, max-over-ranks, gathers (NCCL when N > 1; a gloo
    group for the phases in which only rank 0 works, so that the idle ranks wait on the CPU and
    do not park a spinning collective kernel on their GPU)."""

    def __init__(self, n_gpus: int):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.torch = None
        self.cpu_group = None
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
            try:
                self.cpu_group = dist.new_group(backend="gloo")
            except Exception:
                self.cpu_group = None
        if n_gpus != self.world:
            if self.world == 1 and n_gpus > 1:
                raise SystemExit(f"--gpus {n_gpus} needs torchrun --nproc-per-node {n_gpus}")

    def barrier(self):
        if self.torch is not None:
            self.dist.barrier()

    def cpu_barrier(self):
        """Barrier that leaves the GPUs alone (rank 0 may be driving every GPU meanwhile)."""
        if self.torch is None:
            return
        if self.cpu_group is not None:
            self.dist.barrier(group=self.cpu_group)
        else:
            self.dist.barrier()

    def max(self, v: float) -> float:
        if self.torch is None:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=f"cuda:{self.local_rank}")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather(self, v: float) -> list:
        if self.torch is None:
            return [v]
        t = self.torch.tensor([v], dtype=self.torch.float64, device=f"cuda:{self.local_rank}")
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(x.item()) for x in out]

    def all_true(self, flag: bool) -> bool:
        return self.max(0.0 if flag else 1.0) == 0.0

    def close(self):
        if self.torch is not None:
            self.dist.destroy_process_group()


# ------------------------------------------------------------- CPU baseline ---
# The electrons the REFERENCE runs are plain functions pickled BY VALUE, as a user's electrons are
# ("function body pickled by value"): its remote interpreter imports numpy (ints:
# nothing) and never this repository's package — importing it there would cost the reference ≈ 80 ms
# of its ≈ 400 ms per interpreter.  Same arithmetic as covalent_ssh_plugin_b200/electrons.py.
def _ref_scale_add(x):
    return np.float32(2.0) * x + np.float32(1.0)


def _ref_add_int(a, b):
    return a + b


def _ref_plus_one(x):
    return x + np.float32(1.0)


def _ref_vec_mul(x, y):
    return x * y


def _ref_stack_sum(xs):
    acc = xs[0].copy()
    for x in xs[1:]:
        acc += x
    return acc


def _by_value():
    """cloudpickle pickles functions of ``__main__`` by value; make it so when bench.py is imported."""
    if __name__ != "__main__":
        import cloudpickle

        try:
            cloudpickle.register_pickle_by_value(sys.modules[__name__])
        except Exception:
            pass


def _scale_add_calls(n_sample: int, payload: int):
    return [(_ref_scale_add,
             [np.random.default_rng([0, i]).standard_normal(payload // 4, dtype=np.float32)], {})
            for i in range(n_sample)]


def _reference_runner(tmp: str):
    """(kind, run_pass(calls) coroutine function, cores): the unmodified reference from oracle/_ref
    when it travelled (kind "reference"), else the restated port."""
    from oracle import ref_harness, reference_port as P

    _by_value()
    cores = len(os.sched_getaffinity(0))
    env_threads = {v: "1" for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
    # One BLAS/OpenMP thread per remote interpreter: with `cores` interpreters in flight the default
    # (a pool of `cores` threads EACH) oversubscribes the host and makes the reference ~10x slower
    # (15 vs 150 electrons/s on the 128-core B200 host).  The baseline runs in its best configuration.
    if ref_harness.available():
        saved = {k: os.environ.get(k) for k in env_threads}
        os.environ.update(env_threads)
        try:
            ex = ref_harness.make_executor(tmp, pythonpath=str(REPO))
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v

        async def run_pass(batch):
            out = []
            for lo in range(0, len(batch), cores):
                out += await asyncio.gather(*[
                    ex.run(fn, list(a), dict(k), {"dispatch_id": "cpu", "node_id": lo + j})
                    for j, (fn, a, k) in enumerate(batch[lo:lo + cores])])
            return out

        return "reference", run_pass, cores
    s = P.PortSettings(tmp, pythonpath=str(REPO))
    s.env.update(env_threads)

    async def run_pass(batch):
        return await P.run_many(s, batch, cores)

    return "port", run_pass, cores


def reference_cpu_sample(n_sample: int, payload: int, passes: int = 3, warmup: int = 0, budget_s: float = 120.0):
    """Times the reference's per-electron path on ``n_sample`` electrons of the C3 workload, in waves
    of one coroutine per host core (how Covalent awaits independent electrons), ``passes`` timed
    passes, median reported.

    kind "reference": the UNMODIFIED reference ``SSHExecutor.run`` + ``exec.py`` installed at build
    time into git-ignored ``oracle/_ref`` (oracle/make_ref.py), run under the stub ``covalent`` /
    local-transport ``asyncssh`` of oracle/ref_harness.py — no TCP, no crypto: strictly favourable to
    the reference.  kind "port": oracle/reference_port.py, when the reference did not travel."""
    from oracle import ref_harness

    calls = _scale_add_calls(n_sample, payload)
    times, done = [], []
    with tempfile.TemporaryDirectory(prefix="b2-cpu-") as tmp:
        kind, run_pass, cores = _reference_runner(tmp)
        n_now = n_sample
        out = None
        for it in range(warmup + passes):
            t0 = time.perf_counter()
            out = asyncio.run(run_pass(calls[:n_now]))
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt / n_now)
                done.append(n_now)
            left = warmup + passes - it - 1
            if left and dt * left > budget_s:
                n_now = max(min(cores, n_sample), int(n_now * budget_s / (dt * left)))
        assert np.array_equal(out[0], np.float32(2.0) * calls[0][1][0] + np.float32(1.0))
    per_electron = statistics.median(times)
    per_pass = [1.0 / t for t in times]
    sample = (f"{statistics.median(done):.0f} electrons x {payload} B fp32 scale_add per pass, "
              f"{len(times)} passes (median), through "
              + (f"the unmodified reference ssh.py+exec.py ({ref_harness.REFERENCE_ROOT}) under oracle/ref_harness.py"
                 if kind == "reference" else "oracle/reference_port.py")
              + f" with the local transport, waves of {cores} coroutines, OMP/BLAS threads pinned to 1, "
                "the electron pickled by value (its remote interpreter imports numpy only)")
    return {"value": 1.0 / per_electron, "unit": "electrons/s", "cores": cores, "kind": kind, "sample": sample,
            "per_pass": per_pass, "seconds_per_electron": per_electron}


def reference_other_configs(budget_s: float = 90.0):
    """The reference beside C2, C4 and C5 (BASELINE.json configs[1,3,4]): bounded samples, one wave of
    one coroutine per host core each, through the same runner as ``cpu_baseline``.  The reference has
    no lattice driver (Covalent core resolves dependencies and calls ``run`` once per node with the
    parents' results as arguments), so C4 is emulated exactly that way: root, a sample of the fan-out
    in one wave, a reduce over the sample — and extrapolated to the 1 000 nodes, saying so."""
    out = {}
    t_start = time.perf_counter()
    with tempfile.TemporaryDirectory(prefix="b2-cpu-") as tmp:
        kind, run_pass, cores = _reference_runner(tmp)
        wave = max(8, min(cores, 128))
        out["kind"], out["cores"] = kind, cores

        def timed_pass(calls):
            t0 = time.perf_counter()
            res = asyncio.run(run_pass(calls))
            return time.perf_counter() - t0, res

        # C2: int electrons (nothing but cloudpickle imported in the remote interpreter)
        ints = [(i, 2 * i + 1) for i in range(wave)]
        timed_pass([(_ref_add_int, [a, b], {}) for a, b in ints[:max(8, wave // 4)]])        # warm the page cache
        dt, res = timed_pass([(_ref_add_int, [a, b], {}) for a, b in ints])
        out["c2"] = {"value": len(ints) / dt, "unit": "electrons/s", "exact": res == [a + b for a, b in ints],
                     "sample": f"{len(ints)} add_int electrons, one wave"}

        # C5: one wave per payload size (fewer electrons for the largest: three files per electron)
        rows = []
        for size in (1 * KiB, 4 * KiB, 16 * KiB, 64 * KiB, 256 * KiB, 1 * MiB, 4 * MiB, 16 * MiB):
            if time.perf_counter() - t_start > budget_s:
                rows.append({"payload_bytes": size, "skipped": "time budget of the auxiliary block spent"})
                continue
            m = wave if size <= 1 * MiB else max(8, wave // 8)
            calls = _scale_add_calls(m, size)
            dt, res = timed_pass(calls)
            ok = bool(np.array_equal(res[-1], np.float32(2.0) * calls[-1][1][0] + np.float32(1.0)))
            rows.append({"payload_bytes": size, "electrons": m, "electrons_per_s": m / dt,
                         "gbs_in_plus_out": 2.0 * m * size / dt / 1e9, "exact": ok})
            del calls, res
        out["c5"] = {"rows": rows, "sample": "one wave of scale_add electrons per payload size"}

        # C4: root -> m x vec_mul(r, w_i) -> stack_sum over the m results, 1 MiB fp32 operands
        nel, n_fan = 262144, 998
        m = max(8, min(wave, 64))
        rng = np.random.default_rng(4)
        x0 = rng.random(nel, dtype=np.float32)
        ws = [rng.random(nel, dtype=np.float32) for _ in range(m)]
        t_root, r = timed_pass([(_ref_plus_one, [x0], {})])
        t_fan, prods = timed_pass([(_ref_vec_mul, [r[0], w], {}) for w in ws])
        t_red, total = timed_pass([(_ref_stack_sum, [prods], {})])
        want = _ref_stack_sum([_ref_vec_mul(_ref_plus_one(x0), w) for w in ws])
        waves = -(-n_fan // cores)
        # the reduce = one electron's fixed cost (the root's) + what its m x 1 MiB of arguments added, scaled to 998
        t_red_full = t_root + max(0.0, t_red - t_root) * (n_fan - 1) / max(1, m - 1)
        out["c4"] = {"sampled_nodes": m + 2, "seconds_root": t_root, "seconds_fan_out_wave": t_fan,
                     "seconds_reduce_of_sample": t_red, "exact": bool(np.array_equal(total[0], want)),
                     "extrapolated_ms_1000_nodes": 1e3 * (t_root + waves * t_fan + t_red_full),
                     "sample": f"root + {m} of the 998 fan-out electrons (one wave) + a reduce over those {m}; extrapolated as "
                               f"root + ceil(998 / {cores} cores) fan-out waves + a reduce whose argument-dependent part "
                               f"(reduce minus root) is scaled from {m} to 998 operands"}
    out["seconds"] = time.perf_counter() - t_start
    return out


def t_floor_sample(n_sample: int, payload: int):
    """BASELINE.md §3 T-floor: the same pickle round trips with no files, no subprocess, no
    transport — one worker process per host core.  The strongest CPU restatement there is."""
    import multiprocessing as mp

    cores = len(os.sched_getaffinity(0))
    workers = max(1, min(cores, 64))
    per_worker = max(8, n_sample // workers)
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        pool.map(_floor_worker, [(4, payload, w) for w in range(workers)])          # warm
        t0 = time.perf_counter()
        pool.map(_floor_worker, [(per_worker, payload, w) for w in range(workers)])
        dt = time.perf_counter() - t0
    return {"value": workers * per_worker / dt, "unit": "electrons/s", "workers": workers,
            "sample": f"{workers} processes x {per_worker} electrons, cloudpickle.dumps->loads->fn->dumps->loads in process "
                      "(oracle/reference_port.py::in_process_floor)"}


def _floor_worker(job):
    n, payload, seed = job
    from covalent_ssh_plugin_b200 import electrons
    from oracle import reference_port as P

    x = np.random.default_rng([7, seed]).standard_normal(payload // 4, dtype=np.float32)
    for _ in range(n):
        P.in_process_floor(electrons.scale_add, [x], {})
    return n


def run_reference_arm(args):
    cores = len(os.sched_getaffinity(0))
    n_sample = max(32, min(256, 2 * cores))
    base = reference_cpu_sample(n_sample, PAYLOAD, passes=max(3, min(args.steps, 5)), warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "electrons/sec on 10k-task lattice", "value": base["value"],
        "unit": "electrons/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * base["seconds_per_electron"] * N_ELECTRONS, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C3", "electrons": N_ELECTRONS, "electrons_timed_per_pass": n_sample,
                   "payload_bytes": PAYLOAD, "electron": "scale_add (2x+1) fp32",
                   "note": "bounded sample of the 10 000-electron lattice (the reference's cost per electron does "
                           "not depend on the lattice size); ms_per_step is the sample extrapolated to 10 000"},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample", "per_pass")},
        "e2e": {"value": base["value"], "unit": "electrons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa(index: int):
    """Pin this rank's host threads to the CPUs NVML reports as local to its GPU."""
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
def engine_shape(n_electrons: int, payload: int, max_batch: int):
    """Staging geometry that keeps one whole step in flight each way (results are zero-copy views
    of pinned egress memory, so the egress pool also holds every result the caller still has)."""
    per_seg = max_batch * (((payload + 255) & ~255) + 16) + (1 << 20)
    per_seg = (per_seg + 4095) & ~4095
    n_batches = -(-n_electrons // max_batch)
    n_seg = max(env_int("B2_SEGMENTS", 4), n_batches + 2)
    n_out = max(env_int("B2_OUT_SEGMENTS", 0), 2 * n_batches + 4)
    return per_seg, n_seg, n_out


def run_cuda_arm(args, dist: Dist):
    from covalent_ssh_plugin_b200 import SSHExecutor, _b2lib as L, distributed as D, electrons
    from covalent_ssh_plugin_b200.engine import get_engine, shutdown_all

    dev, world, rank = dist.local_rank, dist.world, dist.rank
    all_cpus = os.sched_getaffinity(0)
    if os.environ.get("B2_BENCH_BIND", "1") != "0":
        bind_to_gpu_numa(dev)          # dispatcher + copy threads next to their GPU; undone before the CPU baseline
    n, payload = args.electrons, args.payload
    count = payload // 4
    total = n * payload
    # electrons per b2_submit: 256 x 64 KiB = 16 MiB per batch keeps the copy / H2D / D2H pipeline busy while the
    # event loop is still staging, and keeps the last batch (pure latency) short; measured flat from 256 to 2048 at
    # 10 000 electrons per dispatcher, 30 % better than 1024 at the 1 250 per dispatcher of the 8-GPU lattice
    max_batch = env_int("B2_MAX_BATCH", 256)
    dev_batch = min(n, 1 << 16)     # device-resident operands need no staging pipeline: one launch per step,
                                    # as engine.run_dag / a lattice level issues it (queue capacity 65 536)
    per_seg, n_seg, n_out = engine_shape(n, payload, max_batch)
    # the ingress mirror must also hold the whole device-resident input set of worker_step
    n_seg = max(n_seg, -(-(n * ((payload + 255) & ~255) + (1 << 20)) // per_seg) + 1)
    n_out = max(n_out, n_seg)
    staging = n_seg * per_seg
    heap = max(2 * total + 768 * MiB, 8 * GiB + 256 * MiB)      # the C5 cells use 2 x 4 GiB of it
    arena = staging + n_out * per_seg + (1 << 20) + heap
    eng = get_engine(dev, arena_bytes=arena, staging_bytes=staging, n_segments=n_seg, n_out_segments=n_out,
                     max_batch=max_batch, queue_cap=1 << 16)
    spec = electrons.lookup(electrons.scale_add)
    peak, peak_src = measured_peak_gbs()
    lib_sha = L.source_hash()

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

    clocks = Clocks(dev) if rank == 0 and env_int("B2_BENCH_CLOCKS_MS", 200) > 0 else None   # ONE sampler per run
    if clocks:
        clocks.start()

    # ============ kernels.worker_step: device-resident loop (weak: every rank its own 10 000) ======
    rng = np.random.default_rng([0, rank])
    slot = (payload + 255) & ~255
    in0 = 0                                             # ingress mirror, contiguous
    out0 = eng.egress_base
    blk = rng.standard_normal(n * count, dtype=np.float32)
    host_in = eng._h_in
    for i in range(n):
        host_in[i * slot:i * slot + payload] = blk[i * count:(i + 1) * count].view(np.uint8)
    L.call("b2_h2d", dev, 0, in0, n * slot)
    L.call("b2_sync", dev)
    tabs = []
    for b0 in range(0, n, dev_batch):
        k = min(dev_batch, n - b0)
        tab = np.zeros(k, dtype=L.TASK_DTYPE)
        tab["fn_id"] = spec.fn_id
        tab["arg0"] = in0 + (b0 + np.arange(k, dtype=np.uint64)) * slot
        tab["out"] = out0 + (b0 + np.arange(k, dtype=np.uint64)) * slot
        tab["count"] = count
        tabs.append(tab)
    ticket = C.c_uint64()

    def device_step():
        for tab in tabs:
            L.call("b2_enqueue", dev, tab.ctypes.data, len(tab), None, 0, None)
            L.call("b2_launch", dev, L.NO_STATUS_COPY, C.byref(ticket))

    device_step()
    got = np.empty(count, dtype=np.float32)
    L.call("b2_read", dev, out0 + (n - 1) * slot, got.ctypes.data, payload)
    assert np.array_equal(got, np.float32(2.0) * blk[(n - 1) * count:] + np.float32(1.0)), \
        "device step result differs from numpy"
    for _ in range(max(3, args.warmup)):
        device_step()
    ms_probe = timed(device_step, 5)
    dev_steps = max(args.steps, int(min(20000, 600.0 / max(ms_probe, 1e-3))))     # >= 0.5 s of device time
    dist.barrier()
    s0 = eng.stats()
    ms_dev = timed(device_step, dev_steps)
    s1 = eng.stats()
    dist.barrier()
    ms_ranks = dist.gather(ms_dev)
    ms_dev_max = max(ms_ranks)
    dev_launches = int(s1.launches_worker - s0.launches_worker)
    worker_gbs = (2.0 * total / 1e9) / (ms_dev_max / 1e3)
    traffic, traffic_src = None, None
    try:   # DRAM bytes per launch from the committed ncu capture (bench.py itself never runs under ncu)
        tj = json.loads((REPO / "profiles" / "traffic.json").read_text())
        traffic = tj["worker_kernel"]["dram_bytes_per_electron"] * n / len(tabs)
        traffic_src = f"profiles/traffic.json ({tj.get('captured_on', 'ncu --set full capture')})"
    except Exception:
        pass
    # the same step from a RESIDENT table (b2_table_launch): no table rebuild / pinned copy / upload per launch
    tids = []
    for tab in tabs:
        tid = C.c_uint32()
        L.call("b2_table_create", dev, tab.ctypes.data, len(tab), None, 0, C.byref(tid))
        tids.append(tid.value)

    def resident_step():
        for tid in tids:
            L.call("b2_table_launch", dev, tid, L.NO_STATUS_COPY, C.byref(ticket))

    L.call("b2_fill", dev, out0, 0, n * slot)
    resident_step()
    L.call("b2_read", dev, out0 + (n - 1) * slot, got.ctypes.data, payload)
    assert np.array_equal(got, np.float32(2.0) * blk[(n - 1) * count:] + np.float32(1.0)), \
        "resident-table step result differs from numpy"
    for _ in range(3):
        resident_step()
    dist.barrier()
    ms_res = dist.max(timed(resident_step, dev_steps))
    dist.barrier()
    for tid in tids:
        L.call("b2_table_destroy", dev, tid)
    res_gbs = (2.0 * total / 1e9) / (ms_res / 1e3)
    kernels = {"worker_resident": {
        "what": "the same step launched from a resident task table (b2_table_launch): no per-launch table work on the host",
        "ms_per_step": ms_res, "steps": dev_steps, "achieved_gbs": res_gbs, "frac": res_gbs / peak,
        "electrons_per_s_per_gpu": n / (ms_res / 1e3)}}
    kernels["worker_step"] = {
        "what": "device-resident step: per step %d x (b2_enqueue + table upload + worker_tma_kernel launch) over "
                "%d electrons already in HBM" % (len(tabs), n),
        "electrons_per_s_per_gpu": n / (ms_dev_max / 1e3), "ms_per_step": ms_dev_max, "steps": dev_steps,
        "ms_per_rank": {"min": min(ms_ranks), "median": statistics.median(ms_ranks), "max": ms_dev_max},
        "achieved_gbs": worker_gbs, "frac": worker_gbs / peak,
        "algorithmic_bytes_per_launch": 2.0 * total / len(tabs), "launches_per_step": len(tabs)}

    # ============ small batches: launch per batch vs the resident worker ==========================
    # The run() path ships max_batch electrons per b2_submit.  With operands resident, the same
    # 10 000 electrons as ceil(n / max_batch) batches per step: one kernel launch each, or published to
    # the resident worker (b2_resident_enable) through its doorbell — no launch, no gap between batches.
    try:
        if os.environ.get("B2_BENCH_SKIP_RESIDENT") == "1":     # profiler runs: ncu serialises kernels, a resident
            raise RuntimeError("skipped (B2_BENCH_SKIP_RESIDENT=1)")   # instance would sit out its lifetime cap per batch
        per_size = {}
        for bsz in sorted({max_batch, 1024}):
            small_tabs = []
            for b0 in range(0, n, bsz):
                k = min(bsz, n - b0)
                tab = np.zeros(k, dtype=L.TASK_DTYPE)
                tab["fn_id"] = spec.fn_id
                tab["arg0"] = in0 + (b0 + np.arange(k, dtype=np.uint64)) * slot
                tab["out"] = out0 + (b0 + np.arange(k, dtype=np.uint64)) * slot
                tab["count"] = count
                small_tabs.append(tab)

            def small_step():
                for tab in small_tabs:
                    L.call("b2_enqueue", dev, tab.ctypes.data, len(tab), None, 0, None)
                    L.call("b2_launch", dev, L.NO_STATUS_COPY, C.byref(ticket))

            for _ in range(3):
                small_step()
            small_steps = max(50, dev_steps // 8)
            ms_small = dist.max(timed(small_step, small_steps))
            r0 = eng.stats()
            L.call("b2_resident_enable", dev, 0)
            L.call("b2_fill", dev, out0, 0, n * slot)
            for _ in range(3):
                small_step()
            L.call("b2_read", dev, out0 + (n - 1) * slot, got.ctypes.data, payload)
            res_exact = bool(np.array_equal(got, np.float32(2.0) * blk[(n - 1) * count:] + np.float32(1.0)))
            ms_small_res = dist.max(timed(small_step, small_steps))
            L.call("b2_resident_disable", dev)
            r1 = eng.stats()
            per_size[str(bsz)] = {
                "batches_per_step": len(small_tabs),
                "launch_per_batch": {"ms_per_step": ms_small, "frac": (2.0 * total / 1e9) / (ms_small / 1e3) / peak},
                "resident_worker": {"ms_per_step": ms_small_res, "frac": (2.0 * total / 1e9) / (ms_small_res / 1e3) / peak,
                                    "exact": res_exact,
                                    "batches": int(r1.resident_batches - r0.resident_batches),
                                    "kernel_instances": int(r1.resident_instances - r0.resident_instances)},
                "steps": small_steps}
        kernels["small_batches"] = {
            "what": "the same 10 000 electrons (device-resident operands) as batches of N per step, driven from this "
                    "Python loop (two ABI calls per batch): one worker launch per batch vs. published to the resident "
                    "worker through its doorbell.  Both are bound by the host's per-batch cost at these sizes.",
            "electrons_per_batch": per_size}
    except Exception as err:           # never lose the line over the optional mode
        try:
            L.call("b2_resident_disable", dev)
        except Exception:
            pass
        kernels["small_batches"] = {"error": repr(err)[:300]}

    # ============ kernels.pack / unpack on the same payload set (rank-local) ======================
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
        for _ in range(10):         # > staging-ring depth, so no slot is sized inside the timed region
            fn()
        ms = timed(fn, max(20, args.steps))
        gbs = (2.0 * total / 1e9) / (ms / 1e3)
        kernels[name] = {"achieved_gbs": gbs, "frac": gbs / peak, "ms": ms,
                         "algorithmic_bytes_per_launch": 2.0 * total}

    # ============ value / e2e: the 10 000-electron lattice through SSHExecutor.run ================
    tmp = tempfile.mkdtemp(prefix="b2-bench-")
    key = Path(tmp) / "id_stub"
    key.touch()
    ex = SSHExecutor(username="bench", hostname=f"cuda:{dev}", ssh_key_file=str(key), cache_dir=tmp)
    mine = list(D.shard(n, rank, world))                                 # electron ids of this rank
    xs = {i: np.random.default_rng([1, i]).standard_normal(count, dtype=np.float32) for i in mine}
    # expected checksum of a step: first, middle and last element of every result, in float64
    probe = (0, count // 2, count - 1)
    want_sum = sum(float(np.float32(2.0) * xs[i][j] + np.float32(1.0)) for i in mine for j in probe)

    async def lattice_step(tag):
        return await asyncio.gather(*[
            ex.run(electrons.scale_add, [xs[i]], {}, {"dispatch_id": tag, "node_id": i}) for i in mine])

    def checksum(out):
        return sum(float(r[j]) for r in out for j in probe)

    e2e_steps = max(1, args.steps)                       # exactly the K steps the caller asked for
    loop = asyncio.new_event_loop()                      # ONE event loop for all steps, like a dispatcher's
    run_step = loop.run_until_complete
    out = run_step(lattice_step("check"))
    for i in mine[:: max(1, len(mine) // 50)]:
        kk = mine.index(i)
        assert np.array_equal(out[kk], np.float32(2.0) * xs[i] + np.float32(1.0)), "e2e result differs from numpy"
    del out
    for w in range(args.warmup):                          # W untimed passes (>= 3), after the correctness pass above
        run_step(lattice_step(f"w{w}"))
    L.call("b2_sync", dev)
    dist.barrier()
    c0, st0 = dict(eng.counters), eng.stats()
    verified = True
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        out = run_step(lattice_step(f"s{s}"))
        verified = verified and abs(checksum(out) - want_sum) <= 1e-9 * max(1.0, abs(want_sum))
        del out
    L.call("b2_sync", dev)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    c1, st1 = dict(eng.counters), eng.stats()
    dist.barrier()
    loop.close()
    e2e_ranks = dist.gather(e2e_s)
    e2e_s = max(e2e_ranks)
    verified = dist.all_true(verified)
    h2d = (c1["bytes_in"] - c0["bytes_in"]) // e2e_steps + 64 * len(mine)
    d2h = (c1["bytes_out"] - c0["bytes_out"]) // e2e_steps + 4 * len(mine)
    e2e_launches = int((st1.launches_worker - st0.launches_worker) + (st1.launches_pack - st0.launches_pack))
    e2e_launches = int(sum(dist.gather(float(e2e_launches))))           # kernels launched by all ranks in the timed steps
    value = n / e2e_s
    if rank == 0:
        print(f"[bench] worker_step {ms_dev_max:.4f} ms ({worker_gbs / peak:.3f} of peak; resident table {res_gbs / peak:.3f}), pack {kernels['pack']['frac']:.3f}, "
              f"unpack {kernels['unpack']['frac']:.3f}; lattice through run(): {value / 1e3:.1f} k electrons/s "
              f"({1e3 * e2e_s:.1f} ms/step, verified={verified})", file=sys.stderr, flush=True)

    clk = clocks.stop() if clocks else None

    # ============ single dispatcher (N > 1): ONE event loop drives cuda:0..N-1 round-robin ========
    single = None
    if world > 1:
        dist.cpu_barrier()
        if rank == 0:
            sp, sn_seg, sn_out = engine_shape(-(-n // world), payload, max_batch)
            exs = [ex]
            for d in range(1, world):
                get_engine(d, arena_bytes=sn_seg * sp + sn_out * sp + 3 * GiB, staging_bytes=sn_seg * sp,
                           n_segments=sn_seg, n_out_segments=sn_out, max_batch=max_batch, queue_cap=1 << 16)
                exs.append(SSHExecutor(username="bench", hostname=f"cuda:{d}", ssh_key_file=str(key), cache_dir=tmp))
            xall = [np.random.default_rng([1, i]).standard_normal(count, dtype=np.float32) for i in range(n)]
            want_all = sum(float(np.float32(2.0) * xall[i][j] + np.float32(1.0)) for i in range(n) for j in probe)

            async def single_step(tag):
                return await asyncio.gather(*[
                    exs[i % world].run(electrons.scale_add, [xall[i]], {}, {"dispatch_id": tag, "node_id": i})
                    for i in range(n)])

            sloop = asyncio.new_event_loop()
            for w in range(2):
                sloop.run_until_complete(single_step(f"sw{w}"))
            ok = True
            sd_steps = max(3, min(e2e_steps, 10))
            t0 = time.perf_counter()
            for s in range(sd_steps):
                out = sloop.run_until_complete(single_step(f"ss{s}"))
                ok = ok and abs(checksum(out) - want_all) <= 1e-9 * max(1.0, abs(want_all))
                del out
            for d in range(world):
                L.call("b2_sync", d)
            dt = (time.perf_counter() - t0) / sd_steps
            sloop.close()
            single = {"value": n / dt, "unit": "electrons/s", "ms_per_step": 1e3 * dt, "steps": sd_steps,
                      "gpus_driven": world, "verified": ok,
                      "what": f"one process, one asyncio loop, {world} executors cuda:0..{world - 1}, electron i -> GPU i mod {world}"}
            del xall
        dist.cpu_barrier()

    # ============ c2: 10 000 scalar-int electrons (rank 0, one GPU) =================================
    c2 = None
    if rank == 0 and not args.quick:
        ints = [(i, 2 * i + 1) for i in range(n)]

        async def int_step(tag):
            return await asyncio.gather(*[
                ex.run(electrons.add_int, [a, b], {}, {"dispatch_id": tag, "node_id": i}) for i, (a, b) in enumerate(ints)])

        asyncio.run(int_step("iw"))
        t0 = time.perf_counter()
        reps = 3
        exact = True
        for s in range(reps):
            out = asyncio.run(int_step(f"i{s}"))
            exact = exact and all(type(v) is int for v in out[::97]) and out == [a + b for a, b in ints]
        dt = (time.perf_counter() - t0) / reps
        c2 = {"workload": "C2: 10 000 independent add_int(i, 2i+1) electrons, one GPU", "value": n / dt,
              "unit": "electrons/s", "ms_per_step": 1e3 * dt, "exact": bool(exact)}
    if world > 1:
        dist.cpu_barrier()

    # ============ c5: payload sweep on one GPU (rank 0) ============================================
    c5 = None
    if rank == 0 and not args.quick:
        try:
            c5 = sweep_c5(L, eng, ex, electrons, spec, timed, peak, dev, heap0, 2 * total + 512 * MiB)
        except Exception as err:          # an auxiliary block never costs the headline line
            c5 = {"error": repr(err)[:300]}
    if world > 1:
        dist.cpu_barrier()

    # ============ peer / bcast (N > 1) ==============================================================
    peer, bcast = None, None
    if world > 1:
        ids = D.exchange_peer_arenas(dev, eng.arena_bytes)
        nxt, prv = (rank + 1) % world, (rank - 1) % world
        m = 2000
        src_payloads = np.random.default_rng([5, rank]).integers(0, 256, (m, payload), dtype=np.uint8)
        L.call("b2_write", dev, heap0, src_payloads.ctypes.data, src_payloads.nbytes)
        order = np.random.default_rng(2).permutation(m).astype(np.uint64)
        rows = np.zeros((m, 3), dtype=np.uint64)
        rows[:, 0] = heap0 + order * payload
        rows[:, 1] = run_off + np.arange(m, dtype=np.uint64) * payload      # contiguous in the peer's heap
        rows[:, 2] = payload

        def do_peer():
            L.call("b2_peer_pack", dev, ids[nxt], rows.ctypes.data, m)

        for _ in range(3):
            do_peer()
        L.call("b2_sync", dev)
        dist.barrier()
        ms_peer = dist.max(timed(do_peer, 20))
        dist.barrier()
        # byte-exact check of what the previous rank packed into THIS rank's arena
        landed = np.empty((m, payload), dtype=np.uint8)
        L.call("b2_read", dev, run_off, landed.ctypes.data, landed.nbytes)
        theirs = np.random.default_rng([5, prv]).integers(0, 256, (m, payload), dtype=np.uint8)
        peer_ok = dist.all_true(bool(np.array_equal(landed, theirs[order.astype(np.int64)])))
        gbs = m * payload / 1e9 / (ms_peer / 1e3)
        peer = {"kernel": "span_move_tma_kernel writing into the next rank's arena (CUDA IPC mapping, NVLink)",
                "gbs_per_direction": gbs, "all_ranks_sending": True, "of_measured_peer_copy_770": gbs / 770.0,
                "of_nominal_900": gbs / 900.0, "payloads": m, "payload_bytes": payload, "verified": peer_ok}
        # shared operand: uploaded on rank 0, ncclBroadcast arena -> arena, used as an electron argument
        shared = np.random.default_rng(11).standard_normal(64 * MiB // 4, dtype=np.float32) if rank == 0 else None
        warm = D.broadcast_shared(eng, shared[:1 << 18] if rank == 0 else None, root=0)   # NCCL setup, not timed
        warm.free()
        dist.barrier()
        bstats = {}
        t0 = time.perf_counter()
        da = D.broadcast_shared(eng, shared, root=0, stats=bstats)
        dist.barrier()
        dt_b = dist.max(time.perf_counter() - t0)
        dt_coll = dist.max(bstats.get("collective_seconds", 0.0))
        # ... and used as an electron argument on every rank (first 4 MiB of it: one electron's operands
        # have to fit a staging segment of this bench's engine)
        use = 4 * MiB
        mine_x = np.random.default_rng([12, rank]).standard_normal(use // 4, dtype=np.float32)
        bc_error = None
        try:
            got = asyncio.run(ex.run(electrons.vec_add, [da.view(0, use, (use // 4,)), mine_x], {},
                                     {"dispatch_id": "bc", "node_id": rank}))
            ref = np.random.default_rng(11).standard_normal(64 * MiB // 4, dtype=np.float32)[:use // 4] + mine_x
            same = bool(np.array_equal(got, ref))
            # the rest of the broadcast bytes: read back and compared on every rank
            whole = eng.get(da)
            same = same and bool(np.array_equal(whole, np.random.default_rng(11).standard_normal(64 * MiB // 4, dtype=np.float32)))
        except Exception as err:
            got = ref = None
            same, bc_error = False, repr(err)[:200]
        bcast = {"bytes": 64 * MiB, "seconds": dt_b, "gbs": 64 * MiB / 1e9 / dt_b,
                 "collective_seconds": dt_coll, "collective_gbs": (64 * MiB / 1e9 / dt_coll) if dt_coll else None,
                 "exact": dist.all_true(same), "error": bc_error,
                 "what": "one 64 MiB operand uploaded on rank 0 (pageable host memory), ncclBroadcast arena->arena, every "
                         "byte read back and compared on every rank, its first 4 MiB used by a vec_add electron on every "
                         "rank; `seconds` = upload + collective, `collective_seconds` = the ncclBroadcast alone (max over ranks)"}
        da.free()
        del got, ref, mine_x

    # ============ c4: 1 000-node diamond, 1 MiB operands, across the N GPUs (rank 0 drives) ========
    c4 = None
    if world > 1:
        dist.cpu_barrier()
    if rank == 0 and not args.quick:
        try:
            c4 = diamond_c4(world, key, tmp)
        except Exception as err:
            c4 = {"error": repr(err)[:300], "exact": False}
    if world > 1:
        dist.cpu_barrier()

    # ============ CPU baseline (rank 0, N = 1 only) =================================================
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)          # the reference gets every host core
        cores = len(os.sched_getaffinity(0))
        base = reference_cpu_sample(max(32, min(128, cores)), payload, passes=3)
        cpu = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample", "per_pass")}
        try:
            cpu["t_floor"] = t_floor_sample(4096, payload)
        except Exception as err:     # never lose the line over the auxiliary number
            cpu["t_floor"] = {"error": repr(err)}
        if not args.quick:
            try:
                cpu["other_configs"] = reference_other_configs()
            except Exception as err:
                cpu["other_configs"] = {"error": repr(err)[:300]}

    if rank == 0:
        line = {
            "metric": "electrons/sec on 10k-task lattice", "value": value,
            "unit": "electrons/s", "n_gpus": world, "steps": e2e_steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * e2e_s, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C3", "electrons": n, "payload_bytes": payload,
                       "electron": "scale_add (2x+1) fp32, 64 KiB in / 64 KiB out", "max_batch": max_batch,
                       "placement": f"electron i -> GPU i mod {world}",
                       "parallelism": (f"one 10 000-electron lattice sharded over {world} dispatcher process(es), one per GPU "
                                       f"({len(mine)} electrons on rank 0)"),
                       "l2": "every step touches 1.31 GB of distinct payload bytes (inputs and outputs larger than L2)",
                       "timed_through": "SSHExecutor.run with host numpy arrays; every result checked by checksum in the timed steps"},
            "e2e": {"value": value, "unit": "electrons/s", "h2d_bytes_per_step": int(h2d) * world,
                    "d2h_bytes_per_step": int(d2h) * world, "ms_per_step": 1e3 * e2e_s, "steps": e2e_steps,
                    "verified": verified,
                    "ms_per_rank": {"min": 1e3 * min(e2e_ranks), "median": 1e3 * statistics.median(e2e_ranks),
                                    "max": 1e3 * e2e_s, "all": [round(1e3 * t, 3) for t in e2e_ranks]}},
            "gpu_launches": e2e_launches,
            "single_dispatcher": single,
            "roofline": {"bound": "hbm", "kernel": "b2::worker_tma_kernel (fused scale_add electron: 64 KiB in, 64 KiB out)",
                         "achieved": worker_gbs, "peak": peak, "unit": "GB/s", "frac": worker_gbs / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": 2.0 * total / len(tabs),
                         "measured_in": "kernels.worker_step (CUDA events on the library's stream, max over ranks)",
                         "launches_timed": dev_launches},
            "kernels": kernels,
            "c2": c2, "c4": c4, "c5": c5,
            "clocks": clk,
            "cpu_baseline": cpu,
            "peer": peer, "bcast": bcast,
            "lib_src_sha": lib_sha,
            "device": eng.info.name.decode(),
        }
        print(json.dumps(line), flush=True)
    shutdown_all()


def sweep_c5(L, eng, ex, electrons, spec, timed, peak, dev, heap0, heap_used):
    """BASELINE config 5 on one GPU: per payload size the worker kernel on device-resident operands
    (GB/s against the HBM copy peak; launched from a resident task table, so that the number is the
    kernel's and not the host's table preparation — `launch_path_gbs` is the b2_enqueue + upload +
    launch path beside it) and the end-to-end rate through SSHExecutor.run."""
    rows = []
    src0 = heap0
    budget = eng.heap_bytes // 2 - 32 * MiB             # operand bytes per cell (in), same again for results
    budget = min(budget, 4 * GiB) & ~(64 * MiB - 1)
    dst0 = heap0 + budget
    noise = np.random.default_rng(0).standard_normal(16 * MiB, dtype=np.float32)        # 64 MiB
    L.call("b2_write", dev, src0, noise.ctypes.data, noise.nbytes)
    reps_fill = budget // noise.nbytes
    spans = np.zeros((reps_fill - 1, 3), dtype=np.uint64)
    spans[:, 0] = src0
    spans[:, 1] = src0 + np.arange(1, reps_fill, dtype=np.uint64) * noise.nbytes
    spans[:, 2] = noise.nbytes
    L.call("b2_move_spans", dev, spans.ctypes.data, len(spans))
    L.call("b2_sync", dev)
    ticket = C.c_uint64()
    qcap = 1 << 16
    slice_bytes = 1 * MiB                                # engine._append slices elementwise electrons >= 2 MiB
    for P in (1 * KiB, 4 * KiB, 16 * KiB, 64 * KiB, 256 * KiB, 1 * MiB, 4 * MiB, 16 * MiB):
        slot = (P + 255) & ~255
        N = int(min(100000, budget // slot))
        per = P if P < 2 * slice_bytes else slice_bytes
        parts = P // per
        n_tasks = N * parts
        tabs = []
        for lo in range(0, n_tasks, qcap):
            k = min(qcap, n_tasks - lo)
            idx = lo + np.arange(k, dtype=np.uint64)
            tab = np.zeros(k, dtype=L.TASK_DTYPE)
            tab["fn_id"] = spec.fn_id
            tab["arg0"] = src0 + (idx // parts) * slot + (idx % parts) * per
            tab["out"] = dst0 + (idx // parts) * slot + (idx % parts) * per
            tab["count"] = per // 4
            tabs.append(tab)

        def launch_path():
            for tab in tabs:
                L.call("b2_enqueue", dev, tab.ctypes.data, len(tab), None, 0, None)
                L.call("b2_launch", dev, L.NO_STATUS_COPY, C.byref(ticket))

        tids = []
        for tab in tabs:
            tid = C.c_uint32()
            L.call("b2_table_create", dev, tab.ctypes.data, len(tab), None, 0, C.byref(tid))
            tids.append(tid.value)

        def resident():
            for tid in tids:
                L.call("b2_table_launch", dev, tid, L.NO_STATUS_COPY, C.byref(ticket))

        resident()
        got = np.empty(P // 4, dtype=np.float32)
        L.call("b2_read", dev, dst0 + (N - 1) * slot, got.ctypes.data, P)
        off = ((N - 1) * slot) % noise.nbytes
        exact = None
        if off + P <= noise.nbytes:
            exact = bool(np.array_equal(got, np.float32(2.0) * noise.view(np.uint8)[off:off + P].view(np.float32)
                                        + np.float32(1.0)))
        resident()
        ms = timed(resident, 10)
        launch_path()
        ms_lp = timed(launch_path, 10)
        for tid in tids:
            L.call("b2_table_destroy", dev, tid)
        gbs = 2.0 * N * P / 1e9 / (ms / 1e3)
        row = {"payload_bytes": P, "electrons": N, "task_records": n_tasks, "worker_gbs": gbs, "worker_frac": gbs / peak,
               "worker_electrons_per_s": N / (ms / 1e3), "exact": exact,
               "launch_path_gbs": 2.0 * N * P / 1e9 / (ms_lp / 1e3),
               "launch_path_frac": 2.0 * N * P / 1e9 / (ms_lp / 1e3) / peak}
        # end to end: bounded by the staging pool of this bench's engine
        n_e2e = int(min(10000, (256 * MiB) // max(P, 1), eng.max_batch * 3 if P > 256 * KiB else 10000))
        if P <= eng.seg_bytes // 2 and n_e2e >= 8:
            xs = [np.resize(noise[i::7], P // 4).copy() for i in range(min(n_e2e, 32))]

            async def go():
                return await asyncio.gather(*[
                    ex.run(electrons.scale_add, [xs[i % len(xs)]], {}, {"dispatch_id": f"c5_{P}", "node_id": i})
                    for i in range(n_e2e)])

            asyncio.run(go())
            t0 = time.perf_counter()
            out = asyncio.run(go())
            L.call("b2_sync", dev)
            dt = time.perf_counter() - t0
            row["e2e_exact"] = bool(np.array_equal(out[-1], np.float32(2.0) * xs[(n_e2e - 1) % len(xs)] + np.float32(1.0)))
            del out
            row.update(e2e_electrons=n_e2e, e2e_electrons_per_s=n_e2e / dt, e2e_gbs_in_plus_out=2.0 * n_e2e * P / dt / 1e9)
        else:
            row["e2e_skipped"] = "operand larger than half a staging segment of this bench's engine (sized for 256 x 64 KiB batches)"
        rows.append(row)
    return {"workload": "C5: payload sweep, one GPU, scale_add electrons; worker on device-resident operands "
                        f"({budget // MiB} MiB of payloads per cell, at most 100 000 electrons; electrons >= 2 MiB go out as "
                        "1 MiB task records, as the engine slices them), e2e through SSHExecutor.run",
            "rows": rows}


def diamond_c4(world: int, key, tmp):
    """BASELINE config 4: root -> 998 fan-out (vec_mul with a private 1 MiB weight each) -> reduce,
    fan-out placed round-robin over the GPUs this process drives; checked bit for bit against the
    electron bodies run in process."""
    from covalent_ssh_plugin_b200 import SSHExecutor, electrons as E, lattice as ct
    from covalent_ssh_plugin_b200.engine import existing_engine, get_engine

    for d in range(world):
        if existing_engine(d) is None:
            get_engine(d, arena_bytes=6 << 30, staging_bytes=256 << 20, n_segments=4, max_batch=1024)
    exs = [SSHExecutor(username="bench", hostname=f"cuda:{d}", ssh_key_file=str(key), cache_dir=tmp)
           for d in range(world)]
    n_fan, nel = 998, 262144
    root = ct.electron(E.plus_one, executor=exs[0])
    mul = [ct.electron(E.vec_mul, executor=exs[i % world]) for i in range(n_fan)]
    reduce_ = ct.electron(E.stack_sum, executor=exs[0])
    rng = np.random.default_rng(4)
    x0 = rng.random(nel, dtype=np.float32)
    ws = [rng.random(nel, dtype=np.float32) for _ in range(n_fan)]

    @ct.lattice
    def wf(x, weights):
        r = root(x)
        return reduce_([mul[i](r, weights[i]) for i in range(n_fan)])

    times, exact, stats = [], True, None
    want = E.stack_sum([E.vec_mul(E.plus_one(x0), w) for w in ws])
    for rep in range(3):
        t0 = time.perf_counter()
        r = ct.get_result(ct.dispatch(wf)(x0, ws))
        times.append(time.perf_counter() - t0)
        ok = r.status == ct.status.COMPLETED and bool(np.array_equal(r.result, want))
        exact = exact and ok
        stats = {k: v for k, v in r.stats.items() if isinstance(v, (int, float, str))}
        if not ok:
            stats["error"] = repr(getattr(r, "error", None))[:300]
    return {"workload": f"C4: 1 000-node diamond (root -> 998 x vec_mul -> stack_sum), 1 MiB fp32 operands, fan-out "
                        f"round-robin over {world} GPU(s) of one process",
            "ms": 1e3 * min(times[1:]) if len(times) > 1 else 1e3 * times[0], "ms_first": 1e3 * times[0],
            "exact": exact, "gpus": world, "stats": stats}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--electrons", type=int, default=N_ELECTRONS)
    ap.add_argument("--payload", type=int, default=PAYLOAD)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="profiling runs: skip the C2 / C4 / C5 blocks")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup
    if args.impl == "reference":
        # no GPU work on this arm: rank 0 alone measures, everybody else leaves
        if int(os.environ.get("RANK", "0")) != 0:
            return
        run_reference_arm(args)
        return
    dist = Dist(args.gpus)
    try:
        run_cuda_arm(args, dist)
    finally:
        dist.close()


if __name__ == "__main__":
    main()
