"""bench.py's reference arm and CPU-baseline helpers (no GPU): the line the driver parses, the rank
protocol under torchrun, and that the reference's electrons travel by value."""

from __future__ import annotations

import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))


def _run_bench(*argv, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, str(REPO / "bench.py"), *argv], capture_output=True, text=True,
                          timeout=600, env=e, cwd=str(REPO))


def test_reference_arm_prints_the_contract_line():
    r = _run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "electrons/s" and d["higher_is_better"] is True
    assert d["metric"] == "electrons/sec on 10k-task lattice" and d["config"]["workload"] == "C3"
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cpu = d["cpu_baseline"]
    assert cpu["kind"] in ("reference", "port") and cpu["cores"] >= 1 and cpu["value"] == d["value"]
    assert "by value" in cpu["sample"]


def test_reference_arm_is_rank_zero_only():
    r = _run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                   env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_electrons_are_pickled_by_value():
    """The reference's remote interpreter must not have to import this repository (bench.py or the
    package): that would be charged to the reference's time."""
    import cloudpickle

    import bench

    bench._by_value()
    blobs = [cloudpickle.dumps(fn) for fn in (bench._ref_scale_add, bench._ref_add_int, bench._ref_plus_one,
                                              bench._ref_vec_mul, bench._ref_stack_sum)]
    assert all(b"covalent_ssh_plugin_b200" not in blob for blob in blobs)
    # an interpreter that cannot see this repository loads and calls them (as exec.py would)
    code = ("import sys, pickle; fns = [pickle.loads(bytes.fromhex(h)) for h in sys.argv[1:]]; "
            "assert 'bench' not in sys.modules and 'covalent_ssh_plugin_b200' not in sys.modules; "
            "print(fns[1](3, 4))")
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    r = subprocess.run([sys.executable, "-c", code, *[b.hex() for b in blobs]], capture_output=True, text=True,
                       cwd="/tmp", env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == "7", r.stderr[-1000:]
    x = np.arange(8, dtype=np.float32)
    assert np.array_equal(bench._ref_scale_add(x), np.float32(2) * x + np.float32(1))
    assert np.array_equal(bench._ref_stack_sum([x, x, x]), x + x + x)


def test_reference_beside_the_other_configs():
    import bench

    o = bench.reference_other_configs(budget_s=60.0)
    assert o["kind"] in ("reference", "port")
    assert o["c2"]["exact"] is True and o["c2"]["value"] > 0
    assert o["c4"]["exact"] is True and o["c4"]["extrapolated_ms_1000_nodes"] > 0
    rows = o["c5"]["rows"]
    assert [r["payload_bytes"] for r in rows] == [1 << 10, 1 << 12, 1 << 14, 1 << 16, 1 << 18, 1 << 20, 1 << 22, 1 << 24]
    assert all(r.get("exact", True) for r in rows) and any("electrons_per_s" in r for r in rows)
