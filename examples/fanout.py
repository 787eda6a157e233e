"""A 10 000-electron fan-out on one B200, written the way a Covalent user would write it
(README.md:52-65 of the reference: construct the executor, decorate, call).

    python examples/fanout.py [n_electrons]
"""
import asyncio
import os
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ.setdefault("HOME", "/root")

import numpy as np

from covalent_ssh_plugin import SSHExecutor                    # same import path as the reference
from covalent_ssh_plugin_b200 import device_electron, lattice as ct


@device_electron("scale_add", alpha=2.0, beta=1.0)             # registers the device callback
def scale_add(x):                                              # the body stays the host definition
    return np.float32(2.0) * x + np.float32(1.0)


def main(n: int) -> None:
    key = Path(tempfile.mkdtemp()) / "id_stub"
    key.touch()
    gpu = SSHExecutor(username="me", hostname="cuda:0", ssh_key_file=str(key))

    # 1. what Covalent's dispatcher does: one awaited run() per electron
    xs = [np.random.default_rng(i).standard_normal(16384, dtype=np.float32) for i in range(n)]

    async def fan_out():
        return await asyncio.gather(*[
            gpu.run(scale_add, [x], {}, {"dispatch_id": "example", "node_id": i}) for i, x in enumerate(xs)])

    asyncio.run(fan_out())                                     # first call opens the engine
    t0 = time.perf_counter()
    ys = asyncio.run(fan_out())
    dt = time.perf_counter() - t0
    assert all(np.array_equal(y, scale_add(x)) for x, y in zip(xs[:50], ys[:50]))
    print(f"{n} electrons x 64 KiB through SSHExecutor.run: {n / dt / 1e3:.1f} k electrons/s")

    # 2. the same through the lattice driver (one device DAG, results gathered with one pack + D2H)
    step = ct.electron(scale_add, executor=gpu)

    @ct.lattice
    def workflow(arrays):
        return [step(step(a)) for a in arrays]                 # two dependent electrons per input

    t0 = time.perf_counter()
    res = ct.get_result(ct.dispatch(workflow)(xs[:2000]))
    dt = time.perf_counter() - t0
    assert res.status == ct.status.COMPLETED, res.error
    assert np.array_equal(res.result[7], scale_add(scale_add(xs[7])))
    print(f"lattice of {2 * 2000} dependent electrons: {dt * 1e3:.0f} ms, {res.stats}")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 10000)
