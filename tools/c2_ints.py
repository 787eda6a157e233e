"""BASELINE config 2: 10 000 independent scalar-int electrons on one GPU — end-to-end rate through
SSHExecutor.run and the device-side rate of the same task table; every result checked exactly."""
import asyncio, ctypes as C, os, sys, tempfile, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
os.environ.setdefault("HOME", "/root")
import numpy as np
from covalent_ssh_plugin_b200 import SSHExecutor, _b2lib as L, electrons
from covalent_ssh_plugin_b200.engine import get_engine

n = 10000
eng = get_engine(0)
tmp = tempfile.mkdtemp(); key = Path(tmp) / "k"; key.touch()
ex = SSHExecutor(username="u", hostname="cuda:0", ssh_key_file=str(key), cache_dir=tmp)

async def step(tag):
    return await asyncio.gather(*[ex.run(electrons.add_int, [i, 2 * i + 1], {}, {"dispatch_id": tag, "node_id": i}) for i in range(n)])

asyncio.run(step("w"))
best = 1e9
for r in range(5):
    t0 = time.perf_counter(); out = asyncio.run(step(f"s{r}")); best = min(best, time.perf_counter() - t0)
assert out == [3 * i + 1 for i in range(n)] and all(type(v) is int for v in out)
print(f"C2 e2e: {n / best / 1e3:.1f} k electrons/s ({best / n * 1e6:.1f} us per electron), results exact")

spec = electrons.lookup(electrons.add_int)
tab = np.zeros(n, dtype=L.TASK_DTYPE)
tab["fn_id"], tab["flags"] = spec.fn_id, L.F_IMM0 | L.F_IMM1
tab["arg0"], tab["arg1"] = np.arange(n), 2 * np.arange(n) + 1
tab["out"], tab["count"] = eng.heap_base + 16 * np.arange(n), 1
ev = [C.c_uint32(), C.c_uint32()]
for e in ev: L.call("b2_event_create", 0, C.byref(e))
t = C.c_uint64()
def launch():
    L.call("b2_enqueue", 0, tab.ctypes.data, n, None, 0, None); L.call("b2_launch", 0, L.NO_STATUS_COPY, C.byref(t))
for _ in range(3): launch()
L.call("b2_sync", 0); L.call("b2_event_record", 0, ev[0].value)
for _ in range(20): launch()
L.call("b2_event_record", 0, ev[1].value); ms = C.c_float(); L.call("b2_event_elapsed_ms", 0, ev[0].value, ev[1].value, C.byref(ms))
got = np.empty(2 * n, dtype=np.int64); L.call("b2_read", 0, eng.heap_base, got.ctypes.data, 16 * n)
assert np.array_equal(got[::2], 3 * np.arange(n) + 1)
print(f"C2 device: {n / (ms.value / 20 / 1e3) / 1e6:.1f} M electrons/s ({ms.value / 20 * 1e3:.1f} us per 10 000-electron launch)")
