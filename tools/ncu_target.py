#!/usr/bin/env python
"""Small, fixed workload for `ncu --set full` captures (one launch of each hot kernel at the
sizes bench.py reports): worker_tma_kernel over 10 000 x 64 KiB scale_add electrons,
worker_tma_small_kernel over 100 000 x 4 KiB, span_move_tma_kernel packing 10 000 x 64 KiB.

    ncu --set full --clock-control none --import-source on -k regex:worker_tma_kernel -s 2 -c 1 \
        -o gpurun_out/prof_worker python tools/ncu_target.py
"""
import ctypes as C
import os
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
os.environ.setdefault("HOME", "/root")

import numpy as np  # noqa: E402

from covalent_ssh_plugin_b200 import _b2lib as L  # noqa: E402

dev, GiB, MiB = 0, 1 << 30, 1 << 20
L.call("b2_init", dev, 6 * GiB, 64 * MiB, 64 * MiB, 1 << 17)
L.call("b2_register", dev, 1, L.K_SCALE_ADD, 2.0, 1.0, 0, 0)
noise = np.random.default_rng(0).standard_normal(16 * MiB, dtype=np.float32)
src0, dst0 = 256 * MiB, 256 * MiB + 2 * GiB
L.call("b2_write", dev, src0, noise.ctypes.data, noise.nbytes)
spans = np.zeros((31, 3), dtype=np.uint64)
spans[:, 0] = src0
spans[:, 1] = src0 + np.arange(1, 32, dtype=np.uint64) * noise.nbytes
spans[:, 2] = noise.nbytes
L.call("b2_move_spans", dev, spans.ctypes.data, len(spans))
L.call("b2_sync", dev)
t = C.c_uint64()
reps = int(os.environ.get("B2_NCU_REPS", 4))
for payload, n in ((65536, 10000), (4096, 100000)):
    tab = np.zeros(n, dtype=L.TASK_DTYPE)
    tab["fn_id"], tab["count"] = 1, payload // 4
    tab["arg0"] = src0 + np.arange(n, dtype=np.uint64) * payload
    tab["out"] = dst0 + np.arange(n, dtype=np.uint64) * payload
    tid = C.c_uint32()
    L.call("b2_table_create", dev, tab.ctypes.data, n, None, 0, C.byref(tid))
    for _ in range(reps):
        L.call("b2_table_launch", dev, tid.value, L.NO_STATUS_COPY, C.byref(t))
    L.call("b2_sync", dev)
n, payload = 10000, 65536
scat = src0 + np.random.default_rng(1).permutation(n).astype(np.uint64) * payload
sizes = np.full(n, payload, dtype=np.uint64)
for _ in range(reps):
    L.call("b2_pack", dev, scat.ctypes.data, sizes.ctypes.data, n, dst0, 16, None, None)
L.call("b2_sync", dev)
L.call("b2_shutdown", dev)
print("ncu target done")
