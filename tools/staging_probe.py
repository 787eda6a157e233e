"""How fast can host threads fill pinned staging memory?  (numpy slice copies release the GIL)"""
import ctypes as C, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("HOME", "/root")
import numpy as np
from covalent_ssh_plugin_b200 import _b2lib as L

L.call("b2_init", 0, 1 << 30, 1 << 30, 16 << 20, 1 << 12)
ptr, n = C.c_void_p(), C.c_uint64()
L.call("b2_host_buffer", 0, 0, C.byref(ptr), C.byref(n))
dst = np.frombuffer((C.c_uint8 * n.value).from_address(ptr.value), dtype=np.uint8)
for size in (64 << 10, 1 << 20, 16 << 20):
    k = (1 << 30) // size
    srcs = [np.random.default_rng(i).integers(0, 255, size, dtype=np.uint8) for i in range(min(k, 256))]
    def work(tid, nt):
        for i in range(tid, k, nt):
            dst[i * size:(i + 1) * size] = srcs[i % len(srcs)]
    for nt in (1, 2, 4, 8, 16):
        best = 0
        for _ in range(3):
            ths = [threading.Thread(target=work, args=(t, nt)) for t in range(nt)]
            t0 = time.perf_counter(); [t.start() for t in ths]; [t.join() for t in ths]
            best = max(best, (1 << 30) / (time.perf_counter() - t0) / 1e9)
        print(f"payload {size >> 10:6d} KiB  threads {nt:2d}: {best:6.1f} GB/s into pinned memory", flush=True)
L.call("b2_shutdown", 0)
