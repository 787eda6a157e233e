"""Property tests (hypothesis) of the host-side allocators and layouts."""

import numpy as np
from hypothesis import given, settings, strategies as st

from covalent_ssh_plugin_b200.engine import ALIGN, Heap
from oracle import device_model as M


@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(st.booleans(), st.integers(0, 50000)), min_size=1, max_size=120))
def test_heap_never_overlaps_and_coalesces(ops):
    heap = Heap(4096, 1 << 22)
    live = {}
    for is_free, n in ops:
        if is_free and live:
            off = sorted(live)[n % len(live)]
            heap.release(off)
            del live[off]
        else:
            try:
                off = heap.alloc(n)
            except MemoryError:
                continue
            size = max((n + ALIGN - 1) & ~(ALIGN - 1), ALIGN)
            assert off % ALIGN == 0 and 4096 <= off and off + size <= 4096 + (1 << 22)
            for o, z in live.items():
                assert off + size <= o or o + z <= off            # no overlap with any live block
            live[off] = size
    for off in list(live):
        heap.release(off)
    assert heap.free == [(4096, 1 << 22)] and heap.bytes_in_use() == 0


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(0, 5000), min_size=0, max_size=60), st.sampled_from([1, 16, 256]),
       st.integers(0, 4096))
def test_pack_layout_is_contiguous_and_aligned(sizes, align, base):
    base = (base + align - 1) & ~(align - 1)
    offs, total = M.pack_layout(sizes, base, align)
    cur = base
    for o, n in zip(offs, sizes):
        assert o == cur and o % align == 0
        cur += (n + align - 1) & ~(align - 1)
    assert total == cur - base


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 2 ** 31), st.lists(st.integers(0, 3000), min_size=1, max_size=40))
def test_pack_unpack_model_round_trip(seed, sizes):
    rng = np.random.default_rng(seed)
    slot = 3072
    n = len(sizes)
    arena = np.zeros(3 * n * slot + 4096, dtype=np.uint8)
    arena[:n * slot] = rng.integers(0, 256, n * slot, dtype=np.uint8)
    src = [int(i) * slot for i in rng.permutation(n)]
    dst = [2 * n * slot + int(i) * slot for i in rng.permutation(n)]
    M.pack(arena, src, sizes, n * slot, 16)
    M.unpack(arena, n * slot, dst, sizes, 16)
    for s, d, z in zip(src, dst, sizes):
        assert np.array_equal(arena[s:s + z], arena[d:d + z])
