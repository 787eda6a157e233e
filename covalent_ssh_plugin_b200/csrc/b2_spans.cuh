// b2_spans.cuh — K1 pack / K2 unpack / K4 peer pack: one batched span mover.
//
// Replaces the byte transport of the reference (asyncssh.scp upload
// ssh.py:360-361, download ssh.py:451): many small payloads are gathered into /
// scattered out of one contiguous HBM run so that a single DMA moves them
// across PCIe or NVLink.
//
// Work decomposition: every span is cut into 16 KiB chunks; chunk c of the
// launch belongs to warp (c mod total_warps).  Consecutive chunks therefore go
// to consecutive warps, which keeps DRAM pages busy in address order.  A warp
// moves its chunk with 128-bit streaming loads/stores, UNROLL independent
// loads in flight per lane before the first store.
//
// HBM-bound: algorithmic bytes = 2 * sum(span bytes) (read once, write once).
#pragma once
#include "b2_device.cuh"

namespace b2 {

// Largest s in [0, n) with base[s] <= c.  base has n + 1 entries, base[0] = 0,
// base[n] = total chunks, so c < base[n] always holds.
__device__ __forceinline__ uint32_t find_span(const uint32_t *__restrict__ base, uint32_t n,
                                              uint32_t c) {
  uint32_t lo = 0, hi = n;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(base + mid) <= c)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

template <int UNROLL>
__device__ __forceinline__ void warp_copy16(int4 *__restrict__ d, const int4 *__restrict__ s,
                                            uint32_t n16, int lane) {
  uint32_t i = lane;
  // full tiles: UNROLL x 32 lanes x 16 B
  for (; i + (UNROLL - 1) * kWarp < n16; i += UNROLL * kWarp) {
    int4 v[UNROLL];
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) v[j] = ld_stream(s + i + j * kWarp);
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) st_stream(d + i + j * kWarp, v[j]);
  }
  for (; i < n16; i += kWarp) st_stream(d + i, ld_stream(s + i));
}

// 256-bit flavour: UNROLL/2 x 32 lanes x 32 B per tile (same bytes in flight as the
// 128-bit loop, half the instructions).  Pointers must be 32 B aligned.
template <int UNROLL>
__device__ __forceinline__ uint32_t warp_copy32(uint8_t *dst, const uint8_t *src, uint32_t len,
                                                int lane) {
  constexpr int U = UNROLL / 2;
  const Vec32 *s = reinterpret_cast<const Vec32 *>(src);
  Vec32 *d = reinterpret_cast<Vec32 *>(dst);
  const uint32_t n32 = len >> 5;
  uint32_t i = lane;
  for (; i + (U - 1) * kWarp < n32; i += U * kWarp) {
    Vec32 v[U];
#pragma unroll
    for (int j = 0; j < U; ++j) v[j] = ld_stream32(s + i + j * kWarp);
#pragma unroll
    for (int j = 0; j < U; ++j) st_stream32(d + i + j * kWarp, v[j]);
  }
  for (; i < n32; i += kWarp) st_stream32(d + i, ld_stream32(s + i));
  return n32 << 5;   // bytes handled
}

// Moves len bytes; any alignment.  16 B aligned src/dst take the vector path
// end to end; equal misalignment peels a byte head; otherwise bytes.
template <int UNROLL, bool WIDE = false>
__device__ __forceinline__ void warp_move(uint8_t *dst, const uint8_t *src, uint32_t len,
                                          int lane) {
  if (WIDE && (((uintptr_t)src | (uintptr_t)dst) & 31u) == 0) {
    const uint32_t done = warp_copy32<UNROLL>(dst, src, len, lane);
    src += done;
    dst += done;
    len -= done;   // < 32 bytes left; falls through to the 128-bit / byte tail below
  }
  const uint32_t ms = (uint32_t)((uintptr_t)src & 15u);
  const uint32_t md = (uint32_t)((uintptr_t)dst & 15u);
  if (ms != md) {  // cold path: correctness only
    for (uint32_t i = lane; i < len; i += kWarp) dst[i] = src[i];
    return;
  }
  uint32_t head = ms ? (16u - ms) : 0u;
  if (head > len) head = len;
  if ((uint32_t)lane < head) dst[lane] = src[lane];
  src += head;
  dst += head;
  len -= head;
  const uint32_t n16 = len >> 4;
  warp_copy16<UNROLL>(reinterpret_cast<int4 *>(dst), reinterpret_cast<const int4 *>(src), n16,
                      lane);
  const uint32_t t = (n16 << 4) + (uint32_t)lane;  // < 16 trailing bytes
  if (t < len) dst[t] = src[t];
}

// uniform_cps > 0: every span has the same byte count and therefore the same
// number of chunks per span — the chunk -> span map is a division.
template <int UNROLL, bool WIDE = false>
__global__ void __launch_bounds__(kMoveThreads, (UNROLL >= 16 ? 2 : (UNROLL >= 8 ? 4 : 8)))
    span_move_kernel(const SpanDev *__restrict__ spans, const uint32_t *__restrict__ chunk_base,
                     uint32_t n_spans, uint32_t n_chunks, uint32_t uniform_cps) {
  const int lane = threadIdx.x & (kWarp - 1);
  const uint32_t warps_per_cta = kMoveThreads / kWarp;
  const uint32_t warp = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
  const uint32_t total_warps = gridDim.x * warps_per_cta;

  for (uint32_t c = warp; c < n_chunks; c += total_warps) {
    uint32_t s, k;
    if (uniform_cps) {
      s = c / uniform_cps;
      k = c - s * uniform_cps;
    } else {
      s = find_span(chunk_base, n_spans, c);
      k = c - __ldg(chunk_base + s);
    }
    // 32 B entry: two 16 B loads, broadcast within the warp by the L1/LSU.
    const int4 e0 = __ldg(reinterpret_cast<const int4 *>(spans + s));
    const int4 e1 = __ldg(reinterpret_cast<const int4 *>(spans + s) + 1);
    const uint8_t *src =
        reinterpret_cast<const uint8_t *>(((uint64_t)(uint32_t)e0.y << 32) | (uint32_t)e0.x);
    uint8_t *dst = reinterpret_cast<uint8_t *>(((uint64_t)(uint32_t)e0.w << 32) | (uint32_t)e0.z);
    const uint64_t bytes = ((uint64_t)(uint32_t)e1.y << 32) | (uint32_t)e1.x;
    const uint64_t off = (uint64_t)k * kChunkBytes;
    const uint64_t rem = bytes - off;
    const uint32_t len = rem < kChunkBytes ? (uint32_t)rem : kChunkBytes;
    warp_move<UNROLL, WIDE>(dst + off, src + off, len, lane);
  }
}

}  // namespace b2
