// b2_spans_tma.cuh — the span mover on the bulk-copy engine (TMA, 1-D):
//   global --cp.async.bulk--> shared --cp.async.bulk--> global
// No payload byte ever touches a register: each pipeline is ONE thread that arms an
// mbarrier with the chunk's byte count, issues the bulk load into a shared-memory
// stage, waits for the transaction to complete, and issues the bulk store from the
// same stage.  kTmaPipes independent pipelines per CTA, kTmaStages stages of one
// 16 KiB chunk each, one CTA per SM (192 KiB of dynamic shared memory).
//
// Eligibility (checked on the host): every span offset and byte count is a multiple of
// 16 — always true for payloads staged by the engine.  Anything else takes the
// register path in b2_spans.cuh.  Same work decomposition (16 KiB chunks, chunk c ->
// pipeline c mod P) and the same algorithmic bytes: 2 * sum(span bytes).
#pragma once
#include "b2_spans.cuh"

namespace b2 {

constexpr int kTmaPipes = 4;
constexpr int kTmaStages = 3;
constexpr int kTmaThreads = kTmaPipes * kWarp;
constexpr uint32_t kTmaSmemBytes = kTmaPipes * kTmaStages * kChunkBytes;   // 192 KiB

__device__ __forceinline__ uint32_t smem_addr(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          dst_smem),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
               "r"(src_smem), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

struct ChunkRef {
  const uint8_t *src;
  uint8_t *dst;
  uint32_t len;
};

__device__ __forceinline__ ChunkRef chunk_ref(const SpanDev *__restrict__ spans,
                                              const uint32_t *__restrict__ chunk_base,
                                              uint32_t n_spans, uint32_t uniform_cps, uint32_t c) {
  uint32_t s, k;
  if (uniform_cps) {
    s = c / uniform_cps;
    k = c - s * uniform_cps;
  } else {
    s = find_span(chunk_base, n_spans, c);
    k = c - __ldg(chunk_base + s);
  }
  const SpanDev e = spans[s];
  const uint64_t off = (uint64_t)k * kChunkBytes;
  const uint64_t rem = e.bytes - off;
  ChunkRef r;
  r.src = e.src + off;
  r.dst = e.dst + off;
  r.len = rem < kChunkBytes ? (uint32_t)rem : kChunkBytes;
  return r;
}

__global__ void __launch_bounds__(kTmaThreads, 1)
    span_move_tma_kernel(const SpanDev *__restrict__ spans, const uint32_t *__restrict__ chunk_base,
                         uint32_t n_spans, uint32_t n_chunks, uint32_t uniform_cps) {
  extern __shared__ __align__(128) uint8_t stage_mem[];
  __shared__ __align__(8) unsigned long long bars[kTmaPipes * kTmaStages];

  const int pipe = threadIdx.x >> 5;
  if ((threadIdx.x & 31) != 0) return;   // one thread per pipeline drives the copy engine

  const uint32_t total_pipes = gridDim.x * kTmaPipes;
  const uint32_t first = blockIdx.x * kTmaPipes + pipe;
  if (first >= n_chunks) return;
  const uint32_t n_my = (n_chunks - first + total_pipes - 1) / total_pipes;

  const uint32_t stage0 = smem_addr(stage_mem) + pipe * kTmaStages * kChunkBytes;
  const uint32_t bar0 = smem_addr(bars) + pipe * kTmaStages * 8;
#pragma unroll
  for (int s = 0; s < kTmaStages; ++s) mbar_init(bar0 + 8 * s, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

  auto load = [&](uint32_t i) {
    const ChunkRef r = chunk_ref(spans, chunk_base, n_spans, uniform_cps, first + i * total_pipes);
    const uint32_t s = i % kTmaStages;
    mbar_expect_tx(bar0 + 8 * s, r.len);
    bulk_g2s(stage0 + s * kChunkBytes, r.src, r.len, bar0 + 8 * s);
  };

  for (uint32_t i = 0; i < n_my && i < (uint32_t)kTmaStages; ++i) load(i);

  for (uint32_t i = 0; i < n_my; ++i) {
    const uint32_t s = i % kTmaStages;
    mbar_wait(bar0 + 8 * s, (i / kTmaStages) & 1u);
    const ChunkRef r = chunk_ref(spans, chunk_base, n_spans, uniform_cps, first + i * total_pipes);
    bulk_s2g(r.dst, stage0 + s * kChunkBytes, r.len);
    // Refill the stage the PREVIOUS store used: allowing one store group (this one) to be
    // outstanding means the previous store has finished reading its stage.
    if (i >= 1) {
      const uint32_t j = i - 1 + kTmaStages;
      if (j < n_my) {
        bulk_wait_read<1>();
        load(j);
      }
    }
  }
  bulk_wait_all();
}

}  // namespace b2
