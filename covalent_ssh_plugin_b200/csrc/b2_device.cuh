// b2_device.cuh — device-side building blocks shared by the span mover (K1/K2/K4)
// and the persistent worker (K3).  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b2exec.h"

namespace b2 {

constexpr int kWarp = 32;
constexpr int kMoveThreads = 256;          // 8 warps per CTA in the span mover
#ifndef B2_WORKER_THREADS
#define B2_WORKER_THREADS 256
#endif
constexpr int kWorkerThreads = B2_WORKER_THREADS;   // worker warps x 32 per CTA in the worker kernel
constexpr int kCtaThreads = kWorkerThreads + kWarp;   // + 1 scheduler warp
constexpr uint32_t kChunkBytes = 16384;    // one warp work unit in the span mover

// ---- streaming 128-bit accessors ------------------------------------------
// Payload bytes are touched exactly once per kernel, so keep them out of L1
// (they would only push useful lines out).
__device__ __forceinline__ int4 ld_stream(const int4 *p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// Coherent variant: operands that another CTA of the SAME launch may have
// produced (dependency edges inside the worker) must not use the .nc path.
__device__ __forceinline__ int4 ld_coherent(const int4 *p) {
  int4 r;
  asm volatile("ld.global.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void st_stream(int4 *p, const int4 &v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// ---- streaming 256-bit accessors (new on sm_100: LDG.E.256 / STG.E.256) -------
struct alignas(32) Vec32 {
  unsigned long long a, b, c, d;
};
__device__ __forceinline__ Vec32 ld_stream32(const Vec32 *p) {
  Vec32 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0,%1,%2,%3}, [%4];"
               : "=l"(r.a), "=l"(r.b), "=l"(r.c), "=l"(r.d)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream32(Vec32 *p, const Vec32 &v) {
  asm volatile("st.global.L1::no_allocate.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(v.a), "l"(v.b),
               "l"(v.c), "l"(v.d)
               : "memory");
}

// ---- 64-bit acquire / release on queue cells --------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---------------------------------------------------------------------------
// Span table entry as the mover kernel reads it (32 B, built by the host
// runtime from b2_span + arena base pointers).
struct SpanDev {
  const uint8_t *src;
  uint8_t *dst;
  uint64_t bytes;
  uint32_t chunk_base;   // index of this span's first 16 KiB chunk
  uint32_t pad;
};

// Callback-table entry (global memory, B2_MAX_FN of them per device).
struct FnEntry {
  uint32_t kind;         // b2_kind
  float alpha, beta;
  uint32_t pad;
  long long ialpha, ibeta;
};

// MPMC queue of ready electrons (bounded: one launch never enqueues more than
// cap electrons, and every electron is enqueued exactly once per launch).
// Producers and consumers both take tickets with one fetch-add, so neither
// side retries under contention:
//   producer: t = P + tail++ ; cells[t] = (epoch << 32) | task_index  (release)
//   consumer: t = head++ ; if t >= n_tasks -> nothing left for this CTA, leave
//             if t < P   -> task t itself (the ready prefix bypasses the cells)
//             spin until cells[t] >> 32 == epoch                      (acquire)
// A ticket below n_tasks is always filled eventually, because exactly n_tasks
// pushes happen per launch; tickets are handed out in order, so every filled
// cell already has an owner that is running it (no lost wake-ups, no waiting
// on a CTA that has not been scheduled).  The last CTA to leave zeroes
// head/tail/cursor; the host bumps the epoch per launch, which is what makes
// stale cells from earlier launches unreadable.
struct Queue {
  unsigned long long *cells;
  uint32_t *head;                // consumer ticket (fetch-add)
  uint32_t *tail;                // producer ticket (fetch-add)
  uint32_t epoch;                // launch number, >= 1
  uint32_t mask;                 // cap - 1
};

// Dependent electrons carry deps = kDepBias + count in the device copy of the
// task table; the releasing decrement that takes it to exactly kDepBias fires
// the electron.  Electrons that are ready at submit carry 0, so the publishing
// prologue can tell "ready at submit" from "released a moment ago" and never
// pushes an electron twice.
constexpr uint32_t kDepBias = 0x80000000u;

struct WorkerArgs {
  b2_task *tasks;                // this batch's task table (deps are counted down in place)
  const uint32_t *succ;          // successor list
  uint32_t *status;              // one b2_task_status per electron
  const FnEntry *fn_table;
  Queue q;
  uint32_t *cursor;              // claim cursor over tasks [n_ready_prefix, n_tasks)
  uint32_t *exited;              // CTAs that have left; the last one re-arms the counters
  uint32_t n_tasks;
  uint32_t n_ready_prefix;       // tasks [0, P) are ready at submit: ticket t < P runs task t
  uint32_t claim_k;              // tickets a CTA takes per fetch-add on the head (>= 1; streaming worker)
  uint32_t pad0;
  uint8_t *arena[B2_MAX_ARENAS];
  uint64_t arena_bytes[B2_MAX_ARENAS];
};

}  // namespace b2
