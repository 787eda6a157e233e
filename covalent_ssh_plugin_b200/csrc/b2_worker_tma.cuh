// b2_worker_tma.cuh — K3, streaming variant: the worker kernel with its operand loads on the
// bulk-copy engine.
//
// Same queue, same callback table, same successor release as worker_kernel (b2_worker.cuh).
// What changes is who issues the loads.  In worker_kernel the worker warps load their own
// operands, so every electron boundary (hand-off barrier, status, successor release) is a
// bubble in that CTA's load stream — at 64 KiB per electron that costs ~15 % of HBM
// bandwidth.  Here the scheduler warp is also the PRODUCER: it walks the electrons it claims
// tile by tile (16 KiB) and issues `cp.async.bulk` loads into a ring of kStages (3) shared-memory stages,
// running ahead ACROSS electron boundaries; the worker warps only consume full stages
// (LDS.128 -> fp32 op -> streaming STG.128) and hand them back.  The load stream of a CTA
// therefore never drains while there is work, whatever the workers are waiting for.
//
//   producer (1 thread)                                   workers (kWorkerThreads)
//   wait empty[s] ── fill hdr[s] ── expect_tx(full[s]) ──► wait full[s]
//   cp.async.bulk operand tile(s) -> stage[s]              read hdr[s], stage[s]; compute; STG
//   next tile / next electron ...                          last tile: status + successors
//                                          ◄── arrive empty[s] (one per worker warp)
//
// Elementwise kinds (COPY, SCALE_ADD, ADD, MUL, AXPY) stream through the ring; SUM,
// STACK_SUM, the integer electrons and the gate electron are handed over as one "legacy"
// stage and run the same callbacks as worker_kernel.
#pragma once
#include "b2_spans_tma.cuh"
#include "b2_worker.cuh"

namespace b2 {

#ifndef B2_TMA_STAGES
#define B2_TMA_STAGES 3
#endif
#ifndef B2_TMA_TILE
#define B2_TMA_TILE 16384
#endif
constexpr int kStages = B2_TMA_STAGES;
constexpr uint32_t kTileBytes = B2_TMA_TILE;          // one stage; binary electrons use two halves
constexpr uint32_t kWorkerSmemBytes = kStages * kTileBytes;
// Small-payload flavour: same shared-memory budget cut into more, smaller stages.  An electron of
// a few KiB occupies one stage whatever the stage size, so with 3 x 16 KiB stages a CTA has at
// most 3 small electrons (a few KiB) in flight — far below what HBM latency needs; 12 x 4 KiB
// stages keep 12 of them in flight per CTA (48 per SM).  Chosen by the host per launch from the
// average operand size (b2exec.cu); larger electrons are still correct there, tile by tile.
#ifndef B2_TMA_SMALL_STAGES
#define B2_TMA_SMALL_STAGES 12
#endif
#ifndef B2_TMA_SMALL_TILE
#define B2_TMA_SMALL_TILE 4096
#endif
constexpr int kSmallStages = B2_TMA_SMALL_STAGES;
constexpr uint32_t kSmallTileBytes = B2_TMA_SMALL_TILE;
constexpr uint32_t kWorkerSmallSmemBytes = kSmallStages * kSmallTileBytes;
constexpr uint32_t kHdrExit = 0xFFFFFFFFu;
constexpr uint32_t kHdrLegacy = 0xFFFFFFFEu;
constexpr uint32_t kHdrFailed = 0xFFFFFFFDu;          // status already decided by the producer
constexpr uint32_t kHdrBegin = 0xFFFFFFFCu;           // a batch starts: its BatchView travels in bview[s]
constexpr uint32_t kHdrEnd = 0xFFFFFFFBu;             // resident worker: this CTA has no more electrons of the batch

struct StageHdr {
  uint8_t *out;              // global destination of this tile
  const uint8_t *tail0;      // global source of the < 16 B tail (last tile only)
  const uint8_t *tail1;
  uint32_t n_vec;            // 128-bit vectors in this tile, per operand
  uint32_t tail_bytes;       // 0..15 bytes after the vectors (last tile only)
  uint32_t kind;             // b2_kind of an elementwise electron, or kHdr*
  uint32_t last;             // 1: last tile of its electron
  uint32_t task;             // task index
  uint32_t status;           // kHdrFailed: the status to report
  float alpha, beta;
  uint32_t succ_begin, succ_count;
  uint32_t owner;            // warp mode: the worker warp that consumes this stage (kOwnerAll: every warp)
  uint32_t pad;
};
constexpr uint32_t kOwnerAll = 0xFFu;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ bool is_streaming_kind(uint32_t k) {
  return k == B2_K_COPY || k == B2_K_SCALE_ADD || k == B2_K_ADD || k == B2_K_MUL || k == B2_K_AXPY;
}

// What both roles need to know about the batch they are working on.  In an ordinary launch there
// is one (built from the kernel arguments); the resident worker gets a new one per doorbell.
struct BatchView {
  b2_task *tasks;
  const uint32_t *succ;
  uint32_t *status;
  Queue q;                 // cells / head / tail / epoch / mask of THIS batch
  uint32_t *cursor;
  uint32_t *finished;      // resident: CTAs that have retired their share of the batch
  uint32_t n_tasks, prefix, claim_k, seq;
};

// ---- resident worker: control blocks -------------------------------------------------------
// The kernel stays on the GPU across batches.  A batch is published by the host with stream-ordered
// writes (task table, payload, ring entry, then the doorbell word) — no kernel launch — and its
// completion is a word the egress stream waits on (cuStreamWaitValue32) before it downloads.
constexpr uint32_t kResRing = 64;            // batches that may be published ahead of the worker
constexpr uint32_t kResNone = 0xFFFFFFFFu;

struct ResidentBatch {     // one ring entry: written by the host into mapped pinned memory, mirrored into ctl->ring by the leader
  b2_task *tasks;
  const uint32_t *succ;
  uint32_t *status;
  uint32_t n_tasks, prefix, claim_k, pad;
  uint64_t pad2[3];
};
struct ResidentQ {         // per-batch queue counters (zero when the slot is handed to a batch)
  uint32_t head, tail, cursor, finished;
  uint32_t pad[12];
};
struct ResidentCtl {       // device memory
  ResidentBatch ring[kResRing];
  ResidentQ q[kResRing];
  uint32_t doorbell;       // batches published so far        (host: stream memory op on the upload stream)
  uint32_t pad0[31];
  uint32_t approved;       // batches every CTA may start     (leader CTA)
  uint32_t exit_at;        // leave when reaching this batch  (leader CTA; kResNone = stay)
  uint32_t pad1[30];
  uint32_t done;           // batches completed               (last CTA of each batch; the egress stream waits on it)
  uint32_t pad2[31];
};
struct ResidentHost {      // mapped pinned host memory: the two words of the exit handshake
  volatile uint32_t submitted;   // host -> kernel: batches the host has submitted (written BEFORE their stream ops)
  volatile uint32_t stop;        // host -> kernel: leave as soon as nothing submitted is pending
  volatile uint32_t state;       // kernel -> host: (instance << 2) | 1 running | 2 leaving
  volatile uint32_t error;       // kernel -> host: 1 = a submitted batch never arrived (lifetime cap hit)
  volatile uint32_t done;        // kernel -> host: batches completed (the host may reuse ring entries below it)
  uint32_t pad[11];
  ResidentBatch ring[kResRing];  // host -> kernel: batch descriptors, plain CPU stores before the doorbell is rung
};
struct ResidentArgs {
  ResidentCtl *ctl;
  ResidentHost *host;      // device pointer of the mapped block
  uint32_t instance;
  uint32_t idle_us;        // leave after this long without a published batch ...
  uint32_t life_ms;        // ... and never stay longer than this (bounded even if the host misbehaves)
  uint32_t pad;
};

__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu_u32(uint32_t *p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---- cooperative producer (small-payload flavour) ------------------------------------------
// With electrons of a few KiB the single producer THREAD is the limiter: ~150 dependent instructions
// per electron (record, callback entry, bounds, header, expect_tx, bulk copy) at one warp's issue
// rate is ~1 us, however fast the workers are.  Here the producer WARP shares the work: a claim of K
// ready-at-submit tickets is handled by K lanes at once — each lane reads its own task record, plans
// its electron, an exclusive scan over the lanes' stage counts hands every lane its range of the
// stage ring, and each lane fills its own stages (it waits only on the ring slots it was dealt, so
// the lanes make progress independently).  Claims that touch released dependents (tickets >= the
// ready prefix) or the end of the table are walked by lane 0 alone, in ticket order.
struct ElectronPlan {
  FnEntry fn;
  const uint8_t *p0, *p1;
  uint8_t *po;
  uint64_t bytes;
  uint32_t kind;      // validated b2_kind, or B2_K__COUNT
  uint32_t mode;      // 0: streams through the ring, 1: one "legacy" stage, 2: status only
  uint32_t status;    // mode 2
  uint32_t need;      // stages it occupies
  bool two;
};

template <uint32_t kTileBytes>
__device__ __forceinline__ ElectronPlan plan_electron(const WorkerArgs &a, const b2_task &t) {
  ElectronPlan c;
  c.fn.kind = 0; c.fn.alpha = 0.f; c.fn.beta = 0.f; c.fn.pad = 0; c.fn.ialpha = 0; c.fn.ibeta = 0;
  c.p0 = c.p1 = nullptr;
  c.po = nullptr;
  c.bytes = 0;
  c.two = false;
  c.need = 1;
  c.status = B2_TS_OK;
  uint32_t kind = B2_K__COUNT;
  if (t.fn_id < B2_MAX_FN) {
    c.fn = a.fn_table[t.fn_id];
    kind = c.fn.kind;
    if (kind >= B2_K__COUNT || (kind == B2_K_NONE && t.fn_id != 0)) kind = B2_K__COUNT;
  }
  c.kind = kind;
  if (t.flags & B2_F_CANCELLED) { c.mode = 2; c.status = B2_TS_CANCELLED; return c; }
  if (kind == B2_K__COUNT) { c.mode = 2; c.status = B2_TS_BAD_FN; return c; }
  if (!is_streaming_kind(kind)) { c.mode = 1; return c; }
  c.two = (kind == B2_K_ADD || kind == B2_K_MUL || kind == B2_K_AXPY);
  c.bytes = (kind == B2_K_COPY) ? t.count : t.count * 4;
  bool ok = true;
  c.p0 = resolve(a, (t.flags >> B2_F_ARENA0_SHIFT) & 15u, t.arg0, c.bytes, &ok);
  c.p1 = c.two ? resolve(a, (t.flags >> B2_F_ARENA1_SHIFT) & 15u, t.arg1, c.bytes, &ok) : nullptr;
  c.po = resolve(a, (t.flags >> B2_F_ARENAO_SHIFT) & 15u, t.out, c.bytes, &ok);
  if ((t.flags & (B2_F_IMM0 | (c.two ? B2_F_IMM1 : 0u))) || !ok || !aligned16(c.p0) || !aligned16(c.po) ||
      (c.two && !aligned16(c.p1))) {
    c.mode = 2;
    c.status = B2_TS_BAD_ARG;
    return c;
  }
  c.mode = 0;
  const uint32_t per_tile = c.two ? kTileBytes / 2 : kTileBytes;
  c.need = c.bytes ? (uint32_t)((c.bytes + per_tile - 1) / per_tile) : 1u;
  return c;
}

template <int kStages, uint32_t kTileBytes>
__device__ __forceinline__ void coop_producer(const WorkerArgs &a, StageHdr *hdr, Slot *legacy, BatchView *bview,
                                              unsigned long long *full_bar, unsigned long long *empty_bar,
                                              uint8_t *stage_mem) {
  const uint32_t lane = threadIdx.x & 31u;
  constexpr uint32_t kWorkerWarps = kWorkerThreads / kWarp;
  constexpr unsigned kFull = 0xffffffffu;
  uint32_t it = 0, deal = 0;   // identical in every lane at the warp-synchronous points

  auto acquire_at = [&](uint32_t i) -> uint32_t {
    const uint32_t s = i % kStages;
    const uint32_t use = i / kStages;
    if (use >= 1) mbar_wait(smem_addr(&empty_bar[s]), (use - 1) & 1u);
    return s;
  };
  // fills the stages [i0, i0 + c.need) for one electron
  auto emit = [&](const b2_task &t, uint32_t ti, const ElectronPlan &c, uint32_t i0, uint32_t owner) {
    if (c.mode != 0) {
      const uint32_t s = acquire_at(i0);
      StageHdr &h = hdr[s];
      h.task = ti;
      h.last = 1;
      h.succ_begin = t.succ_begin;
      h.succ_count = t.succ_count;
      if (c.mode == 2) {
        h.kind = kHdrFailed;
        h.status = c.status;
        h.owner = owner;
      } else {
        h.kind = kHdrLegacy;
        h.owner = kOwnerAll;
        legacy[s].t = t;
        legacy[s].fn = c.fn;
        legacy[s].index = ti;
        legacy[s].kind = c.kind;
      }
      mbar_arrive(smem_addr(&full_bar[s]));
      return;
    }
    const uint32_t per_tile = c.two ? kTileBytes / 2 : kTileBytes;
    uint64_t off = 0;
    uint32_t i = i0;
    do {
      const uint64_t left = c.bytes - off;
      const uint32_t chunk = left < per_tile ? (uint32_t)left : per_tile;
      const bool last = (off + chunk == c.bytes);
      const uint32_t b16 = chunk & ~15u;
      const uint32_t s = acquire_at(i);
      StageHdr &h = hdr[s];
      h.out = c.po + off;
      h.tail0 = c.p0 + off + b16;
      h.tail1 = c.two ? c.p1 + off + b16 : nullptr;
      h.n_vec = b16 >> 4;
      h.tail_bytes = chunk - b16;
      h.kind = c.kind;
      h.owner = owner;
      h.last = last ? 1u : 0u;
      h.task = ti;
      h.alpha = c.fn.alpha;
      h.beta = c.fn.beta;
      h.succ_begin = t.succ_begin;
      h.succ_count = t.succ_count;
      const uint32_t fb = smem_addr(&full_bar[s]);
      if (b16) {
        const uint32_t dst = smem_addr(stage_mem) + s * kTileBytes;
        mbar_expect_tx(fb, c.two ? 2 * b16 : b16);
        bulk_g2s(dst, c.p0 + off, b16, fb);
        if (c.two) bulk_g2s(dst + kTileBytes / 2, c.p1 + off, b16, fb);
      } else {
        mbar_arrive(fb);
      }
      ++i;
      off += chunk;
    } while (off < c.bytes);
  };

  BatchView v;
  v.tasks = a.tasks;
  v.succ = a.succ;
  v.status = a.status;
  v.q = a.q;
  v.cursor = a.cursor;
  v.finished = nullptr;
  v.n_tasks = a.n_tasks;
  v.prefix = a.n_ready_prefix;
  v.claim_k = a.claim_k;
  v.seq = 0;
  if (lane == 0) {
    const uint32_t s = acquire_at(it);
    hdr[s].kind = kHdrBegin;
    hdr[s].owner = kOwnerAll;
    bview[s] = v;
    mbar_arrive(smem_addr(&full_bar[s]));
  }
  ++it;
  __syncwarp();

  const uint32_t P = v.prefix, N = v.n_tasks;
  // A claim is at most as wide as the stage ring: lanes that fill stages in parallel must each own a
  // DISTINCT ring slot.  (mbarrier parity waits only tell "the previous phase" from "the current one":
  // a lane two uses ahead of another on the same slot would sail through its wait.)
  const uint32_t kmax = kStages < 32 ? (uint32_t)kStages : 32u;
  const uint32_t K = v.claim_k < 1u ? 1u : (v.claim_k > kmax ? kmax : v.claim_k);
  const uint32_t lim = P < N ? P : N;
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = queue_claim_n(v.q, K);
    base = __shfl_sync(kFull, base, 0);
    if (base >= N) break;
    if (base + K <= lim) {
      // every ticket of this claim is an electron that was ready at submit: one lane each
      const bool act = lane < K;
      b2_task t = {};
      ElectronPlan c;
      c.need = 0;
      c.mode = 2;
      if (act) {
        t = v.tasks[base + lane];
        c = plan_electron<kTileBytes>(a, t);
      }
      uint32_t incl = c.need;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(kFull, incl, o);
        if ((int)lane >= o) incl += up;
      }
      const uint32_t total = __shfl_sync(kFull, incl, 31);
      if (total <= (uint32_t)kStages) {
        // the whole claim fits the ring once: every lane fills its own slots, all at the same time
        if (act) emit(t, base + lane, c, it + (incl - c.need), (deal + lane) % kWorkerWarps);
        __syncwarp();
      } else {
        // multi-tile electrons in the claim: the lanes take turns, in ticket order (planning — the
        // part with the memory latency in it — has still been done by all of them at once)
        for (uint32_t l = 0; l < K; ++l) {
          if (lane == l) emit(t, base + lane, c, it + (incl - c.need), (deal + lane) % kWorkerWarps);
          __syncwarp();
        }
      }
      it += total;
      deal += K;
    } else {
      if (lane == 0) {
        for (uint32_t j = 0; j < K; ++j) {
          const uint32_t ticket = base + j;
          if (ticket >= N) break;
          const uint32_t ti = ticket < P ? ticket : queue_take(v.q, ticket, N);
          // a released dependent reads what other SMs just wrote with ordinary stores: order the
          // acquire above before the copy engine's (async proxy) reads below
          asm volatile("fence.proxy.async.global;" ::: "memory");
          const b2_task t = v.tasks[ti];
          const ElectronPlan c = plan_electron<kTileBytes>(a, t);
          emit(t, ti, c, it, deal % kWorkerWarps);
          it += c.need;
          ++deal;
        }
      }
      it = __shfl_sync(kFull, it, 0);
      deal = __shfl_sync(kFull, deal, 0);
    }
  }
  if (lane == 0) {
    const uint32_t s = acquire_at(it);
    hdr[s].kind = kHdrExit;
    hdr[s].owner = kOwnerAll;
    mbar_arrive(smem_addr(&full_bar[s]));
  }
}

// kWarpMode (small-payload flavour): a streaming electron is consumed by ONE worker warp — stages are
// dealt to the eight warps round-robin, per electron — instead of by all 256 worker threads.  An
// electron of a few KiB is 64-256 vectors: spread over 256 threads it costs eight warps a barrier
// round trip, a status hand-off and ~100 instructions each; owned by one warp it costs that once, the
// other seven only look at the stage header and hand the stage back.  CTA-wide work (SUM, STACK_SUM,
// integer electrons, batch markers) is still done by all warps together.
template <int kStages, uint32_t kTileBytes, bool kResident, bool kWarpMode>
__device__ __forceinline__ void worker_tma_body(const WorkerArgs &a, const ResidentArgs &ra) {
  extern __shared__ __align__(128) uint8_t stage_mem[];   // kStages x kTileBytes, dynamic
  __shared__ StageHdr hdr[kStages];
  __shared__ Slot legacy[kStages];
  __shared__ BatchView bview[kStages];                    // travels with kHdrBegin
  __shared__ BatchView wv;                                // the batch the worker warps are on
  __shared__ __align__(8) unsigned long long full_bar[kStages];
  __shared__ __align__(8) unsigned long long empty_bar[kStages];
  __shared__ uint32_t s_base;
  __shared__ uint32_t s_status;
  constexpr int kWorkerWarps = kWorkerThreads / kWarp;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_addr(&full_bar[s]), 1);
      mbar_init(smem_addr(&empty_bar[s]), kWorkerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (threadIdx.x >= kWorkerThreads) {
    // ----------------------------------------------------------- producer warp --
    if (kWarpMode && !kResident) {
      coop_producer<kStages, kTileBytes>(a, hdr, legacy, bview, full_bar, empty_bar, stage_mem);
    } else if ((threadIdx.x & 31) == 0) {
      uint32_t it = 0;   // stages handed out so far
      auto acquire = [&]() -> uint32_t {
        const uint32_t s = it % kStages;
        const uint32_t use = it / kStages;
        if (use >= 1) mbar_wait(smem_addr(&empty_bar[s]), (use - 1) & 1u);
        return s;
      };
      auto marker = [&](uint32_t kind, const BatchView *v) {
        const uint32_t s = acquire();
        hdr[s].kind = kind;
        hdr[s].owner = kOwnerAll;
        if (v) bview[s] = *v;
        mbar_arrive(smem_addr(&full_bar[s]));
        ++it;
      };
      uint32_t deal = 0;   // warp mode: next warp to be dealt an electron

      // Streams every electron this CTA claims from batch v (tickets in increasing order).
      auto produce = [&](const BatchView &v) {
        const uint32_t P = v.prefix;
        // The producer always holds one claimed ticket beyond the electron it is streaming,
        // and — when that ticket lies in the ready prefix — its task record is already on its
        // way into registers: the fetch-add and the record's L2 round trips overlap the tile
        // issue of the current electron instead of sitting between two electrons.
        // Tickets are claimed claim_k at a time (one fetch-add on the shared head per K electrons:
        // with electrons of a few KiB the single-address atomic is otherwise the rate limiter).  A CTA
        // runs its tickets in increasing order, so a ticket that has to wait for its cell never sits
        // in front of a published one: cells are published in ticket order.
        const uint32_t K = v.claim_k ? v.claim_k : 1u;
        uint32_t claimed_end = 0;            // one past the last ticket of the current claim
        auto next_ticket = [&](uint32_t after) -> uint32_t {
          if (after + 1 < claimed_end) return after + 1;
          const uint32_t base = queue_claim_n(v.q, K);
          claimed_end = base + K;
          return base;
        };
        uint32_t ticket = queue_claim_n(v.q, K);
        claimed_end = ticket + K;
        b2_task t_next = {};
        if (ticket < P) t_next = v.tasks[ticket];
        for (;;) {
          const uint32_t cur = ticket;
          uint32_t ti;
          b2_task t;
          if (cur < P) {
            ti = cur;
            t = t_next;
          } else {
            ti = queue_take(v.q, cur, v.n_tasks);
            // a released dependent reads what other SMs just wrote with ordinary stores: order
            // the acquire above before the copy engine's (async proxy) reads below
            asm volatile("fence.proxy.async.global;" ::: "memory");
            if (ti != kNoTask) t = v.tasks[ti];
          }
          if (ti == kNoTask) return;
          ticket = next_ticket(cur);
          if (ticket < P) t_next = v.tasks[ticket];
          uint32_t kind = B2_K__COUNT;
          FnEntry fn;
          fn.kind = 0; fn.alpha = 0.f; fn.beta = 0.f; fn.pad = 0; fn.ialpha = 0; fn.ibeta = 0;
          if (t.fn_id < B2_MAX_FN) {
            fn = a.fn_table[t.fn_id];
            kind = fn.kind;
            if (kind >= B2_K__COUNT || (kind == B2_K_NONE && t.fn_id != 0)) kind = B2_K__COUNT;
          }
          if (t.flags & B2_F_CANCELLED) {
            // tombstone (b2_cancel): nothing runs, the status says so, successors are still released
            const uint32_t s = acquire();
            StageHdr &h = hdr[s];
            h.kind = kHdrFailed;
            h.status = B2_TS_CANCELLED;
            h.owner = kWarpMode ? (deal++ % kWorkerWarps) : kOwnerAll;
            h.task = ti;
            h.last = 1;
            h.succ_begin = t.succ_begin;
            h.succ_count = t.succ_count;
            mbar_arrive(smem_addr(&full_bar[s]));
            ++it;
            continue;
          }
          if (kind == B2_K__COUNT || !is_streaming_kind(kind)) {
            // one stage carries the whole electron: bad fn_id, or a callback that reads global
            // memory itself (SUM, STACK_SUM, integer electrons, the gate)
            const uint32_t s = acquire();
            StageHdr &h = hdr[s];
            h.task = ti;
            h.last = 1;
            h.succ_begin = t.succ_begin;
            h.succ_count = t.succ_count;
            h.owner = kOwnerAll;
            if (kind == B2_K__COUNT) {
              h.kind = kHdrFailed;
              h.status = B2_TS_BAD_FN;
            } else {
              h.kind = kHdrLegacy;
              legacy[s].t = t;
              legacy[s].fn = fn;
              legacy[s].index = ti;
              legacy[s].kind = kind;
            }
            mbar_arrive(smem_addr(&full_bar[s]));
            ++it;
            continue;
          }
          // ---- streaming electron: resolve once, then tile after tile -----------------
          const bool two = (kind == B2_K_ADD || kind == B2_K_MUL || kind == B2_K_AXPY);
          const uint64_t bytes = (kind == B2_K_COPY) ? t.count : t.count * 4;
          bool ok = true;
          const uint8_t *p0 = resolve(a, (t.flags >> B2_F_ARENA0_SHIFT) & 15u, t.arg0, bytes, &ok);
          const uint8_t *p1 = two ? resolve(a, (t.flags >> B2_F_ARENA1_SHIFT) & 15u, t.arg1, bytes, &ok)
                                  : nullptr;
          uint8_t *po = resolve(a, (t.flags >> B2_F_ARENAO_SHIFT) & 15u, t.out, bytes, &ok);
          if ((t.flags & (B2_F_IMM0 | (two ? B2_F_IMM1 : 0u))) || !ok || !aligned16(p0) ||
              !aligned16(po) || (two && !aligned16(p1))) {
            const uint32_t s = acquire();
            StageHdr &h = hdr[s];
            h.kind = kHdrFailed;
            h.status = B2_TS_BAD_ARG;
            h.owner = kOwnerAll;
            h.task = ti;
            h.last = 1;
            h.succ_begin = t.succ_begin;
            h.succ_count = t.succ_count;
            mbar_arrive(smem_addr(&full_bar[s]));
            ++it;
            continue;
          }
          const uint32_t per_tile = two ? kTileBytes / 2 : kTileBytes;
          const uint32_t owner = kWarpMode ? (deal++ % kWorkerWarps) : kOwnerAll;   // every tile of it
          uint64_t off = 0;
          do {
            const uint64_t left = bytes - off;
            const uint32_t chunk = left < per_tile ? (uint32_t)left : per_tile;
            const bool last = (off + chunk == bytes);
            const uint32_t b16 = chunk & ~15u;
            const uint32_t s = acquire();
            StageHdr &h = hdr[s];
            h.out = po + off;
            h.tail0 = p0 + off + b16;
            h.tail1 = two ? p1 + off + b16 : nullptr;
            h.n_vec = b16 >> 4;
            h.tail_bytes = chunk - b16;
            h.kind = kind;
            h.owner = owner;
            h.last = last ? 1u : 0u;
            h.task = ti;
            h.alpha = fn.alpha;
            h.beta = fn.beta;
            h.succ_begin = t.succ_begin;
            h.succ_count = t.succ_count;
            const uint32_t fb = smem_addr(&full_bar[s]);
            if (b16) {
              const uint32_t dst = smem_addr(stage_mem) + s * kTileBytes;
              mbar_expect_tx(fb, two ? 2 * b16 : b16);
              bulk_g2s(dst, p0 + off, b16, fb);
              if (two) bulk_g2s(dst + kTileBytes / 2, p1 + off, b16, fb);
            } else {
              mbar_arrive(fb);
            }
            ++it;
            off += chunk;
          } while (off < bytes);
        }
      };

      if (!kResident) {
        BatchView v;
        v.tasks = a.tasks;
        v.succ = a.succ;
        v.status = a.status;
        v.q = a.q;
        v.cursor = a.cursor;
        v.finished = nullptr;
        v.n_tasks = a.n_tasks;
        v.prefix = a.n_ready_prefix;
        v.claim_k = a.claim_k;
        v.seq = 0;
        marker(kHdrBegin, &v);
        produce(v);
        marker(kHdrExit, nullptr);
      } else {
        // ------------------------------------------------------- resident: batch after batch --
        ResidentCtl *ctl = ra.ctl;
        const unsigned long long born = global_ns();
        const unsigned long long idle_ns = (unsigned long long)ra.idle_us * 1000ull;
        const unsigned long long life_ns = (unsigned long long)ra.life_ms * 1000000ull;
        // Every CTA of an instance starts at the same batch: none can complete before all CTAs
        // of the instance have passed through it.
        uint32_t b = ld_acquire_gpu_u32(&ctl->done);
        for (;;) {
          if (blockIdx.x == 0) {
            // The leader decides, for every CTA alike, whether batch b runs or the instance leaves.
            unsigned long long t0 = global_ns();
            uint32_t spins = 0;
            ra.host->state = (ra.instance << 2) | 1u;
            for (;;) {
              const uint32_t published = ld_acquire_sys_u32(&ctl->doorbell);
              if ((int32_t)(published - b) > 0) {
                // at most two batches are in flight (two sets of queue cells): b starts once b-2 is done
                const uint32_t dn = ld_acquire_gpu_u32(&ctl->done);
                if ((int32_t)(b - dn) <= 1) {
                  // The descriptor sits in mapped host memory (the host wrote it with ordinary stores
                  // before it rang the doorbell): one 64-byte read over PCIe by the leader, mirrored
                  // into device memory for everybody else.
                  const volatile uint64_t *src = reinterpret_cast<const volatile uint64_t *>(&ra.host->ring[b % kResRing]);
                  uint64_t *dst = reinterpret_cast<uint64_t *>(&ctl->ring[b % kResRing]);
#pragma unroll
                  for (int w = 0; w < (int)(sizeof(ResidentBatch) / 8); ++w) dst[w] = src[w];
                  st_release_gpu_u32(&ctl->approved, b + 1);
                  break;
                }
                __nanosleep(128);
              } else {
                const unsigned long long now = global_ns();
                const bool expired = now - born > life_ns;
                const bool asked = ((++spins & 15u) == 0u) && ra.host->stop != 0u;   // b2_resident_disable / quiesce
                if (now - t0 > idle_ns || expired || asked) {
                  // exit handshake with the host (Dekker): announce, fence, then look at what the host
                  // has submitted.  The host writes `submitted` before it issues a batch's stream
                  // operations and reads `state` afterwards — one of the two sees the other.
                  ra.host->state = (ra.instance << 2) | 2u;
                  __threadfence_system();
                  const uint32_t sub = ra.host->submitted;
                  if ((int32_t)(sub - b) > 0 && !expired) {
                    ra.host->state = (ra.instance << 2) | 1u;     // a batch is on its way: stay
                    t0 = now;
                  } else {
                    if ((int32_t)(sub - b) > 0) ra.host->error = 1u;   // lifetime cap with work pending
                    __threadfence_system();
                    st_release_gpu_u32(&ctl->exit_at, b);
                    break;
                  }
                }
                __nanosleep(256);
              }
            }
          }
          bool leave = false;
          for (uint32_t backoff = 32;;) {
            if ((int32_t)(ld_acquire_gpu_u32(&ctl->approved) - b) > 0) break;
            if (ld_acquire_gpu_u32(&ctl->exit_at) == b) {
              leave = true;
              break;
            }
            __nanosleep(backoff);
            if (backoff < 512) backoff <<= 1;
          }
          if (leave) break;
          // the ring entry, the table and the payload were written by the copy engine before the
          // doorbell: the acquire chain (doorbell -> approved) orders them before these reads
          asm volatile("fence.proxy.async.global;" ::: "memory");
          const volatile ResidentBatch *rb = &ctl->ring[b % kResRing];
          ResidentQ *rq = &ctl->q[b % kResRing];
          BatchView v;
          v.tasks = rb->tasks;
          v.succ = rb->succ;
          v.status = rb->status;
          v.q.cells = a.q.cells + (size_t)(b & 1u) * ((size_t)a.q.mask + 1);
          v.q.head = &rq->head;
          v.q.tail = &rq->tail;
          v.q.epoch = b + 1u;
          v.q.mask = a.q.mask;
          v.cursor = &rq->cursor;
          v.finished = &rq->finished;
          v.n_tasks = rb->n_tasks;
          v.prefix = rb->prefix;
          v.claim_k = rb->claim_k;
          v.seq = b;
          marker(kHdrBegin, &v);
          produce(v);
          marker(kHdrEnd, nullptr);
          ++b;
        }
        marker(kHdrExit, nullptr);
      }
    }
  } else {
    // --------------------------------------------------------------- worker warps --
    const int lane = threadIdx.x & 31;
    for (uint32_t it = 0;; ++it) {
      const uint32_t s = it % kStages;
      mbar_wait(smem_addr(&full_bar[s]), (it / kStages) & 1u);
      const StageHdr &h = hdr[s];
      const uint32_t kind = h.kind;
      if (kind == kHdrExit) break;
      if (kind == kHdrBegin) {
        worker_sync();                                  // nobody still reads the previous batch's view
        if (threadIdx.x == 0) wv = bview[s];
        worker_sync();
        const uint32_t P = wv.prefix;
        if (P < wv.n_tasks) {   // publish ready electrons that sit behind a dependent one in the table
          for (;;) {
            if (threadIdx.x == 0) s_base = P + atomicAdd(wv.cursor, (uint32_t)kWorkerThreads);
            worker_sync();
            const uint32_t base = s_base;
            worker_sync();
            if (base >= wv.n_tasks) break;
            const uint32_t i = base + threadIdx.x;
            const bool ready = (i < wv.n_tasks) && (wv.tasks[i].deps == 0);
            queue_push_warp(wv.q, P, ready, i);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_addr(&empty_bar[s]));
        continue;
      }
      if (kResident && kind == kHdrEnd) {
        // resident: this CTA has retired its share of the batch.  The last CTA to say so completes
        // the batch: re-arms the slot's counters and publishes `done` for the egress stream.
        worker_sync();                                  // every store of this CTA for the batch is issued
        if (threadIdx.x == 0) {
          __threadfence_system();                       // ... and visible to the copy engine
          if (atomicAdd(wv.finished, 1u) == gridDim.x - 1) {
            *wv.q.head = 0;
            *wv.q.tail = 0;
            *wv.cursor = 0;
            *wv.finished = 0;
            __threadfence_system();
            st_release_sys_u32(&ra.ctl->done, wv.seq + 1u);
            ra.host->done = wv.seq + 1u;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_addr(&empty_bar[s]));
        continue;
      }
      const uint32_t owner = kWarpMode ? h.owner : kOwnerAll;
      if (kWarpMode && owner != kOwnerAll && owner != (uint32_t)(threadIdx.x >> 5)) {
        __syncwarp();                                   // somebody else's electron: hand the stage back
        if (lane == 0) mbar_arrive(smem_addr(&empty_bar[s]));
        continue;
      }
      const bool solo = kWarpMode && owner != kOwnerAll;   // this warp alone runs the electron
      const uint32_t tid = solo ? (uint32_t)lane : threadIdx.x;
      const uint32_t nthr = solo ? (uint32_t)kWarp : (uint32_t)kWorkerThreads;
      uint32_t status = B2_TS_OK;
      if (kind == kHdrFailed) {
        status = h.status;
      } else if (kind == kHdrLegacy) {
        const Slot &sl = legacy[s];
        const uint32_t lk = sl.kind;
        Ctx c;
        bool ok = true;
        const uint64_t arg0 = sl.t.arg0, arg1 = sl.t.arg1, count = sl.t.count;
        const uint32_t flags = sl.t.flags, nrows = sl.t.nrows;
        const bool is_int = (lk == B2_K_IADD || lk == B2_K_IMUL || lk == B2_K_ISCALE_ADD);
        uint64_t in_bytes = count * 4;
        if (lk == B2_K_STACK_SUM && nrows) in_bytes += (uint64_t)(nrows - 1) * arg1;
        const uint64_t out_bytes = (lk == B2_K_SUM) ? 4 : (is_int ? 8 : count * 4);
        c.imm0 = arg0;
        c.imm1 = arg1;
        c.a0 = ((flags & B2_F_IMM0) || is_int || lk == B2_K_NONE)
                   ? nullptr
                   : resolve(a, (flags >> B2_F_ARENA0_SHIFT) & 15u, arg0, in_bytes, &ok);
        c.a1 = nullptr;
        c.out = (lk == B2_K_NONE)
                    ? nullptr
                    : resolve(a, (flags >> B2_F_ARENAO_SHIFT) & 15u, sl.t.out, out_bytes, &ok);
        c.count = count;
        c.nrows = nrows;
        c.fn = &sl.fn;
        status = ok ? g_callbacks[lk](c) : (uint32_t)B2_TS_BAD_ARG;
      } else {
        // one tile of an elementwise electron, straight out of shared memory
        const int4 *x = reinterpret_cast<const int4 *>(stage_mem + s * kTileBytes);
        const int4 *y = x + (kTileBytes / 2) / 16;
        int4 *o = reinterpret_cast<int4 *>(h.out);
        const uint32_t n = h.n_vec;
        const float al = h.alpha, be = h.beta;
        for (uint32_t i = tid; i < n; i += nthr) {
          int4 v = x[i];
          if (kind != B2_K_COPY) {
            float4 f = *reinterpret_cast<float4 *>(&v);
            if (kind == B2_K_SCALE_ADD) {
              f.x = __fadd_rn(__fmul_rn(al, f.x), be); f.y = __fadd_rn(__fmul_rn(al, f.y), be);
              f.z = __fadd_rn(__fmul_rn(al, f.z), be); f.w = __fadd_rn(__fmul_rn(al, f.w), be);
            } else {
              int4 w = y[i];
              const float4 g = *reinterpret_cast<float4 *>(&w);
              if (kind == B2_K_ADD) {
                f.x = __fadd_rn(f.x, g.x); f.y = __fadd_rn(f.y, g.y);
                f.z = __fadd_rn(f.z, g.z); f.w = __fadd_rn(f.w, g.w);
              } else if (kind == B2_K_MUL) {
                f.x = __fmul_rn(f.x, g.x); f.y = __fmul_rn(f.y, g.y);
                f.z = __fmul_rn(f.z, g.z); f.w = __fmul_rn(f.w, g.w);
              } else {  // AXPY
                f.x = __fadd_rn(__fmul_rn(al, f.x), g.x); f.y = __fadd_rn(__fmul_rn(al, f.y), g.y);
                f.z = __fadd_rn(__fmul_rn(al, f.z), g.z); f.w = __fadd_rn(__fmul_rn(al, f.w), g.w);
              }
            }
            v = *reinterpret_cast<int4 *>(&f);
          }
          st_stream(o + i, v);
        }
        // < 16 trailing bytes of the electron: read from global, element by element
        const uint32_t tb = h.tail_bytes;
        if (tb && tid < (kind == B2_K_COPY ? tb : tb / 4)) {
          uint8_t *ot = h.out + (n << 4);
          if (kind == B2_K_COPY) {
            ot[tid] = h.tail0[tid];
          } else {
            const float xv = reinterpret_cast<const float *>(h.tail0)[tid];
            float r;
            if (kind == B2_K_SCALE_ADD) {
              r = __fadd_rn(__fmul_rn(al, xv), be);
            } else {
              const float yv = reinterpret_cast<const float *>(h.tail1)[tid];
              r = (kind == B2_K_ADD) ? __fadd_rn(xv, yv)
                  : (kind == B2_K_MUL) ? __fmul_rn(xv, yv)
                                       : __fadd_rn(__fmul_rn(al, xv), yv);
            }
            reinterpret_cast<float *>(ot)[tid] = r;
          }
        }
      }

      if (h.last) {
        const uint32_t ti = h.task, succ_begin = h.succ_begin, succ_count = h.succ_count;
        if (solo) {
          __syncwarp();   // the warp's result stores of this electron are ordered before what follows
          if (lane == 0) wv.status[ti] = status;
        } else {
          if (threadIdx.x == 0) s_status = status;
          worker_sync();  // every result store of this electron is ordered before what follows
          if (threadIdx.x == 0) wv.status[ti] = s_status;
        }
        for (uint32_t base = 0; base < succ_count; base += nthr) {
          const uint32_t e = base + tid;
          bool fire = false;
          uint32_t sidx = 0;
          if (e < succ_count) {
            sidx = wv.succ[succ_begin + e];
            if (sidx < wv.n_tasks) {
              __threadfence();
              fire = (atomicSub(&wv.tasks[sidx].deps, 1u) == kDepBias + 1u);
            }
          }
          queue_push_warp(wv.q, wv.prefix, fire, sidx);
        }
      }
      // hand the stage back: one arrival per worker warp, after the whole warp has read it
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_addr(&empty_bar[s]));
    }
  }

  __syncthreads();
  if (!kResident && threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(a.exited, 1u) == gridDim.x - 1) {
      *a.q.head = 0;
      *a.q.tail = 0;
      *a.cursor = 0;
      *a.exited = 0;
      __threadfence();
    }
  }
}

__global__ void __launch_bounds__(kCtaThreads, B2_WORKER_MIN_CTAS)
    worker_tma_kernel(const __grid_constant__ WorkerArgs a) {
  worker_tma_body<kStages, kTileBytes, false, false>(a, ResidentArgs{});
}

__global__ void __launch_bounds__(kCtaThreads, B2_WORKER_MIN_CTAS)
    worker_tma_small_kernel(const __grid_constant__ WorkerArgs a) {
  worker_tma_body<kSmallStages, kSmallTileBytes, false, true>(a, ResidentArgs{});
}

// The resident worker: one CTA set that outlives its batches (b2_resident_enable).  `a` supplies
// what does not change between batches (callback table, arenas, the two sets of queue cells);
// everything per batch comes through the ring in `ra.ctl`.
__global__ void __launch_bounds__(kCtaThreads, B2_WORKER_MIN_CTAS)
    worker_resident_kernel(const __grid_constant__ WorkerArgs a, const __grid_constant__ ResidentArgs ra) {
  worker_tma_body<kStages, kTileBytes, true, false>(a, ra);
}

}  // namespace b2
