// b2exec.cu — host runtime + C ABI (include/b2exec.h) of the B200 electron
// executor.  One translation unit: the kernels live in b2_spans.cuh (K1/K2/K4)
// and b2_worker.cuh (K3).
//
// Build (see covalent_ssh_plugin_b200/build.py):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
//        -Xcompiler -fPIC -shared -o libb2exec.so b2exec.cu
//
// Per device the runtime owns: one compute stream (every kernel of the library
// runs on it, in call order), the device arena, two pinned staging buffers,
// the task table / successor list / status array, the MPMC queue cells and the
// callback table.  Host-side tables that a kernel reads are staged through
// small event-guarded pinned rings so that calls never block on the GPU unless
// the ring wraps.
#include <cuda.h>   // types and flags of the stream memory operations (entry points are looked up at run time)
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only; ranges are no-ops unless a profiler is attached

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "b2_hostcopy.h"
#include "b2_spans.cuh"
#include "b2_spans_tma.cuh"
#include "b2_worker.cuh"
#include "b2_worker_tma.cuh"

namespace {

using namespace b2;

thread_local std::string g_err = "";

int fail(int code, const char *fmt, ...) {
  char buf[768];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(B2_E_CUDA, "%s -> %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__,    \
                  __LINE__);                                                                  \
  } while (0)

#ifndef B2_SRC_HASH
#define B2_SRC_HASH "unknown"
#endif

constexpr int kMaxDevices = 16;
constexpr int kRing = 8;          // staging ring depth
constexpr int kTickets = 4096;    // completion events kept alive
constexpr int kMaxUserEvents = 256;
constexpr size_t kCopyPieceBytes = 256 * 1024;   // large operands are split across the copy threads

// One queued b2_submit (the caller keeps jobs / tasks alive until the ticket completes).
struct Submission {
  b2_batch b;
  uint64_t ticket = 0;
  std::vector<std::pair<uint32_t, uint32_t>> cancels;   // tombstones that arrived while it was queued
};

// Where the batch behind a ticket is, for b2_cancel.
enum TicketPhase : uint8_t { kPhaseNone = 0, kPhaseQueued = 1, kPhaseTabled = 2, kPhaseLaunched = 3 };

// A pinned host buffer with a device mirror, guarded by an event so it is not
// rewritten while an earlier async copy may still be reading it.
struct Staging {
  uint8_t *h = nullptr;
  uint8_t *d = nullptr;
  size_t cap = 0;
  cudaEvent_t ev = nullptr;   // last consumer of this slot has finished
  cudaEvent_t up = nullptr;   // upload of this slot has finished (task tables)
  bool in_flight = false;
  // slot last used by a batch of the resident worker that queued nothing else: reusable once the worker's
  // `done` word (mirrored in mapped host memory) has passed this sequence number — no event involved
  volatile const uint32_t *res_done = nullptr;
  uint32_t res_seq = 0;
};

struct Device {
  std::mutex mu;
  bool up = false;
  int dev = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;         // compute: every kernel and payload copy, in call order
  cudaStream_t upload = nullptr;         // task tables only, so table k+1 uploads while kernel k runs
  cudaStream_t egress = nullptr;         // device->host copies: D2H of batch k overlaps H2D of batch k+1
  cudaEvent_t to_egress = nullptr;       // compute -> egress ordering point
  bool egress_dirty = false;             // copies queued on the egress stream since the last ticket

  uint8_t *h_in = nullptr, *h_out = nullptr;
  uint64_t h_in_bytes = 0, h_out_bytes = 0;

  // arena 0 is local; 1.. are mapped peers
  uint8_t *arena[B2_MAX_ARENAS] = {};
  uint64_t arena_bytes[B2_MAX_ARENAS] = {};
  int arena_owner_dev[B2_MAX_ARENAS] = {};
  bool arena_ipc[B2_MAX_ARENAS] = {};
  int n_arenas = 0;

  // queue + callback table
  unsigned long long *d_cells = nullptr;
  uint32_t *d_head = nullptr, *d_tail = nullptr, *d_cursor = nullptr, *d_exited = nullptr;
  uint32_t qcap = 0;
  uint32_t epoch = 0;
  FnEntry *d_fn = nullptr;

  // pending batch: [qcap tasks][succ_cap successor indices]
  Staging task_ring[kRing];
  int task_slot = 0;
  uint32_t pend_tasks = 0, pend_succ = 0, succ_cap = 0;
  uint32_t pend_prefix = 0;      // leading electrons with deps == 0
  bool pend_prefix_open = true;
  bool pend_external = false;    // the pending batch was opened by b2_enqueue (not by the submission thread)
  uint64_t pend_bytes = 0;       // operand bytes of the pending batch (picks the worker flavour at launch)
  uint8_t h_fn_kind[B2_MAX_FN] = {};   // host mirror of the callback kinds (b2_register)

  Staging span_ring[kRing];
  int span_slot = 0;

  int worker_ctas_per_sm = 0;
  int worker_small_ctas_per_sm = 0;
  int worker_small_mode = -1;    // B2_WORKER_SMALL: -1 auto (by average operand size), 0 never, 1 always
  int claim_k_override = 0;      // B2_CLAIM_K
  // resident worker (b2_resident_enable): one kernel instance outlives its batches
  bool res_enabled = false;
  bool res_live = false;           // an instance has been launched and may still be running
  cudaStream_t res_stream = nullptr;
  ResidentCtl *d_ctl = nullptr;
  ResidentHost *h_res = nullptr, *d_res = nullptr;   // mapped pinned block: host / device address
  uint32_t *h_const = nullptr;                       // pinned constants for stream-ordered word writes
  uint32_t res_submitted = 0;      // sequence number of the next batch
  uint32_t res_instance = 0;
  uint32_t res_idle_us = 200, res_life_ms = 2000;
  int res_ctas_per_sm = 0;         // every CTA of an instance must be co-resident (they wait for each other)
  uint64_t res_batches = 0, res_launches = 0;
  bool compute_dirty = false;      // real work sits on the compute stream that a later batch may depend on
  uint32_t res_unwaited = 0;       // != 0: the compute stream has not been ordered after resident batch (this - 1) yet
  uint32_t res_last_pure = 0;      // set by resident_publish for launch_locked: seq + 1 of the batch just published
  struct Table {                 // resident task table (b2_table_create)
    uint8_t *d_pristine = nullptr;   // [n tasks][n_succ successor indices] as uploaded
    uint8_t *d_work = nullptr;       // working copy of the tasks (dependency counters are consumed)
    uint32_t *d_status = nullptr;
    uint32_t n = 0, n_succ = 0, prefix = 0;
    uint64_t bytes = 0;
    bool has_deps = false;
    bool live = false;
  };
  std::vector<Table> tables;
  int mover_ctas_per_sm = 0;
  int mover_unroll = 8;
  bool mover_tma = true;         // bulk-copy-engine kernel for 16 B aligned spans (B2_MOVER=ldg|wide to disable)
  bool mover_wide = false;       // 256-bit loads/stores where 32 B aligned (B2_MOVER=wide)
  bool worker_tma = true;        // streaming worker, operand loads on the copy engine (B2_WORKER=classic to disable)

  cudaEvent_t tickets[kTickets] = {};
  uint64_t next_ticket = 1;
  // per ticket slot (index = ticket % kTickets), guarded by mu
  bool ticket_recorded[kTickets] = {};     // its event has been recorded (async tickets are recorded late)
  int ticket_err[kTickets] = {};           // failure of the asynchronous pipeline, reported by poll / wait
  std::string ticket_msg[kTickets];
  uint8_t ticket_phase[kTickets] = {};
  int8_t ticket_slot[kTickets] = {};       // task-ring slot holding the batch's table (-1: none)
  uint32_t ticket_ntasks[kTickets] = {};
  uint32_t ticket_res_seq[kTickets] = {};  // != 0: no event — complete when the resident worker's done word reaches this
  std::condition_variable ticket_cv;       // signalled when an async ticket gets recorded
  cudaEvent_t user_ev[kMaxUserEvents] = {};
  uint32_t n_user_ev = 0;

  // asynchronous submissions (b2_submit): one thread per device issues them in order
  std::thread submitter;
  std::mutex qmu;
  std::condition_variable qcv, qidle;
  std::deque<Submission> queue;
  bool q_busy = false, q_stop = false;
  Submission *q_current = nullptr;   // the one the submission thread is working on

  b2_stats stats = {};
};

Device g_dev[kMaxDevices];

int resident_quiesce(Device *D);
int compute_after_resident(Device *D);

Device *get(int dev) {
  if (dev < 0 || dev >= kMaxDevices) return nullptr;
  return &g_dev[dev];
}

#define DEV_OR_FAIL(D, dev)                                              \
  Device *D = get(dev);                                                  \
  if (!D) return fail(B2_E_ARG, "device ordinal %d out of range", dev);  \
  std::lock_guard<std::mutex> lock_(D->mu);                              \
  if (!D->up) return fail(B2_E_STATE, "device %d is not initialised (call b2_init)", dev); \
  CU(cudaSetDevice(dev))

// Calls whose effect is ordered against queued b2_submit batches first wait until the
// submission thread has issued everything queued so far (program order is kept).
void drain_submissions(Device *D) {
  std::unique_lock<std::mutex> lk(D->qmu);
  D->qidle.wait(lk, [&] { return D->queue.empty() && !D->q_busy; });
}

#define DEV_ORDERED(D, dev)                                              \
  Device *D = get(dev);                                                  \
  if (!D) return fail(B2_E_ARG, "device ordinal %d out of range", dev);  \
  if (std::this_thread::get_id() != D->submitter.get_id()) drain_submissions(D); \
  std::lock_guard<std::mutex> lock_(D->mu);                              \
  if (!D->up) return fail(B2_E_STATE, "device %d is not initialised (call b2_init)", dev); \
  CU(cudaSetDevice(dev))

// ... and, for calls that queue work on the compute stream, after every batch published to the
// resident worker (see compute_after_resident).
#define DEV_ORDERED_SYNCED(D, dev) \
  DEV_ORDERED(D, dev);              \
  {                                 \
    int rc_res_ = compute_after_resident(D); \
    if (rc_res_) return rc_res_;    \
  }

int staging_reserve(Staging &s, size_t bytes) {
  if (s.res_done) {
    for (int spins = 0; (int32_t)(*s.res_done - s.res_seq) < 0; ++spins) {
      if (spins > 50000000) return fail(B2_E_CUDA, "resident worker never completed the batch that holds this slot");
      if (spins > 256) std::this_thread::yield();
    }
    s.res_done = nullptr;
  }
  if (s.in_flight) {
    CU(cudaEventSynchronize(s.ev));
    s.in_flight = false;
  }
  if (!s.ev) CU(cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
  if (bytes > s.cap) {
    size_t ncap = std::max<size_t>(bytes, s.cap * 2);
    ncap = (ncap + 4095) & ~size_t(4095);
    if (s.h) cudaFreeHost(s.h);
    if (s.d) cudaFree(s.d);
    s.h = nullptr;
    s.d = nullptr;
    s.cap = 0;
    CU(cudaHostAlloc(reinterpret_cast<void **>(&s.h), ncap, cudaHostAllocDefault));
    CU(cudaMalloc(reinterpret_cast<void **>(&s.d), ncap));
    s.cap = ncap;
  }
  return B2_OK;
}

void staging_free(Staging &s) {
  if (s.h) cudaFreeHost(s.h);
  if (s.d) cudaFree(s.d);
  if (s.ev) cudaEventDestroy(s.ev);
  if (s.up) cudaEventDestroy(s.up);
  s = Staging();
}

template <int UNROLL>
void launch_mover(Device *D, const SpanDev *spans, const uint32_t *base, uint32_t n, uint32_t chunks,
                  uint32_t cps, int grid) {
  span_move_kernel<UNROLL><<<grid, kMoveThreads, 0, D->stream>>>(spans, base, n, chunks, cps);
}

// Builds the device span table for n spans (src in arena src_id, dst in arena
// dst_id), uploads it and launches the mover.  Caller holds the device lock.
int move_spans_locked(Device *D, const b2_span *spans, uint32_t n, uint32_t src_id,
                      uint32_t dst_id) {
  if (n == 0) return B2_OK;
  if (!spans) return fail(B2_E_ARG, "spans is NULL");
  {   // another kernel needs the SMs: ask a resident instance to leave now rather than at its idle timeout
    int rc_ = resident_quiesce(D);
    if (rc_) return rc_;
  }
  if (src_id >= (uint32_t)D->n_arenas || dst_id >= (uint32_t)D->n_arenas)
    return fail(B2_E_ARG, "arena id out of range (src %u, dst %u, have %d)", src_id, dst_id,
                D->n_arenas);
  const size_t table_bytes = (size_t)n * sizeof(SpanDev);
  const size_t base_off = (table_bytes + 255) & ~size_t(255);
  const size_t total = base_off + ((size_t)n + 1) * sizeof(uint32_t);
  Staging &s = D->span_ring[D->span_slot];
  if (total > s.cap) {
    // grow every slot of the ring in the same call: later calls of this size find
    // all of them ready and never allocate (allocation synchronises the device)
    for (auto &other : D->span_ring) {
      int rc_ = staging_reserve(other, total + total / 4);
      if (rc_) return rc_;
    }
  }
  int rc = staging_reserve(s, total);
  if (rc) return rc;
  D->span_slot = (D->span_slot + 1) % kRing;

  SpanDev *tab = reinterpret_cast<SpanDev *>(s.h);
  uint32_t *base = reinterpret_cast<uint32_t *>(s.h + base_off);
  uint64_t chunks = 0, moved = 0;
  bool uniform = true;
  uint64_t misalign = 0;
  const uint64_t b0 = spans[0].bytes;
  for (uint32_t i = 0; i < n; ++i) {
    const b2_span &sp = spans[i];
    if (sp.src_off + sp.bytes > D->arena_bytes[src_id] || sp.src_off + sp.bytes < sp.src_off)
      return fail(B2_E_ARG, "span %u: source [%llu, +%llu) outside arena %u (%llu bytes)", i,
                  (unsigned long long)sp.src_off, (unsigned long long)sp.bytes, src_id,
                  (unsigned long long)D->arena_bytes[src_id]);
    if (sp.dst_off + sp.bytes > D->arena_bytes[dst_id] || sp.dst_off + sp.bytes < sp.dst_off)
      return fail(B2_E_ARG, "span %u: destination [%llu, +%llu) outside arena %u (%llu bytes)", i,
                  (unsigned long long)sp.dst_off, (unsigned long long)sp.bytes, dst_id,
                  (unsigned long long)D->arena_bytes[dst_id]);
    tab[i].src = D->arena[src_id] + sp.src_off;
    tab[i].dst = D->arena[dst_id] + sp.dst_off;
    tab[i].bytes = sp.bytes;
    tab[i].chunk_base = (uint32_t)chunks;
    tab[i].pad = 0;
    base[i] = (uint32_t)chunks;
    chunks += (sp.bytes + kChunkBytes - 1) / kChunkBytes;
    moved += sp.bytes;
    misalign |= sp.src_off | sp.dst_off | sp.bytes;
    uniform = uniform && (sp.bytes == b0);
    if (chunks > 0xFFFFFFF0ull) return fail(B2_E_NOMEM, "too many chunks in one span move");
  }
  base[n] = (uint32_t)chunks;
  if (chunks == 0) return B2_OK;

  // like the task tables, the span table uploads on its own stream so that it
  // overlaps whatever the compute stream is still running
  if (!s.up) CU(cudaEventCreateWithFlags(&s.up, cudaEventDisableTiming));
  CU(cudaMemcpyAsync(s.d, s.h, total, cudaMemcpyHostToDevice, D->upload));
  CU(cudaEventRecord(s.up, D->upload));
  CU(cudaStreamWaitEvent(D->stream, s.up, 0));

  const uint32_t cps = uniform ? (uint32_t)((b0 + kChunkBytes - 1) / kChunkBytes) : 0u;
  const int warps_per_cta = kMoveThreads / kWarp;
  const uint64_t want = (chunks + warps_per_cta - 1) / warps_per_cta;
  const int cap = D->sm_count * std::max(1, D->mover_ctas_per_sm);
  const int grid = (int)std::min<uint64_t>(want, (uint64_t)cap);
  const SpanDev *dtab = reinterpret_cast<const SpanDev *>(s.d);
  const uint32_t *dbase = reinterpret_cast<const uint32_t *>(s.d + base_off);
  if (D->mover_tma && (misalign & 15u) == 0) {
    const uint64_t want_tma = (chunks + kTmaPipes - 1) / kTmaPipes;
    const int grid_tma = (int)std::min<uint64_t>(want_tma, (uint64_t)D->sm_count);
    span_move_tma_kernel<<<grid_tma, kTmaThreads, kTmaSmemBytes, D->stream>>>(dtab, dbase, n,
                                                                              (uint32_t)chunks, cps);
  } else if (D->mover_wide) {
    span_move_kernel<8, true><<<grid, kMoveThreads, 0, D->stream>>>(dtab, dbase, n, (uint32_t)chunks, cps);
  } else
  switch (D->mover_unroll) {
    case 4: launch_mover<4>(D, dtab, dbase, n, (uint32_t)chunks, cps, grid); break;
    case 16: launch_mover<16>(D, dtab, dbase, n, (uint32_t)chunks, cps, grid); break;
    default: launch_mover<8>(D, dtab, dbase, n, (uint32_t)chunks, cps, grid); break;
  }
  CU(cudaGetLastError());
  CU(cudaEventRecord(s.ev, D->stream));
  s.in_flight = true;
  D->compute_dirty = true;
  D->stats.launches_pack += 1;
  D->stats.bytes_packed += moved;
  return B2_OK;
}

// Everything issued on the compute stream so far happens before whatever is queued on the
// egress stream next.
int egress_after_compute(Device *D) {
  {
    int rc_ = compute_after_resident(D);
    if (rc_) return rc_;
  }
  CU(cudaEventRecord(D->to_egress, D->stream));
  CU(cudaStreamWaitEvent(D->egress, D->to_egress, 0));
  return B2_OK;
}

// NVTX range around the host side of an ABI call (visible to ncu --nvtx / nsys).
struct Range {
  explicit Range(const char *name) { nvtxRangePushA(name); }
  ~Range() { nvtxRangePop(); }
};

int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  if (!v || !*v) return dflt;
  return atoi(v);
}

// ------------------------------------------------------------ resident worker --
// Stream memory operations (driver API) looked up through the runtime, so that the library does
// not link against libcuda: the doorbell is a 32-bit word the upload stream WRITES after the batch's
// table / payload / ring entry, completion is a word the compute stream WAITS on.
typedef CUresult (*stream_value32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
stream_value32_fn g_write32 = nullptr, g_wait32 = nullptr;

int load_stream_memops() {
  if (g_write32 && g_wait32) return B2_OK;
  cudaDriverEntryPointQueryResult q1, q2;
  void *w = nullptr, *v = nullptr;
  CU(cudaGetDriverEntryPoint("cuStreamWriteValue32", &w, cudaEnableDefault, &q1));
  CU(cudaGetDriverEntryPoint("cuStreamWaitValue32", &v, cudaEnableDefault, &q2));
  if (!w || !v || q1 != cudaDriverEntryPointSuccess || q2 != cudaDriverEntryPointSuccess)
    return fail(B2_E_UNSUPPORTED, "stream memory operations are not available in this driver");
  g_write32 = reinterpret_cast<stream_value32_fn>(w);
  g_wait32 = reinterpret_cast<stream_value32_fn>(v);
  return B2_OK;
}

#define CUD(call)                                                                      \
  do {                                                                                 \
    CUresult r_ = (call);                                                              \
    if (r_ != CUDA_SUCCESS)                                                            \
      return fail(B2_E_CUDA, "%s -> CUresult %d (%s:%d)", #call, (int)r_, __FILE__, __LINE__); \
  } while (0)

void fill_worker_consts(Device *D, WorkerArgs &a) {
  memset(&a, 0, sizeof a);
  a.fn_table = D->d_fn;
  a.q.cells = D->d_cells;
  a.q.head = D->d_head;
  a.q.tail = D->d_tail;
  a.q.mask = D->qcap - 1;
  a.cursor = D->d_cursor;
  a.exited = D->d_exited;
  for (int i = 0; i < B2_MAX_ARENAS; ++i) {
    a.arena[i] = D->arena[i];
    a.arena_bytes[i] = D->arena_bytes[i];
  }
}

// Launches one instance of the resident kernel (it leaves by itself when idle; see the exit
// handshake in b2_worker_tma.cuh).  Stream-ordered after the previous instance.
int resident_launch_instance(Device *D) {
  D->res_instance += 1;
  D->h_res->error = 0;             // whatever made the previous instance give up is that instance's story
  D->h_const[0] = kResNone;
  CU(cudaMemcpyAsync(&D->d_ctl->exit_at, &D->h_const[0], 4, cudaMemcpyHostToDevice, D->res_stream));
  WorkerArgs a;
  fill_worker_consts(D, a);
  a.q.cells = D->d_cells + D->qcap;          // the resident worker's two cell sets
  ResidentArgs ra;
  memset(&ra, 0, sizeof ra);
  ra.ctl = D->d_ctl;
  ra.host = D->d_res;
  ra.instance = D->res_instance;
  ra.idle_us = D->res_idle_us;
  ra.life_ms = D->res_life_ms;
  const int grid = D->sm_count * std::max(1, D->res_ctas_per_sm);   // all of them resident at once, by construction
  worker_resident_kernel<<<grid, kCtaThreads, kWorkerSmemBytes, D->res_stream>>>(a, ra);
  CU(cudaGetLastError());
  D->res_live = true;
  D->res_launches += 1;
  D->stats.launches_worker += 1;
  return B2_OK;
}

// Publishes one batch to the resident worker instead of launching a kernel for it.
int resident_publish(Device *D, b2_task *d_tasks, const uint32_t *d_succ, uint32_t *d_status, uint32_t n,
                     uint32_t prefix, uint64_t operand_bytes) {
  const uint32_t seq = D->res_submitted;
  const uint32_t slot = seq % kResRing;
  // the ring entry being replaced belongs to batch seq - kResRing: wait (on the host) until it is done.
  // Only a producer that runs a whole ring ahead of the GPU ever waits here.
  for (int spins = 0; seq >= kResRing && (int32_t)(D->h_res->done - (seq - kResRing + 1)) < 0; ++spins) {
    if (D->h_res->error || spins > 4000000) return fail(B2_E_CUDA, "resident worker stopped consuming batches");
    if (spins > 64) std::this_thread::yield();
  }
  const uint64_t avg = n ? operand_bytes / n : 0;
  const int grid = D->sm_count * std::max(1, D->res_ctas_per_sm);
  uint32_t k = 1;
  if (operand_bytes) k = avg <= 2048 ? 8u : avg <= 8192 ? 4u : avg <= 32768 ? 2u : 1u;
  while (k > 1 && (uint64_t)grid * k * 4 > n) k >>= 1;
  if (D->claim_k_override > 0) k = (uint32_t)D->claim_k_override;
  ResidentBatch rb;
  memset(&rb, 0, sizeof rb);
  rb.tasks = d_tasks;
  rb.succ = d_succ;
  rb.status = d_status;
  rb.n_tasks = n;
  rb.prefix = prefix;
  rb.claim_k = k;
  memcpy(const_cast<ResidentBatch *>(&D->h_res->ring[slot]), &rb, sizeof rb);   // plain stores; visible before the doorbell
  // exit handshake, host half: announce the batch BEFORE any of its stream operations, then look
  // at the kernel's state.  If the instance we launched last is leaving, queue another one behind it.
  D->h_res->submitted = seq + 1;
  __sync_synchronize();
  const uint32_t st = D->h_res->state;
  if (!D->res_live || ((st & 3u) == 2u && (st >> 2) == D->res_instance)) {
    int rc = resident_launch_instance(D);
    if (rc) return rc;
  }
  if (D->compute_dirty) {   // e.g. a b2_h2d this batch reads from: the doorbell must not overtake it
    CU(cudaEventRecord(D->to_egress, D->stream));
    CU(cudaStreamWaitEvent(D->upload, D->to_egress, 0));
    D->compute_dirty = false;
  }
  CUD(g_write32(D->upload, (CUdeviceptr)(uintptr_t)&D->d_ctl->doorbell, seq + 1, CU_STREAM_WRITE_VALUE_DEFAULT));
  // The compute stream is ordered after this batch LAZILY (compute_after_resident): only when something
  // is about to be queued behind it — a status / result download, a span move, a ticket event.  A pure
  // stream of launches costs two stream operations per batch: the table upload and the doorbell.
  D->res_unwaited = seq + 1;
  D->res_last_pure = seq + 1;
  D->res_submitted = seq + 1;
  D->res_batches += 1;
  D->stats.electrons += n;
  return B2_OK;
}

// Orders the compute stream after every batch published to the resident worker so far.
int compute_after_resident(Device *D) {
  if (!D->res_unwaited) return B2_OK;
  CUD(g_wait32(D->stream, (CUdeviceptr)(uintptr_t)&D->d_ctl->done, D->res_unwaited, CU_STREAM_WAIT_VALUE_GEQ));
  D->res_unwaited = 0;
  return B2_OK;
}

// Makes the resident instance leave now (it finishes what has been submitted first).
int resident_quiesce(Device *D) {
  if (!D->res_live) return B2_OK;
  D->h_res->stop = 1;
  __sync_synchronize();
  CU(cudaStreamSynchronize(D->res_stream));
  D->h_res->stop = 0;
  D->res_live = false;
  D->res_unwaited = 0;             // everything published has completed
  return B2_OK;
}

}  // namespace

// =============================================================== C ABI =====
extern "C" {

static void submit_loop(Device *D);
static int resident_setup(Device *D, uint32_t idle_us);

const char *b2_last_error(void) { return g_err.c_str(); }
int b2_abi_version(void) { return (int)B2_ABI_VERSION; }
const char *b2_source_hash(void) { return B2_SRC_HASH; }

int b2_device_count(int *count) {
  if (!count) return fail(B2_E_ARG, "count is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    return fail(B2_E_CUDA, "cudaGetDeviceCount -> %s", cudaGetErrorString(e));
  }
  *count = n;
  return B2_OK;
}

int b2_is_initialized(int dev) {
  Device *D = get(dev);
  if (!D) return 0;
  std::lock_guard<std::mutex> lock_(D->mu);
  return D->up ? 1 : 0;
}

int b2_init(int dev, uint64_t arena_bytes, uint64_t staging_in_bytes, uint64_t staging_out_bytes,
            uint32_t queue_cap) {
  Device *D = get(dev);
  if (!D) return fail(B2_E_ARG, "device ordinal %d out of range", dev);
  std::lock_guard<std::mutex> lock_(D->mu);
  if (D->up) return fail(B2_E_STATE, "device %d is already initialised", dev);
  if (queue_cap < 2 || (queue_cap & (queue_cap - 1)))
    return fail(B2_E_ARG, "queue_cap must be a power of two >= 2 (got %u)", queue_cap);
  if (arena_bytes < 4096 || (arena_bytes & 255))
    return fail(B2_E_ARG, "arena_bytes must be a multiple of 256 and >= 4096");
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (dev >= ndev) return fail(B2_E_ARG, "no CUDA device %d (have %d)", dev, ndev);
  CU(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, dev));
  if (prop.major < 10)
    return fail(B2_E_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only",
                dev, prop.major, prop.minor);
  D->dev = dev;
  D->sm_count = prop.multiProcessorCount;
  CU(cudaStreamCreateWithFlags(&D->stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&D->upload, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&D->egress, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&D->to_egress, cudaEventDisableTiming));
  D->egress_dirty = false;

  CU(cudaMalloc(reinterpret_cast<void **>(&D->arena[0]), arena_bytes));
  D->arena_bytes[0] = arena_bytes;
  D->arena_owner_dev[0] = dev;
  D->n_arenas = 1;
  if (staging_in_bytes) {
    CU(cudaHostAlloc(reinterpret_cast<void **>(&D->h_in), staging_in_bytes, cudaHostAllocDefault));
    D->h_in_bytes = staging_in_bytes;
  }
  if (staging_out_bytes) {
    CU(cudaHostAlloc(reinterpret_cast<void **>(&D->h_out), staging_out_bytes,
                     cudaHostAllocDefault));
    D->h_out_bytes = staging_out_bytes;
  }

  D->qcap = queue_cap;
  D->succ_cap = queue_cap * 4u;
  // three sets of cells: ordinary launches use the first, the resident worker alternates between the
  // other two (at most two of its batches are in flight); epochs never mix between the two users
  CU(cudaMalloc(reinterpret_cast<void **>(&D->d_cells), (size_t)queue_cap * 8 * 3));
  CU(cudaMalloc(reinterpret_cast<void **>(&D->d_head), 256));
  D->d_tail = D->d_head + 16;     // separate 64 B sectors
  D->d_cursor = D->d_head + 32;
  D->d_exited = D->d_head + 48;
  CU(cudaMalloc(reinterpret_cast<void **>(&D->d_fn), sizeof(FnEntry) * B2_MAX_FN));
  CU(cudaMemsetAsync(D->d_fn, 0, sizeof(FnEntry) * B2_MAX_FN, D->stream));
  queue_init_kernel<<<(3 * queue_cap + 255) / 256, 256, 0, D->stream>>>(D->d_cells, 3 * queue_cap,
                                                                       D->d_head);
  CU(cudaGetLastError());
  D->stats = b2_stats();
  D->stats.launches_other = 1;
  D->epoch = 0;
  D->pend_tasks = D->pend_succ = 0;
  D->pend_prefix = 0;
  D->pend_prefix_open = true;
  D->pend_external = false;
  D->task_slot = D->span_slot = 0;
  D->next_ticket = 1;
  D->n_user_ev = 0;

  {
    const char *wk = getenv("B2_WORKER");
    // default: the streaming worker (operand loads on the copy engine); B2_WORKER=classic
    // selects worker_kernel, whose worker warps load their own operands
    D->worker_tma = !(wk && std::string(wk) == "classic");
    CU(cudaFuncSetAttribute(worker_tma_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CU(cudaFuncSetAttribute(worker_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)kWorkerSmemBytes));
  }
  CU(cudaFuncSetAttribute(worker_tma_small_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  CU(cudaFuncSetAttribute(worker_tma_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)kWorkerSmallSmemBytes));
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D->worker_small_ctas_per_sm, worker_tma_small_kernel,
                                                   kCtaThreads, kWorkerSmallSmemBytes));
  D->worker_small_mode = env_int("B2_WORKER_SMALL", -1);
  D->claim_k_override = env_int("B2_CLAIM_K", 0);
  memset(D->h_fn_kind, 0, sizeof D->h_fn_kind);
  D->pend_bytes = 0;
  if (D->worker_tma)
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D->worker_ctas_per_sm, worker_tma_kernel,
                                                     kCtaThreads, kWorkerSmemBytes));
  else
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D->worker_ctas_per_sm, worker_kernel,
                                                     kCtaThreads, 0));
  D->mover_unroll = env_int("B2_MOVER_UNROLL", 8);
  if (D->mover_unroll != 4 && D->mover_unroll != 16) D->mover_unroll = 8;
  switch (D->mover_unroll) {
    case 4:
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D->mover_ctas_per_sm, span_move_kernel<4>,
                                                       kMoveThreads, 0));
      break;
    case 16:
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D->mover_ctas_per_sm,
                                                       span_move_kernel<16>, kMoveThreads, 0));
      break;
    default:
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D->mover_ctas_per_sm, span_move_kernel<8>,
                                                       kMoveThreads, 0));
  }
  {
    const char *mv = getenv("B2_MOVER");
    // default: the bulk-copy-engine kernel whenever every span is 16 B aligned (measured
    // +2..3 % over the register path at every payload size, profiles/r01_mover_variants.log)
    D->mover_wide = mv && std::string(mv) == "wide";
    D->mover_tma = !(mv && (std::string(mv) == "ldg" || std::string(mv) == "wide"));
    CU(cudaFuncSetAttribute(span_move_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)kTmaSmemBytes));
  }
  const int wcap = env_int("B2_WORKER_CTAS_PER_SM", 0);
  if (wcap > 0) D->worker_ctas_per_sm = std::min(D->worker_ctas_per_sm, wcap);
  if (wcap > 0) D->worker_small_ctas_per_sm = std::min(D->worker_small_ctas_per_sm, wcap);
  const int mcap = env_int("B2_MOVER_CTAS_PER_SM", 0);
  if (mcap > 0) D->mover_ctas_per_sm = std::min(D->mover_ctas_per_sm, mcap);

  // Size every staging slot up front so that steady-state calls never allocate
  // (cudaHostAlloc / cudaMalloc synchronise the device).
  for (auto &s : D->task_ring) {
    int rc = staging_reserve(s, (size_t)queue_cap * sizeof(b2_task) + (size_t)D->succ_cap * 4 +
                                    (size_t)queue_cap * 4);
    if (rc) return rc;
  }
  for (auto &s : D->span_ring) {
    int rc = staging_reserve(s, (size_t)1 << 20);
    if (rc) return rc;
  }
  // fn_id 0 stays kind NONE: the gate electron, a registered no-op used to fan
  // one dependency out to many successors.
  CU(cudaStreamSynchronize(D->stream));
  for (int i = 0; i < kTickets; ++i) {
    D->ticket_recorded[i] = false;
    D->ticket_err[i] = 0;
    D->ticket_phase[i] = kPhaseNone;
    D->ticket_slot[i] = -1;
    D->ticket_ntasks[i] = 0;
  }
  {
    std::lock_guard<std::mutex> ql(D->qmu);
    D->queue.clear();
    D->q_busy = false;
    D->q_stop = false;
    D->q_current = nullptr;
  }
  D->res_enabled = false;
  D->res_live = false;
  D->compute_dirty = false;
  if (env_int("B2_RESIDENT", 0) == 1) {
    int rc = resident_setup(D, (uint32_t)std::max(0, env_int("B2_RESIDENT_IDLE_US", 0)));
    if (rc) return rc;
  }
  D->submitter = std::thread(submit_loop, D);
  D->up = true;
  return B2_OK;
}

int b2_shutdown(int dev) {
  Device *D = get(dev);
  if (!D) return fail(B2_E_ARG, "device ordinal %d out of range", dev);
  {
    std::lock_guard<std::mutex> lk(D->mu);
    if (!D->up) return fail(B2_E_STATE, "device %d is not initialised", dev);
  }
  // queued submissions are issued, then the submission thread leaves (it needs the device
  // lock to finish, so this happens before we take it)
  {
    std::lock_guard<std::mutex> ql(D->qmu);
    D->q_stop = true;
  }
  D->qcv.notify_all();
  if (D->submitter.joinable()) D->submitter.join();
  std::lock_guard<std::mutex> lock_(D->mu);
  if (!D->up) return fail(B2_E_STATE, "device %d is not initialised", dev);
  cudaSetDevice(dev);
  D->res_enabled = false;
  resident_quiesce(D);
  cudaStreamSynchronize(D->upload);
  cudaStreamSynchronize(D->stream);
  cudaStreamSynchronize(D->egress);
  if (D->d_ctl) {
    cudaStreamDestroy(D->res_stream);
    cudaFree(D->d_ctl);
    cudaFreeHost(D->h_res);
    D->d_ctl = nullptr;
    D->h_res = D->d_res = nullptr;
    D->h_const = nullptr;
    D->res_stream = nullptr;
  }
  for (int i = 1; i < D->n_arenas; ++i)
    if (D->arena_ipc[i] && D->arena[i]) cudaIpcCloseMemHandle(D->arena[i]);
  for (auto &s : D->task_ring) staging_free(s);
  for (auto &s : D->span_ring) staging_free(s);
  for (auto &t : D->tables) {
    if (t.d_pristine) cudaFree(t.d_pristine);
    if (t.d_work) cudaFree(t.d_work);
    if (t.d_status) cudaFree(t.d_status);
  }
  D->tables.clear();
  for (auto &e : D->tickets)
    if (e) { cudaEventDestroy(e); e = nullptr; }
  for (uint32_t i = 0; i < D->n_user_ev; ++i)
    if (D->user_ev[i]) { cudaEventDestroy(D->user_ev[i]); D->user_ev[i] = nullptr; }
  cudaFree(D->arena[0]);
  if (D->h_in) cudaFreeHost(D->h_in);
  if (D->h_out) cudaFreeHost(D->h_out);
  cudaFree(D->d_cells);
  cudaFree(D->d_head);
  cudaFree(D->d_fn);
  cudaStreamDestroy(D->stream);
  cudaStreamDestroy(D->upload);
  cudaStreamDestroy(D->egress);
  if (D->to_egress) { cudaEventDestroy(D->to_egress); D->to_egress = nullptr; }
  for (int i = 0; i < B2_MAX_ARENAS; ++i) {
    D->arena[i] = nullptr;
    D->arena_bytes[i] = 0;
    D->arena_ipc[i] = false;
  }
  D->h_in = D->h_out = nullptr;
  D->h_in_bytes = D->h_out_bytes = 0;
  D->n_arenas = 0;
  D->up = false;
  D->ticket_cv.notify_all();
  return B2_OK;
}

int b2_get_info(int dev, b2_device_info *out) {
  if (!out) return fail(B2_E_ARG, "out is NULL");
  DEV_OR_FAIL(D, dev);
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, dev));
  memset(out, 0, sizeof *out);
  out->cc_major = prop.major;
  out->cc_minor = prop.minor;
  out->sm_count = prop.multiProcessorCount;
  out->l2_bytes_mb = prop.l2CacheSize >> 20;
  out->total_mem = prop.totalGlobalMem;
  out->arena_bytes = D->arena_bytes[0];
  out->staging_in_bytes = D->h_in_bytes;
  out->staging_out_bytes = D->h_out_bytes;
  out->queue_cap = D->qcap;
  out->abi_version = B2_ABI_VERSION;
  strncpy(out->name, prop.name, sizeof(out->name) - 1);
  return B2_OK;
}

int b2_get_stats(int dev, b2_stats *out) {
  if (!out) return fail(B2_E_ARG, "out is NULL");
  DEV_OR_FAIL(D, dev);
  *out = D->stats;
  out->resident_instances = D->res_launches;
  out->resident_batches = D->res_batches;
  return B2_OK;
}

// ------------------------------------------------------------------ memory --
int b2_host_buffer(int dev, int which, void **ptr, uint64_t *bytes) {
  if (!ptr || !bytes) return fail(B2_E_ARG, "ptr/bytes is NULL");
  DEV_OR_FAIL(D, dev);
  if (which == 0) {
    *ptr = D->h_in;
    *bytes = D->h_in_bytes;
  } else if (which == 1) {
    *ptr = D->h_out;
    *bytes = D->h_out_bytes;
  } else {
    return fail(B2_E_ARG, "which must be 0 (ingress) or 1 (egress)");
  }
  return B2_OK;
}

int b2_arena_base(int dev, uint32_t arena_id, void **ptr, uint64_t *bytes) {
  if (!ptr || !bytes) return fail(B2_E_ARG, "ptr/bytes is NULL");
  DEV_OR_FAIL(D, dev);
  if (arena_id >= (uint32_t)D->n_arenas) return fail(B2_E_ARG, "unknown arena id %u", arena_id);
  *ptr = D->arena[arena_id];
  *bytes = D->arena_bytes[arena_id];
  return B2_OK;
}

int b2_detach_host_buffer(int dev, int which) {
  DEV_ORDERED(D, dev);
  if (which == 0) {
    D->h_in = nullptr;
    D->h_in_bytes = 0;
  } else if (which == 1) {
    D->h_out = nullptr;
    D->h_out_bytes = 0;
  } else {
    return fail(B2_E_ARG, "which must be 0 (ingress) or 1 (egress)");
  }
  return B2_OK;
}

int b2_h2d(int dev, uint64_t host_off, uint64_t dev_off, uint64_t bytes) {
  DEV_ORDERED_SYNCED(D, dev);
  if (host_off + bytes > D->h_in_bytes || host_off + bytes < host_off)
    return fail(B2_E_ARG, "h2d: host range [%llu, +%llu) outside ingress staging (%llu bytes)",
                (unsigned long long)host_off, (unsigned long long)bytes,
                (unsigned long long)D->h_in_bytes);
  if (dev_off + bytes > D->arena_bytes[0] || dev_off + bytes < dev_off)
    return fail(B2_E_ARG, "h2d: device range outside arena");
  if (!bytes) return B2_OK;
  CU(cudaMemcpyAsync(D->arena[0] + dev_off, D->h_in + host_off, bytes, cudaMemcpyHostToDevice,
                     D->stream));
  D->compute_dirty = true;
  D->stats.bytes_h2d += bytes;
  return B2_OK;
}

int b2_d2h(int dev, uint64_t dev_off, uint64_t host_off, uint64_t bytes) {
  DEV_ORDERED(D, dev);
  if (host_off + bytes > D->h_out_bytes || host_off + bytes < host_off)
    return fail(B2_E_ARG, "d2h: host range [%llu, +%llu) outside egress staging (%llu bytes)",
                (unsigned long long)host_off, (unsigned long long)bytes,
                (unsigned long long)D->h_out_bytes);
  if (dev_off + bytes > D->arena_bytes[0] || dev_off + bytes < dev_off)
    return fail(B2_E_ARG, "d2h: device range outside arena");
  if (!bytes) return B2_OK;
  int rc = egress_after_compute(D);
  if (rc) return rc;
  CU(cudaMemcpyAsync(D->h_out + host_off, D->arena[0] + dev_off, bytes, cudaMemcpyDeviceToHost,
                     D->egress));
  D->egress_dirty = true;
  D->stats.bytes_d2h += bytes;
  return B2_OK;
}

int b2_write(int dev, uint64_t dev_off, const void *src, uint64_t bytes) {
  DEV_ORDERED_SYNCED(D, dev);
  if (!src && bytes) return fail(B2_E_ARG, "src is NULL");
  if (dev_off + bytes > D->arena_bytes[0] || dev_off + bytes < dev_off)
    return fail(B2_E_ARG, "write: device range outside arena");
  if (!bytes) return B2_OK;
  CU(cudaMemcpyAsync(D->arena[0] + dev_off, src, bytes, cudaMemcpyHostToDevice, D->stream));
  CU(cudaStreamSynchronize(D->stream));
  return B2_OK;
}

int b2_read(int dev, uint64_t dev_off, void *dst, uint64_t bytes) {
  DEV_ORDERED_SYNCED(D, dev);
  if (!dst && bytes) return fail(B2_E_ARG, "dst is NULL");
  if (dev_off + bytes > D->arena_bytes[0] || dev_off + bytes < dev_off)
    return fail(B2_E_ARG, "read: device range outside arena");
  if (!bytes) return B2_OK;
  CU(cudaMemcpyAsync(dst, D->arena[0] + dev_off, bytes, cudaMemcpyDeviceToHost, D->stream));
  CU(cudaStreamSynchronize(D->stream));
  return B2_OK;
}

int b2_fill(int dev, uint64_t dev_off, int byte_value, uint64_t bytes) {
  DEV_ORDERED_SYNCED(D, dev);
  if (dev_off + bytes > D->arena_bytes[0] || dev_off + bytes < dev_off)
    return fail(B2_E_ARG, "fill: device range outside arena");
  if (!bytes) return B2_OK;
  CU(cudaMemsetAsync(D->arena[0] + dev_off, byte_value, bytes, D->stream));
  D->compute_dirty = true;
  return B2_OK;
}

// ----------------------------------------------------------- pack / unpack --
int b2_move_spans(int dev, const b2_span *spans, uint32_t n) {
  Range r_("b2_move_spans");
  DEV_ORDERED(D, dev);
  return move_spans_locked(D, spans, n, 0, 0);
}

static int pack_like(int dev, bool gather, const uint64_t *scattered, const uint64_t *sizes,
                     uint32_t n, uint64_t run_off, uint32_t align, uint64_t *out_offs,
                     uint64_t *out_total) {
  if (n && (!scattered || !sizes)) return fail(B2_E_ARG, "offsets/sizes is NULL");
  if (align == 0 || (align & (align - 1))) return fail(B2_E_ARG, "align must be a power of two");
  Range r_(gather ? "b2_pack" : "b2_unpack");
  DEV_ORDERED(D, dev);
  std::vector<b2_span> spans(n);
  uint64_t cur = run_off;
  for (uint32_t i = 0; i < n; ++i) {
    if (gather) {
      spans[i].src_off = scattered[i];
      spans[i].dst_off = cur;
    } else {
      spans[i].src_off = cur;
      spans[i].dst_off = scattered[i];
    }
    spans[i].bytes = sizes[i];
    if (out_offs) out_offs[i] = cur;
    cur += (sizes[i] + align - 1) & ~(uint64_t)(align - 1);
  }
  if (out_total) *out_total = cur - run_off;
  return move_spans_locked(D, spans.data(), n, 0, 0);
}

int b2_pack(int dev, const uint64_t *src_offs, const uint64_t *sizes, uint32_t n, uint64_t dst_off,
            uint32_t align, uint64_t *out_offs, uint64_t *out_total) {
  return pack_like(dev, true, src_offs, sizes, n, dst_off, align, out_offs, out_total);
}

int b2_unpack(int dev, uint64_t src_off, const uint64_t *dst_offs, const uint64_t *sizes, uint32_t n,
              uint32_t align, uint64_t *out_offs, uint64_t *out_total) {
  return pack_like(dev, false, dst_offs, sizes, n, src_off, align, out_offs, out_total);
}

// -------------------------------------------------- callback table + queue --
int b2_register(int dev, uint32_t fn_id, int kind, double alpha, double beta, int64_t ialpha,
                int64_t ibeta) {
  DEV_ORDERED_SYNCED(D, dev);
  if (fn_id >= B2_MAX_FN) return fail(B2_E_ARG, "fn_id %u >= B2_MAX_FN", fn_id);
  if (kind < 0 || kind >= B2_K__COUNT) return fail(B2_E_ARG, "unknown callback kind %d", kind);
  if (fn_id == 0 && kind != B2_K_NONE)
    return fail(B2_E_ARG, "fn_id 0 is reserved for the gate electron");
  FnEntry e;
  memset(&e, 0, sizeof e);
  e.kind = (uint32_t)kind;
  e.alpha = (float)alpha;
  e.beta = (float)beta;
  e.ialpha = ialpha;
  e.ibeta = ibeta;
  // 32-byte entry: passing it by value through a synchronous copy keeps the
  // host copy alive for the duration.
  CU(cudaMemcpyAsync(D->d_fn + fn_id, &e, sizeof e, cudaMemcpyHostToDevice, D->stream));
  CU(cudaStreamSynchronize(D->stream));
  D->h_fn_kind[fn_id] = (uint8_t)kind;
  return B2_OK;
}

// Appends n task records to the pending batch in the current ring slot.  One pass over the
// table: each record is copied and fixed up (successor range made batch-relative, dependency
// count biased) in the same loop that tracks the ready prefix; a batch the caller declares
// independent (plain fan-out) is a straight memcpy.  Caller holds the device lock.
static int enqueue_locked(Device *D, const b2_task *tasks, uint32_t n, const uint32_t *succ,
                          uint32_t n_succ, uint32_t *first_index, bool independent) {
  if (n && !tasks) return fail(B2_E_ARG, "tasks is NULL");
  if (n_succ && !succ) return fail(B2_E_ARG, "succ is NULL but n_succ = %u", n_succ);
  if ((uint64_t)D->pend_tasks + n > D->qcap)
    return fail(B2_E_NOMEM, "batch of %u + %u electrons exceeds queue capacity %u", D->pend_tasks,
                n, D->qcap);
  if ((uint64_t)D->pend_succ + n_succ > D->succ_cap)
    return fail(B2_E_NOMEM, "successor list of %u + %u exceeds capacity %u", D->pend_succ, n_succ,
                D->succ_cap);
  Staging &s = D->task_ring[D->task_slot];
  const size_t succ_off = (size_t)D->qcap * sizeof(b2_task);
  if (D->pend_tasks == 0 && D->pend_succ == 0) {
    int rc = staging_reserve(s, succ_off + (size_t)D->succ_cap * 4 + (size_t)D->qcap * 4);
    if (rc) return rc;
  }
  if (first_index) *first_index = D->pend_tasks;
  b2_task *ht = reinterpret_cast<b2_task *>(s.h) + D->pend_tasks;
  if (independent) {
    memcpy(ht, tasks, (size_t)n * sizeof(b2_task));
    if (D->pend_prefix_open) D->pend_prefix += n;
  } else {
    bool open = D->pend_prefix_open;
    uint32_t prefix = D->pend_prefix;
    uint64_t bytes = 0;
    for (uint32_t i = 0; i < n; ++i) {
      b2_task t = tasks[i];
      bytes += (t.fn_id < B2_MAX_FN && D->h_fn_kind[t.fn_id] == B2_K_COPY) ? t.count : t.count * 4;
      if (t.succ_count) {
        if ((uint64_t)t.succ_begin + t.succ_count > n_succ)
          return fail(B2_E_ARG, "task %u: successor range [%u, +%u) outside succ[%u]", i,
                      t.succ_begin, t.succ_count, n_succ);
        t.succ_begin += D->pend_succ;   // ranges are given relative to this call's succ array
      }
      if (t.deps) {
        if (t.deps >= kDepBias) return fail(B2_E_ARG, "task %u: deps too large", i);
        t.deps += kDepBias;  // see kDepBias in b2_device.cuh
        open = false;
      } else if (open) {
        ++prefix;
      }
      ht[i] = t;
    }
    D->pend_prefix_open = open;
    D->pend_prefix = prefix;
    D->pend_bytes += bytes;
  }
  if (n_succ)
    memcpy(reinterpret_cast<uint32_t *>(s.h + succ_off) + D->pend_succ, succ, (size_t)n_succ * 4);
  D->pend_tasks += n;
  D->pend_succ += n_succ;
  return B2_OK;
}

int b2_enqueue(int dev, const b2_task *tasks, uint32_t n, const uint32_t *succ, uint32_t n_succ,
               uint32_t *first_index) {
  DEV_ORDERED(D, dev);
  int rc = enqueue_locked(D, tasks, n, succ, n_succ, first_index, false);
  if (rc == B2_OK && D->pend_tasks) D->pend_external = true;
  return rc;
}

// ---- tickets ---------------------------------------------------------------
// A ticket is a sequence number; slot ticket % kTickets holds its event and state.
static uint64_t alloc_ticket(Device *D, TicketPhase phase) {
  const uint64_t t = D->next_ticket++;
  const int i = (int)(t % kTickets);
  D->ticket_recorded[i] = false;
  D->ticket_err[i] = 0;
  D->ticket_phase[i] = phase;
  D->ticket_slot[i] = -1;
  D->ticket_ntasks[i] = 0;
  D->ticket_res_seq[i] = 0;
  return t;
}

// Ticket of a batch that went to the resident worker and queued nothing else (no status copy, no
// download): it is complete when the worker's `done` word — mirrored into mapped host memory — has
// reached the batch's sequence number.  No CUDA event is recorded for it.
static void resident_ticket(Device *D, uint64_t t, uint32_t seq_plus_one) {
  const int i = (int)(t % kTickets);
  D->ticket_res_seq[i] = seq_plus_one;
  D->ticket_recorded[i] = true;
}

// Records ticket t's event after everything issued so far (compute stream and, if copies are
// queued there, the egress stream).
static int record_ticket_event(Device *D, uint64_t t) {
  {
    int rc_ = compute_after_resident(D);
    if (rc_) return rc_;
  }
  const int i = (int)(t % kTickets);
  cudaEvent_t &ev = D->tickets[i];
  if (!ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  if (D->egress_dirty) {
    // the ticket covers the compute stream AND the copies queued on the egress stream
    int rc = egress_after_compute(D);
    if (rc) return rc;
    CU(cudaEventRecord(ev, D->egress));
    D->egress_dirty = false;
  } else {
    CU(cudaEventRecord(ev, D->stream));
  }
  D->ticket_recorded[i] = true;
  return B2_OK;
}

static int record_ticket(Device *D, uint64_t *ticket) {
  const uint64_t t = alloc_ticket(D, kPhaseNone);
  int rc = record_ticket_event(D, t);
  if (rc) return rc;
  if (ticket) *ticket = t;
  return B2_OK;
}

// Launches the worker over a task table that is already (or will be, in stream order) on the
// device.  Picks the streaming flavour from the average operand size: electrons of a few KiB run
// on the 12 x 4 KiB-stage kernel and claim several tickets per fetch-add.
static int launch_worker(Device *D, b2_task *d_tasks, const uint32_t *d_succ, uint32_t *d_status,
                         uint32_t n, uint32_t prefix, uint64_t operand_bytes) {
  if (D->res_enabled && D->worker_tma) return resident_publish(D, d_tasks, d_succ, d_status, n, prefix, operand_bytes);
  WorkerArgs a;
  fill_worker_consts(D, a);
  a.tasks = d_tasks;
  a.succ = d_succ;
  // status words live in the slot's own device buffer: the copy of batch k's statuses runs on
  // the egress stream while kernel k+1 (another slot) is already writing its own
  a.status = d_status;
  if (++D->epoch == 0) D->epoch = 1;
  a.q.epoch = D->epoch;
  a.n_tasks = n;
  a.n_ready_prefix = prefix;
  const uint64_t avg = n ? operand_bytes / n : 0;
  bool small = D->worker_tma && operand_bytes && avg <= 8192;
  if (D->worker_small_mode == 0) small = false;
  if (D->worker_small_mode == 1) small = D->worker_tma;
  const int per_sm = small ? D->worker_small_ctas_per_sm : D->worker_ctas_per_sm;
  const int cap = D->sm_count * std::max(1, per_sm);
  const int grid = (int)std::min<uint32_t>(n, (uint32_t)cap);
  uint32_t k = 1;
  if (operand_bytes) k = avg <= 2048 ? 8u : avg <= 8192 ? 4u : avg <= 32768 ? 2u : 1u;
  // the small flavour's producer warp plans one electron per lane; a claim never exceeds its 12-slot ring
  // (24 x 2 KiB stages with claims of 16 were tried for <= 2 KiB electrons: no gain at 1 KiB, 0.40 vs 0.41)
  if (small) k = 8u;
  while (k > 1 && (uint64_t)grid * k * 4 > n) k >>= 1;   // keep >= 4 claims per CTA: the tail stays short
  if (D->claim_k_override > 0) k = (uint32_t)D->claim_k_override;
  a.claim_k = k;
  if (!D->worker_tma)
    worker_kernel<<<grid, kCtaThreads, 0, D->stream>>>(a);
  else if (small)
    worker_tma_small_kernel<<<grid, kCtaThreads, kWorkerSmallSmemBytes, D->stream>>>(a);
  else
    worker_tma_kernel<<<grid, kCtaThreads, kWorkerSmemBytes, D->stream>>>(a);
  CU(cudaGetLastError());
  D->stats.launches_worker += 1;
  D->stats.electrons += n;
  return B2_OK;
}

// Uploads the pending batch's table, launches the worker over it and queues the status copy.
// `t` (0 = none) is the ticket that will cover this launch: its slot is remembered for
// b2_cancel.  Caller holds the device lock and records the ticket afterwards.
static int launch_locked(Device *D, uint64_t status_host_off, uint64_t t) {
  const uint32_t n = D->pend_tasks;
  if (n == 0) return B2_OK;
  if (status_host_off != UINT64_MAX &&
      (status_host_off + (uint64_t)n * 4 > D->h_out_bytes || (status_host_off & 3)))
    return fail(B2_E_ARG, "status_host_off %llu (+%u words) outside egress staging",
                (unsigned long long)status_host_off, n);
  Staging &s = D->task_ring[D->task_slot];
  const size_t succ_off = (size_t)D->qcap * sizeof(b2_task);
  // The table travels on its own stream: the upload for this launch overlaps
  // whatever the compute stream is still running; the kernel waits for it.
  if (!s.up) CU(cudaEventCreateWithFlags(&s.up, cudaEventDisableTiming));
  CU(cudaMemcpyAsync(s.d, s.h, (size_t)n * sizeof(b2_task), cudaMemcpyHostToDevice, D->upload));
  if (D->pend_succ)
    CU(cudaMemcpyAsync(s.d + succ_off, s.h + succ_off, (size_t)D->pend_succ * 4,
                       cudaMemcpyHostToDevice, D->upload));
  if (!(D->res_enabled && D->worker_tma)) {   // resident worker: the doorbell follows the table on the upload stream
    CU(cudaEventRecord(s.up, D->upload));
    CU(cudaStreamWaitEvent(D->stream, s.up, 0));
  }

  uint32_t *slot_status = reinterpret_cast<uint32_t *>(s.d + succ_off + (size_t)D->succ_cap * 4);
  D->res_last_pure = 0;
  {
    int rc_ = launch_worker(D, reinterpret_cast<b2_task *>(s.d), reinterpret_cast<const uint32_t *>(s.d + succ_off),
                            slot_status, n, D->pend_prefix, D->pend_bytes);
    if (rc_) return rc_;
  }
  CU(cudaGetLastError());
  // the slot (pinned table + device mirror + status words) is reusable once this kernel — and the
  // status copy below, if any — is done
  if (status_host_off == UINT64_MAX) {
    if (D->res_last_pure) {
      // resident batch with nothing else queued behind it: the slot is held by the batch's sequence
      // number (the worker's done word in mapped host memory), no event
      s.res_done = &D->h_res->done;
      s.res_seq = D->res_last_pure;
    } else {
      CU(cudaEventRecord(s.ev, D->stream));
      s.in_flight = true;
    }
  } else {
    D->res_last_pure = 0;                   // a download follows: the ticket is an event on the egress stream
    int rc_ = egress_after_compute(D);      // (also orders the compute stream after the resident batch)
    if (rc_) return rc_;
    CU(cudaMemcpyAsync(D->h_out + status_host_off, slot_status, (size_t)n * 4, cudaMemcpyDeviceToHost,
                       D->egress));
    CU(cudaEventRecord(s.ev, D->egress));
    s.in_flight = true;
    D->egress_dirty = true;
    D->stats.bytes_d2h += (uint64_t)n * 4;
  }
  D->stats.bytes_h2d += (uint64_t)n * sizeof(b2_task) + (uint64_t)D->pend_succ * 4;
  if (t) {
    const int ti = (int)(t % kTickets);
    D->ticket_slot[ti] = (int8_t)D->task_slot;
    D->ticket_ntasks[ti] = n;
    D->ticket_phase[ti] = kPhaseLaunched;
  }
  D->task_slot = (D->task_slot + 1) % kRing;
  D->pend_tasks = D->pend_succ = 0;
  D->pend_prefix = 0;
  D->pend_prefix_open = true;
  D->pend_external = false;
  D->pend_bytes = 0;
  return B2_OK;
}

int b2_launch(int dev, uint64_t status_host_off, uint64_t *ticket) {
  Range r_("b2_launch");
  DEV_ORDERED(D, dev);
  const uint64_t t = alloc_ticket(D, kPhaseNone);
  D->res_last_pure = 0;
  int rc = launch_locked(D, status_host_off, t);
  if (rc == B2_OK) {
    if (D->res_last_pure) resident_ticket(D, t, D->res_last_pure);
    else rc = record_ticket_event(D, t);
  }
  if (rc) {
    // the ticket exists but will never complete normally: make poll / wait report the failure
    const int i = (int)(t % kTickets);
    D->ticket_err[i] = rc;
    D->ticket_msg[i] = g_err;
    D->ticket_recorded[i] = true;
    D->pend_tasks = D->pend_succ = 0;
    D->pend_prefix = 0;
    D->pend_prefix_open = true;
    D->pend_external = false;
    return rc;
  }
  if (ticket) *ticket = t;
  return B2_OK;
}

// ---- resident task tables ------------------------------------------------------
// A lattice that is dispatched again and again (same graph, same arena offsets) need not have
// its table rebuilt, copied into pinned memory and uploaded per launch: b2_table_create uploads
// it once, b2_table_launch re-arms the dependency counters device-side (one D2D copy, skipped for
// plain fan-outs) and launches.  The per-launch host cost becomes independent of the table size.
constexpr uint32_t kMaxTables = 256;

int b2_table_create(int dev, const b2_task *tasks, uint32_t n, const uint32_t *succ, uint32_t n_succ,
                    uint32_t *table_id) {
  if (!table_id) return fail(B2_E_ARG, "table_id is NULL");
  if (!n || !tasks) return fail(B2_E_ARG, "empty table");
  if (n_succ && !succ) return fail(B2_E_ARG, "succ is NULL but n_succ = %u", n_succ);
  DEV_ORDERED_SYNCED(D, dev);
  if (n > D->qcap) return fail(B2_E_NOMEM, "table of %u electrons exceeds queue capacity %u", n, D->qcap);
  uint32_t id = 0;
  for (; id < D->tables.size(); ++id)
    if (!D->tables[id].live) break;
  if (id == D->tables.size()) {
    if (id >= kMaxTables) return fail(B2_E_NOMEM, "too many resident tables");
    D->tables.emplace_back();
  }
  std::vector<b2_task> fixed(tasks, tasks + n);
  uint32_t prefix = 0;
  bool open = true, has_deps = false;
  uint64_t bytes = 0;
  for (uint32_t i = 0; i < n; ++i) {
    b2_task &t = fixed[i];
    bytes += (t.fn_id < B2_MAX_FN && D->h_fn_kind[t.fn_id] == B2_K_COPY) ? t.count : t.count * 4;
    if (t.succ_count && (uint64_t)t.succ_begin + t.succ_count > n_succ)
      return fail(B2_E_ARG, "task %u: successor range [%u, +%u) outside succ[%u]", i, t.succ_begin,
                  t.succ_count, n_succ);
    if (t.deps) {
      if (t.deps >= kDepBias) return fail(B2_E_ARG, "task %u: deps too large", i);
      t.deps += kDepBias;
      open = false;
      has_deps = true;
    } else if (open) {
      ++prefix;
    }
  }
  Device::Table tb;
  const size_t tab_bytes = (size_t)n * sizeof(b2_task);
  CU(cudaMalloc(reinterpret_cast<void **>(&tb.d_pristine), tab_bytes + (size_t)std::max<uint32_t>(n_succ, 1) * 4));
  if (has_deps) CU(cudaMalloc(reinterpret_cast<void **>(&tb.d_work), tab_bytes));
  CU(cudaMalloc(reinterpret_cast<void **>(&tb.d_status), (size_t)n * 4));
  CU(cudaMemcpyAsync(tb.d_pristine, fixed.data(), tab_bytes, cudaMemcpyHostToDevice, D->stream));
  if (n_succ)
    CU(cudaMemcpyAsync(tb.d_pristine + tab_bytes, succ, (size_t)n_succ * 4, cudaMemcpyHostToDevice, D->stream));
  CU(cudaStreamSynchronize(D->stream));   // `fixed` and `succ` are pageable: they must outlive the copies
  tb.n = n;
  tb.n_succ = n_succ;
  tb.prefix = prefix;
  tb.bytes = bytes;
  tb.has_deps = has_deps;
  tb.live = true;
  D->tables[id] = tb;
  *table_id = id;
  return B2_OK;
}

int b2_table_launch(int dev, uint32_t table_id, uint64_t status_host_off, uint64_t *ticket) {
  Range r_("b2_table_launch");
  DEV_ORDERED(D, dev);
  if (table_id >= D->tables.size() || !D->tables[table_id].live)
    return fail(B2_E_ARG, "unknown table %u", table_id);
  Device::Table &tb = D->tables[table_id];
  if (status_host_off != UINT64_MAX &&
      (status_host_off + (uint64_t)tb.n * 4 > D->h_out_bytes || (status_host_off & 3)))
    return fail(B2_E_ARG, "status_host_off %llu (+%u words) outside egress staging",
                (unsigned long long)status_host_off, tb.n);
  const size_t tab_bytes = (size_t)tb.n * sizeof(b2_task);
  b2_task *d_tasks = reinterpret_cast<b2_task *>(tb.d_pristine);
  if (tb.has_deps) {   // the worker counts the dependency words down: every launch starts from the pristine copy
    {   // the previous launch of this table must be done with the working copy before it is re-armed
      int rc_ = compute_after_resident(D);
      if (rc_) return rc_;
    }
    CU(cudaMemcpyAsync(tb.d_work, tb.d_pristine, tab_bytes, cudaMemcpyDeviceToDevice, D->stream));
    D->compute_dirty = true;
    d_tasks = reinterpret_cast<b2_task *>(tb.d_work);
  }
  D->res_last_pure = 0;
  int rc = launch_worker(D, d_tasks, reinterpret_cast<const uint32_t *>(tb.d_pristine + tab_bytes), tb.d_status,
                         tb.n, tb.prefix, tb.bytes);
  if (rc) return rc;
  if (status_host_off == UINT64_MAX && D->res_last_pure) {   // resident batch, nothing else queued: no event
    const uint64_t t = alloc_ticket(D, kPhaseNone);
    resident_ticket(D, t, D->res_last_pure);
    if (ticket) *ticket = t;
    return B2_OK;
  }
  if (status_host_off != UINT64_MAX) {
    rc = egress_after_compute(D);
    if (rc) return rc;
    CU(cudaMemcpyAsync(D->h_out + status_host_off, tb.d_status, (size_t)tb.n * 4, cudaMemcpyDeviceToHost,
                       D->egress));
    D->egress_dirty = true;
    D->stats.bytes_d2h += (uint64_t)tb.n * 4;
  }
  return record_ticket(D, ticket);
}

int b2_table_destroy(int dev, uint32_t table_id) {
  DEV_ORDERED_SYNCED(D, dev);
  if (table_id >= D->tables.size() || !D->tables[table_id].live)
    return fail(B2_E_ARG, "unknown table %u", table_id);
  CU(cudaStreamSynchronize(D->stream));
  CU(cudaStreamSynchronize(D->egress));
  Device::Table &tb = D->tables[table_id];
  cudaFree(tb.d_pristine);
  if (tb.d_work) cudaFree(tb.d_work);
  cudaFree(tb.d_status);
  tb = Device::Table();
  return B2_OK;
}

// ---- resident worker ------------------------------------------------------------
static int resident_setup(Device *D, uint32_t idle_us) {
  int rc = load_stream_memops();
  if (rc) return rc;
  if (!D->worker_tma) return fail(B2_E_UNSUPPORTED, "the resident worker is the streaming worker (unset B2_WORKER=classic)");
  if (!D->d_ctl) {
    CU(cudaStreamCreateWithFlags(&D->res_stream, cudaStreamNonBlocking));
    CU(cudaMalloc(reinterpret_cast<void **>(&D->d_ctl), sizeof(ResidentCtl)));
    CU(cudaMemset(D->d_ctl, 0, sizeof(ResidentCtl)));
    uint8_t *blk = nullptr;
    const size_t blk_bytes = ((sizeof(ResidentHost) + 255) & ~size_t(255)) + 256;
    CU(cudaHostAlloc(reinterpret_cast<void **>(&blk), blk_bytes, cudaHostAllocMapped));
    memset(blk, 0, blk_bytes);
    D->h_res = reinterpret_cast<ResidentHost *>(blk);
    D->h_const = reinterpret_cast<uint32_t *>(blk + blk_bytes - 256);
    CU(cudaHostGetDevicePointer(reinterpret_cast<void **>(&D->d_res), D->h_res, 0));
    CU(cudaFuncSetAttribute(worker_resident_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CU(cudaFuncSetAttribute(worker_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)kWorkerSmemBytes));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&D->res_ctas_per_sm, worker_resident_kernel, kCtaThreads,
                                                     kWorkerSmemBytes));
    if (D->res_ctas_per_sm < 1) return fail(B2_E_UNSUPPORTED, "the resident worker does not fit on an SM");
    D->res_ctas_per_sm = std::min(D->res_ctas_per_sm, std::max(1, D->worker_ctas_per_sm));
    D->res_submitted = 0;
    D->res_instance = 0;
  }
  if (idle_us) D->res_idle_us = idle_us;
  D->res_life_ms = (uint32_t)std::max(100, env_int("B2_RESIDENT_LIFE_MS", 2000));
  D->res_enabled = true;
  return B2_OK;
}

int b2_resident_enable(int dev, uint32_t idle_us) {
  DEV_ORDERED(D, dev);
  return resident_setup(D, idle_us);
}

int b2_resident_disable(int dev) {
  DEV_ORDERED(D, dev);
  D->res_enabled = false;
  return resident_quiesce(D);
}

int b2_mark(int dev, uint64_t *ticket) {
  DEV_ORDERED(D, dev);
  return record_ticket(D, ticket);
}

// Ticket state for poll / wait.  Returns B2_OK when its event may be queried (copied to *ev),
// 1 when the ticket is so old that it has certainly completed, B2_E_NOTREADY when the
// submission thread has not recorded it yet, or the pipeline's error.
static int ticket_lookup(Device *D, uint64_t ticket, cudaEvent_t *ev) {
  if (ticket == 0 || ticket >= D->next_ticket)
    return fail(B2_E_ARG, "unknown ticket %llu", (unsigned long long)ticket);
  if (D->next_ticket - ticket >= (uint64_t)kTickets) return 1;  // event slot recycled: long done
  const int i = (int)(ticket % kTickets);
  if (!D->ticket_recorded[i]) return B2_E_NOTREADY;
  if (D->ticket_err[i]) return fail(D->ticket_err[i], "%s", D->ticket_msg[i].c_str());
  if (D->ticket_res_seq[i]) {          // resident batch: no event, a sequence number
    if (!D->h_res || (int32_t)(D->h_res->done - D->ticket_res_seq[i]) >= 0) return 1;
    *ev = nullptr;
    return 2;
  }
  *ev = D->tickets[i];
  return B2_OK;
}

int b2_poll(int dev, uint64_t ticket) {
  DEV_OR_FAIL(D, dev);
  cudaEvent_t ev = nullptr;
  const int rc = ticket_lookup(D, ticket, &ev);
  if (rc == 1) return B2_OK;
  if (rc == 2) return B2_E_NOTREADY;
  if (rc != B2_OK) return rc;
  cudaError_t e = cudaEventQuery(ev);
  if (e == cudaSuccess) return B2_OK;
  if (e == cudaErrorNotReady) return B2_E_NOTREADY;
  return fail(B2_E_CUDA, "cudaEventQuery -> %s", cudaGetErrorString(e));
}

int b2_wait(int dev, uint64_t ticket) {
  cudaEvent_t ev = nullptr;
  {
    Device *D = get(dev);
    if (!D) return fail(B2_E_ARG, "device ordinal %d out of range", dev);
    std::unique_lock<std::mutex> lk(D->mu);
    if (!D->up) return fail(B2_E_STATE, "device %d is not initialised (call b2_init)", dev);
    int rc;
    for (;;) {
      rc = ticket_lookup(D, ticket, &ev);
      if (rc != B2_E_NOTREADY) break;
      D->ticket_cv.wait(lk);          // an asynchronous ticket: wait until it has been issued
      if (!D->up) return fail(B2_E_STATE, "device %d was shut down while waiting", dev);
    }
    if (rc == 1) return B2_OK;
    if (rc == 2) {
      // a resident batch: wait for the worker's done word (mapped host memory) outside the lock
      const volatile uint32_t *done = &D->h_res->done;
      const volatile uint32_t *err = &D->h_res->error;
      const uint32_t want = D->ticket_res_seq[ticket % kTickets];
      lk.unlock();
      for (uint64_t spins = 0; (int32_t)(*done - want) < 0; ++spins) {
        if (*err) return fail(B2_E_CUDA, "resident worker left with a submitted batch pending (lifetime cap)");
        if (spins > 2000) std::this_thread::sleep_for(std::chrono::microseconds(20));
        else if (spins > 200) std::this_thread::yield();
      }
      return B2_OK;
    }
    if (rc != B2_OK) return rc;
    CU(cudaSetDevice(dev));
  }
  // Block outside the device lock so other threads can keep submitting.
  Device *D = get(dev);
  if (D->h_res && D->h_res->error) {
    // a resident instance hit its lifetime cap while a submitted batch had not arrived: nothing
    // will ever release the streams waiting for it — say so instead of blocking forever
    for (int spins = 0; cudaEventQuery(ev) == cudaErrorNotReady; ++spins) {
      if (spins > 20000) return fail(B2_E_CUDA, "resident worker left with a submitted batch pending (lifetime cap)");
      std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
    return B2_OK;
  }
  CU(cudaEventSynchronize(ev));
  return B2_OK;
}

int b2_sync(int dev) {
  cudaStream_t st, up, eg;
  {
    DEV_ORDERED_SYNCED(D, dev);
    st = D->stream;
    up = D->upload;
    eg = D->egress;
  }
  CU(cudaStreamSynchronize(up));
  CU(cudaStreamSynchronize(st));
  CU(cudaStreamSynchronize(eg));
  return B2_OK;
}

// ------------------------------------------------- asynchronous submissions --
// Splits copy jobs into pieces for the host copy pool.
static int run_copy_jobs(Device *D, const b2_copy_job *jobs, uint32_t n) {
  if (n == 0) return B2_OK;
  if (!jobs) return fail(B2_E_ARG, "jobs is NULL");
  std::vector<b2host::CopyPiece> pieces;
  pieces.reserve(n);
  uint8_t *base;
  uint64_t cap;
  {
    std::lock_guard<std::mutex> lk(D->mu);
    base = D->h_in;
    cap = D->h_in_bytes;
  }
  for (uint32_t i = 0; i < n; ++i) {
    const b2_copy_job &j = jobs[i];
    if (j.bytes == 0) continue;
    if (!j.src) return fail(B2_E_ARG, "copy job %u: src is NULL", i);
    if (j.dst_off + j.bytes > cap || j.dst_off + j.bytes < j.dst_off)
      return fail(B2_E_ARG, "copy job %u: [%llu, +%llu) outside ingress staging (%llu bytes)", i,
                  (unsigned long long)j.dst_off, (unsigned long long)j.bytes, (unsigned long long)cap);
    const uint8_t *src = reinterpret_cast<const uint8_t *>(static_cast<uintptr_t>(j.src));
    for (uint64_t off = 0; off < j.bytes; off += kCopyPieceBytes)
      pieces.push_back({src + off, base + j.dst_off + off,
                        (size_t)std::min<uint64_t>(kCopyPieceBytes, j.bytes - off)});
  }
  b2host::copy_pieces(pieces.data(), pieces.size());
  return B2_OK;
}

int b2_stage_copy(int dev, const b2_copy_job *jobs, uint32_t n) {
  Device *D = get(dev);
  if (!D) return fail(B2_E_ARG, "device ordinal %d out of range", dev);
  {
    std::lock_guard<std::mutex> lk(D->mu);
    if (!D->up) return fail(B2_E_STATE, "device %d is not initialised (call b2_init)", dev);
  }
  return run_copy_jobs(D, jobs, n);
}

// Sets B2_F_CANCELLED on records [first, first+count) of the table in ring slot `slot`;
// when the table is already on the device the flag words are re-uploaded (4 bytes each,
// from the pinned host copy, on the upload stream — not behind the running kernel).
static int tombstone_locked(Device *D, int slot, uint32_t n_tasks, uint32_t first, uint32_t count,
                            bool on_device) {
  if (slot < 0 || first >= n_tasks) return B2_OK;
  Staging &s = D->task_ring[slot];
  b2_task *ht = reinterpret_cast<b2_task *>(s.h);
  const uint32_t end = (uint32_t)std::min<uint64_t>((uint64_t)first + count, n_tasks);
  for (uint32_t i = first; i < end; ++i) ht[i].flags |= B2_F_CANCELLED;
  if (on_device) {
    for (uint32_t i = first; i < end; ++i) {
      const size_t off = (size_t)i * sizeof(b2_task) + offsetof(b2_task, flags);
      CU(cudaMemcpyAsync(s.d + off, s.h + off, 4, cudaMemcpyHostToDevice, D->upload));
    }
  }
  return B2_OK;
}

// The per-device submission thread: issues queued batches in order.
static void submit_fail(Device *D, uint64_t t, int rc) {
  const int i = (int)(t % kTickets);
  D->ticket_err[i] = rc ? rc : B2_E_CUDA;
  D->ticket_msg[i] = g_err;
  D->ticket_recorded[i] = true;
  D->pend_tasks = D->pend_succ = 0;
  D->pend_prefix = 0;
  D->pend_prefix_open = true;
}

static void submit_one(Device *D, Submission &sub) {
  const b2_batch &b = sub.b;
  const uint64_t t = sub.ticket;
  const int ti = (int)(t % kTickets);
  int rc = B2_OK;
  // 1. make sure the ring slot is free WITHOUT holding the device lock (poll / wait stay responsive)
  {
    cudaEvent_t wait_ev = nullptr;
    {
      std::lock_guard<std::mutex> lk(D->mu);
      Staging &s = D->task_ring[D->task_slot];
      if (s.in_flight) wait_ev = s.ev;
    }
    if (wait_ev) {
      cudaSetDevice(D->dev);
      cudaEventSynchronize(wait_ev);
    }
  }
  // 2. table into the ring slot (tombstones that arrived while queued are applied to it)
  {
    std::lock_guard<std::mutex> lk(D->mu);
    cudaSetDevice(D->dev);
    if (D->pend_tasks != 0) {   // cannot happen: ordered calls drain us first, b2_submit refuses an open batch
      rc = fail(B2_E_STATE, "b2_submit while a b2_enqueue batch is open");
    } else if (b.n_tasks) {
      rc = enqueue_locked(D, b.tasks, b.n_tasks, nullptr, 0, nullptr,
                          (b.flags & B2_BATCH_INDEPENDENT) != 0);
    }
    if (rc == B2_OK) {
      if (b.flags & B2_BATCH_INDEPENDENT)   // no table scan: the operand bytes are what gets uploaded
        D->pend_bytes = b.in_bytes ? b.in_bytes : b.n_tasks;
      D->ticket_slot[ti] = (int8_t)D->task_slot;
      D->ticket_ntasks[ti] = b.n_tasks;
      {
        std::lock_guard<std::mutex> ql(D->qmu);
        D->ticket_phase[ti] = kPhaseTabled;
        for (auto &c : sub.cancels) tombstone_locked(D, D->task_slot, b.n_tasks, c.first, c.second, false);
      }
    } else {
      submit_fail(D, t, rc);
      D->ticket_cv.notify_all();
      return;
    }
  }
  // 3. operand bytes into pinned ingress memory (copy pool; no lock held)
  rc = run_copy_jobs(D, b.jobs, b.n_jobs);
  // 4. upload, worker, downloads, ticket
  {
    std::lock_guard<std::mutex> lk(D->mu);
    cudaSetDevice(D->dev);
    auto cu = [&](cudaError_t e, const char *what) {
      if (e != cudaSuccess && rc == B2_OK)
        rc = fail(B2_E_CUDA, "%s -> %s (b2_submit)", what, cudaGetErrorString(e));
    };
    if (rc == B2_OK && b.in_bytes) {
      if (b.in_off + b.in_bytes > D->h_in_bytes || b.in_off + b.in_bytes > D->arena_bytes[0])
        rc = fail(B2_E_ARG, "b2_submit: ingress range [%llu, +%llu) out of bounds",
                  (unsigned long long)b.in_off, (unsigned long long)b.in_bytes);
      else {
        // resident worker: the payload travels on the upload stream, in front of the table and the doorbell
        cu(cudaMemcpyAsync(D->arena[0] + b.in_off, D->h_in + b.in_off, b.in_bytes,
                           cudaMemcpyHostToDevice, D->res_enabled ? D->upload : D->stream), "cudaMemcpyAsync(H2D)");
        D->stats.bytes_h2d += b.in_bytes;
      }
    }
    if (rc == B2_OK) rc = launch_locked(D, b.n_tasks ? b.status_host_off : UINT64_MAX, t);
    if (rc == B2_OK && b.out_bytes) {
      if (b.out_host_off + b.out_bytes > D->h_out_bytes || b.out_dev_off + b.out_bytes > D->arena_bytes[0])
        rc = fail(B2_E_ARG, "b2_submit: egress range out of bounds");
      else {
        rc = egress_after_compute(D);
        if (rc == B2_OK) {
          cu(cudaMemcpyAsync(D->h_out + b.out_host_off, D->arena[0] + b.out_dev_off, b.out_bytes,
                             cudaMemcpyDeviceToHost, D->egress), "cudaMemcpyAsync(D2H)");
          D->egress_dirty = true;
          D->stats.bytes_d2h += b.out_bytes;
        }
      }
    }
    if (rc == B2_OK) rc = record_ticket_event(D, t);
    if (rc != B2_OK) submit_fail(D, t, rc);
  }
  D->ticket_cv.notify_all();
}

static void submit_loop(Device *D) {
  for (;;) {
    Submission sub;
    {
      std::unique_lock<std::mutex> lk(D->qmu);
      D->qcv.wait(lk, [&] { return D->q_stop || !D->queue.empty(); });
      if (D->queue.empty()) return;   // q_stop and nothing left
      sub = std::move(D->queue.front());
      D->queue.pop_front();
      D->q_busy = true;
      D->q_current = &sub;
    }
    submit_one(D, sub);
    {
      std::lock_guard<std::mutex> lk(D->qmu);
      D->q_busy = false;
      D->q_current = nullptr;
    }
    D->qidle.notify_all();
  }
}

int b2_submit(int dev, b2_batch *batch) {
  if (!batch) return fail(B2_E_ARG, "batch is NULL");
  if (batch->n_tasks && !batch->tasks) return fail(B2_E_ARG, "batch->tasks is NULL");
  if (batch->n_jobs && !batch->jobs) return fail(B2_E_ARG, "batch->jobs is NULL");
  Device *D = get(dev);
  if (!D) return fail(B2_E_ARG, "device ordinal %d out of range", dev);
  Submission sub;
  {
    std::lock_guard<std::mutex> lk(D->mu);
    if (!D->up) return fail(B2_E_STATE, "device %d is not initialised (call b2_init)", dev);
    if (batch->n_tasks > D->qcap)
      return fail(B2_E_NOMEM, "batch of %u electrons exceeds queue capacity %u", batch->n_tasks, D->qcap);
    if (D->pend_external) return fail(B2_E_STATE, "b2_submit while a b2_enqueue batch is open");
    sub.ticket = alloc_ticket(D, kPhaseQueued);
    D->ticket_ntasks[sub.ticket % kTickets] = batch->n_tasks;
  }
  batch->ticket = sub.ticket;
  sub.b = *batch;
  {
    std::lock_guard<std::mutex> lk(D->qmu);
    D->queue.push_back(std::move(sub));
  }
  D->qcv.notify_one();
  return B2_OK;
}

int b2_cancel(int dev, uint64_t ticket, uint32_t first, uint32_t count) {
  DEV_OR_FAIL(D, dev);
  if (ticket == 0 || ticket >= D->next_ticket)
    return fail(B2_E_ARG, "unknown ticket %llu", (unsigned long long)ticket);
  if (D->next_ticket - ticket >= (uint64_t)kTickets || count == 0) return B2_OK;   // long done
  const int i = (int)(ticket % kTickets);
  std::lock_guard<std::mutex> ql(D->qmu);
  switch (D->ticket_phase[i]) {
    case kPhaseQueued:
      for (auto &sub : D->queue)
        if (sub.ticket == ticket) {
          sub.cancels.emplace_back(first, count);
          return B2_OK;
        }
      // taken by the submission thread but not tabled yet: it applies the list when it tables
      // the batch (under qmu, which we hold)
      if (D->q_current && D->q_current->ticket == ticket) D->q_current->cancels.emplace_back(first, count);
      return B2_OK;
    case kPhaseTabled:
      return tombstone_locked(D, D->ticket_slot[i], D->ticket_ntasks[i], first, count, false);
    case kPhaseLaunched:
      return tombstone_locked(D, D->ticket_slot[i], D->ticket_ntasks[i], first, count, true);
    default:
      return B2_OK;
  }
}

// ------------------------------------------------------------ peer routing --
int b2_peer_enable(int dev, int peer_dev, uint32_t *arena_id) {
  if (dev == peer_dev) return fail(B2_E_ARG, "peer_dev equals dev");
  Device *P = get(peer_dev);
  if (!P) return fail(B2_E_ARG, "peer ordinal %d out of range", peer_dev);
  uint8_t *pbase;
  uint64_t pbytes;
  {
    std::lock_guard<std::mutex> lk(P->mu);
    if (!P->up) return fail(B2_E_STATE, "peer device %d is not initialised", peer_dev);
    pbase = P->arena[0];
    pbytes = P->arena_bytes[0];
  }
  DEV_OR_FAIL(D, dev);
  for (int i = 1; i < D->n_arenas; ++i)
    if (!D->arena_ipc[i] && D->arena_owner_dev[i] == peer_dev) {
      if (arena_id) *arena_id = (uint32_t)i;
      return B2_OK;
    }
  if (D->n_arenas >= B2_MAX_ARENAS) return fail(B2_E_NOMEM, "arena table full");
  int can = 0;
  CU(cudaDeviceCanAccessPeer(&can, dev, peer_dev));
  if (!can) return fail(B2_E_UNSUPPORTED, "device %d cannot access peer %d", dev, peer_dev);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_dev, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
  } else if (e != cudaSuccess) {
    return fail(B2_E_CUDA, "cudaDeviceEnablePeerAccess(%d) -> %s", peer_dev, cudaGetErrorString(e));
  }
  const int id = D->n_arenas++;
  D->arena[id] = pbase;
  D->arena_bytes[id] = pbytes;
  D->arena_owner_dev[id] = peer_dev;
  D->arena_ipc[id] = false;
  if (arena_id) *arena_id = (uint32_t)id;
  return B2_OK;
}

int b2_ipc_export(int dev, void *handle64) {
  if (!handle64) return fail(B2_E_ARG, "handle64 is NULL");
  DEV_OR_FAIL(D, dev);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, D->arena[0]));
  memcpy(handle64, &h, 64);
  return B2_OK;
}

int b2_ipc_open(int dev, const void *handle64, uint64_t peer_arena_bytes, uint32_t *arena_id) {
  if (!handle64) return fail(B2_E_ARG, "handle64 is NULL");
  DEV_OR_FAIL(D, dev);
  if (D->n_arenas >= B2_MAX_ARENAS) return fail(B2_E_NOMEM, "arena table full");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void *p = nullptr;
  CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  const int id = D->n_arenas++;
  D->arena[id] = static_cast<uint8_t *>(p);
  D->arena_bytes[id] = peer_arena_bytes;
  D->arena_owner_dev[id] = -1;
  D->arena_ipc[id] = true;
  if (arena_id) *arena_id = (uint32_t)id;
  return B2_OK;
}

int b2_peer_copy(int dev, uint32_t dst_arena_id, uint64_t src_off, uint64_t dst_off,
                 uint64_t bytes) {
  DEV_ORDERED_SYNCED(D, dev);
  if (dst_arena_id >= (uint32_t)D->n_arenas) return fail(B2_E_ARG, "unknown arena id %u", dst_arena_id);
  if (src_off + bytes > D->arena_bytes[0] || dst_off + bytes > D->arena_bytes[dst_arena_id])
    return fail(B2_E_ARG, "peer_copy: range outside arena");
  if (!bytes) return B2_OK;
  CU(cudaMemcpyAsync(D->arena[dst_arena_id] + dst_off, D->arena[0] + src_off, bytes,
                     cudaMemcpyDefault, D->stream));
  D->compute_dirty = true;
  return B2_OK;
}

int b2_peer_pack(int dev, uint32_t dst_arena_id, const b2_span *spans, uint32_t n) {
  Range r_("b2_peer_pack");
  DEV_ORDERED(D, dev);
  return move_spans_locked(D, spans, n, 0, dst_arena_id);
}

// ------------------------------------------------------------- measurement --
int b2_event_create(int dev, uint32_t *event_id) {
  if (!event_id) return fail(B2_E_ARG, "event_id is NULL");
  DEV_OR_FAIL(D, dev);
  if (D->n_user_ev >= (uint32_t)kMaxUserEvents) return fail(B2_E_NOMEM, "too many events");
  CU(cudaEventCreate(&D->user_ev[D->n_user_ev]));
  *event_id = D->n_user_ev++;
  return B2_OK;
}

int b2_event_record(int dev, uint32_t event_id) {
  DEV_ORDERED_SYNCED(D, dev);
  if (event_id >= D->n_user_ev) return fail(B2_E_ARG, "unknown event %u", event_id);
  CU(cudaEventRecord(D->user_ev[event_id], D->stream));
  return B2_OK;
}

int b2_event_elapsed_ms(int dev, uint32_t start_id, uint32_t stop_id, float *ms) {
  if (!ms) return fail(B2_E_ARG, "ms is NULL");
  cudaEvent_t a, b;
  {
    DEV_OR_FAIL(D, dev);
    if (start_id >= D->n_user_ev || stop_id >= D->n_user_ev)
      return fail(B2_E_ARG, "unknown event id");
    a = D->user_ev[start_id];
    b = D->user_ev[stop_id];
  }
  CU(cudaEventSynchronize(b));
  CU(cudaEventElapsedTime(ms, a, b));
  return B2_OK;
}

int b2_flush_l2(int dev, uint64_t scratch_off, uint64_t bytes) {
  DEV_ORDERED_SYNCED(D, dev);
  if (scratch_off + bytes > D->arena_bytes[0]) return fail(B2_E_ARG, "flush range outside arena");
  static int v = 0;
  CU(cudaMemsetAsync(D->arena[0] + scratch_off, (++v) & 0xff, bytes, D->stream));
  D->compute_dirty = true;
  return B2_OK;
}

}  // extern "C"
