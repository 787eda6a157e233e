// b2_worker.cuh — K3: persistent worker kernel, device-side MPMC queue and the
// on-device callback table.
//
// Replaces the reference's per-electron remote interpreter
// (ssh.py:377-383 `python exec_<id>.py` -> exec.py:29-46 unpickle, call,
// pickle): a registered electron is a 64-byte b2_task whose fn_id indexes the
// callback table; one CTA pops the task from the queue, calls the callback
// through a device function pointer, publishes the status word, and releases
// the electron's successors (pushing those whose dependency count reaches zero
// back into the same queue — CTAs are consumers and producers at once).
#pragma once
#include "b2_device.cuh"

namespace b2 {

// ------------------------------------------------------------------ queue ---
// Warp-aggregated push: lanes with want != 0 each enqueue one index with a
// single fetch-add on the tail per warp.  Must be called by all 32 lanes.
__device__ __forceinline__ void queue_push_warp(const Queue &q, uint32_t first_ticket, bool want,
                                                uint32_t task_index) {
  const unsigned m = __ballot_sync(0xffffffffu, want);
  if (m == 0) return;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(q.tail, (uint32_t)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (want) {
    const uint32_t t = first_ticket + base + (uint32_t)__popc(m & ((1u << lane) - 1u));
    st_release_u64(q.cells + (t & q.mask), ((unsigned long long)q.epoch << 32) | task_index);
  }
}

constexpr uint32_t kNoTask = 0xFFFFFFFFu;

// One thread per CTA: claim the next ticket (or the next n tickets: returns the first).
__device__ __forceinline__ uint32_t queue_claim(const Queue &q) { return atomicAdd(q.head, 1u); }
__device__ __forceinline__ uint32_t queue_claim_n(const Queue &q, uint32_t n) { return atomicAdd(q.head, n); }

// One thread per CTA: wait until the claimed ticket has been published.
__device__ __forceinline__ uint32_t queue_take(const Queue &q, uint32_t ticket, uint32_t n_tasks) {
  if (ticket >= n_tasks) return kNoTask;
  const unsigned long long *cell = q.cells + (ticket & q.mask);
  uint32_t backoff = 20;
  for (;;) {
    const unsigned long long v = ld_acquire_u64(cell);
    if ((uint32_t)(v >> 32) == q.epoch) return (uint32_t)v;
    __nanosleep(backoff);
    if (backoff < 640) backoff <<= 1;
  }
}

// -------------------------------------------------------------- callbacks ---
// Barrier over the 8 worker warps only (the scheduler warp never joins it).
__device__ __forceinline__ void worker_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(kWorkerThreads) : "memory");
}

struct Ctx {
  const uint8_t *a0;     // operand 0 (nullptr when immediate)
  const uint8_t *a1;     // operand 1
  uint8_t *out;
  uint64_t imm0, imm1;   // raw arg0 / arg1 words
  uint64_t count;
  uint32_t nrows;
  const FnEntry *fn;
};
typedef uint32_t (*callback_t)(const Ctx &);   // returns a b2_task_status; called by all threads

#ifndef B2_WORKER_VEC
#define B2_WORKER_VEC 4
#endif
#ifndef B2_WORKER_MIN_CTAS
#define B2_WORKER_MIN_CTAS 4
#endif
constexpr int kVec = B2_WORKER_VEC;  // 128-bit vectors in flight per thread

template <typename F>
__device__ __forceinline__ void cta_map1(const Ctx &c, F f) {
  const uint64_t n4 = c.count >> 2;
  const int4 *x = reinterpret_cast<const int4 *>(c.a0);
  int4 *o = reinterpret_cast<int4 *>(c.out);
  uint64_t i = threadIdx.x;
  for (; i + (uint64_t)(kVec - 1) * kWorkerThreads < n4; i += (uint64_t)kVec * kWorkerThreads) {
    int4 v[kVec];
#pragma unroll
    for (int j = 0; j < kVec; ++j) v[j] = ld_coherent(x + i + (uint64_t)j * kWorkerThreads);
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
      float4 t = *reinterpret_cast<float4 *>(&v[j]);
      t.x = f(t.x); t.y = f(t.y); t.z = f(t.z); t.w = f(t.w);
      st_stream(o + i + (uint64_t)j * kWorkerThreads, *reinterpret_cast<int4 *>(&t));
    }
  }
  for (; i < n4; i += kWorkerThreads) {
    int4 v = ld_coherent(x + i);
    float4 t = *reinterpret_cast<float4 *>(&v);
    t.x = f(t.x); t.y = f(t.y); t.z = f(t.z); t.w = f(t.w);
    st_stream(o + i, *reinterpret_cast<int4 *>(&t));
  }
  const uint64_t r = (n4 << 2) + threadIdx.x;  // < 4 trailing elements
  if (r < c.count)
    reinterpret_cast<float *>(c.out)[r] = f(reinterpret_cast<const float *>(c.a0)[r]);
}

template <typename F>
__device__ __forceinline__ void cta_map2(const Ctx &c, F f) {
  const uint64_t n4 = c.count >> 2;
  const int4 *x = reinterpret_cast<const int4 *>(c.a0);
  const int4 *y = reinterpret_cast<const int4 *>(c.a1);
  int4 *o = reinterpret_cast<int4 *>(c.out);
  constexpr int V = kVec / 2;
  uint64_t i = threadIdx.x;
  for (; i + (uint64_t)(V - 1) * kWorkerThreads < n4; i += (uint64_t)V * kWorkerThreads) {
    int4 a[V], b[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      a[j] = ld_coherent(x + i + (uint64_t)j * kWorkerThreads);
      b[j] = ld_coherent(y + i + (uint64_t)j * kWorkerThreads);
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float4 s = *reinterpret_cast<float4 *>(&a[j]);
      const float4 t = *reinterpret_cast<float4 *>(&b[j]);
      s.x = f(s.x, t.x); s.y = f(s.y, t.y); s.z = f(s.z, t.z); s.w = f(s.w, t.w);
      st_stream(o + i + (uint64_t)j * kWorkerThreads, *reinterpret_cast<int4 *>(&s));
    }
  }
  for (; i < n4; i += kWorkerThreads) {
    int4 a = ld_coherent(x + i), b = ld_coherent(y + i);
    float4 s = *reinterpret_cast<float4 *>(&a);
    const float4 t = *reinterpret_cast<float4 *>(&b);
    s.x = f(s.x, t.x); s.y = f(s.y, t.y); s.z = f(s.z, t.z); s.w = f(s.w, t.w);
    st_stream(o + i, *reinterpret_cast<int4 *>(&s));
  }
  const uint64_t r = (n4 << 2) + threadIdx.x;
  if (r < c.count)
    reinterpret_cast<float *>(c.out)[r] = f(reinterpret_cast<const float *>(c.a0)[r],
                                            reinterpret_cast<const float *>(c.a1)[r]);
}

__device__ __forceinline__ bool aligned16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

__device__ uint32_t cb_none(const Ctx &) { return B2_TS_OK; }  // gate electron: only releases successors

__device__ uint32_t cb_copy(const Ctx &c) {  // count = bytes
  if (!c.a0 || !aligned16(c.a0) || !aligned16(c.out)) return B2_TS_BAD_ARG;
  const uint64_t n16 = c.count >> 4;
  const int4 *x = reinterpret_cast<const int4 *>(c.a0);
  int4 *o = reinterpret_cast<int4 *>(c.out);
  uint64_t i = threadIdx.x;
  for (; i + (uint64_t)(kVec - 1) * kWorkerThreads < n16; i += (uint64_t)kVec * kWorkerThreads) {
    int4 v[kVec];
#pragma unroll
    for (int j = 0; j < kVec; ++j) v[j] = ld_coherent(x + i + (uint64_t)j * kWorkerThreads);
#pragma unroll
    for (int j = 0; j < kVec; ++j) st_stream(o + i + (uint64_t)j * kWorkerThreads, v[j]);
  }
  for (; i < n16; i += kWorkerThreads) st_stream(o + i, ld_coherent(x + i));
  const uint64_t r = (n16 << 4) + threadIdx.x;
  if (r < c.count) c.out[r] = c.a0[r];
  return B2_TS_OK;
}

// fp32 electrons.  __fmul_rn/__fadd_rn forbid FMA contraction so that
// alpha*x + beta rounds twice, exactly as numpy does on the host (exec.py:37
// running the Python body).
__device__ uint32_t cb_scale_add(const Ctx &c) {
  if (!c.a0 || !aligned16(c.a0) || !aligned16(c.out)) return B2_TS_BAD_ARG;
  const float a = c.fn->alpha, b = c.fn->beta;
  cta_map1(c, [a, b](float x) { return __fadd_rn(__fmul_rn(a, x), b); });
  return B2_TS_OK;
}
__device__ uint32_t cb_add(const Ctx &c) {
  if (!c.a0 || !c.a1 || !aligned16(c.a0) || !aligned16(c.a1) || !aligned16(c.out))
    return B2_TS_BAD_ARG;
  cta_map2(c, [](float x, float y) { return __fadd_rn(x, y); });
  return B2_TS_OK;
}
__device__ uint32_t cb_mul(const Ctx &c) {
  if (!c.a0 || !c.a1 || !aligned16(c.a0) || !aligned16(c.a1) || !aligned16(c.out))
    return B2_TS_BAD_ARG;
  cta_map2(c, [](float x, float y) { return __fmul_rn(x, y); });
  return B2_TS_OK;
}
__device__ uint32_t cb_axpy(const Ctx &c) {
  if (!c.a0 || !c.a1 || !aligned16(c.a0) || !aligned16(c.a1) || !aligned16(c.out))
    return B2_TS_BAD_ARG;
  const float a = c.fn->alpha;
  cta_map2(c, [a](float x, float y) { return __fadd_rn(__fmul_rn(a, x), y); });
  return B2_TS_OK;
}

// out[0] = sum(x): per-thread double partials, shuffle tree, one fp32 rounding.
// Order differs from numpy's pairwise float32 sum; the result is at least as
// accurate (parity bar: 1e-6 relative, tests/test_kernels_gpu.py).
__device__ uint32_t cb_sum(const Ctx &c) {
  if (!c.a0 || !aligned16(c.a0)) return B2_TS_BAD_ARG;
  __shared__ double warp_part[kWorkerThreads / kWarp];
  const uint64_t n4 = c.count >> 2;
  const int4 *x = reinterpret_cast<const int4 *>(c.a0);
  double acc = 0.0;
  for (uint64_t i = threadIdx.x; i < n4; i += kWorkerThreads) {
    int4 v = ld_coherent(x + i);
    const float4 t = *reinterpret_cast<float4 *>(&v);
    acc += ((double)t.x + (double)t.y) + ((double)t.z + (double)t.w);
  }
  const uint64_t r = (n4 << 2) + threadIdx.x;
  if (r < c.count) acc += (double)reinterpret_cast<const float *>(c.a0)[r];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
  worker_sync();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kWorkerThreads / kWarp; ++w) s += warp_part[w];
    *reinterpret_cast<float *>(c.out) = (float)s;
  }
  return B2_TS_OK;
}

// out[i] = ((x0[i] + x1[i]) + x2[i]) + ...  over nrows rows, stride imm1 bytes.
// Sequential fp32 accumulation in row order: the order numpy uses for
// np.sum(stack, axis=0) and Python uses for sum(list_of_arrays).
__device__ uint32_t cb_stack_sum(const Ctx &c) {
  if (!c.a0 || !aligned16(c.a0) || !aligned16(c.out) || (c.imm1 & 15u) || c.nrows == 0)
    return B2_TS_BAD_ARG;
  const uint64_t stride16 = c.imm1 >> 4;
  const uint64_t n4 = c.count >> 2;
  const int4 *x = reinterpret_cast<const int4 *>(c.a0);
  int4 *o = reinterpret_cast<int4 *>(c.out);
  for (uint64_t i = threadIdx.x; i < n4; i += kWorkerThreads) {
    int4 v = ld_coherent(x + i);
    float4 acc = *reinterpret_cast<float4 *>(&v);
    uint32_t r = 1;
    for (; r + 3 < c.nrows; r += 4) {  // 4 independent loads in flight, adds stay in order
      int4 w0 = ld_coherent(x + i + (uint64_t)(r + 0) * stride16);
      int4 w1 = ld_coherent(x + i + (uint64_t)(r + 1) * stride16);
      int4 w2 = ld_coherent(x + i + (uint64_t)(r + 2) * stride16);
      int4 w3 = ld_coherent(x + i + (uint64_t)(r + 3) * stride16);
      const float4 t0 = *reinterpret_cast<float4 *>(&w0), t1 = *reinterpret_cast<float4 *>(&w1);
      const float4 t2 = *reinterpret_cast<float4 *>(&w2), t3 = *reinterpret_cast<float4 *>(&w3);
      acc.x = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc.x, t0.x), t1.x), t2.x), t3.x);
      acc.y = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc.y, t0.y), t1.y), t2.y), t3.y);
      acc.z = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc.z, t0.z), t1.z), t2.z), t3.z);
      acc.w = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc.w, t0.w), t1.w), t2.w), t3.w);
    }
    for (; r < c.nrows; ++r) {
      int4 w = ld_coherent(x + i + (uint64_t)r * stride16);
      const float4 t = *reinterpret_cast<float4 *>(&w);
      acc.x = __fadd_rn(acc.x, t.x); acc.y = __fadd_rn(acc.y, t.y);
      acc.z = __fadd_rn(acc.z, t.z); acc.w = __fadd_rn(acc.w, t.w);
    }
    st_stream(o + i, *reinterpret_cast<int4 *>(&acc));
  }
  const uint64_t e = (n4 << 2) + threadIdx.x;
  if (e < c.count) {
    const float *xs = reinterpret_cast<const float *>(c.a0);
    float acc = xs[e];
    for (uint32_t r = 1; r < c.nrows; ++r) acc = __fadd_rn(acc, xs[e + (uint64_t)r * (c.imm1 >> 2)]);
    reinterpret_cast<float *>(c.out)[e] = acc;
  }
  return B2_TS_OK;
}

// int64 electrons (Python-int add/mul of BASELINE config 2).  Python ints are
// unbounded, so an overflow is reported, never wrapped: the host then takes the
// reference's failure path (_on_ssh_fail, ssh.py:181-208).
__device__ __forceinline__ bool add_ovf(long long a, long long b, long long *r) {
  const unsigned long long u = (unsigned long long)a + (unsigned long long)b;
  *r = (long long)u;
  return (((a ^ *r) & (b ^ *r)) < 0);
}
__device__ __forceinline__ bool mul_ovf(long long a, long long b, long long *r) {
  const long long lo = (long long)((unsigned long long)a * (unsigned long long)b);
  const long long hi = __mul64hi(a, b);
  *r = lo;
  return hi != (lo >> 63);
}
__device__ uint32_t cb_iadd(const Ctx &c) {
  uint32_t st = B2_TS_OK;
  if (threadIdx.x == 0) {
    long long r;
    if (add_ovf((long long)c.imm0, (long long)c.imm1, &r)) st = B2_TS_OVERFLOW;
    *reinterpret_cast<long long *>(c.out) = r;
  }
  return st;  // only thread 0's value is consumed
}
__device__ uint32_t cb_imul(const Ctx &c) {
  uint32_t st = B2_TS_OK;
  if (threadIdx.x == 0) {
    long long r;
    if (mul_ovf((long long)c.imm0, (long long)c.imm1, &r)) st = B2_TS_OVERFLOW;
    *reinterpret_cast<long long *>(c.out) = r;
  }
  return st;
}
__device__ uint32_t cb_iscale_add(const Ctx &c) {
  uint32_t st = B2_TS_OK;
  if (threadIdx.x == 0) {
    long long p, r;
    bool o = mul_ovf(c.fn->ialpha, (long long)c.imm0, &p);
    o |= add_ovf(p, c.fn->ibeta, &r);
    if (o) st = B2_TS_OVERFLOW;
    *reinterpret_cast<long long *>(c.out) = r;
  }
  return st;
}

// The callback table proper: b2_kind -> device function pointer.
__device__ callback_t g_callbacks[B2_K__COUNT] = {
    cb_none, cb_copy, cb_scale_add, cb_add, cb_mul, cb_axpy,
    cb_sum,  cb_stack_sum, cb_iadd, cb_imul, cb_iscale_add};

// ----------------------------------------------------------------- worker ---
__device__ __forceinline__ uint8_t *resolve(const WorkerArgs &a, uint32_t arena_id, uint64_t off,
                                            uint64_t bytes, bool *ok) {
  if (arena_id >= B2_MAX_ARENAS || a.arena[arena_id] == nullptr ||
      off + bytes > a.arena_bytes[arena_id] || off + bytes < off) {
    *ok = false;
    return nullptr;
  }
  return a.arena[arena_id] + off;
}

// Named barriers (barrier 0 is __syncthreads over the whole CTA, barrier 1 is
// worker_sync over the 8 worker warps).
enum : int { kBarFull0 = 2, kBarFull1 = 3, kBarEmpty0 = 4, kBarEmpty1 = 5 };
__device__ __forceinline__ void bar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int n) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}

// What the scheduler warp hands to the worker warps through shared memory.
struct Slot {
  b2_task t;
  FnEntry fn;
  uint32_t index;   // task index, or kNoTask: nothing left for this CTA
  uint32_t kind;    // validated b2_kind, or B2_K__COUNT for "bad fn_id"
};

// One CTA = B2_WORKER_THREADS / 32 (8) worker warps + 1 scheduler warp, 4 CTAs per SM.  The scheduler runs one electron
// ahead: while the workers stream electron k it claims the next ticket, waits
// for that cell to be published if it is a released dependent, and stages the
// task record and its callback-table entry in shared memory, so the L2 round
// trips of fetching work never sit between two electrons.
__global__ void __launch_bounds__(kCtaThreads, B2_WORKER_MIN_CTAS)
    worker_kernel(const __grid_constant__ WorkerArgs a) {
  __shared__ Slot slots[2];
  __shared__ uint32_t s_base;
  __shared__ uint32_t s_status;
  const uint32_t P = a.n_ready_prefix;

  if (threadIdx.x >= kWorkerThreads) {
    // ------------------------------------------------------- scheduler warp --
    const int lane = threadIdx.x & 31;
    uint32_t k = 0;
    for (;; ++k) {
      const int s = k & 1;
      if (k >= 2) bar_sync(kBarEmpty0 + s, kCtaThreads);   // workers are done with this slot
      uint32_t ti = kNoTask;
      if (lane == 0) {
        const uint32_t ticket = queue_claim(a.q);
        // Tickets below the ready prefix ARE task indices: electrons that are
        // ready at submit never touch the queue cells.
        ti = (ticket < P) ? ticket : queue_take(a.q, ticket, a.n_tasks);
        Slot &sl = slots[s];
        sl.index = ti;
        if (ti != kNoTask) {
          sl.t = a.tasks[ti];
          uint32_t kind = B2_K__COUNT;
          if (sl.t.fn_id < B2_MAX_FN) {
            sl.fn = a.fn_table[sl.t.fn_id];
            kind = sl.fn.kind;
            if (kind >= B2_K__COUNT || (kind == B2_K_NONE && sl.t.fn_id != 0)) kind = B2_K__COUNT;
          }
          sl.kind = kind;
        }
        __threadfence_block();
      }
      ti = __shfl_sync(0xffffffffu, ti, 0);
      bar_arrive(kBarFull0 + s, kCtaThreads);
      if (ti == kNoTask) break;
    }
    // drain the arrivals the workers posted for slots that are never refilled
    if (k >= 1) bar_sync(kBarEmpty0 + ((k - 1) & 1), kCtaThreads);
  } else {
    // --------------------------------------------------------- worker warps --
    // Electrons that are ready at submit but sit behind a dependent one in the
    // table (index >= P) are published here; blocks of 256 indices are claimed
    // through a cursor so that whichever CTAs are resident do all of it.
    if (P < a.n_tasks) {
      for (;;) {
        if (threadIdx.x == 0) s_base = P + atomicAdd(a.cursor, (uint32_t)kWorkerThreads);
        worker_sync();
        const uint32_t base = s_base;
        worker_sync();
        if (base >= a.n_tasks) break;
        const uint32_t i = base + threadIdx.x;
        const bool ready = (i < a.n_tasks) && (a.tasks[i].deps == 0);
        queue_push_warp(a.q, P, ready, i);
      }
    }

    for (uint32_t k = 0;; ++k) {
      const int s = k & 1;
      bar_sync(kBarFull0 + s, kCtaThreads);
      const Slot &sl = slots[s];
      const uint32_t ti = sl.index;
      if (ti == kNoTask) break;
      const uint32_t kind = sl.kind;
      uint32_t status;
      if (sl.t.flags & B2_F_CANCELLED) {
        status = B2_TS_CANCELLED;   // tombstone (b2_cancel): successors are still released below
      } else if (kind >= B2_K__COUNT) {
        status = B2_TS_BAD_FN;
      } else {
        Ctx c;
        bool ok = true;
        const uint64_t arg0 = sl.t.arg0, arg1 = sl.t.arg1, count = sl.t.count;
        const uint32_t flags = sl.t.flags, nrows = sl.t.nrows;
        const uint64_t esz = (kind == B2_K_COPY) ? 1 : 4;  // bytes per element of array operands
        const bool is_int = (kind == B2_K_IADD || kind == B2_K_IMUL || kind == B2_K_ISCALE_ADD);
        uint64_t in_bytes = count * esz;
        if (kind == B2_K_STACK_SUM && nrows) in_bytes += (uint64_t)(nrows - 1) * arg1;
        const uint64_t out_bytes = (kind == B2_K_SUM) ? 4 : (is_int ? 8 : count * esz);
        c.imm0 = arg0;
        c.imm1 = arg1;
        c.a0 = ((flags & B2_F_IMM0) || is_int || kind == B2_K_NONE)
                   ? nullptr
                   : resolve(a, (flags >> B2_F_ARENA0_SHIFT) & 15u, arg0, in_bytes, &ok);
        const bool two = (kind == B2_K_ADD || kind == B2_K_MUL || kind == B2_K_AXPY);
        c.a1 = (two && !(flags & B2_F_IMM1))
                   ? resolve(a, (flags >> B2_F_ARENA1_SHIFT) & 15u, arg1, in_bytes, &ok)
                   : nullptr;
        c.out = (kind == B2_K_NONE)
                    ? nullptr
                    : resolve(a, (flags >> B2_F_ARENAO_SHIFT) & 15u, sl.t.out, out_bytes, &ok);
        c.count = count;
        c.nrows = nrows;
        c.fn = &sl.fn;
        status = ok ? g_callbacks[kind](c) : (uint32_t)B2_TS_BAD_ARG;
      }
      const uint32_t succ_begin = sl.t.succ_begin, succ_count = sl.t.succ_count;
      if (threadIdx.x == 0) s_status = status;
      worker_sync();  // every result store of this CTA is ordered before what follows

      if (threadIdx.x == 0) a.status[ti] = s_status;
      // Release successors, one lane per edge.  Each releasing lane fences first:
      // the barrier above ordered the whole CTA's result stores before it, the
      // fence makes them visible device-wide before the decrement another SM
      // may observe.
      for (uint32_t base = 0; base < succ_count; base += kWorkerThreads) {
        const uint32_t e = base + threadIdx.x;
        bool fire = false;
        uint32_t sidx = 0;
        if (e < succ_count) {
          sidx = a.succ[succ_begin + e];
          if (sidx < a.n_tasks) {
            __threadfence();
            fire = (atomicSub(&a.tasks[sidx].deps, 1u) == kDepBias + 1u);
          }
        }
        queue_push_warp(a.q, P, fire, sidx);
      }
      __threadfence_block();
      bar_arrive(kBarEmpty0 + s, kCtaThreads);  // the scheduler may refill this slot
    }
  }

  // Last CTA out re-arms the counters for the next launch (no memset node).
  __syncthreads();
  if (threadIdx.x == 0) {
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

// Zeroes the queue cells once at init (epoch 0 is never used by a launch).
__global__ void queue_init_kernel(unsigned long long *cells, uint32_t cap, uint32_t *counters) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap) cells[i] = 0ull;
  if (i < 64) counters[i] = 0u;
}

}  // namespace b2
