/*
 * b2exec.h — C ABI of the B200 electron executor ("libb2exec.so").
 *
 * This is the drop-in boundary underneath the Python host
 * (covalent_ssh_plugin_b200/ssh.py :: SSHExecutor).  The reference plugin has
 * no FFI of its own: every step below is an asyncssh call in
 * /root/reference/covalent_ssh_plugin/ssh.py.  Each entry point names the
 * reference step it replaces (file:line).
 *
 *   reference step                                   replaced by
 *   ------------------------------------------------ ---------------------------
 *   asyncssh.connect + retry     ssh.py:237-282      b2_init / b2_peer_enable / b2_ipc_*
 *   conn.close/wait_closed       ssh.py:585-587      b2_shutdown
 *   pickle.dump -> function.pkl  ssh.py:147-150      host staging buffer (b2_host_buffer) + b2_task records
 *   asyncssh.scp upload x2       ssh.py:360-361      b2_h2d  (+ b2_unpack into arg slots)
 *   conn.run("python exec.py")   ssh.py:377-383      b2_enqueue + b2_launch  (worker kernel + callback table)
 *   exec.py fn(*args, **kwargs)  exec.py:29-46       b2_register (callback table entry)
 *   conn.run("ls result.pkl")    ssh.py:401-406      b2_poll / b2_wait
 *   asyncssh.scp download        ssh.py:451          b2_pack of result slots + b2_d2h
 *   rm x3                        ssh.py:313-315      arena offsets are simply reused by the host
 *   (one dispatcher per host)    README.md:52-65     b2_peer_copy / b2_peer_pack over NVLink
 *   the whole of run(), per batch ssh.py:466-591      b2_submit  (copy pool -> H2D -> worker -> D2H -> ticket, one call)
 *   cancel() (unsupported)       ssh.py:460-464      b2_cancel / B2_F_CANCELLED tombstones
 *   re-upload per electron       ssh.py:360-361      b2_table_create / b2_table_launch (resident task tables)
 *   a fresh interpreter per run  ssh.py:377-383      b2_resident_enable (the worker kernel outlives its batches)
 *
 * Conventions
 *   - every function returns 0 (B2_OK) on success or a negative b2_err code;
 *     b2_last_error() returns a thread-local, NUL-terminated description.
 *   - plain C types only: pointers, sizes, fixed-width integers.  Python calls
 *     these through ctypes (which releases the GIL for the duration).
 *   - "dev" is a CUDA device ordinal.  All device memory is one arena per
 *     device; every device address in this ABI is a BYTE OFFSET into that arena.
 *   - the caller owns host buffers; the library owns the device arena, the
 *     pinned staging buffers, streams, events and the device-side queue.
 *   - work is stream-ordered per device: calls enqueue asynchronously in call
 *     order; b2_sync / b2_wait / b2_poll are the only blocking/observing calls.
 *     Device->host copies (b2_d2h, the status words of b2_launch) run on their
 *     own stream so that the download of one batch overlaps the upload of the
 *     next: such a copy starts after everything issued before it, but calls
 *     issued AFTER it do not wait for it.  A ticket (b2_launch / b2_mark) taken
 *     after the copy completes only when the copy has — do not overwrite the
 *     copy's source before such a ticket has completed.
 */
#ifndef B2EXEC_H
#define B2EXEC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2_ABI_VERSION 2u

/* ---- error codes -------------------------------------------------------- */
enum b2_err {
  B2_OK = 0,
  B2_E_CUDA = -1,      /* a CUDA runtime call failed (text in b2_last_error) */
  B2_E_ARG = -2,       /* bad argument (range, alignment, null)              */
  B2_E_STATE = -3,     /* device not initialised / already initialised       */
  B2_E_NOMEM = -4,     /* arena / staging / queue capacity exceeded          */
  B2_E_NOTREADY = -5,  /* b2_poll: ticket not complete yet (not an error)    */
  B2_E_UNSUPPORTED = -6
};

/* ---- per-electron status written by the worker kernel -------------------- */
enum b2_task_status {
  B2_TS_PENDING = 0,
  B2_TS_OK = 1,
  B2_TS_OVERFLOW = 2,  /* int64 electron overflowed: host must take the _on_ssh_fail path */
  B2_TS_BAD_FN = 3,    /* fn_id has no callback registered                               */
  B2_TS_BAD_ARG = 4,   /* operand out of arena range / misaligned for the callback        */
  B2_TS_CANCELLED = 5  /* tombstoned before the worker reached it (B2_F_CANCELLED / b2_cancel) */
};

/* ---- built-in callbacks (the on-device stand-in for exec.py:37) ---------- */
enum b2_kind {
  B2_K_NONE = 0,
  B2_K_COPY = 1,       /* out[i]  = x[i]                     (bytes, count = bytes)   */
  B2_K_SCALE_ADD = 2,  /* out[i]  = alpha * x[i] + beta      fp32, two roundings      */
  B2_K_ADD = 3,        /* out[i]  = x[i] + y[i]              fp32                     */
  B2_K_MUL = 4,        /* out[i]  = x[i] * y[i]              fp32                     */
  B2_K_AXPY = 5,       /* out[i]  = alpha * x[i] + y[i]      fp32, two roundings      */
  B2_K_SUM = 6,        /* out[0]  = sum_i x[i]               fp32 (tree order)        */
  B2_K_STACK_SUM = 7,  /* out[i]  = ((x0[i]+x1[i])+x2[i])..  fp32, K rows, row order  */
  B2_K_IADD = 8,       /* out[0]  = arg0 + arg1              int64 immediates, checked */
  B2_K_IMUL = 9,       /* out[0]  = arg0 * arg1              int64 immediates, checked */
  B2_K_ISCALE_ADD = 10,/* out[0]  = ialpha*arg0 + ibeta      int64, checked           */
  B2_K__COUNT = 11
};

/* b2_task.flags */
#define B2_F_IMM0 0x1u          /* arg0 is an immediate, not an arena offset   */
#define B2_F_IMM1 0x2u          /* arg1 is an immediate                        */
#define B2_F_CANCELLED 0x4u     /* tombstone: do not run, report B2_TS_CANCELLED (successors are still released) */
#define B2_F_ARENA0_SHIFT 8     /* bits 8..11 : arena id of arg0 (0 = local)   */
#define B2_F_ARENA1_SHIFT 12    /* bits 12..15: arena id of arg1               */
#define B2_F_ARENAO_SHIFT 16    /* bits 16..19: arena id of out                */
#define B2_NO_SUCC 0xFFFFFFFFu
#define B2_MAX_ARENAS 16        /* local + 15 mapped peers (4-bit arena id): same-process and IPC mappings of the same GPU count separately */
#define B2_MAX_FN 1024          /* callback-table capacity                     */

/*
 * One electron, as the worker kernel sees it (64 bytes, little endian).
 * This record is the binary replacement of the pickled (fn, args, kwargs)
 * tuple (ssh.py:150) for registered electrons: fn -> fn_id, ndarray args ->
 * arena offsets of their raw buffers, Python ints -> immediates.
 */
typedef struct b2_task {
  uint32_t fn_id;       /* callback-table index (b2_register)                    */
  uint32_t flags;       /* B2_F_*                                                */
  uint64_t arg0;        /* arena byte offset (16 B aligned) or immediate         */
  uint64_t arg1;        /* second operand; for B2_K_STACK_SUM: row stride, bytes */
  uint64_t out;         /* arena byte offset of the result (16 B aligned)        */
  uint64_t count;       /* elements per operand (bytes for B2_K_COPY)            */
  uint32_t nrows;       /* B2_K_STACK_SUM: number of stacked rows, else 0        */
  uint32_t deps;        /* predecessors inside this batch still to finish        */
  uint32_t succ_begin;  /* first index into the batch's successor list           */
  uint32_t succ_count;  /* number of successors to release on completion         */
  uint64_t tag;         /* opaque to the device; echoed nowhere (host bookkeeping) */
} b2_task;

/* One payload move for pack / unpack / peer routing (24 bytes). */
typedef struct b2_span {
  uint64_t src_off;     /* arena byte offset on the source device      */
  uint64_t dst_off;     /* arena byte offset on the destination device */
  uint64_t bytes;
} b2_span;

/* One host->pinned-ingress copy of a batch (b2_stage_copy / b2_submit): the
 * replacement of writing function_<id>.pkl and SCP-ing it (ssh.py:147-150, 360). */
typedef struct b2_copy_job {
  uint64_t src;         /* host virtual address of the operand bytes (caller keeps it alive) */
  uint64_t dst_off;     /* byte offset inside the pinned ingress buffer                      */
  uint64_t bytes;
} b2_copy_job;

/* b2_batch.flags */
#define B2_BATCH_INDEPENDENT 0x1u  /* no task has deps or successors (plain fan-out): skip the table scan */

/* One whole batch of electrons for b2_submit: everything ssh.py:466-591 does for
 * ONE electron, for n_tasks of them, in one asynchronous call. */
typedef struct b2_batch {
  const b2_copy_job *jobs;   /* operand copies into the pinned ingress buffer (may be NULL)     */
  const b2_task *tasks;      /* task records; offsets refer to the arena                        */
  uint32_t n_jobs;
  uint32_t n_tasks;
  uint32_t flags;            /* B2_BATCH_*                                                      */
  uint32_t reserved;
  uint64_t in_off;           /* ingress range to upload after the copies: host off == arena off */
  uint64_t in_bytes;
  uint64_t status_host_off;  /* egress offset receiving n_tasks status words                    */
  uint64_t out_dev_off;      /* arena range holding the results ...                             */
  uint64_t out_host_off;     /* ... downloaded to this egress offset                            */
  uint64_t out_bytes;
  uint64_t ticket;           /* OUT: completion ticket for b2_poll / b2_wait                    */
} b2_batch;

typedef struct b2_device_info {
  int32_t  cc_major, cc_minor;
  int32_t  sm_count;
  int32_t  l2_bytes_mb;
  uint64_t total_mem;
  uint64_t arena_bytes;
  uint64_t staging_in_bytes;
  uint64_t staging_out_bytes;
  uint32_t queue_cap;
  uint32_t abi_version;
  char     name[64];
} b2_device_info;

typedef struct b2_stats {
  uint64_t launches_pack;     /* span-move launches (pack, unpack, peer_pack) */
  uint64_t launches_worker;
  uint64_t launches_other;    /* queue reset etc.                              */
  uint64_t electrons;         /* tasks retired by the worker kernel           */
  uint64_t bytes_packed;      /* sum of span bytes moved                      */
  uint64_t bytes_h2d;
  uint64_t bytes_d2h;
  uint64_t resident_instances;  /* resident-worker kernel instances launched (b2_resident_enable) */
  uint64_t resident_batches;    /* batches published to a resident worker through the doorbell     */
} b2_stats;

/* ---- lifecycle ----------------------------------------------------------- */
const char *b2_last_error(void);
int b2_abi_version(void);
/* First 16 hex digits of the SHA-256 over every file of csrc/, this header and the compiler
 * flags the library was built from (covalent_ssh_plugin_b200/build.py); "unknown" for a build
 * that did not pass -DB2_SRC_HASH. */
const char *b2_source_hash(void);
int b2_device_count(int *count);

/* Opens the "connection" to one GPU (replaces asyncssh.connect, ssh.py:263-268):
 * creates streams, allocates the device arena, two pinned host staging buffers
 * (ingress = host->device, egress = device->host), the task table and the MPMC
 * queue (queue_cap cells, power of two), and installs the built-in callbacks. */
int b2_init(int dev, uint64_t arena_bytes, uint64_t staging_in_bytes,
            uint64_t staging_out_bytes, uint32_t queue_cap);
int b2_shutdown(int dev);                       /* ssh.py:585-587 */
int b2_is_initialized(int dev);
int b2_get_info(int dev, b2_device_info *out);
int b2_get_stats(int dev, b2_stats *out);

/* ---- memory -------------------------------------------------------------- */
/* which: 0 = ingress (pinned, host writes / device reads), 1 = egress. */
int b2_host_buffer(int dev, int which, void **ptr, uint64_t *bytes);
/* Raw device pointer and size of an arena (0 = local, else a mapped peer) so that
 * plumbing outside this library — an NCCL broadcast of a shared operand — can
 * address arena memory.  The memory stays owned by the library. */
int b2_arena_base(int dev, uint32_t arena_id, void **ptr, uint64_t *bytes);
/* Ownership of one pinned staging buffer passes to the caller: the library
 * forgets it and b2_shutdown will not free it.  Used when result arrays that
 * alias the egress buffer (zero-copy results) outlive the engine. */
int b2_detach_host_buffer(int dev, int which);
/* async copies on the device's compute stream; host_off is relative to the
 * pinned staging buffer named by the direction (ssh.py:360-361 / ssh.py:451). */
int b2_h2d(int dev, uint64_t host_off, uint64_t dev_off, uint64_t bytes);
int b2_d2h(int dev, uint64_t dev_off, uint64_t host_off, uint64_t bytes);
/* synchronous debug/test helpers for arbitrary (pageable) host memory */
int b2_write(int dev, uint64_t dev_off, const void *src, uint64_t bytes);
int b2_read(int dev, uint64_t dev_off, void *dst, uint64_t bytes);
int b2_fill(int dev, uint64_t dev_off, int byte_value, uint64_t bytes);

/* ---- pack / unpack (K1 / K2) --------------------------------------------- */
/* Generic span move inside one arena: for every i, arena[dst_off .. +bytes)
 * = arena[src_off .. +bytes).  Spans must not overlap each other's
 * destinations.  Fast path needs 16 B aligned offsets; any alignment works. */
int b2_move_spans(int dev, const b2_span *spans, uint32_t n);
/* pack: gather n scattered payloads into one contiguous run starting at
 * dst_off; payload i lands at dst_off + prefix(round_up(bytes, align)).
 * unpack: the inverse (contiguous source, scattered destinations).
 * If out_offs != NULL it receives the n contiguous-side offsets. */
int b2_pack(int dev, const uint64_t *src_offs, const uint64_t *sizes, uint32_t n,
            uint64_t dst_off, uint32_t align, uint64_t *out_offs, uint64_t *out_total);
int b2_unpack(int dev, uint64_t src_off, const uint64_t *dst_offs, const uint64_t *sizes,
              uint32_t n, uint32_t align, uint64_t *out_offs, uint64_t *out_total);

/* ---- callback table + queue (K3) ----------------------------------------- */
/* Binds fn_id to a built-in callback and its registered constants; the
 * on-device replacement of "unpickle fn" (exec.py:29-30). */
int b2_register(int dev, uint32_t fn_id, int kind, double alpha, double beta,
                int64_t ialpha, int64_t ibeta);
/* Appends n electrons (and their successor list) to the device's pending
 * batch.  succ may be NULL when no task has successors.  Returns the index of
 * the first appended task in *first_index (successor indices are batch-relative). */
int b2_enqueue(int dev, const b2_task *tasks, uint32_t n, const uint32_t *succ,
               uint32_t n_succ, uint32_t *first_index);
/* Uploads the pending batch's task table, pushes ready electrons into the
 * device MPMC queue and runs the persistent worker kernel until every electron
 * of the batch has retired.  status_host_off: egress offset that receives one
 * uint32 b2_task_status per electron (pass UINT64_MAX to skip the copy).
 * Returns a ticket for b2_poll / b2_wait. */
int b2_launch(int dev, uint64_t status_host_off, uint64_t *ticket);
/* The whole per-batch pipeline in ONE asynchronous call (replaces, per electron,
 * ssh.py:147-150 pickle + 360-361 upload + 377-383 run + 451 download): a submission
 * thread inside the library copies the operands into pinned ingress memory (b2_copy_job,
 * several host threads, non-temporal stores), uploads [in_off, +in_bytes), runs the
 * worker over the task records, downloads statuses and results, and records the ticket.
 * Returns as soon as the batch is queued; batch->ticket is valid on return.  jobs, tasks
 * and the operand memory must stay valid until the ticket has completed.  Batches run in
 * submission order; every other stream-ordered call of this ABI first waits until queued
 * submissions have been issued, so program order is kept.  Do not interleave with an open
 * b2_enqueue .. b2_launch pair.  A failure inside the pipeline is reported by
 * b2_poll / b2_wait on the ticket. */
int b2_submit(int dev, b2_batch *batch);
/* The host-copy stage alone, synchronously (same thread pool). */
int b2_stage_copy(int dev, const b2_copy_job *jobs, uint32_t n);
/* Tombstones electrons [first, first+count) of the batch behind `ticket` (b2_submit or
 * b2_launch) that the worker has not reached yet: they report B2_TS_CANCELLED instead of
 * running.  Best effort once the batch is on the device; exact before.  (The reference
 * cannot cancel at all: ssh.py:460-464.) */
int b2_cancel(int dev, uint64_t ticket, uint32_t first, uint32_t count);
/* Resident task tables: a lattice that is dispatched repeatedly (same graph, same arena offsets)
 * uploads its table ONCE; every b2_table_launch then re-arms the dependency counters on the
 * device and launches the worker — no table rebuild, no pinned copy, no upload per launch (the
 * steady-state analogue of keeping function_<id>.pkl on the remote host instead of re-sending it,
 * which the reference cannot do: ssh.py:360-361 uploads per electron).  Same record and successor
 * conventions as b2_enqueue.  status_host_off as in b2_launch. */
int b2_table_create(int dev, const b2_task *tasks, uint32_t n, const uint32_t *succ, uint32_t n_succ,
                    uint32_t *table_id);
int b2_table_launch(int dev, uint32_t table_id, uint64_t status_host_off, uint64_t *ticket);
int b2_table_destroy(int dev, uint32_t table_id);
/* Resident worker: with this on, the worker kernel is launched ONCE and stays on the GPU across
 * batches (north_star: "a persistent per-GPU worker kernel that pulls task descriptors from a
 * device-side MPMC queue").  b2_launch / b2_submit / b2_table_launch then publish their batch
 * through a descriptor ring in device memory — table, payload and ring entry travel on the upload
 * stream, followed by a stream-ordered write of the doorbell word; completion is a word the other
 * streams wait on — instead of launching a kernel per batch, so consecutive batches run with no
 * launch gap and overlap tail-to-head.  The instance leaves by itself after idle_us microseconds
 * without a published batch (0 = keep the current setting, default 200; a Dekker handshake over
 * two words of mapped pinned memory makes "leave" and "a batch is being submitted" mutually
 * visible), and a later batch simply starts a new instance: the SMs are never held while there is
 * nothing to do, so other kernels of the process (the span mover, NCCL) cannot deadlock behind it.
 * Also switched on by B2_RESIDENT=1 at b2_init. */
int b2_resident_enable(int dev, uint32_t idle_us);
int b2_resident_disable(int dev);
/* Records a completion ticket after whatever has been queued so far. */
int b2_mark(int dev, uint64_t *ticket);
int b2_poll(int dev, uint64_t ticket);          /* B2_OK | B2_E_NOTREADY  (ssh.py:388-406) */
int b2_wait(int dev, uint64_t ticket);          /* blocks (ssh.py:408-432) */
int b2_sync(int dev);

/* ---- inter-GPU routing (K4) ---------------------------------------------- */
/* same-process peers */
int b2_peer_enable(int dev, int peer_dev, uint32_t *arena_id);
/* cross-process peers (one process per GPU under torchrun): export this
 * device's arena as a 64-byte CUDA IPC handle, map a peer's handle. */
int b2_ipc_export(int dev, void *handle64);
int b2_ipc_open(int dev, const void *handle64, uint64_t peer_arena_bytes, uint32_t *arena_id);
/* copy engine path (cudaMemcpyPeerAsync) and kernel path (remote stores). */
int b2_peer_copy(int dev, uint32_t dst_arena_id, uint64_t src_off, uint64_t dst_off, uint64_t bytes);
int b2_peer_pack(int dev, uint32_t dst_arena_id, const b2_span *spans, uint32_t n);

/* ---- measurement --------------------------------------------------------- */
/* CUDA events on the device's compute stream (the stream every kernel of this
 * library is launched on). */
int b2_event_create(int dev, uint32_t *event_id);
int b2_event_record(int dev, uint32_t event_id);
int b2_event_elapsed_ms(int dev, uint32_t start_id, uint32_t stop_id, float *ms);
/* writes arena[off .. off+bytes) with a rolling pattern; used to evict L2. */
int b2_flush_l2(int dev, uint64_t scratch_off, uint64_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* B2EXEC_H */
