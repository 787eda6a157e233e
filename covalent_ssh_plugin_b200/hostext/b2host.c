/*
 * _b2host — CPython helper for the per-electron host path of SSHExecutor.run.
 *
 * What the reference does per electron before anything leaves the machine is
 * cloudpickle.dumps((fn, args, kwargs)) into function_<id>.pkl
 * (/root/reference/covalent_ssh_plugin/ssh.py:147-150).  Here the "pickle" of a
 * registered electron is (a) one 64-byte b2_task record and (b) one b2_copy_job per
 * ndarray operand (source address, destination offset in pinned ingress memory, bytes);
 * the bytes themselves are moved later, off the event-loop thread, by libb2exec's copy
 * pool (b2_submit).  This module writes both records for one electron in a single C call,
 * so that the Python cost of staging an electron is one method call instead of a dozen
 * numpy/struct calls.
 *
 * No CUDA here, and no computation on payload bytes: this is host book-keeping only.
 * Layouts (b2_task, b2_copy_job) come from include/b2exec.h.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

#include "../../include/b2exec.h"

#define B2H_ALIGN 256u
#define B2H_SPLIT_BYTES (2u << 20)  /* elementwise electrons this large are sliced by the Python path */
#define B2H_IDX_BITS 20

static inline uint64_t round_up(uint64_t n) { return (n + (B2H_ALIGN - 1)) & ~(uint64_t)(B2H_ALIGN - 1); }

typedef struct {
  PyObject_HEAD
  PyTypeObject *ndarray_type; /* exact type accepted as an array operand */
  b2_copy_job *jobs;
  b2_task *tasks;
  uint32_t n_jobs, cap_jobs;
  uint32_t n, cap_tasks;
  uint64_t in_base, in_used, in_cap;   /* ingress segment: offset (host off == arena off), fill, size */
  uint64_t out_base, out_used, out_cap; /* egress segment data area: ARENA offset, fill, size      */
  int open;
} Stager;

/* result codes of the stage_* methods (negative ints) */
#define B2H_FULL (-1)       /* does not fit into the open batch: ship it, open another, retry   */
#define B2H_GENERAL (-2)    /* not a case this module handles: take the general Python path      */

static int Stager_init(Stager *self, PyObject *args, PyObject *kwds) {
  PyObject *nd = NULL;
  if (!PyArg_ParseTuple(args, "O", &nd)) return -1;
  if (!PyType_Check(nd)) {
    PyErr_SetString(PyExc_TypeError, "Stager(ndarray_type)");
    return -1;
  }
  Py_INCREF(nd);
  self->ndarray_type = (PyTypeObject *)nd;
  self->open = 0;
  return 0;
}

static void Stager_dealloc(Stager *self) {
  Py_XDECREF((PyObject *)self->ndarray_type);
  Py_TYPE(self)->tp_free((PyObject *)self);
}

/* begin(jobs_addr, cap_jobs, tasks_addr, cap_tasks, in_base, in_cap, out_base, out_cap) */
static PyObject *Stager_begin(Stager *self, PyObject *const *a, Py_ssize_t n) {
  if (n != 8) {
    PyErr_SetString(PyExc_TypeError, "begin() takes 8 integers");
    return NULL;
  }
  unsigned long long v[8];
  for (int i = 0; i < 8; ++i) {
    v[i] = PyLong_AsUnsignedLongLong(a[i]);
    if (v[i] == (unsigned long long)-1 && PyErr_Occurred()) return NULL;
  }
  self->jobs = (b2_copy_job *)(uintptr_t)v[0];
  self->cap_jobs = (uint32_t)v[1];
  self->tasks = (b2_task *)(uintptr_t)v[2];
  self->cap_tasks = (uint32_t)v[3];
  self->in_base = v[4];
  self->in_cap = v[5];
  self->out_base = v[6];
  self->out_cap = v[7];
  self->n = self->n_jobs = 0;
  self->in_used = self->out_used = 0;
  self->open = 1;
  Py_RETURN_NONE;
}

static PyObject *Stager_end(Stager *self, PyObject *Py_UNUSED(ignored)) {
  self->open = 0;
  Py_RETURN_NONE;
}

/* state() -> (n, n_jobs, in_used, out_used) */
static PyObject *Stager_state(Stager *self, PyObject *Py_UNUSED(ignored)) {
  return Py_BuildValue("(IIKK)", self->n, self->n_jobs, (unsigned long long)self->in_used,
                       (unsigned long long)self->out_used);
}

/* set_state(n, n_jobs, in_used, out_used): the general Python path reserves space itself */
static PyObject *Stager_set_state(Stager *self, PyObject *const *a, Py_ssize_t n) {
  if (n != 4) {
    PyErr_SetString(PyExc_TypeError, "set_state() takes 4 integers");
    return NULL;
  }
  unsigned long long v[4];
  for (int i = 0; i < 4; ++i) {
    v[i] = PyLong_AsUnsignedLongLong(a[i]);
    if (v[i] == (unsigned long long)-1 && PyErr_Occurred()) return NULL;
  }
  if (v[0] > self->cap_tasks || v[1] > self->cap_jobs || v[2] > self->in_cap || v[3] > self->out_cap) {
    PyErr_SetString(PyExc_ValueError, "set_state() beyond the batch's capacity");
    return NULL;
  }
  self->n = (uint32_t)v[0];
  self->n_jobs = (uint32_t)v[1];
  self->in_used = v[2];
  self->out_used = v[3];
  Py_RETURN_NONE;
}

/* add_job(src_addr, dst_off, bytes): a copy the general Python path wants done by the pool */
static PyObject *Stager_add_job(Stager *self, PyObject *const *a, Py_ssize_t n) {
  if (n != 3) {
    PyErr_SetString(PyExc_TypeError, "add_job() takes 3 integers");
    return NULL;
  }
  unsigned long long v[3];
  for (int i = 0; i < 3; ++i) {
    v[i] = PyLong_AsUnsignedLongLong(a[i]);
    if (v[i] == (unsigned long long)-1 && PyErr_Occurred()) return NULL;
  }
  if (!self->open || self->n_jobs >= self->cap_jobs) return PyLong_FromLong(B2H_FULL);
  b2_copy_job *j = &self->jobs[self->n_jobs++];
  j->src = v[0];
  j->dst_off = v[1];
  j->bytes = v[2];
  return PyLong_FromLong(0);
}

/* Raw float32 operand: pointer and byte count of a C-contiguous float32 ndarray (exact type).
 * Returns 0 on success, B2H_GENERAL when the object is anything else. */
static int f32_operand(Stager *self, PyObject *x, const void **ptr, uint64_t *bytes, Py_buffer *view_out) {
  if (Py_TYPE(x) != self->ndarray_type) return B2H_GENERAL;
  Py_buffer view;
  if (PyObject_GetBuffer(x, &view, PyBUF_FORMAT | PyBUF_ND | PyBUF_C_CONTIGUOUS) != 0) {
    PyErr_Clear();
    return B2H_GENERAL;
  }
  const char *f = view.format;
  int ok = view.itemsize == 4 && view.ndim >= 1 && f != NULL;
  if (ok) {
    if (f[0] == '<' || f[0] == '=' || f[0] == '@') ++f;   /* native little endian only */
    ok = (f[0] == 'f' && f[1] == '\0');
  }
  if (!ok) {
    PyBuffer_Release(&view);
    return B2H_GENERAL;
  }
  *ptr = view.buf;
  *bytes = (uint64_t)view.len;
  if (view_out) {
    *view_out = view; /* caller releases */
  } else {
    PyBuffer_Release(&view);
  }
  return 0;
}

static inline PyObject *packed(uint64_t out_off, uint32_t idx) {
  return PyLong_FromUnsignedLongLong((out_off << B2H_IDX_BITS) | idx);
}

static inline void write_task(b2_task *t, uint32_t fn_id, uint32_t flags, uint64_t a0, uint64_t a1, uint64_t out,
                              uint64_t count) {
  t->fn_id = fn_id;
  t->flags = flags;
  t->arg0 = a0;
  t->arg1 = a1;
  t->out = out;
  t->count = count;
  t->nrows = 0;
  t->deps = 0;
  t->succ_begin = 0;
  t->succ_count = 0;
  t->tag = 0;
}

/* stage1(fn_id, kind, x) -> (out_off << 20 | index) or a negative code.
 * kind: B2_K_COPY, B2_K_SCALE_ADD (result = same bytes) or B2_K_SUM (result = 4 bytes). */
static PyObject *Stager_stage1(Stager *self, PyObject *const *a, Py_ssize_t n) {
  if (n != 3) {
    PyErr_SetString(PyExc_TypeError, "stage1(fn_id, kind, x)");
    return NULL;
  }
  const long fn_id = PyLong_AsLong(a[0]);
  const long kind = PyLong_AsLong(a[1]);
  if ((fn_id == -1 || kind == -1) && PyErr_Occurred()) return NULL;
  const void *p;
  uint64_t bytes;
  if (f32_operand(self, a[2], &p, &bytes, NULL) != 0) return PyLong_FromLong(B2H_GENERAL);
  if (kind != B2_K_COPY && kind != B2_K_SCALE_ADD && kind != B2_K_SUM) return PyLong_FromLong(B2H_GENERAL);
  uint64_t out_bytes = (kind == B2_K_SUM) ? 4 : bytes;
  if (kind != B2_K_SUM && bytes >= B2H_SPLIT_BYTES) return PyLong_FromLong(B2H_GENERAL);
  const uint64_t need_in = round_up(bytes);
  const uint64_t need_out = round_up(out_bytes < 16 ? 16 : out_bytes);
  if (!self->open || self->n >= self->cap_tasks || self->n_jobs >= self->cap_jobs ||
      self->in_used + need_in > self->in_cap || self->out_used + need_out > self->out_cap)
    return PyLong_FromLong(B2H_FULL);
  const uint64_t in_off = self->in_base + self->in_used;
  if (bytes) {
    b2_copy_job *j = &self->jobs[self->n_jobs++];
    j->src = (uint64_t)(uintptr_t)p;
    j->dst_off = in_off;
    j->bytes = bytes;
  }
  const uint32_t idx = self->n++;
  const uint64_t out_off = self->out_used;
  write_task(&self->tasks[idx], (uint32_t)fn_id, 0u, in_off, 0, self->out_base + out_off,
             kind == B2_K_COPY ? bytes : bytes / 4);
  self->in_used += need_in;
  self->out_used += need_out;
  return packed(out_off, idx);
}

/* stage2(fn_id, x, y) -> like stage1, for ADD / MUL / AXPY: two float32 operands of one shape. */
static PyObject *Stager_stage2(Stager *self, PyObject *const *a, Py_ssize_t n) {
  if (n != 3) {
    PyErr_SetString(PyExc_TypeError, "stage2(fn_id, x, y)");
    return NULL;
  }
  const long fn_id = PyLong_AsLong(a[0]);
  if (fn_id == -1 && PyErr_Occurred()) return NULL;
  const void *p0, *p1;
  uint64_t b0, b1;
  Py_buffer v0, v1;
  if (f32_operand(self, a[1], &p0, &b0, &v0) != 0) return PyLong_FromLong(B2H_GENERAL);
  if (f32_operand(self, a[2], &p1, &b1, &v1) != 0) {
    PyBuffer_Release(&v0);
    return PyLong_FromLong(B2H_GENERAL);
  }
  int same = (v0.ndim == v1.ndim) && (b0 == b1);
  for (int i = 0; same && i < v0.ndim; ++i) same = (v0.shape[i] == v1.shape[i]);
  PyBuffer_Release(&v0);
  PyBuffer_Release(&v1);
  if (!same || b0 >= B2H_SPLIT_BYTES) return PyLong_FromLong(B2H_GENERAL);
  const uint64_t need_in = 2 * round_up(b0);
  const uint64_t need_out = round_up(b0 < 16 ? 16 : b0);
  if (!self->open || self->n >= self->cap_tasks || self->n_jobs + 2 > self->cap_jobs ||
      self->in_used + need_in > self->in_cap || self->out_used + need_out > self->out_cap)
    return PyLong_FromLong(B2H_FULL);
  const uint64_t in0 = self->in_base + self->in_used;
  const uint64_t in1 = in0 + round_up(b0);
  if (b0) {
    b2_copy_job *j = &self->jobs[self->n_jobs];
    j[0].src = (uint64_t)(uintptr_t)p0;
    j[0].dst_off = in0;
    j[0].bytes = b0;
    j[1].src = (uint64_t)(uintptr_t)p1;
    j[1].dst_off = in1;
    j[1].bytes = b0;
    self->n_jobs += 2;
  }
  const uint32_t idx = self->n++;
  const uint64_t out_off = self->out_used;
  write_task(&self->tasks[idx], (uint32_t)fn_id, 0u, in0, in1, self->out_base + out_off, b0 / 4);
  self->in_used += need_in;
  self->out_used += need_out;
  return packed(out_off, idx);
}

/* stage_int(fn_id, a, b) -> like stage1, for the int64 electrons (operands travel as immediates). */
static PyObject *Stager_stage_int(Stager *self, PyObject *const *a, Py_ssize_t n) {
  if (n != 3) {
    PyErr_SetString(PyExc_TypeError, "stage_int(fn_id, a, b)");
    return NULL;
  }
  const long fn_id = PyLong_AsLong(a[0]);
  if (fn_id == -1 && PyErr_Occurred()) return NULL;
  if (!PyLong_CheckExact(a[1]) || !PyLong_CheckExact(a[2])) return PyLong_FromLong(B2H_GENERAL);
  int o0 = 0, o1 = 0;
  const long long x = PyLong_AsLongLongAndOverflow(a[1], &o0);
  const long long y = PyLong_AsLongLongAndOverflow(a[2], &o1);
  if (o0 || o1) return PyLong_FromLong(B2H_GENERAL); /* beyond int64: the general path reports it */
  if (!self->open || self->n >= self->cap_tasks || self->out_used + B2H_ALIGN > self->out_cap)
    return PyLong_FromLong(B2H_FULL);
  const uint32_t idx = self->n++;
  const uint64_t out_off = self->out_used;
  write_task(&self->tasks[idx], (uint32_t)fn_id, B2_F_IMM0 | B2_F_IMM1, (uint64_t)x, (uint64_t)y,
             self->out_base + out_off, 1);
  self->out_used += B2H_ALIGN;
  return packed(out_off, idx);
}

static PyMethodDef Stager_methods[] = {
    {"begin", (PyCFunction)(void (*)(void))Stager_begin, METH_FASTCALL, "start a batch"},
    {"end", (PyCFunction)Stager_end, METH_NOARGS, "close the batch"},
    {"state", (PyCFunction)Stager_state, METH_NOARGS, "(n, n_jobs, in_used, out_used)"},
    {"set_state", (PyCFunction)(void (*)(void))Stager_set_state, METH_FASTCALL, "set the counters"},
    {"add_job", (PyCFunction)(void (*)(void))Stager_add_job, METH_FASTCALL, "append one copy job"},
    {"stage1", (PyCFunction)(void (*)(void))Stager_stage1, METH_FASTCALL, "stage a one-array electron"},
    {"stage2", (PyCFunction)(void (*)(void))Stager_stage2, METH_FASTCALL, "stage a two-array electron"},
    {"stage_int", (PyCFunction)(void (*)(void))Stager_stage_int, METH_FASTCALL, "stage an int64 electron"},
    {NULL, NULL, 0, NULL}};

static PyTypeObject StagerType = {
    PyVarObject_HEAD_INIT(NULL, 0).tp_name = "_b2host.Stager",
    .tp_basicsize = sizeof(Stager),
    .tp_flags = Py_TPFLAGS_DEFAULT,
    .tp_doc = "Writes b2_task / b2_copy_job records for one batch of electrons.",
    .tp_methods = Stager_methods,
    .tp_init = (initproc)Stager_init,
    .tp_dealloc = (destructor)Stager_dealloc,
    .tp_new = PyType_GenericNew,
};

static struct PyModuleDef b2host_module = {PyModuleDef_HEAD_INIT, "_b2host",
                                           "host-side staging helper of covalent_ssh_plugin_b200", -1, NULL};

PyMODINIT_FUNC PyInit__b2host(void) {
  if (PyType_Ready(&StagerType) < 0) return NULL;
  PyObject *m = PyModule_Create(&b2host_module);
  if (!m) return NULL;
  Py_INCREF(&StagerType);
  if (PyModule_AddObject(m, "Stager", (PyObject *)&StagerType) < 0) {
    Py_DECREF(&StagerType);
    Py_DECREF(m);
    return NULL;
  }
  PyModule_AddIntConstant(m, "FULL", B2H_FULL);
  PyModule_AddIntConstant(m, "GENERAL", B2H_GENERAL);
  PyModule_AddIntConstant(m, "IDX_BITS", B2H_IDX_BITS);
  PyModule_AddIntConstant(m, "TASK_BYTES", (long)sizeof(b2_task));
  PyModule_AddIntConstant(m, "JOB_BYTES", (long)sizeof(b2_copy_job));
  return m;
}
