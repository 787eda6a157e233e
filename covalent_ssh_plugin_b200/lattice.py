"""A small lattice driver for when the Covalent SDK is not installed — and the
place where whole dependency graphs are compiled for the device.

``electron`` / ``lattice`` / ``dispatch`` / ``get_result`` mirror the calls the
reference's functional tests make (/root/reference/tests/functional_tests/
basic_workflow_test.py:9-29, svm_workflow.py:11-46) so those workflows run
unchanged.  Covalent core is otherwise out of scope: there is no server, no
database, no UI — ``dispatch`` executes the lattice in this process.

What is new compared to calling ``executor.run`` once per electron: connected
groups of *registered* electrons that target GPU executors are compiled into a
single task table with dependency counts and successor lists and run by ONE
worker launch per GPU (``DeviceEngine.run_dag``).  Intermediate results stay in
the arena heap (no D2H -> H2D per edge, which is what the reference does with
a pickle round trip per edge, ssh.py:147-150 / 451-456); the rows of a fan-in
reduction are laid out back to back so the reducer reads them in place; wide
reductions are split by column behind a gate electron so every SM takes part.
Across GPUs of one process the graph runs level by level and consumers read
their operands straight out of the producer GPU's arena over NVLink.
"""

from __future__ import annotations

import asyncio
import uuid
from typing import Any, Callable, Dict, List, Optional, Sequence

import numpy as np

from . import _b2lib as L
from .electrons import K_STACK_SUM, NotDeviceEligible, decode, lookup, plan_call
from .engine import DeviceArray, DeviceEngine, get_engine

__all__ = ["electron", "lattice", "dispatch", "get_result", "status", "Result", "DepsPip", "executor"]


class status:                       # ct.status look-alike
    COMPLETED = "COMPLETED"
    FAILED = "FAILED"


class _ExecutorNamespace:
    """``ct.executor.SSHExecutor`` (README.md:55-60 of the reference); resolved on first use because
    ``ssh.py`` imports this package."""

    def __getattr__(self, name):
        if name == "SSHExecutor":
            from .ssh import SSHExecutor

            return SSHExecutor
        raise AttributeError(name)


executor = _ExecutorNamespace()


class DepsPip:
    """``ct.DepsPip(packages=[...])`` as the reference's svm workflow builds it (svm_workflow.py:7,20).
    Recorded on the electron, never applied: installing packages is Covalent core's job on a remote
    host and there is no remote host here (the electron runs in this box's interpreter)."""

    def __init__(self, packages=None, reqs_path: str = ""):
        self.packages = list(packages or [])
        self.reqs_path = reqs_path


class Result:
    def __init__(self, dispatch_id: str):
        self.dispatch_id = dispatch_id
        self.status = None
        self.result = None
        self.error: Optional[BaseException] = None
        self.stats: Dict[str, Any] = {}

    def __repr__(self):
        return f"<Result {self.dispatch_id} {self.status}>"


class NodeRef:
    """Stands for the output of electron ``node_id`` while a lattice is being traced."""

    __slots__ = ("graph", "node_id")

    def __init__(self, graph, node_id):
        self.graph, self.node_id = graph, node_id


class _Node:
    __slots__ = ("id", "fn", "args", "kwargs", "executor", "deps")

    def __init__(self, nid, fn, args, kwargs, executor):
        self.id, self.fn, self.args, self.kwargs, self.executor = nid, fn, args, kwargs, executor
        self.deps: List[int] = []


def _refs(value) -> List[NodeRef]:
    if isinstance(value, NodeRef):
        return [value]
    if isinstance(value, (list, tuple)):
        return [r for v in value for r in _refs(v)]
    if isinstance(value, dict):
        return [r for v in value.values() for r in _refs(v)]
    return []


class _Graph:
    def __init__(self):
        self.nodes: List[_Node] = []

    def add(self, el: "Electron", args, kwargs) -> NodeRef:
        node = _Node(len(self.nodes), el.fn, list(args), dict(kwargs), el.executor)
        node.deps = sorted({r.node_id for r in _refs(node.args) + _refs(node.kwargs)})
        self.nodes.append(node)
        return NodeRef(self, node.id)


_tracing: List[_Graph] = []


class Electron:
    def __init__(self, fn: Callable, executor=None):
        self.fn, self.executor = fn, executor
        self.__name__ = getattr(fn, "__name__", "electron")
        self.__wrapped__ = fn

    def __call__(self, *args, **kwargs):
        if not _tracing:                       # outside a lattice an electron is a plain call
            return self.fn(*args, **kwargs)
        return _tracing[-1].add(self, args, kwargs)


def electron(_fn: Optional[Callable] = None, *, executor=None, **_unused):
    """``@electron`` / ``@electron(executor=...)`` (README.md:43-65 of the reference)."""
    if _fn is not None:
        return Electron(_fn, executor)
    return lambda fn: Electron(fn, executor)


class Lattice:
    def __init__(self, fn: Callable):
        self.fn = fn
        self.__name__ = getattr(fn, "__name__", "lattice")

    def __call__(self, *args, **kwargs):
        return self.fn(*args, **kwargs)


def lattice(_fn: Optional[Callable] = None, **_unused):
    if _fn is not None:
        return Lattice(_fn)
    return lambda fn: Lattice(fn)


_results: Dict[str, Result] = {}


def dispatch(lat: Lattice) -> Callable[..., str]:
    """``dispatch(workflow)(*args)`` -> dispatch id; the lattice has run when it returns."""

    def submit(*args, **kwargs) -> str:
        dispatch_id = uuid.uuid4().hex
        res = _results[dispatch_id] = Result(dispatch_id)
        graph = _Graph()
        _tracing.append(graph)
        try:
            out = lat.fn(*args, **kwargs)
        except BaseException as err:
            res.status, res.error = status.FAILED, err
            return dispatch_id
        finally:
            _tracing.pop()
        values: Dict[int, Any] = {}
        try:
            _run_sync(_execute(graph, dispatch_id, out, res, values))
            res.result = _materialise(_substitute(out, values))
            res.status = status.COMPLETED
        except BaseException as err:      # a failed electron fails the lattice (basic_workflow_test.py:32-49)
            res.status, res.error = status.FAILED, err
        finally:
            for v in values.values():     # device-resident results die with the dispatch
                if isinstance(v, DeviceArray):
                    v.free()
        return dispatch_id

    return submit


def get_result(dispatch_id: str, wait: bool = True) -> Result:
    return _results[dispatch_id]


def _run_sync(coro):
    try:
        asyncio.get_running_loop()
    except RuntimeError:
        return asyncio.run(coro)
    raise RuntimeError("dispatch() must be called from synchronous code")


# ---------------------------------------------------------------- execution ---
def _substitute(value, values: Dict[int, Any]):
    if isinstance(value, NodeRef):
        return values[value.node_id]
    if isinstance(value, list):
        return [_substitute(v, values) for v in value]
    if isinstance(value, tuple):
        return tuple(_substitute(v, values) for v in value)
    if isinstance(value, dict):
        return {k: _substitute(v, values) for k, v in value.items()}
    return value


def _materialise(value):
    """Device-resident results become host arrays where the host needs them: all of them
    in one pack launch + one D2H per staging segment (K1), then substituted in place."""
    found: List[DeviceArray] = []

    def collect(v):
        if isinstance(v, DeviceArray):
            found.append(v)
        elif isinstance(v, (list, tuple)):
            for x in v:
                collect(x)
        elif isinstance(v, dict):
            for x in v.values():
                collect(x)

    collect(value)
    if len(found) <= 1:
        return _materialise_one(value, {})
    host: Dict[int, np.ndarray] = {}
    by_engine: Dict[int, List[DeviceArray]] = {}
    for da in found:
        by_engine.setdefault(id(da.engine), []).append(da)
    for das in by_engine.values():
        for da, arr in zip(das, das[0].engine.get_many(das)):
            host[id(da)] = arr
    return _materialise_one(value, host)


def _materialise_one(value, host):
    if isinstance(value, DeviceArray):
        return host[id(value)] if id(value) in host else value.numpy()
    if isinstance(value, list):
        return [_materialise_one(v, host) for v in value]
    if isinstance(value, tuple):
        return tuple(_materialise_one(v, host) for v in value)
    if isinstance(value, dict):
        return {k: _materialise_one(v, host) for k, v in value.items()}
    return value


def _resolve_executor(ex):
    """Executor object for an electron's ``executor=`` argument (string aliases
    resolve the way Covalent's executor manager does: ``"ssh"`` -> config)."""
    if ex is None or ex in ("local", "dask"):
        return None
    if isinstance(ex, str):
        if ex == "ssh":
            from .ssh import SSHExecutor, get_config

            return SSHExecutor(username=get_config("executors.ssh.username"),
                               hostname=get_config("executors.ssh.hostname"))
        raise ValueError(f"unknown executor alias {ex!r}")
    return ex


def _device_of(ex) -> Optional[int]:
    from .ssh import SSHExecutor, resolve_device

    if isinstance(ex, SSHExecutor):
        try:
            return resolve_device(ex.hostname)
        except OSError:
            return None
    return None


async def _execute(graph: _Graph, dispatch_id: str, outputs, res: Result, values: Dict[int, Any]) -> None:
    nodes = graph.nodes
    execs = {n.id: _resolve_executor(n.executor) for n in nodes}
    device = {n.id: (_device_of(execs[n.id]) if lookup(n.fn) is not None else None) for n in nodes}
    consumers: Dict[int, List[int]] = {n.id: [] for n in nodes}
    for n in nodes:
        for d in n.deps:
            consumers[d].append(n.id)
    needed_on_host = {r.node_id for r in _refs(outputs)}
    res.stats.update(device_batches=0, device_electrons=0, host_electrons=0)

    remaining = {n.id for n in nodes}
    while remaining:
        progressed = False
        # 1. host electrons whose inputs are ready
        for nid in sorted(remaining):
            n = nodes[nid]
            if device[nid] is not None or any(d in remaining for d in n.deps):
                continue
            args = _materialise(_substitute(n.args, values))
            kwargs = _materialise(_substitute(n.kwargs, values))
            ex = execs[nid]
            if ex is None:
                values[nid] = n.fn(*args, **kwargs)
            else:
                values[nid] = await ex.run(n.fn, args, kwargs, {"dispatch_id": dispatch_id, "node_id": nid})
            res.stats["host_electrons"] += 1
            remaining.discard(nid)
            progressed = True
        # 2. the largest set of device electrons closed under "every unfinished predecessor
        #    is in the set and leaves an array in the arena" (scalars come back to the host
        #    first: Python ints travel as immediates, never as device operands)
        group = [nid for nid in sorted(remaining) if device[nid] is not None]
        changed = True
        while changed:
            changed = False
            gs = set(group)
            for nid in list(group):
                if any(d in remaining and (d not in gs or lookup(nodes[d].fn).kind in _SCALAR_KINDS)
                       for d in nodes[nid].deps):
                    group.remove(nid)
                    changed = True
        if group:
            gs = set(group)
            keep = {nid for nid in group
                    if nid in needed_on_host or any(c not in gs for c in consumers[nid])}
            try:
                _run_device_group(nodes, group, device, values, keep, res)
                for nid in group:
                    remaining.discard(nid)
            except (NotDeviceEligible, _DeviceGroupFailed) as why:
                # The DAG path is an optimisation: when it cannot take these electrons they
                # go one by one through executor.run, which applies the reference's own
                # failure decision (_on_ssh_fail: RuntimeError, or a local run if opted in).
                res.stats.setdefault("demoted", []).append(str(why))
                for nid in group:
                    device[nid] = None
            progressed = True
        if not progressed:
            raise RuntimeError("lattice has a dependency cycle")


class _DeviceGroupFailed(Exception):
    pass


# ------------------------------------------------------- device compilation ---
from .electrons import K_IADD, K_IMUL, K_ISCALE_ADD, K_SUM  # noqa: E402

_SCALAR_KINDS = (K_SUM, K_IADD, K_IMUL, K_ISCALE_ADD)
SLICE_ELEMS = 1024           # column width of one slice of a split fan-in (4 KiB per row segment)
SPLIT_MIN_BYTES = 8 << 20    # fan-ins that read less than this stay one electron


def _meta_of(value):
    if isinstance(value, DeviceArray):
        return ("array", value.shape, value.order)
    if type(value) is np.ndarray and value.dtype == np.float32:
        order = "F" if (value.flags.f_contiguous and not value.flags.c_contiguous) else "C"
        return ("array", value.shape, order)
    if type(value) is int:
        return ("int", value)
    raise NotDeviceEligible(f"operand of type {type(value).__name__} is not device-eligible")


class _Planned:
    __slots__ = ("node", "spec", "plan", "operands", "out", "dev", "level")


def _run_device_group(nodes, group: List[int], device, values, keep, res: Result) -> None:
    """Compile ``group`` (device electrons, topologically sorted ids) and run it:
    one DAG launch if it lives on one GPU, else level by level."""
    engines: Dict[int, DeviceEngine] = {}
    for nid in group:
        d = device[nid]
        if d not in engines:
            engines[d] = get_engine(d)
    gset = set(group)
    planned: Dict[int, _Planned] = {}
    metas: Dict[int, tuple] = {}
    temporaries: List[DeviceArray] = []

    def operand_meta(v):
        if isinstance(v, NodeRef):
            if v.node_id in gset:
                return metas[v.node_id]
            return _meta_of(values[v.node_id])
        if isinstance(v, (list, tuple)):
            return [operand_meta(x) for x in v]
        return _meta_of(v)

    def settle(v):
        """Outputs of electrons that finished earlier become their values."""
        if isinstance(v, NodeRef):
            return v if v.node_id in gset else values[v.node_id]
        if isinstance(v, (list, tuple)):
            return [settle(x) for x in v]
        return v

    # ---- pass 1: shapes ------------------------------------------------------------
    for nid in group:
        n = nodes[nid]
        if n.kwargs:
            raise NotDeviceEligible("keyword arguments are not supported on the device DAG path")
        n = _Node(n.id, n.fn, [settle(a) for a in n.args], {}, n.executor)
        n.deps = nodes[nid].deps
        p = _Planned()
        p.node, p.spec, p.dev = n, lookup(n.fn), device[nid]
        p.plan = plan_call(p.spec, [operand_meta(a) for a in n.args])
        p.level = 1 + max([planned[d].level for d in n.deps if d in gset], default=-1)
        p.out = None
        planned[nid] = p
        rm = p.plan["result_meta"]
        metas[nid] = ("array", rm[2], rm[3]) if rm[0] == "array" else ("scalar", rm)

    # ---- pass 2: placement.  Rows of a fan-in are laid out back to back so the reducer
    # reads them in place (pack by construction); everything else gets its own slot. -----
    consumers_in: Dict[int, int] = {nid: 0 for nid in group}
    for nid in group:
        for r in _refs(planned[nid].node.args):
            if r.node_id in gset:
                consumers_in[r.node_id] += 1
    blocks: Dict[int, DeviceArray] = {}
    for nid in group:
        p = planned[nid]
        if p.spec.kind != K_STACK_SUM:
            continue
        rows = p.node.args[0]
        row_bytes = 4 * p.plan["count"]
        stride = (row_bytes + 255) & ~255
        eng = engines[p.dev]
        block = eng.alloc(stride * len(rows), (len(rows), stride // 4))
        temporaries.append(block)
        blocks[nid] = block
        for r, src in enumerate(rows):
            in_place = (isinstance(src, NodeRef) and src.node_id in gset and planned[src.node_id].out is None
                        and consumers_in[src.node_id] == 1 and src.node_id not in keep
                        and planned[src.node_id].dev == p.dev)
            if in_place:
                rm = planned[src.node_id].plan["result_meta"]
                planned[src.node_id].out = block.view(r * stride, row_bytes, rm[2])
    for nid in group:
        p = planned[nid]
        if p.out is None:
            rm = p.plan["result_meta"]
            shape = rm[2] if rm[0] == "array" else (max(p.plan["out_bytes"], 4) // 4,)
            p.out = engines[p.dev].alloc(p.plan["out_bytes"], shape, rm[3] if rm[0] == "array" else "C")
            if nid not in keep:
                temporaries.append(p.out)

    # ---- pass 3: task records ---------------------------------------------------------
    # Host operands of the whole group are uploaded first, one H2D + one unpack launch per
    # staging segment (K2) instead of one synchronous copy per array.
    uploaded: Dict[tuple, DeviceArray] = {}
    pending: Dict[int, Dict[int, np.ndarray]] = {d: {} for d in engines}

    def host_operands(v, dev):
        if isinstance(v, (list, tuple)):
            for x in v:
                host_operands(x, dev)
        elif type(v) is np.ndarray:
            pending[dev].setdefault(id(v), v)

    for nid in group:
        for a in planned[nid].node.args:
            host_operands(a, planned[nid].dev)
    for dev, arrs in pending.items():
        if arrs:
            das = engines[dev].put_many(list(arrs.values()))
            for key, da in zip(arrs, das):
                uploaded[(dev, key)] = da
                temporaries.append(da)

    def where(v, eng: DeviceEngine) -> DeviceArray:
        """Device location of an array operand."""
        if isinstance(v, NodeRef):
            v = planned[v.node_id].out if v.node_id in gset else values[v.node_id]
        if isinstance(v, DeviceArray):
            return v
        da = uploaded.get((eng.ordinal, id(v)))
        if da is None:
            da = eng.put(v)
            temporaries.append(da)
        return da

    per_dev: Dict[int, List[dict]] = {d: [] for d in engines}
    index_of: Dict[int, tuple] = {}              # node id -> (device, task index of its LAST task)
    first_of: Dict[int, tuple] = {}              # node id -> (device, task index that waits for its deps)

    def add(dev, **t) -> int:
        rec = dict(fn_id=0, flags=0, arg0=0, arg1=0, out=0, count=0, nrows=0, deps=0, succ=[], tag=0)
        rec.update(t)
        per_dev[dev].append(rec)
        return len(per_dev[dev]) - 1

    def operand(dev, eng, v, slot):
        da = where(v, eng)
        if da.ordinal == dev:
            return da.offset, 0
        shift = (L.F_ARENA0_SHIFT, L.F_ARENA1_SHIFT)[slot]
        return da.offset, eng.peer_arena_id(da.ordinal) << shift

    for nid in group:
        p = planned[nid]
        n, dev, eng = p.node, p.dev, engines[p.dev]
        k = p.spec.kind
        if p.plan["result_meta"][0] == "int":
            imm = [a if type(a) is int else None for a in n.args]
            if any(v is None for v in imm):
                raise NotDeviceEligible("integer electrons cannot take device-resident operands")
            t = add(dev, fn_id=p.spec.fn_id, flags=L.F_IMM0 | L.F_IMM1, arg0=imm[0] & (2 ** 64 - 1),
                    arg1=(imm[1] if len(imm) > 1 else 0) & (2 ** 64 - 1), out=p.out.offset, count=1)
            first_of[nid] = index_of[nid] = (dev, t)
            continue
        if k == K_STACK_SUM:
            block, rows = blocks[nid], n.args[0]
            row_bytes = 4 * p.plan["count"]
            stride = block.shape[1] * 4
            copies = []
            for r, src in enumerate(rows):
                row = block.view(r * stride, row_bytes, (p.plan["count"],))
                if isinstance(src, NodeRef) and src.node_id in gset and planned[src.node_id].out.offset == row.offset \
                        and planned[src.node_id].out.engine is eng:
                    continue                                  # producer writes this row itself
                da = where(src, eng)
                off, flg = operand(dev, eng, da, 0)           # gather (K1 as an electron): copy the row in
                c = add(dev, fn_id=_copy_fn_id(), flags=flg, arg0=off, out=row.offset, count=row_bytes)
                copies.append((c, src))
            total = p.plan["count"]
            if p.plan["nrows"] * row_bytes >= SPLIT_MIN_BYTES and total > SLICE_ELEMS:
                gate = add(dev, fn_id=0)
                slices = []
                for c0 in range(0, total, SLICE_ELEMS):
                    w = min(SLICE_ELEMS, total - c0)
                    slices.append(add(dev, fn_id=p.spec.fn_id, arg0=block.offset + 4 * c0, arg1=stride,
                                      out=p.out.offset + 4 * c0, count=w, nrows=p.plan["nrows"], deps=1))
                per_dev[dev][gate]["succ"] = list(slices)
                join = add(dev, fn_id=0, deps=len(slices))
                for s_ in slices:
                    per_dev[dev][s_]["succ"] = [join]
                first_of[nid], index_of[nid] = (dev, gate), (dev, join)
            else:
                t = add(dev, fn_id=p.spec.fn_id, arg0=block.offset, arg1=stride, out=p.out.offset,
                        count=total, nrows=p.plan["nrows"])
                first_of[nid] = index_of[nid] = (dev, t)
            # the fan-in waits for its row copies as well as for its producers
            for c, src in copies:
                per_dev[dev][c]["succ"] = [first_of[nid][1]]
                per_dev[dev][first_of[nid][1]]["deps"] += 1
                if isinstance(src, NodeRef) and src.node_id in gset:
                    _edge(per_dev, index_of, src.node_id, dev, c)
            for src in rows:
                if isinstance(src, NodeRef) and src.node_id in gset and \
                        not any(s is src for _, s in copies):
                    _edge(per_dev, index_of, src.node_id, dev, first_of[nid][1])
            continue
        a0, f0 = operand(dev, eng, n.args[0], 0)
        a1, f1 = (operand(dev, eng, n.args[1], 1) if len(n.args) > 1 else (0, 0))
        t = add(dev, fn_id=p.spec.fn_id, flags=f0 | f1, arg0=a0, arg1=a1, out=p.out.offset,
                count=p.plan["count"])
        first_of[nid] = index_of[nid] = (dev, t)
        for a in n.args:
            if isinstance(a, NodeRef) and a.node_id in gset:
                _edge(per_dev, index_of, a.node_id, dev, t)

    # ---- pass 4: launch ------------------------------------------------------------------
    single = len(engines) == 1
    import time as _time

    t_launch = _time.perf_counter()
    if single:
        dev = next(iter(engines))
        st = _launch(engines[dev], per_dev[dev], None)
        _check(st, per_dev[dev])
        res.stats["device_batches"] += 1
    else:
        # several GPUs: edges that cross devices cannot be released on the device, so run
        # level by level; tasks of one level are independent, consumers read peers' arenas.
        levels: Dict[int, Dict[int, List[int]]] = {}
        lvl_of: Dict[tuple, int] = {}
        for dev, recs in per_dev.items():
            for i, _ in enumerate(recs):
                lvl_of[(dev, i)] = 0
        changed = True
        while changed:                      # longest-path levels over the task graph
            changed = False
            for dev, recs in per_dev.items():
                for i, rec in enumerate(recs):
                    for (sd, si) in rec.get("xsucc", []) + [(dev, s_) for s_ in rec["succ"]]:
                        if lvl_of[(sd, si)] < lvl_of[(dev, i)] + 1:
                            lvl_of[(sd, si)] = lvl_of[(dev, i)] + 1
                            changed = True
        for (dev, i), lv in lvl_of.items():
            levels.setdefault(lv, {}).setdefault(dev, []).append(i)
        for lv in sorted(levels):
            # every GPU works on its share of the level at the same time
            handles = [(dev, sorted(idxs), _launch(engines[dev], per_dev[dev], sorted(idxs), wait=False))
                       for dev, idxs in levels[lv].items()]
            for dev, idxs, h in handles:
                st = engines[dev].finish_dag(h)
                _check(st, [per_dev[dev][i] for i in idxs])
                res.stats["device_batches"] += 1
    res.stats["device_electrons"] += len(group)
    res.stats["device_tasks"] = res.stats.get("device_tasks", 0) + sum(len(v) for v in per_dev.values())
    res.stats["dag_seconds"] = res.stats.get("dag_seconds", 0.0) + (_time.perf_counter() - t_launch)

    # ---- results ---------------------------------------------------------------------------
    for nid in group:
        p = planned[nid]
        rm = p.plan["result_meta"]
        if rm[0] == "array":
            values[nid] = p.out if nid in keep else None
        else:
            raw = np.empty(max(p.plan["out_bytes"], 16), dtype=np.uint8)
            L.call("b2_read", p.dev, p.out.offset, raw.ctypes.data, p.plan["out_bytes"])
            values[nid] = decode(rm, raw)
    for t in temporaries:
        t.free()


def _edge(per_dev, index_of, producer: int, dev: int, consumer_task: int) -> None:
    pdev, ptask = index_of[producer]
    if pdev == dev:
        per_dev[pdev][ptask]["succ"].append(consumer_task)
    else:
        per_dev[pdev][ptask].setdefault("xsucc", []).append((dev, consumer_task))
    per_dev[dev][consumer_task]["deps"] += 1


_COPY_ID: List[int] = []


def _copy_fn_id() -> int:
    """Callback-table entry used for row gathers inside a DAG (registered lazily)."""
    if not _COPY_ID:
        from .electrons import device_electron

        @device_electron("copy", name="b200.lattice_row_copy")
        def _row_copy(x):
            return x

        _COPY_ID.append(lookup(_row_copy).fn_id)
    return _COPY_ID[0]


def _launch(eng: DeviceEngine, recs: List[dict], subset: Optional[List[int]], wait: bool = True):
    """Task table for ``recs`` (all of them with on-device dependency release, or an
    independent ``subset`` of one level), ready electrons first."""
    if subset is None:
        order = sorted(range(len(recs)), key=lambda i: (recs[i]["deps"] != 0, i))
        remap = {old: new for new, old in enumerate(order)}
        table = np.zeros(len(order), dtype=L.TASK_DTYPE)
        succ: List[int] = []
        for new, old in enumerate(order):
            r = recs[old]
            table[new] = (r["fn_id"], r["flags"], r["arg0"], r["arg1"], r["out"], r["count"], r["nrows"],
                          r["deps"], len(succ), len(r["succ"]), 0)
            succ += [remap[s] for s in r["succ"]]
        st = eng.run_dag(table, np.array(succ, dtype=np.uint32) if succ else None)
        out = np.zeros(len(recs), dtype=np.uint32)
        out[order] = st
        return out
    table = np.zeros(len(subset), dtype=L.TASK_DTYPE)
    for new, old in enumerate(subset):
        r = recs[old]
        table[new] = (r["fn_id"], r["flags"], r["arg0"], r["arg1"], r["out"], r["count"], r["nrows"], 0, 0, 0, 0)
    return eng.run_dag(table, None) if wait else eng.submit_dag(table, None)


def _check(st: np.ndarray, recs: Sequence[dict]) -> None:
    bad = np.nonzero(st != L.TS_OK)[0]
    if len(bad):
        i = int(bad[0])
        names = {L.TS_OVERFLOW: "integer result does not fit int64", L.TS_BAD_FN: "no device callback registered",
                 L.TS_BAD_ARG: "operand rejected by the device callback", L.TS_PENDING: "electron never ran"}
        raise _DeviceGroupFailed(f"device electron {i} of the lattice failed: "
                                 f"{names.get(int(st[i]), 'device error')} (status {int(st[i])})")
