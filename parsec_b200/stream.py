"""Python mirror of the streaming C ABI (include/pb2_stream.h): ctypes plumbing for tests and bench only."""
import ctypes as C

import numpy as np

from . import _lib as L


class StreamParams(C.Structure):
    _fields_ = [("cmd_slots", C.c_int32), ("max_tiles", C.c_int32), ("idle_us", C.c_int32), ("dry_run", C.c_int32),
                ("timeout_ms", C.c_int32), ("part_bytes", C.c_int32), ("max_workers", C.c_int32), ("trace", C.c_int32)]


class Retire(C.Structure):
    _fields_ = [("cookie", C.c_uint64), ("result", C.c_uint64), ("seen_version", C.c_uint32 * L.MAX_FLOWS),
                ("ticket", C.c_int32), ("status", C.c_int32),
                ("t_start_ns", C.c_uint64), ("t_end_ns", C.c_uint64), ("smid", C.c_uint32), ("pad", C.c_uint32)]


class StreamStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("submitted", "retired", "bytes_h2d", "bytes_d2d", "bytes_d2h", "stage_ins",
                                          "body_errors", "kernel_launches", "edges", "edges_late", "released_on_device")]


STREAM_SYMBOLS = ["pb2_stream_create", "pb2_stream_destroy", "pb2_stream_last_error", "pb2_stream_set_tile",
                  "pb2_stream_submit", "pb2_stream_add_edge", "pb2_stream_kick", "pb2_stream_poll",
                  "pb2_stream_quiesce", "pb2_stream_stats", "pb2_stream_inflight"]

_bound = False


def lib():
    global _bound
    l = L.load()
    if not _bound:
        vp, i32 = C.c_void_p, C.c_int32
        l.pb2_stream_create.argtypes = [vp, C.POINTER(StreamParams), C.POINTER(vp)]
        l.pb2_stream_destroy.argtypes = [vp]
        l.pb2_stream_last_error.argtypes = [vp]
        l.pb2_stream_last_error.restype = C.c_char_p
        l.pb2_stream_set_tile.argtypes = [vp, i32, vp]
        l.pb2_stream_submit.argtypes = [vp, vp, C.c_uint64, C.POINTER(i32)]
        l.pb2_stream_add_edge.argtypes = [vp, i32, i32]
        l.pb2_stream_kick.argtypes = [vp]
        l.pb2_stream_poll.argtypes = [vp, C.POINTER(Retire), i32]
        l.pb2_stream_quiesce.argtypes = [vp]
        l.pb2_stream_stats.argtypes = [vp, C.POINTER(StreamStats)]
        l.pb2_stream_inflight.argtypes = [vp]
        for n in STREAM_SYMBOLS:
            if n != "pb2_stream_last_error":
                getattr(l, n).restype = C.c_int
        _bound = True
    return l


class Stream:
    """One streaming ring + persistent kernel (engine=None only for dry_run)."""

    def __init__(self, engine=None, **kw):
        self._l = lib()
        self._h = C.c_void_p()
        p = StreamParams(**kw)
        rc = self._l.pb2_stream_create(engine._h if engine is not None else None, C.byref(p), C.byref(self._h))
        if rc != L.PB2_SUCCESS:
            self._h = C.c_void_p()
            raise L.Pb2Error(rc, "pb2_stream_create")
        self._buf = (Retire * 4096)()

    def _check(self, rc, what):
        if rc != L.PB2_SUCCESS:
            raise L.Pb2Error(rc, what, (self._l.pb2_stream_last_error(self._h) or b"").decode())

    def set_tile(self, tile_id, tile_row):
        row = np.ascontiguousarray(tile_row)
        self._check(self._l.pb2_stream_set_tile(self._h, tile_id, row.ctypes.data_as(C.c_void_p)), "pb2_stream_set_tile")

    def submit(self, task_row, cookie=0, allow_full=False):
        row = np.ascontiguousarray(task_row)
        tk = C.c_int32(-1)
        rc = self._l.pb2_stream_submit(self._h, row.ctypes.data_as(C.c_void_p), cookie, C.byref(tk))
        if allow_full and rc == L.PB2_ERR_OUT_OF_RESOURCE:
            return None
        self._check(rc, "pb2_stream_submit")
        return tk.value

    def add_edge(self, pred, succ):
        self._check(self._l.pb2_stream_add_edge(self._h, pred, succ), "pb2_stream_add_edge")

    def kick(self):
        self._check(self._l.pb2_stream_kick(self._h), "pb2_stream_kick")

    def poll(self, maxn=4096):
        n = self._l.pb2_stream_poll(self._h, self._buf, min(maxn, 4096))
        if n < 0:
            self._check(n, "pb2_stream_poll")
        return [(r.cookie, r.result, tuple(r.seen_version), r.ticket, r.status) for r in self._buf[:n]]

    def quiesce(self):
        self._check(self._l.pb2_stream_quiesce(self._h), "pb2_stream_quiesce")

    def inflight(self):
        return self._l.pb2_stream_inflight(self._h)

    def stats(self):
        s = StreamStats()
        self._check(self._l.pb2_stream_stats(self._h, C.byref(s)), "pb2_stream_stats")
        return {f[0]: getattr(s, f[0]) for f in StreamStats._fields_}

    def close(self):
        if self._h:
            self._l.pb2_stream_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def run_dag(stream, dag, tiles, mode="lookahead", poll_budget=10_000_000):
    """Drive a whole oracle-format DAG (tasks, CSR succ, tiles, ready) through a stream.

    mode "lookahead": every task is submitted up front with dep_goal = in-degree and one device edge per
    in-edge, so successors are released on the device; mode "host": only ready tasks are submitted and the
    host releases successors when it polls a retirement (one host round trip per edge, what a device module
    without look-ahead does).  Returns per-task result / seen_version arrays and the retire order."""
    n = len(dag.tasks)
    for i in range(len(tiles)):
        stream.set_tile(i, tiles[i:i + 1])
    succ_of = [[int(x) & 0x07FFFFFF for x in dag.succ[t["succ_begin"]:t["succ_begin"] + t["succ_count"]]] for t in dag.tasks]
    indeg = np.zeros(n, np.int64)
    preds = [[] for _ in range(n)]
    for u in range(n):
        for v in succ_of[u]:
            indeg[v] += 1
            preds[v].append(u)
    result = np.zeros(n, np.uint64)
    seen = np.zeros((n, L.MAX_FLOWS), np.uint32)
    order = []
    ticket = {}

    def drain(block_until_empty):
        budget = poll_budget
        while True:
            recs = stream.poll()
            for cookie, res, sv, tk, status in recs:
                assert status == 0
                t = int(cookie)
                result[t] = res
                seen[t] = sv
                order.append(t)
                if mode == "host":
                    for v in succ_of[t]:
                        indeg[v] -= 1
                        if indeg[v] == 0:
                            push(v, 0)
            if not block_until_empty or stream.inflight() == 0:
                return
            budget -= 1
            if budget <= 0:
                raise TimeoutError("stream did not drain")

    def push(t, goal):
        row = dag.tasks[t:t + 1].copy()
        row["dep_goal"] = goal
        while True:
            tk = stream.submit(row, cookie=t, allow_full=True)
            if tk is not None:
                ticket[t] = tk
                return tk
            stream.kick()
            drain(False)

    if mode == "lookahead":
        # topological order: a task's predecessors are always submitted (and unpolled) before it
        order_sub, deg, q = [], indeg.copy(), [int(r) for r in dag.ready]
        while q:
            u = q.pop(0)
            order_sub.append(u)
            for v in succ_of[u]:
                deg[v] -= 1
                if deg[v] == 0:
                    q.append(v)
        assert len(order_sub) == n
        for t in order_sub:
            row = dag.tasks[t:t + 1].copy()
            while True:
                # a predecessor whose retirement has been polled has lost its ticket: it counts as satisfied
                done = set(order)
                live = [p for p in preds[t] if p not in done]
                row["dep_goal"] = len(live)
                tk = stream.submit(row, cookie=t, allow_full=True)
                if tk is not None:
                    break
                stream.kick()
                drain(False)
            ticket[t] = tk
            for p in live:
                stream.add_edge(ticket[p], tk)
    else:
        for r in dag.ready:
            push(int(r), 0)
    stream.kick()
    drain(True)
    return {"result": result, "seen_version": seen, "retire_order": np.array(order, np.int64)}
