"""ORACLE (test infrastructure): numpy restatement of the reference's task graphs as engine windows.

Each builder restates the dataflow of one reference JDF / DTD program as the arrays the engine
consumes: ``pb2_task_t[]``, CSR successor lists, the ids of the startup tasks and the tiles each flow
touches.  The product path builds the same graphs with the C++ DSL shims (parsec_b200/csrc/pb2_dsl.cpp);
the tests cross-check those, and the oracle's own DTD rule (orc_dtd.c), against these builders.

Dependency-goal conventions follow what ``parsec-ptgpp`` emits:
  * PTG task classes use the *mask* mode (``PARSEC_USE_DEPS_MASK``): one bit per input flow that is
    fed by a task, ``dependencies_goal`` = OR of those bits (parsec/parsec.c:1656-1720); tasks whose
    active inputs all come from memory / NEW are startup tasks (parsec.c:1372-1446).
  * DTD tasks use a counter of unsatisfied flows (``flow_count``, insert_function.c:2962-2976).
"""
from dataclasses import dataclass, field

import numpy as np

from . import orc as L


@dataclass
class Dag:
    tasks: np.ndarray            # TASK_DTYPE[ntasks]
    succ: np.ndarray             # uint32[nsucc]
    ready: np.ndarray            # int32[nready]
    ntiles: int
    tile_bytes: int
    kind: int = 0                # 0 = HBM bodies, 1 = GEMM bodies
    name: str = ""
    meta: dict = field(default_factory=dict)

    @property
    def ntasks(self):
        return len(self.tasks)

    def edges(self):
        """(src, dst, dst_flow) arrays of every dependency edge."""
        cnt = self.tasks["succ_count"].astype(np.int64)
        src = np.repeat(np.arange(self.ntasks, dtype=np.int64), cnt)
        dst = (self.succ & np.uint32(0x07FFFFFF)).astype(np.int64)
        flow = (self.succ >> np.uint32(27)).astype(np.int64)
        return src, dst, flow


def _new_tasks(n):
    t = np.zeros(n, dtype=L.TASK_DTYPE)
    t["tile"][:] = -1
    return t


def _csr_from_edges(ntasks, src, dst, flow):
    """Stable CSR (edges of one source keep their given order = iterate_successors order)."""
    src = np.asarray(src, np.int64)
    order = np.argsort(src, kind="stable")
    src, dst, flow = src[order], np.asarray(dst, np.int64)[order], np.asarray(flow, np.int64)[order]
    count = np.bincount(src, minlength=ntasks).astype(np.int32)
    begin = np.zeros(ntasks, np.int32)
    np.cumsum(count[:-1], out=begin[1:])
    succ = ((flow.astype(np.uint32) << np.uint32(27)) | dst.astype(np.uint32)).astype(np.uint32)
    return begin, count, succ


def ex05_broadcast(K, NB=14, tile_bytes=256 * 256 * 4):
    """examples/Ex05_Broadcast.jdf:24-58 with tile-sized data (BASELINE config 2).

    TaskBcast(k), k=0..K-1:   RW  A <- mydata(k)            -> A TaskRecv(k, 0..NB..2);  body: A[:] = k
    TaskRecv(k,n), n=0..NB..2: READ A <- A TaskBcast(k);                                  body: check A == k
    Task ids: Bcast(k) = k ; Recv(k, n) = K + k*F + n/2 with F = NB/2 + 1.
    """
    F = NB // 2 + 1
    n = K + K * F
    t = _new_tasks(n)
    k = np.arange(K, dtype=np.int32)
    # TaskBcast
    b = t[:K]
    b["body"] = L.BODY_FILL_I32
    b["nb_flows"] = 1
    b["flags"] = L.TASK_DEPS_MASK
    b["class_id"] = 0
    b["dep_goal"] = 0                      # A <- mydata(k): memory input, startup task
    b["tile"][:, 0] = k
    b["access"][:, 0] = L.ACCESS_RW
    b["iparam"][:, 0] = k
    b["locals"][:, 0] = k
    # TaskRecv
    r = t[K:]
    kk = np.repeat(k, F)
    nn = np.tile(np.arange(F, dtype=np.int32) * 2, K)
    r["body"] = L.BODY_CHECK_I32
    r["nb_flows"] = 1
    r["flags"] = L.TASK_DEPS_MASK
    r["class_id"] = 1
    r["dep_goal"] = 0x1                    # flow A (index 0) is fed by TaskBcast
    r["tile"][:, 0] = kk
    r["access"][:, 0] = L.ACCESS_READ
    r["iparam"][:, 0] = kk
    r["locals"][:, 0] = kk
    r["locals"][:, 1] = nn
    src = kk.astype(np.int64)
    dst = K + np.arange(K * F, dtype=np.int64)
    begin, count, succ = _csr_from_edges(n, src, dst, np.zeros(K * F, np.int64))
    t["succ_begin"], t["succ_count"] = begin, count
    return Dag(t, succ, k.copy(), ntiles=K, tile_bytes=tile_bytes, name="ex05_broadcast",
               meta={"K": K, "NB": NB, "F": F})


def ex02_chain(NB, tile_bytes=4):
    """examples/Ex02_Chain.jdf:27-50: Task(k), k=0..NB; RW A <- (k==0)? NEW : A Task(k-1) -> A Task(k+1).
    Body: k==0 ? A=0 : A+=1.  One NEW (arena) datum circulates: a single tile, never staged in
    (NEW + version 0 => no transfer, device_gpu.c:2049-2052)."""
    n = NB + 1
    t = _new_tasks(n)
    k = np.arange(n, dtype=np.int32)
    t["body"] = L.BODY_INCR_I32
    t["body"][0] = L.BODY_FILL_I32
    t["iparam"][:, 0] = 1
    t["iparam"][0, 0] = 0
    t["nb_flows"] = 1
    t["flags"] = L.TASK_DEPS_MASK
    t["dep_goal"] = 0x1
    t["dep_goal"][0] = 0
    t["tile"][:, 0] = 0
    t["access"][:, 0] = L.ACCESS_RW
    t["access"][0, 0] = L.ACCESS_WRITE     # NEW: nothing to read
    t["locals"][:, 0] = k
    begin, count, succ = _csr_from_edges(n, k[:-1], k[1:], np.zeros(NB, np.int64))
    t["succ_begin"], t["succ_count"] = begin, count
    return Dag(t, succ, np.array([0], np.int32), ntiles=1, tile_bytes=tile_bytes, name="ex02_chain",
               meta={"NB": NB})


def rtt_chain(NT, FRAGS=1, tile_bytes=1024 * 1024 * 4, body=L.BODY_INCR_F32, pushout_last=True):
    """tests/apps/pingpong/rtt.jdf:26-47: PING(k,f), k=0..NT-1, f=0..FRAGS-1, on A(f, k % WS).
    RW T <- (k==0) ? A(f,0) : T PING(k-1,f) -> (k<NT-1) ? T PING(k+1,f) : A(f, k%WS).
    Single-GPU restatement: FRAGS independent chains, tile f; body T[:] += 1 (BASELINE config 4)."""
    n = NT * FRAGS
    t = _new_tasks(n)
    k = np.repeat(np.arange(NT, dtype=np.int32), FRAGS)
    f = np.tile(np.arange(FRAGS, dtype=np.int32), NT)
    t["body"] = body
    if body == L.BODY_INCR_I32:
        t["iparam"][:, 0] = 1
    else:
        t["fparam"] = 1.0
    t["nb_flows"] = 1
    t["flags"] = L.TASK_DEPS_MASK
    t["dep_goal"] = np.where(k == 0, 0, 0x1)
    t["tile"][:, 0] = f
    t["access"][:, 0] = L.ACCESS_RW
    if pushout_last:
        t["access"][k == NT - 1, 0] = L.ACCESS_RW | L.FLOW_PUSHOUT   # -> A(f, k % WS): write back home
    t["locals"][:, 0] = k
    t["locals"][:, 1] = f
    ids = np.arange(n, dtype=np.int64)
    m = k < NT - 1
    begin, count, succ = _csr_from_edges(n, ids[m], ids[m] + FRAGS, np.zeros(m.sum(), np.int64))
    t["succ_begin"], t["succ_count"] = begin, count
    return Dag(t, succ, np.arange(FRAGS, dtype=np.int32), ntiles=FRAGS, tile_bytes=tile_bytes,
               name="rtt_chain", meta={"NT": NT, "FRAGS": FRAGS})


def ep(NT, DEPTH):
    """tests/runtime/scheduling/ep.jdf:15-40 (schedmicro): INIT(0) -CTL-> TASK(1..NT,1); TASK(i,l) -CTL-> TASK(i,l+1).
    Empty bodies, CTL flows only (no data): pure scheduling cost.  ids: INIT=0, TASK(i,l)=1+(l-1)*NT+(i-1)."""
    n = 1 + NT * DEPTH
    t = _new_tasks(n)
    t["body"] = L.BODY_NOP
    t["nb_flows"] = 1
    t["flags"] = L.TASK_DEPS_MASK
    t["dep_goal"] = 0x1
    t["dep_goal"][0] = 0
    t["class_id"][1:] = 1
    ids = np.arange(1, n, dtype=np.int64)
    t["locals"][1:, 0] = (ids - 1) % NT + 1
    t["locals"][1:, 1] = (ids - 1) // NT + 1
    src = [np.zeros(NT if DEPTH >= 1 else 0, np.int64)]
    dst = [np.arange(1, 1 + (NT if DEPTH >= 1 else 0), dtype=np.int64)]
    inner = ids[ids + NT < n]
    src.append(inner)
    dst.append(inner + NT)
    src, dst = np.concatenate(src), np.concatenate(dst)
    begin, count, succ = _csr_from_edges(n, src, dst, np.zeros(len(src), np.int64))
    t["succ_begin"], t["succ_count"] = begin, count
    return Dag(t, succ, np.array([0], np.int32), ntiles=0, tile_bytes=0, name="ep",
               meta={"NT": NT, "DEPTH": DEPTH})


def dtd_gemm(NT, tile=512, elem_bytes=2):
    """tests/dsl/dtd/dtd_test_simple_gemm.c:675-696: for i, for j, for k: GEMM(A(i,k) IN, B(k,j) IN, C(i,j) INOUT),
    last k PUSHOUT.  DTD dependency rule (insert_function.c:3006-3260, overlap_strategies.c:139-):
    per tile, a writer waits for the previous writer and all readers in between; consecutive readers
    wait only for the previous writer.  A and B are never written => only the C(i,j) chain over k.
    ids follow insertion order ((i*NT + j)*NT + k); tiles: A(i,k) = i*NT+k, B(k,j) = NT^2 + k*NT+j,
    C(i,j) = 2*NT^2 + i*NT+j.  flow_count counter mode."""
    n = NT ** 3
    t = _new_tasks(n)
    ids = np.arange(n, dtype=np.int64)
    i, j, k = ids // (NT * NT), (ids // NT) % NT, ids % NT
    t["body"] = L.BODY_GEMM_BF16
    t["nb_flows"] = 3
    t["flags"] = 0
    t["dep_goal"] = np.where(k == 0, 0, 1)
    t["tile"][:, 0] = i * NT + k
    t["tile"][:, 1] = NT * NT + k * NT + j
    t["tile"][:, 2] = 2 * NT * NT + i * NT + j
    t["access"][:, 0] = L.ACCESS_READ
    t["access"][:, 1] = L.ACCESS_READ
    t["access"][:, 2] = np.where(k == NT - 1, L.ACCESS_RW | L.FLOW_PUSHOUT, L.ACCESS_RW)
    t["iparam"][:] = tile
    # priority exactly as the reference passes it: mt*nt*kt - i*nt + j
    t["priority"] = NT * NT * NT - i * NT + j
    t["locals"][:, 0] = i
    t["locals"][:, 1] = j
    m = k < NT - 1
    begin, count, succ = _csr_from_edges(n, ids[m], ids[m] + 1, np.full(m.sum(), 2, np.int64))
    t["succ_begin"], t["succ_count"] = begin, count
    return Dag(t, succ, ids[k == 0].astype(np.int32), ntiles=3 * NT * NT,
               tile_bytes=tile * tile * elem_bytes, kind=1, name="dtd_gemm", meta={"NT": NT, "tile": tile})


def check_execution(dag, res):
    """Dependency-order parity checks on one window run (all integer, all exact).

    1. every task retired exactly once;  2. the retire log is a linear extension of the DAG;
    3. for every edge u->v: end_seq[u] < start_seq[v] in the single global event order.
    Returns a dict of violation counts (all zero == pass)."""
    n = dag.ntasks
    order = res["retire_order"].astype(np.int64)
    out = {}
    out["not_permutation"] = int(n - len(np.unique(order))) if n else 0
    pos = np.empty(n, np.int64)
    pos[order] = np.arange(n)
    src, dst, _ = dag.edges()
    out["retire_order_violations"] = int(np.sum(pos[src] >= pos[dst]))
    out["event_order_violations"] = int(np.sum(res["end_seq"][src].astype(np.int64) >= res["start_seq"][dst].astype(np.int64)))
    out["start_after_end"] = int(np.sum(res["start_seq"].astype(np.int64) >= res["end_seq"].astype(np.int64)))
    return out


def ptg_pingpong(NB_TOKEN, tile_bytes=None):
    """tests/runtime/cuda/ptg_pingpong.jdf:44-165: one token tile T hops INIT -> TOKEN_CPU(k) -> TOKEN_GPU(k,0)
    -> TOKEN_GPU(k,1) -> TOKEN_CPU(k+1) ... -> CHECK.  INIT: tile[i] = i; TOKEN_CPU(k): tile[2k] += 2k,
    tile[2k+1] += 2k+1; TOKEN_GPU(k,l): tile[2k+l] += 2k+l (ping_kernel.cu:15); CHECK: tile[i] == 3*i.
    LOAD_DIST(r,d) anchors (READ D) are placement only and omitted on one device.
    ids: INIT=0, CPU(k)=1+4k, CPUb(k)=2+4k (second element), GPU(k,0)=3+4k, GPU(k,1)=4+4k, CHECK=last.
    TOKEN_CPU touches two elements: restated as two chained single-element updates."""
    n = 1 + 4 * NB_TOKEN + 1
    tb = tile_bytes or 2 * NB_TOKEN * 4
    t = _new_tasks(n)
    t["nb_flows"] = 1
    t["flags"] = L.TASK_DEPS_MASK
    t["tile"][:, 0] = 0
    t["access"][:, 0] = L.ACCESS_RW
    t["dep_goal"] = 0x1
    t["body"] = L.BODY_ADD_AT_I32
    t["body"][0] = L.BODY_IOTA_I32
    t["access"][0, 0] = L.ACCESS_WRITE       # WRITE T <- NEW
    t["dep_goal"][0] = 0
    for k in range(NB_TOKEN):
        base = 1 + 4 * k
        for off, idx in ((0, 2 * k), (1, 2 * k + 1), (2, 2 * k), (3, 2 * k + 1)):
            t["iparam"][base + off, 0] = idx
            t["iparam"][base + off, 1] = idx
            t["locals"][base + off, 0] = k
            t["locals"][base + off, 1] = off
        t["class_id"][base:base + 2] = 1
        t["class_id"][base + 2:base + 4] = 2
    t["body"][n - 1] = L.BODY_NOP                # CHECK runs on the host in the reference: the test checks
    t["access"][n - 1, 0] = L.ACCESS_READ        # the tile the last GPU task pushed out (successor is a CPU task)
    t["access"][n - 2, 0] = L.ACCESS_RW | L.FLOW_PUSHOUT
    t["class_id"][n - 1] = 3
    ids = np.arange(n - 1, dtype=np.int64)
    begin, count, succ = _csr_from_edges(n, ids, ids + 1, np.zeros(n - 1, np.int64))
    t["succ_begin"], t["succ_count"] = begin, count
    return Dag(t, succ, np.array([0], np.int32), ntiles=1, tile_bytes=tb, name="ptg_pingpong",
               meta={"NB_TOKEN": NB_TOKEN})


def dtd_new_tile(nb_tiles, nb_elems):
    """tests/dsl/dtd/dtd_test_new_tile.c + dtd_test_new_tile_cuda_kernels.cu:14-56: per NEW tile,
    init (A[i] = i) -> multiply_by_two (A[i] *= 2) -> sum_add / check (A[i] == 2*i).  DTD counter mode:
    each task depends on the previous writer of its tile."""
    n = 3 * nb_tiles
    t = _new_tasks(n)
    tile = np.repeat(np.arange(nb_tiles, dtype=np.int32), 3)
    stage = np.tile(np.arange(3, dtype=np.int32), nb_tiles)
    t["nb_flows"] = 1
    t["tile"][:, 0] = tile
    t["body"] = np.choose(stage, [L.BODY_IOTA_I32, L.BODY_SCALE_I32, L.BODY_NOP])
    t["iparam"][:, 0] = 2
    t["access"][:, 0] = np.choose(stage, [L.ACCESS_WRITE, L.ACCESS_RW, L.ACCESS_READ | 0])
    t["dep_goal"] = np.where(stage == 0, 0, 1)
    t["locals"][:, 0] = tile
    t["locals"][:, 1] = stage
    ids = np.arange(n, dtype=np.int64)
    m = stage < 2
    begin, count, succ = _csr_from_edges(n, ids[m], ids[m] + 1, np.zeros(m.sum(), np.int64))
    t["succ_begin"], t["succ_count"] = begin, count
    return Dag(t, succ, ids[stage == 0].astype(np.int32), ntiles=nb_tiles, tile_bytes=nb_elems * 4,
               name="dtd_new_tile", meta={"nb_tiles": nb_tiles, "nb_elems": nb_elems})
