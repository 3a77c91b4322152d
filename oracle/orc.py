"""ctypes front-end of the CPU oracle (oracle/liboracle.so) and of oracle/_ref (real reference pieces).

ORACLE = test infrastructure.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; nothing under parsec_b200/ does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
REF_ZONE_PATH = os.path.join(_HERE, "_ref", "libzone_ref.so")
REF_DATA_PATH = os.path.join(_HERE, "_ref", "libdata_ref.so")
REF_TWODBC_PATH = os.path.join(_HERE, "_ref", "libtwodbc_ref.so")
REF_SELECT_PATH = os.path.join(_HERE, "_ref", "libselect_ref.so")

# flow access bits / bodies / flags: same values as include/pb2_engine.h (restated, not imported)
ACCESS_NONE, ACCESS_READ, ACCESS_WRITE, ACCESS_RW, FLOW_PUSHOUT = 0x00, 0x04, 0x08, 0x0C, 0x40
BODY_NOP, BODY_FILL_I32, BODY_CHECK_I32, BODY_INCR_I32, BODY_ADD_IOTA_I32 = 0, 1, 2, 3, 4
BODY_SCALE_I32, BODY_IOTA_I32, BODY_COPY, BODY_FILL_F32, BODY_CHECK_F32 = 5, 6, 7, 8, 9
BODY_INCR_F32, BODY_AXPY_F32, BODY_MEMSET_U8, BODY_ADD_AT_I32, BODY_GEMM_BF16 = 10, 11, 12, 13, 16
TASK_DEPS_MASK = 0x01
TILE_INVALID, TILE_STAGING, TILE_VALID = 0, 1, 2
SRC_HOST, SRC_PEER = 0, 1
MAX_FLOWS = 4
DTD_INPUT, DTD_OUTPUT, DTD_INOUT = 1, 2, 3

TASK_DTYPE = np.dtype([
    ("dep_goal", "<i4"), ("succ_begin", "<i4"), ("succ_count", "<i4"), ("priority", "<i4"),
    ("body", "u1"), ("nb_flows", "u1"), ("flags", "u1"), ("class_id", "u1"),
    ("tile", "<i4", (4,)), ("access", "u1", (4,)),
    ("iparam", "<i4", (3,)), ("fparam", "<f4"), ("locals", "<i4", (2,)),
], align=False)
TILE_DTYPE = np.dtype([
    ("dev_ptr", "<u8"), ("src_ptr", "<u8"), ("bytes", "<u4"), ("state", "<i4"),
    ("version", "<u4"), ("src_kind", "<i4"),
], align=False)


class OrcStats(C.Structure):
    _fields_ = [("tasks_retired", C.c_uint64), ("bytes_h2d", C.c_uint64), ("bytes_d2d", C.c_uint64),
                ("bytes_d2h", C.c_uint64), ("stage_ins", C.c_uint64), ("body_errors", C.c_uint64)]


class TwoDBC(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("myrank", "mb", "nb", "lm", "ln", "i", "j", "m", "n", "P", "Q", "kp", "kq", "ip", "jq",
                                       "lmt", "lnt", "mt", "nt")] + [("bsiz", C.c_int64)] + \
               [(n, C.c_int) for n in ("rrank", "crank", "nb_elem_r", "nb_elem_c", "nb_local_tiles", "llm", "lln")]


class OrcCopy(C.Structure):
    _fields_ = [("present", C.c_int32), ("coherency_state", C.c_int32), ("data_transfer_status", C.c_int32),
                ("readers", C.c_int32), ("version", C.c_uint32), ("flags", C.c_int32)]


class OrcData(C.Structure):
    _fields_ = [("owner_device", C.c_int32), ("preferred_device", C.c_int32), ("nb_devices", C.c_int32),
                ("new_data", C.c_int32), ("copy", OrcCopy * 16)]


class SelDev(C.Structure):
    _fields_ = [("is_gpu", C.c_int32), ("is_recursive", C.c_int32), ("enabled", C.c_int32),
                ("device_load", C.c_int64), ("time_estimate", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} missing: run `make -C oracle`")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.orc_run_window.argtypes = [vp, C.c_int32, vp, C.c_int32, vp, C.c_int32, vp, C.c_int32, vp, vp, vp, vp, vp, C.POINTER(OrcStats)]
        L.orc_run_window.restype = C.c_int
        L.orc_cpu_sched_run.argtypes = [vp, C.c_int32, vp, C.c_int32, vp, C.c_int32, vp, C.c_int32, C.c_int, vp, C.POINTER(C.c_uint64)]
        L.orc_cpu_sched_run.restype = C.c_double
        L.orc_twodbc_init.argtypes = [C.POINTER(TwoDBC)]
        L.orc_twodbc_rank_of.argtypes = [C.POINTER(TwoDBC), C.c_int, C.c_int]
        L.orc_twodbc_rank_of.restype = C.c_uint32
        L.orc_twodbc_position.argtypes = [C.POINTER(TwoDBC), C.c_int, C.c_int]
        L.orc_twodbc_key.argtypes = [C.POINTER(TwoDBC), C.c_int, C.c_int]
        L.orc_twodbc_key.restype = C.c_uint64
        L.orc_twodbc_key2coords.argtypes = [C.POINTER(TwoDBC), C.c_uint64, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_twodbc_tile_offset_elems.argtypes = [C.POINTER(TwoDBC), C.c_int, C.c_int]
        L.orc_twodbc_tile_offset_elems.restype = C.c_int64
        L.orc_rnd64_jump.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_rnd64_jump.restype = C.c_uint64
        L.orc_rnd64_step.argtypes = [C.c_uint64]
        L.orc_rnd64_step.restype = C.c_uint64
        L.orc_lcg_tile.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32]
        L.orc_zone_init.argtypes = [C.c_int, C.c_size_t]
        L.orc_zone_init.restype = vp
        L.orc_zone_fini.argtypes = [vp]
        L.orc_zone_malloc.argtypes = [vp, C.c_size_t]
        L.orc_zone_free.argtypes = [vp, C.c_int]
        L.orc_zone_in_use.argtypes = [vp]
        L.orc_zone_in_use.restype = C.c_size_t
        L.orc_zone_free_profile.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_data_create.argtypes = [C.POINTER(OrcData), C.c_int, C.c_int]
        L.orc_data_start_transfer_ownership.argtypes = [C.POINTER(OrcData), C.c_int, C.c_int]
        L.orc_data_end_transfer_ownership.argtypes = [C.POINTER(OrcData), C.c_int, C.c_int]
        L.orc_gpu_stage_in.argtypes = [C.POINTER(OrcData), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.orc_gpu_stage_in_complete.argtypes = [C.POINTER(OrcData), C.c_int, C.c_int]
        L.orc_gpu_task_complete.argtypes = [C.POINTER(OrcData), C.c_int, C.c_int, C.c_int]
        L.orc_gpu_w2r_complete.argtypes = [C.POINTER(OrcData), C.c_int]
        L.orc_dtd_build.argtypes = [C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, vp]
        L.orc_select_best_device.argtypes = [C.POINTER(SelDev), C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def run_window(tasks, succ, tiles_spec, ready, host=None):
    """Run one window through the sequential oracle.

    tiles_spec: list/array of (bytes, state, version, src_kind, src_offset) or a TILE_DTYPE array whose
    dev_ptr/src_ptr are ignored: the oracle owns its "device" memory; `host` (numpy, modified in place by
    pushout) is the home of the tiles, tile i living at byte offset src_ptr[i] (taken as an OFFSET here).
    Returns dict(results..., stats, device=list of numpy byte arrays)."""
    L = lib()
    tasks = np.ascontiguousarray(tasks, dtype=TASK_DTYPE)
    succ = np.ascontiguousarray(succ, dtype=np.uint32)
    ready = np.ascontiguousarray(ready, dtype=np.int32)
    tiles = np.array(tiles_spec, dtype=TILE_DTYPE, copy=True)
    nt, n = len(tiles), len(tasks)
    dev = [np.zeros(max(int(b), 1), np.uint8) for b in tiles["bytes"]]
    host_u8 = host.view(np.uint8).reshape(-1) if host is not None else None
    for i in range(nt):
        off = int(tiles["src_ptr"][i])
        tiles["dev_ptr"][i] = dev[i].ctypes.data
        tiles["src_ptr"][i] = (host_u8.ctypes.data + off) if host_u8 is not None else 0
    out = {
        "retire_order": np.full(n, -1, np.int32), "start_seq": np.zeros(n, np.uint32), "end_seq": np.zeros(n, np.uint32),
        "seen_version": np.zeros((n, 4), np.uint32), "result": np.zeros(n, np.uint64),
    }
    st = OrcStats()
    rc = L.orc_run_window(_p(tasks), n, _p(succ), len(succ), _p(tiles), nt, _p(ready), len(ready),
                          _p(out["retire_order"]), _p(out["start_seq"]), _p(out["end_seq"]),
                          _p(out["seen_version"]), _p(out["result"]), C.byref(st))
    out["rc"] = rc
    out["stats"] = {f[0]: getattr(st, f[0]) for f in OrcStats._fields_}
    out["tiles"] = tiles
    out["device"] = dev
    return out


def run_window_raw(tasks, succ, tiles, ready, policy=0, seed=1):
    """orc_run_window on descriptors whose dev_ptr / src_ptr are real host addresses owned by the caller (several
    descriptors may alias one buffer: that is how a window split over ranks is replayed as one merged DAG)."""
    L = lib()
    tasks = np.ascontiguousarray(tasks, dtype=TASK_DTYPE)
    succ = np.ascontiguousarray(succ, dtype=np.uint32)
    ready = np.ascontiguousarray(ready, dtype=np.int32)
    tiles = np.array(tiles, dtype=TILE_DTYPE, copy=True)
    n = len(tasks)
    out = {
        "retire_order": np.full(n, -1, np.int32), "start_seq": np.zeros(n, np.uint32), "end_seq": np.zeros(n, np.uint32),
        "seen_version": np.zeros((n, 4), np.uint32), "result": np.zeros(n, np.uint64),
    }
    st = OrcStats()
    L.orc_set_policy(policy, seed)
    try:
        rc = L.orc_run_window(_p(tasks), n, _p(succ), len(succ), _p(tiles), len(tiles), _p(ready), len(ready),
                              _p(out["retire_order"]), _p(out["start_seq"]), _p(out["end_seq"]),
                              _p(out["seen_version"]), _p(out["result"]), C.byref(st))
    finally:
        L.orc_set_policy(0, 1)
    out["rc"] = rc
    out["stats"] = {f[0]: getattr(st, f[0]) for f in OrcStats._fields_}
    out["tiles"] = tiles
    return out


def cpu_sched_run(tasks, succ, tiles, ready, nthreads):
    """Multi-threaded CPU scheduler port (the CPU baseline).  tiles[].dev_ptr must be host pointers."""
    L = lib()
    tasks = np.ascontiguousarray(tasks, dtype=TASK_DTYPE)
    succ = np.ascontiguousarray(succ, dtype=np.uint32)
    ready = np.ascontiguousarray(ready, dtype=np.int32)
    tiles = np.ascontiguousarray(tiles, dtype=TILE_DTYPE)
    per_thread = np.zeros(nthreads, np.uint64)
    errs = C.c_uint64(0)
    secs = L.orc_cpu_sched_run(_p(tasks), len(tasks), _p(succ), len(succ), _p(tiles), len(tiles),
                               _p(ready), len(ready), nthreads, _p(per_thread), C.byref(errs))
    return secs, per_thread, errs.value


def twodbc(myrank=0, mb=1, nb=1, lm=1, ln=1, i=0, j=0, m=None, n=None, P=1, Q=1, kp=1, kq=1, ip=0, jq=0):
    d = TwoDBC()
    d.myrank, d.mb, d.nb, d.lm, d.ln, d.i, d.j = myrank, mb, nb, lm, ln, i, j
    d.m, d.n = (lm if m is None else m), (ln if n is None else n)
    d.P, d.Q, d.kp, d.kq, d.ip, d.jq = P, Q, kp, kq, ip, jq
    lib().orc_twodbc_init(C.byref(d))
    return d


def dtd_build(nb_flows, flow_tile, flow_op, ntiles):
    """flow_tile/flow_op: int32 [ntasks, 4].  Returns (src, dst, flow, dep_count)."""
    L = lib()
    nb_flows = np.ascontiguousarray(nb_flows, np.int32)
    flow_tile = np.ascontiguousarray(flow_tile, np.int32)
    flow_op = np.ascontiguousarray(flow_op, np.int32)
    n = len(nb_flows)
    cap = max(16, 8 * n * 4)
    while True:
        src, dst, fl = np.empty(cap, np.int32), np.empty(cap, np.int32), np.empty(cap, np.int32)
        dep = np.zeros(n, np.int32)
        ne = L.orc_dtd_build(n, _p(nb_flows), _p(flow_tile), _p(flow_op), ntiles, _p(src), _p(dst), _p(fl), cap, _p(dep))
        if ne >= 0:
            return src[:ne].copy(), dst[:ne].copy(), fl[:ne].copy(), dep
        cap *= 4


_ref_twodbc = None


def ref_twodbc():
    """The reference's own two_dim_rectangle_cyclic.c / grid_2Dcyclic.c / matrix.c, built by Makefile.ref."""
    global _ref_twodbc
    if _ref_twodbc is None:
        L = C.CDLL(REF_TWODBC_PATH)
        L.ref_twodbc_new.restype = C.c_void_p
        L.ref_twodbc_new.argtypes = [C.c_int] * 15
        L.ref_twodbc_free.argtypes = [C.c_void_p]
        L.ref_twodbc_rank_of.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_twodbc_rank_of.restype = C.c_uint32
        L.ref_twodbc_key.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_twodbc_key.restype = C.c_uint64
        L.ref_twodbc_rank_of_key.argtypes = [C.c_void_p, C.c_uint64]
        L.ref_twodbc_rank_of_key.restype = C.c_uint32
        L.ref_twodbc_info.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        _ref_twodbc = L
    return _ref_twodbc


_ref_select = None


def ref_select():
    """The reference's own parsec/mca/device/device.c (parsec_select_best_device), built by Makefile.ref."""
    global _ref_select
    if _ref_select is None:
        L = C.CDLL(REF_SELECT_PATH)
        L.ref_sel_init.argtypes = [C.c_int, C.c_int]
        L.ref_sel_add_device.argtypes = [C.c_int, C.c_int64, C.c_int64]
        L.ref_sel_set_load.argtypes = [C.c_int, C.c_int64, C.c_int64]
        L.ref_sel_select.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.POINTER(C.c_int64)]
        _ref_select = L
    return _ref_select
