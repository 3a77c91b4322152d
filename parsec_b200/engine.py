"""Thin Python mirror of the L0 engine C ABI (include/pb2_engine.h).

Everything here is plumbing around ``libparsec_b200.so``: arrays are numpy structured arrays with
the exact C layouts, pointers are plain integers.  No compute happens in Python.
"""
import ctypes as C

import numpy as np

from . import _lib as L


def _check(rc, what, engine=None):
    if rc != L.PB2_SUCCESS:
        detail = ""
        if engine is not None and engine._h:
            detail = (L.load().pb2_engine_last_error(engine._h) or b"").decode()
        raise L.Pb2Error(rc, what, detail)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    """One engine per GPU (the reference's parsec_device_cuda_module_t, device_cuda.h:43-48)."""

    def __init__(self, cuda_device=0, workers_per_sm=0, threads=0, max_workers=0, stage_mode=0,
                 queue_policy=0, timeout_ms=0, gemm_mode=0, part_bytes=0):
        self._lib = L.load()
        self._h = C.c_void_p()
        p = L.EngineParams(workers_per_sm, threads, max_workers, stage_mode, queue_policy, timeout_ms, gemm_mode, part_bytes)
        rc = self._lib.pb2_engine_create(C.byref(self._h), cuda_device, C.byref(p))
        if rc != L.PB2_SUCCESS:
            self._h = C.c_void_p()
            raise L.Pb2Error(rc, "pb2_engine_create")
        self._allocs = []

    def info(self):
        i = L.EngineInfo()
        _check(self._lib.pb2_engine_info(self._h, C.byref(i)), "pb2_engine_info", self)
        return {f[0]: getattr(i, f[0]) for f in L.EngineInfo._fields_ if f[0] != "reserved"}

    def malloc(self, nbytes):
        p = C.c_void_p()
        _check(self._lib.pb2_engine_malloc(self._h, nbytes, C.byref(p)), "pb2_engine_malloc", self)
        self._allocs.append(p.value)
        return p.value

    def free(self, ptr):
        _check(self._lib.pb2_engine_free(self._h, C.c_void_p(ptr)), "pb2_engine_free", self)
        if ptr in self._allocs:
            self._allocs.remove(ptr)

    def host_register(self, arr):
        """cudaHostRegister a numpy array (memory_register, device_cuda_module.c:183); returns the device alias."""
        alias = C.c_void_p()
        _check(self._lib.pb2_engine_host_register(self._h, _ptr(arr), arr.nbytes, C.byref(alias)),
               "pb2_engine_host_register", self)
        return alias.value

    def host_unregister(self, arr):
        _check(self._lib.pb2_engine_host_unregister(self._h, _ptr(arr)), "pb2_engine_host_unregister", self)

    def h2d(self, dev_ptr, arr):
        arr = np.ascontiguousarray(arr)
        _check(self._lib.pb2_engine_memcpy_h2d(self._h, C.c_void_p(dev_ptr), _ptr(arr), arr.nbytes), "h2d", self)
        self.synchronize()

    def d2h(self, arr, dev_ptr):
        assert arr.flags["C_CONTIGUOUS"]
        _check(self._lib.pb2_engine_memcpy_d2h(self._h, _ptr(arr), C.c_void_p(dev_ptr), arr.nbytes), "d2h", self)
        return arr

    def use_stream(self, cuda_stream):
        """Enqueue engine work on a caller-owned stream (int cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream)."""
        if not cuda_stream:
            raise ValueError("use_stream needs a real stream handle (the legacy default stream is 0: create a torch.cuda.Stream)")
        _check(self._lib.pb2_engine_set_stream(self._h, C.c_void_p(cuda_stream)), "pb2_engine_set_stream", self)

    def set_shared_windows(self, on=True, next_rs_begin=None):
        """next_rs_begin: int32 remote out-edge CSR of the NEXT window (kept alive by the caller until it is created)."""
        self._rs_keep = None if next_rs_begin is None else np.ascontiguousarray(next_rs_begin, np.int32)
        _check(self._lib.pb2_engine_set_shared_windows(self._h, 1 if on else 0, _ptr(self._rs_keep)), "set_shared_windows", self)

    def set_part_bytes(self, part_bytes):
        _check(self._lib.pb2_engine_set_part_bytes(self._h, part_bytes), "set_part_bytes", self)

    def ipc_export(self, dev_ptr):
        h = (C.c_ubyte * 64)()
        _check(self._lib.pb2_engine_ipc_export(self._h, C.c_void_p(dev_ptr), h), "ipc_export", self)
        return bytes(h)

    def ipc_open(self, handle):
        h = (C.c_ubyte * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        _check(self._lib.pb2_engine_ipc_open(self._h, h, C.byref(p)), "ipc_open", self)
        return p.value

    def synchronize(self):
        _check(self._lib.pb2_engine_synchronize(self._h), "pb2_engine_synchronize", self)

    def window(self, kind, tasks, succ, tiles, ready):
        return Window(self, kind, tasks, succ, tiles, ready)

    def close(self):
        if self._h:
            for p in list(self._allocs):
                self._lib.pb2_engine_free(self._h, C.c_void_p(p))
            self._allocs = []
            self._lib.pb2_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Window:
    """One DAG window resident on the GPU: create once, launch/wait any number of times."""

    def __init__(self, engine, kind, tasks, succ, tiles, ready):
        self.engine = engine
        self._lib = engine._lib
        tasks = np.ascontiguousarray(tasks, dtype=L.TASK_DTYPE)
        succ = np.ascontiguousarray(succ, dtype=np.uint32)
        tiles = np.ascontiguousarray(tiles, dtype=L.TILE_DTYPE)
        ready = np.ascontiguousarray(ready, dtype=np.int32)
        self.ntasks, self.ntiles = len(tasks), len(tiles)
        self._h = C.c_void_p()
        rc = self._lib.pb2_window_create(engine._h, C.byref(self._h), kind,
                                         _ptr(tasks), len(tasks), _ptr(succ), len(succ),
                                         _ptr(tiles), len(tiles), _ptr(ready), len(ready))
        if rc != L.PB2_SUCCESS:
            self._h = C.c_void_p()
        _check(rc, "pb2_window_create", engine)

    def launch(self):
        _check(self._lib.pb2_window_launch(self._h), "pb2_window_launch", self.engine)

    def arm(self):
        _check(self._lib.pb2_window_arm(self._h), "pb2_window_arm", self.engine)

    def start(self):
        _check(self._lib.pb2_window_start(self._h), "pb2_window_start", self.engine)

    def export(self):
        h = L.WindowHandle()
        _check(self._lib.pb2_window_export(self._h, C.byref(h)), "pb2_window_export", self.engine)
        return bytes(h)

    def task_entries(self):
        out = np.empty(self.ntasks, np.int32)
        _check(self._lib.pb2_window_task_entries(self._h, _ptr(out)), "pb2_window_task_entries", self.engine)
        return out

    def set_push(self, ps_begin, push):
        """Producer-side pushes (pb2_window_set_push): call after set_remote."""
        ps_begin = np.ascontiguousarray(ps_begin, np.int32)
        push = np.ascontiguousarray(push, L.PUSH_DTYPE)
        _check(self._lib.pb2_window_set_push(self._h, _ptr(ps_begin), _ptr(push), len(push)), "pb2_window_set_push", self.engine)

    def set_remote(self, my_rank, handles, rs_begin, rs_rank, rs_target):
        """handles: list of bytes (one exported WindowHandle per rank)."""
        arr = (L.WindowHandle * len(handles))(*[L.WindowHandle.from_buffer_copy(h) for h in handles])
        rs_begin = np.ascontiguousarray(rs_begin, np.int32)
        rs_rank = np.ascontiguousarray(rs_rank, np.int32)
        rs_target = np.ascontiguousarray(rs_target, np.uint32)
        _check(self._lib.pb2_window_set_remote(self._h, my_rank, len(handles), arr, _ptr(rs_begin), _ptr(rs_rank),
                                               _ptr(rs_target), len(rs_rank)), "pb2_window_set_remote", self.engine)

    def wait(self):
        st = L.WindowStats()
        rc = self._lib.pb2_window_wait(self._h, C.byref(st))
        self.stats = {f[0]: getattr(st, f[0]) for f in L.WindowStats._fields_}
        _check(rc, "pb2_window_wait", self.engine)
        return self.stats

    def run(self):
        self.launch()
        return self.wait()

    def results(self):
        n = self.ntasks
        out = {
            "retire_order": np.empty(n, np.int32), "start_seq": np.empty(n, np.uint32),
            "end_seq": np.empty(n, np.uint32), "seen_version": np.empty((n, L.MAX_FLOWS), np.uint32),
            "result": np.empty(n, np.uint64), "worker": np.empty(n, np.int32),
            "tiles": np.empty(self.ntiles, L.TILE_DTYPE),
        }
        _check(self._lib.pb2_window_results(self._h, _ptr(out["retire_order"]), _ptr(out["start_seq"]),
                                            _ptr(out["end_seq"]), _ptr(out["seen_version"]),
                                            _ptr(out["result"]), _ptr(out["worker"]), _ptr(out["tiles"])),
               "pb2_window_results", self.engine)
        return out

    def close(self):
        if self._h:
            self._lib.pb2_window_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
