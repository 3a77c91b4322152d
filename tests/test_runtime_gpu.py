"""GPU parity tests through the reference-shaped host API (C ABI of include/pb2_parsec.h), checked against the
oracle and the known answers of the reference's own GPU tests."""
import ctypes as C

import numpy as np
import pytest

from oracle import orc, orc_dags as dags
from parsec_b200 import _lib as L
from parsec_b200 import runtime as R
from parsec_b200.bf16 import bf16_bits_to_f32, f32_to_bf16_bits, round_to_bf16

pytestmark = pytest.mark.gpu


def test_ex05_through_ptg_front_end():
    """Ex05_Broadcast.jdf on the device module: every TaskRecv observes k; one H2D per tile; all successors
    released on the device; host-visible versions / coherency as the reference leaves them."""
    K, NB, tb = 64, 14, 256 * 256 * 4
    host = np.full(K * tb // 4, -3, np.int32)
    with R.Context(cuda_devices=(0,)) as ctx:
        dc = ctx.block_cyclic(4, tb // 4, 1, K * tb // 4, 1, mat=host)
        tp = C.c_void_p(ctx.l.pb2_ptg_ex05_broadcast_new(ctx.h, dc, K, NB))
        ctx.wait()
        F = NB // 2 + 1
        info = ctx.task_info(tp)
        recv = info["class_id"] == 1
        assert np.all((info["result"][recv] >> np.uint64(32)) == 0)
        assert np.array_equal(info["result"][recv] & np.uint64(0xFFFFFFFF), info["locals"][recv, 0].astype(np.uint64))
        assert np.all(info["seen_version"][~recv, 0] == 0) and np.all(info["seen_version"][recv, 0] == 1)
        st = ctx.stats(ctx.devices[0])
        assert st["executed_tasks"] == K * (1 + F) and st["windows_launched"] == 1
        assert st["tasks_released_on_device"] == K * F
        assert st["data_in_from_device"][0] == K * tb == st["nb_data_faults"]
        for k in (0, K - 1):
            d = C.c_void_p(ctx.l.pb2_dc_data_of(dc, k, 0))
            g = ctx.copy_state(d, 2)
            assert (g["coherency"], g["version"], g["readers"]) == (R.COHERENCY_OWNED, 1, 0)
            assert ctx.l.pb2_data_owner_device(d) == 2
        # flush: the dirty replicas go home (parsec_device_flush_lru), host memory then holds k
        assert ctx.l.pb2_device_memory_release(ctx.devices[0]) == 0
        assert np.array_equal(host.reshape(K, -1)[:, 0], np.arange(K)) and np.all(host.reshape(K, -1)[:, -1] == np.arange(K))


@pytest.mark.parametrize("NB_TOKEN", [1, 8])
def test_ptg_pingpong(NB_TOKEN):
    """ptg_pingpong.jdf:144-149: CPU(+i) -> GPU(+i) -> GPU ... every element ends at 3*i."""
    with R.Context(cuda_devices=(0,)) as ctx:
        nb_err = C.c_int32(-1)
        tp = C.c_void_p(ctx.l.pb2_ptg_pingpong_new(ctx.h, None, NB_TOKEN, C.byref(nb_err)))
        ctx.wait()
        assert nb_err.value == 0
        t, dev = ctx.trace(tp)
        info = ctx.task_info(tp)
        assert np.all(dev[np.isin(t, np.where(info["class_id"] == 2)[0])] == 2)     # TOKEN_GPU ran on the GPU
        assert np.all(dev[np.isin(t, np.where(info["class_id"] == 1)[0])] == 0)     # TOKEN_CPU on the host
        assert ctx.stats(ctx.devices[0])["executed_tasks"] == 2 * NB_TOKEN


def test_get_best_device_two_modules():
    """get_best_device_check.jdf: placement (n*nt+m) % ngpu and B == 0x01010101 (two modules on one GPU)."""
    nt, ngpu, mb = 4, 2, 16
    with R.Context(cuda_devices=(0, 0), mca={"device_cuda_memory_number_of_blocks": 64}) as ctx:
        host = np.zeros(nt * nt * mb * mb, np.float64)
        dc = ctx.block_cyclic(8, mb, mb, nt * mb, nt * mb, mat=host)
        n = nt * (nt + 1) // 2
        info = np.full(n + 1, -1, np.int32)
        tp = C.c_void_p(ctx.l.pb2_ptg_get_best_device_new(ctx.h, dc, info.ctypes.data_as(C.c_void_p)))
        ctx.wait()
        idx = 0
        for m in range(nt):
            for nn in range(m + 1):
                assert info[idx] == 2 + (nn * nt + m) % ngpu
                idx += 1
        assert info[n] == 0


def test_dtd_simple_gemm_app():
    """dtd_test_simple_gemm.c simple_gemm(): NT^3 GEMM tasks through the DTD front end; C equals the oracle's
    per-task bf16 chain within 2 bf16 ulps of the running magnitude (accumulation order differs)."""
    NT, T = 3, 256
    rng = np.random.default_rng(1901)
    mats = [round_to_bf16(rng.uniform(-0.5, 0.5, (NT * NT, T, T)).astype(np.float32)) for _ in range(3)]
    bits = [f32_to_bf16_bits(m).reshape(-1).copy() for m in mats]
    with R.Context(cuda_devices=(0,)) as ctx:
        # TILE storage is column-major over tiles: tile (m, n) sits at position n*NT + m of the local array
        dcs = [ctx.block_cyclic(2, T, T, NT * T, NT * T, mat=b) for b in bits]
        secs = C.c_double()
        tp = C.c_void_p()
        assert ctx.l.pb2_app_dtd_simple_gemm(ctx.h, dcs[0], dcs[1], dcs[2], R.DEV_CUDA, C.byref(secs), C.byref(tp)) == 0
        st = ctx.stats(ctx.devices[0])
        assert st["executed_tasks"] == NT ** 3 and st["windows_launched"] == 1
        assert st["tasks_released_on_device"] == NT ** 3 - NT * NT
        info = ctx.task_info(tp)
        assert np.array_equal(info["seen_version"][:, 2], np.tile(np.arange(NT), NT * NT))
    A, B, Cm = mats
    pos = lambda m, n: n * NT + m
    got = bf16_bits_to_f32(bits[2]).reshape(NT * NT, T, T)
    for i in range(NT):
        for j in range(NT):
            c = Cm[pos(i, j)].copy(); mag = np.abs(c)
            for k in range(NT):
                c = round_to_bf16(c + A[pos(i, k)] @ B[pos(k, j)].T); mag = np.maximum(mag, np.abs(c))
            assert np.all(np.abs(got[pos(i, j)] - c) <= 2.0 ** -7 * np.maximum(mag, 1.0)), (i, j)


def test_eviction_and_write_back_under_memory_pressure():
    """Device heap smaller than the working set: the module cuts windows, evicts clean replicas and writes
    dirty ones back (transfer_gpu.c W2R); results stay exact and the statistics show it."""
    K, NB, tb = 40, 2, 64 * 1024
    host = np.zeros(K * tb // 4, np.int32)
    with R.Context(cuda_devices=(0,), mca={"device_cuda_memory_number_of_blocks": 12, "device_cuda_memory_block_size": tb}) as ctx:
        dc = ctx.block_cyclic(4, tb // 4, 1, K * tb // 4, 1, mat=host)
        tp = C.c_void_p(ctx.l.pb2_ptg_ex05_broadcast_new(ctx.h, dc, K, NB))
        ctx.wait()
        info = ctx.task_info(tp)
        recv = info["class_id"] == 1
        assert np.all((info["result"][recv] >> np.uint64(32)) == 0)
        st = ctx.stats(ctx.devices[0])
        assert st["executed_tasks"] == K * 3 and st["windows_launched"] > 1
        assert st["nb_evictions"] >= K - 12 and st["data_out_to_host"] >= (K - 12) * tb
        assert ctx.l.pb2_device_memory_release(ctx.devices[0]) == 0
        assert np.array_equal(host.reshape(K, -1)[:, 7], np.arange(K))


def test_rtt_chain_writes_back_to_the_last_owner_tile():
    """rtt.jdf:33: the final PING writes T to A(f, (NT-1) % WS); T == T0 + NT (config 4 on one GPU)."""
    NT, FRAGS, WS, mb = 10, 3, 4, 32
    tile = mb * mb
    host = np.arange(FRAGS * WS * tile, dtype=np.float32)
    t0 = host.copy().reshape(WS, FRAGS, tile)                  # tile (f, w) at position w*FRAGS + f
    with R.Context(cuda_devices=(0,)) as ctx:
        dc = ctx.block_cyclic(4, mb, mb, FRAGS * mb, WS * mb, mat=host)
        tp = C.c_void_p(ctx.l.pb2_ptg_rtt_new(ctx.h, dc, NT, FRAGS, WS))
        ctx.wait()
        out = host.reshape(WS, FRAGS, tile)
        for f in range(FRAGS):
            assert np.array_equal(out[(NT - 1) % WS, f], t0[0, f] + NT)
        assert ctx.stats(ctx.devices[0])["executed_tasks"] == NT * FRAGS


def test_dtd_two_rounds_and_new_tiles():
    """dtd_test_new_tile: init (i), x2, check 2*i on NEW tiles; a second round of inserts after a wait chains
    behind the completed tasks (parent not alive => take its output directly)."""
    nb = 1000
    with R.Context(cuda_devices=(0,)) as ctx:
        tp = C.c_void_p(ctx.l.pb2_dtd_taskpool_new(ctx.h))
        op = np.array([R.INOUT], np.int32)
        mk = lambda body: (lambda tc: (ctx.l.pb2_dtd_task_class_add_chore(tp, tc, R.DEV_CUDA, body, None), tc)[1])(
            C.c_void_p(ctx.l.pb2_dtd_create_task_class(tp, b"k", 1, op.ctypes.data_as(C.c_void_p))))
        init, mul, chk = mk(L.BODY_IOTA_I32), mk(L.BODY_SCALE_I32), mk(L.BODY_ADD_IOTA_I32)
        tiles = [C.c_void_p(ctx.l.pb2_dtd_tile_new(tp, nb * 4)) for _ in range(6)]
        two = np.array([2, 0, 0], np.int32)
        for t in tiles:
            arr = (C.c_void_p * 1)(t)
            ctx.l.pb2_dtd_insert_task_with_task_class(tp, init, 0, R.DEV_CUDA, arr, np.array([R.OUTPUT], np.int32).ctypes.data_as(C.c_void_p), None, 0.0)
            ctx.l.pb2_dtd_insert_task_with_task_class(tp, mul, 0, R.DEV_CUDA, arr, None, two.ctypes.data_as(C.c_void_p), 0.0)
        ctx.wait()
        for t in tiles:                                         # second round: += i, pushed out to the host
            arr = (C.c_void_p * 1)(t)
            ctx.l.pb2_dtd_insert_task_with_task_class(tp, chk, 0, R.DEV_CUDA, arr, np.array([R.INOUT | R.PUSHOUT], np.int32).ctypes.data_as(C.c_void_p), None, 0.0)
        ctx.wait()
        for t in tiles:
            d = C.c_void_p(ctx.l.pb2_dtd_tile_data(t))
            hc = ctx.copy_state(d, 0)
            assert hc["version"] == 3
            cp = C.cast(ctx.l.pb2_data_get_copy(d, 0), C.POINTER(C.c_char))
        assert ctx.stats(ctx.devices[0])["executed_tasks"] == 18 and ctx.stats(ctx.devices[0])["windows_launched"] == 2


def test_cholesky_shaped_dag_completes():
    """Config-5 shape on one GPU at small NT: every task runs exactly once in a dependency-respecting order."""
    NT, nb = 5, 128
    bits = f32_to_bf16_bits(np.full(NT * NT * nb * nb, 0.001, np.float32))
    with R.Context(cuda_devices=(0,)) as ctx:
        dc = ctx.block_cyclic(2, nb, nb, NT * nb, NT * nb, mat=bits)
        tp = C.c_void_p(ctx.l.pb2_ptg_cholesky_shape_new(ctx.h, dc, NT))
        n = ctx.l.pb2_taskpool_nb_tasks(tp)
        assert n == NT + 2 * (NT * (NT - 1) // 2) + NT * (NT - 1) * (NT - 2) // 6
        ctx.wait()
        t, dev = ctx.trace(tp)
        assert sorted(t.tolist()) == list(range(n)) and np.all(dev == 2)
        assert np.all(np.isfinite(bf16_bits_to_f32(bits)))


def _cudart():
    import ctypes.util
    for name in ("libcudart.so.12", "libcudart.so", ctypes.util.find_library("cudart") or ""):
        try:
            if name:
                return C.CDLL(name)
        except OSError:
            continue
    import glob, os, torch                                            # torch ships its own runtime
    for p in glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*")):
        return C.CDLL(p)
    raise RuntimeError("libcudart not found")


def test_user_submit_chore_runs_between_engine_windows():
    """A GPU chore that is an opaque user `submit(device, gpu_task, stream)` function (device_gpu.h:49-51, the kind of
    body dtd_test_simple_gemm.c and every generated BODY [type=CUDA] use): the host lane stages the flows in, calls
    it on the stream, and the task chains with in-engine bodies before and after it on the same tiles.
      tile: FILL 7 (engine) -> user memset 0x01 per byte (submit) -> CHECK == 0x01010101 (engine), then flushed home."""
    import torch  # noqa: F401  (loads the CUDA runtime this process uses)
    rt = _cudart()
    rt.cudaMemsetAsync.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]
    nb, ntiles = 4096, 5
    calls = []
    with R.Context(cuda_devices=(0,)) as ctx:
        def submit(dev, gtask, stream):
            ptr = ctx.l.pb2_gpu_task_flow_ptr(dev, gtask, 0)
            nbytes = ctx.l.pb2_gpu_task_flow_bytes(gtask, 0)
            val = ctx.l.pb2_gpu_task_iparam(gtask)[0]
            calls.append((ptr, nbytes, val))
            return 0 if rt.cudaMemsetAsync(ptr, val, nbytes, stream) == 0 else -5
        cb = R.GPU_SUBMIT(submit)
        tp = C.c_void_p(ctx.l.pb2_dtd_taskpool_new(ctx.h))
        op = np.array([R.INOUT], np.int32)
        new_class = lambda: C.c_void_p(ctx.l.pb2_dtd_create_task_class(tp, b"k", 1, op.ctypes.data_as(C.c_void_p)))
        fill, user, chk = new_class(), new_class(), new_class()
        assert ctx.l.pb2_dtd_task_class_add_chore(tp, fill, R.DEV_CUDA, L.BODY_FILL_I32, None) == 0
        assert ctx.l.pb2_dtd_task_class_add_submit(tp, user, cb) == 0
        assert ctx.l.pb2_dtd_task_class_add_chore(tp, chk, R.DEV_CUDA, L.BODY_CHECK_I32, None) == 0
        tiles = [C.c_void_p(ctx.l.pb2_dtd_tile_new(tp, nb * 4)) for _ in range(ntiles)]
        p = lambda *v: np.array(v, np.int32).ctypes.data_as(C.c_void_p)
        ids = []
        for t in tiles:
            arr = (C.c_void_p * 1)(t)
            ctx.l.pb2_dtd_insert_task_with_task_class(tp, fill, 0, R.DEV_CUDA, arr, p(R.OUTPUT), p(7, 0, 0), 0.0)
            ctx.l.pb2_dtd_insert_task_with_task_class(tp, user, 0, R.DEV_CUDA, arr, p(R.INOUT), p(1, 0, 0), 0.0)
            ids.append(ctx.l.pb2_dtd_insert_task_with_task_class(tp, chk, 0, R.DEV_CUDA, arr, p(R.INPUT), p(0x01010101, 0, 0), 0.0))
        ctx.wait()
        info = ctx.task_info(tp)
        st = ctx.stats(ctx.devices[0])
        assert st["executed_tasks"] == 3 * ntiles
        assert len(calls) == ntiles and all(c[1] == nb * 4 and c[2] == 1 for c in calls)
        for i in ids:                                               # the engine's CHECK saw the user kernel's bytes
            assert info["result"][i] == 0x01010101, hex(int(info["result"][i]))
            assert info["seen_version"][i, 0] == 2                  # FILL -> 1, user submit -> 2
        assert np.all(info["seen_version"][[i - 1 for i in ids], 0] == 1)
        assert st["windows_launched"] == 3                          # engine window, submit lane, engine window
