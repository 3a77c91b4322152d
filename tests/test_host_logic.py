"""Host-side logic (no GPU): C-ABI exports, 2D block-cyclic map, device heap, DTD / PTG graph construction and
window building, device selection, coherency replay -- each checked against the oracle."""
import ctypes as C
import random
import re

import numpy as np
import pytest

from oracle import orc, orc_dags as dags
from parsec_b200 import _lib as L
from parsec_b200 import runtime as R


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads without a GPU and exports every symbol include/*.h declares."""
    lib = L.load()
    declared = set()
    for hdr in ("include/pb2_engine.h", "include/pb2_parsec.h", "include/pb2_stream.h"):
        txt = open(hdr).read()
        declared |= set(re.findall(r"\b(pb2_[a-z0-9_]+)\s*\(", txt))
    declared -= {"pb2_cpu_hook_t"}
    assert declared, "no declarations found"
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    from parsec_b200.stream import STREAM_SYMBOLS
    assert set(L.ENGINE_SYMBOLS) | set(R.PARSEC_SYMBOLS) | set(STREAM_SYMBOLS) >= declared


def test_no_gpu_means_loud_failure():
    """Without a CUDA device the product path refuses to run (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from parsec_b200.engine import Engine
    with pytest.raises(L.Pb2Error) as ei:
        Engine(0)
    assert ei.value.rc == L.PB2_ERR_DEVICE
    with pytest.raises(L.Pb2Error):
        R.Context(cuda_devices=(0,), dry_run=False)


@pytest.mark.parametrize("P,Q,kp,kq,ip,jq", [(1, 1, 1, 1, 0, 0), (2, 2, 1, 1, 0, 0), (2, 4, 1, 1, 1, 3), (2, 3, 2, 3, 0, 0), (3, 2, 2, 2, 1, 1)])
def test_block_cyclic_matches_oracle(P, Q, kp, kq, ip, jq):
    mb, nb, lm, ln = 4, 6, 4 * 11 + 1, 6 * 7
    with R.Context(cuda_devices=(), dry_run=True) as ctx:
        for rank in range(P * Q):
            dc = ctx.block_cyclic(4, mb, nb, lm, ln, P=P, Q=Q, myrank=rank, kp=kp, kq=kq, ip=ip, jq=jq)
            o = orc.twodbc(rank, mb, nb, lm, ln, P=P, Q=Q, kp=kp, kq=kq, ip=ip, jq=jq)
            info = (C.c_int64 * 8)()
            ctx.l.pb2_dc_info(dc, info)
            assert list(info)[:7] == [o.lmt, o.lnt, o.mt, o.nt, o.nb_elem_r, o.nb_elem_c, o.nb_local_tiles]
            for m in range(o.mt):
                for n in range(o.nt):
                    assert ctx.l.pb2_dc_rank_of(dc, m, n) == orc.lib().orc_twodbc_rank_of(C.byref(o), m, n)
                    assert ctx.l.pb2_dc_position(dc, m, n) == orc.lib().orc_twodbc_position(C.byref(o), m, n)
                    assert ctx.l.pb2_dc_data_key(dc, m, n) == orc.lib().orc_twodbc_key(C.byref(o), m, n)
            ctx.l.pb2_data_collection_free(dc)


def test_device_heap_matches_oracle():
    """The module's slot allocator makes the same decisions as the reference's zone_malloc (via the oracle that
    is itself pinned to the real zone_malloc.c build)."""
    with R.Context(cuda_devices=(0,), dry_run=True, mca={"device_cuda_memory_number_of_blocks": 96, "device_cuda_memory_block_size": 512}) as ctx:
        dev = ctx.devices[0]
        O = orc.lib()
        base = None
        for seed in range(10):
            rnd = random.Random(seed)
            zo = O.orc_zone_init(96, 512)
            live = []
            for _ in range(1200):
                if live and rnd.random() < 0.45:
                    ptr, tid = live.pop(rnd.randrange(len(live)))
                    assert ctx.l.pb2_device_zone_free(dev, ptr) == 0 and O.orc_zone_free(zo, tid) == 0
                else:
                    size = rnd.choice([1, 256, 512, 513, 1024, 2000, 4096, 8192])
                    ptr, tid = ctx.l.pb2_device_zone_malloc(dev, size), O.orc_zone_malloc(zo, size)
                    assert (ptr is None) == (tid < 0)
                    if ptr is not None:
                        if base is None:
                            base = ptr - tid * 512
                        assert (ptr - base) // 512 == tid
                        live.append((ptr, tid))
                assert ctx.l.pb2_device_zone_in_use(dev) == O.orc_zone_in_use(zo)
            for ptr, tid in live:
                ctx.l.pb2_device_zone_free(dev, ptr)
            assert ctx.l.pb2_device_zone_in_use(dev) == 0
            assert ctx.l.pb2_device_zone_free(dev, base) == L.PB2_ERR_EXISTS      # double free is reported
            O.orc_zone_fini(zo)


def _window_equal(win, dag):
    """The exported window, with tasks mapped back to taskpool order, equals the oracle's graph."""
    ids = win["task_ids"]
    n = len(ids)
    assert n == dag.ntasks and sorted(ids.tolist()) == list(range(n))
    inv = np.empty(n, np.int64); inv[np.arange(n)] = ids          # window index -> pool id
    for fld in ("body", "nb_flows", "flags", "dep_goal", "priority"):
        assert np.array_equal(win["tasks"][fld], dag.tasks[fld][ids]), fld
    assert np.array_equal(win["tasks"]["iparam"], dag.tasks["iparam"][ids])
    assert np.array_equal(win["tasks"]["access"], dag.tasks["access"][ids])
    cnt = win["tasks"]["succ_count"].astype(np.int64)
    src = np.repeat(ids.astype(np.int64), cnt)
    dst = ids[(win["succ"] & np.uint32(0x07FFFFFF)).astype(np.int64)].astype(np.int64)
    fl = (win["succ"] >> np.uint32(27)).astype(np.int64)
    es, ed, ef = dag.edges()
    assert sorted(zip(src.tolist(), dst.tolist(), fl.tolist())) == sorted(zip(es.tolist(), ed.tolist(), ef.tolist()))
    assert sorted(ids[win["ready"]].tolist()) == sorted(dag.ready.tolist())


def test_ptg_ex05_window_equals_oracle_graph():
    K, NB, tb = 12, 6, 1024
    dag = dags.ex05_broadcast(K, NB, tb)
    host = np.zeros(K * tb // 4, np.int32)
    with R.Context(cuda_devices=(0,), dry_run=True) as ctx:
        dc = ctx.block_cyclic(4, tb // 4, 1, K * tb // 4, 1, mat=host)          # K tiles of tb bytes, 1-D
        tp = C.c_void_p(ctx.l.pb2_ptg_ex05_broadcast_new(ctx.h, dc, K, NB))
        win = ctx.export_window(tp, ctx.devices[0])
        _window_equal(win, dag)
        # tiles: TaskBcast(k) and its receivers share tile k, all INVALID with the host alias as source
        assert len(win["tiles"]) == K and np.all(win["tiles"]["state"] == L.TILE_INVALID)
        assert np.all(win["tiles"]["bytes"] == tb)
        t = win["tasks"]
        for i in range(len(t)):
            k = t["iparam"][i, 0]
            assert win["tiles"]["src_ptr"][t["tile"][i, 0]] == host.ctypes.data + k * tb


def test_ptg_chain_rtt_ep_windows_equal_oracle_graphs():
    with R.Context(cuda_devices=(0,), dry_run=True) as ctx:
        tp = C.c_void_p(ctx.l.pb2_ptg_ex02_chain_new(ctx.h, 25))
        _window_equal(ctx.export_window(tp, ctx.devices[0]), dags.ex02_chain(25))
    with R.Context(cuda_devices=(0,), dry_run=True) as ctx:
        host = np.zeros(3 * 4 * 16, np.float32)
        dc = ctx.block_cyclic(4, 4, 4, 4 * 3, 4 * 4, P=1, Q=1, mat=host)        # FRAGS x WS tiles of 4x4 floats
        tp = C.c_void_p(ctx.l.pb2_ptg_rtt_new(ctx.h, dc, 9, 3, 4))
        _window_equal(ctx.export_window(tp, ctx.devices[0]), dags.rtt_chain(9, 3, 64))
    with R.Context(cuda_devices=(0,), dry_run=True) as ctx:
        tp = C.c_void_p(ctx.l.pb2_ptg_ep_new(ctx.h, None, 16, 4))
        _window_equal(ctx.export_window(tp, ctx.devices[0]), dags.ep(16, 4))


def test_dtd_gemm_window_equals_oracle_graph_and_rule():
    NT, T = 3, 64
    dag = dags.dtd_gemm(NT, T)
    with R.Context(cuda_devices=(0,), dry_run=True) as ctx:
        mats = [np.zeros(NT * NT * T * T, np.uint16) for _ in range(3)]
        dcs = [ctx.block_cyclic(2, T, T, NT * T, NT * T, mat=m) for m in mats]
        tp = C.c_void_p(ctx.l.pb2_dtd_taskpool_new(ctx.h))
        ops = np.array([R.INPUT, R.INPUT, R.INOUT | R.AFFINITY], np.int32)
        tc = C.c_void_p(ctx.l.pb2_dtd_create_task_class(tp, b"GEMM", 3, ops.ctypes.data_as(C.c_void_p)))
        assert ctx.l.pb2_dtd_task_class_add_chore(tp, tc, R.DEV_CUDA, L.BODY_GEMM_BF16, None) == 0
        dims = np.array([T, T, T], np.int32)
        for i in range(NT):
            for j in range(NT):
                for k in range(NT):
                    tiles = (C.c_void_p * 3)(ctx.l.pb2_dtd_tile_of(tp, dcs[0], ctx.l.pb2_dc_data_key(dcs[0], i, k)),
                                             ctx.l.pb2_dtd_tile_of(tp, dcs[1], ctx.l.pb2_dc_data_key(dcs[1], k, j)),
                                             ctx.l.pb2_dtd_tile_of(tp, dcs[2], ctx.l.pb2_dc_data_key(dcs[2], i, j)))
                    fo = np.array([R.INPUT, R.INPUT, (R.INOUT | R.PUSHOUT) if k == NT - 1 else R.INOUT], np.int32)
                    tid = ctx.l.pb2_dtd_insert_task_with_task_class(tp, tc, NT ** 3 - i * NT + j, R.DEV_CUDA, tiles,
                                                                    fo.ctypes.data_as(C.c_void_p), dims.ctypes.data_as(C.c_void_p), 0.0)
                    assert tid == (i * NT + j) * NT + k
        win = ctx.export_window(tp, ctx.devices[0])
        # tile numbering differs (first-touch order); compare everything else, then the tile structure
        ids = win["task_ids"]
        for fld in ("body", "nb_flows", "flags", "dep_goal", "priority"):
            assert np.array_equal(win["tasks"][fld], dag.tasks[fld][ids]), fld
        assert np.array_equal(win["tasks"]["access"], dag.tasks["access"][ids])
        cnt = win["tasks"]["succ_count"].astype(np.int64)
        src = np.repeat(ids.astype(np.int64), cnt)
        dst = ids[(win["succ"] & np.uint32(0x07FFFFFF)).astype(np.int64)]
        es, ed, _ = dag.edges()
        assert sorted(zip(src.tolist(), dst.tolist())) == sorted(zip(es.tolist(), ed.tolist()))
        # same tile sharing pattern: a bijection between window tile ids and oracle tile ids
        m = {}
        for wi, pid in enumerate(ids):
            for f in range(3):
                a, b = int(win["tasks"]["tile"][wi, f]), int(dag.tasks["tile"][pid, f])
                assert m.setdefault(a, b) == b
        assert len(set(m.values())) == len(m) == 3 * NT * NT


def test_dtd_rule_matches_oracle_on_random_programs():
    """Random insert_task programs: the front end builds exactly the edges of the oracle's DTD rule."""
    rnd = random.Random(7)
    for trial in range(20):
        ntiles, ntasks = 5, 40
        with R.Context(cuda_devices=(0,), dry_run=True) as ctx:
            tp = C.c_void_p(ctx.l.pb2_dtd_taskpool_new(ctx.h))
            tiles = [C.c_void_p(ctx.l.pb2_dtd_tile_new(tp, 64)) for _ in range(ntiles)]
            ops3 = np.array([R.INOUT, R.INOUT, R.INOUT], np.int32)
            tcs = {n: C.c_void_p(ctx.l.pb2_dtd_create_task_class(tp, b"t", n, ops3.ctypes.data_as(C.c_void_p))) for n in (1, 2, 3)}
            for tc in tcs.values():
                ctx.l.pb2_dtd_task_class_add_chore(tp, tc, R.DEV_CUDA, L.BODY_NOP, None)
            nbf, ft, fo = np.zeros(ntasks, np.int32), np.full((ntasks, 4), -1, np.int32), np.zeros((ntasks, 4), np.int32)
            for t in range(ntasks):
                nf = rnd.choice([1, 2, 3])
                sel = rnd.sample(range(ntiles), nf)
                kinds = [rnd.choice([orc.DTD_INPUT, orc.DTD_INPUT, orc.DTD_INOUT, orc.DTD_OUTPUT]) for _ in range(nf)]
                nbf[t] = nf; ft[t, :nf] = sel; fo[t, :nf] = kinds
                arr = (C.c_void_p * nf)(*[tiles[s] for s in sel])
                opv = np.array([{1: R.INPUT, 2: R.OUTPUT, 3: R.INOUT}[k] for k in kinds], np.int32)
                assert ctx.l.pb2_dtd_insert_task_with_task_class(tp, tcs[nf], 0, R.DEV_CUDA, arr, opv.ctypes.data_as(C.c_void_p), None, 0.0) == t
            src, dst, fl, dep = orc.dtd_build(nbf, ft, fo, ntiles)
            win = ctx.export_window(tp, ctx.devices[0])
            ids = win["task_ids"]
            assert sorted(ids.tolist()) == list(range(ntasks))            # everything is reachable on one device
            cnt = win["tasks"]["succ_count"].astype(np.int64)
            wsrc = np.repeat(ids.astype(np.int64), cnt)
            wdst = ids[(win["succ"] & np.uint32(0x07FFFFFF)).astype(np.int64)]
            wfl = (win["succ"] >> np.uint32(27)).astype(np.int64)
            assert sorted(zip(wsrc.tolist(), wdst.tolist(), wfl.tolist())) == sorted(zip(src.tolist(), dst.tolist(), fl.tolist()))
            assert np.array_equal(win["tasks"]["dep_goal"], dep[ids])


def test_dry_run_execution_replays_reference_coherency():
    """Run Ex05 end to end in dry-run mode (windows retire in order, nothing computes): the host-visible state
    the module leaves behind equals the oracle's replay of the reference protocol, flow by flow."""
    K, NB, tb = 6, 4, 256
    host = np.zeros(K * tb // 4, np.int32)
    with R.Context(cuda_devices=(0,), dry_run=True) as ctx:
        dc = ctx.block_cyclic(4, tb // 4, 1, K * tb // 4, 1, mat=host)
        tp = C.c_void_p(ctx.l.pb2_ptg_ex05_broadcast_new(ctx.h, dc, K, NB))
        ctx.wait()
        st = ctx.stats(ctx.devices[0])
        F = NB // 2 + 1
        assert st["executed_tasks"] == K * (1 + F)
        assert st["data_in_from_device"][0] == K * tb                     # each tile H2D exactly once
        assert st["required_data_in"] == K * (1 + F) * tb
        assert st["windows_launched"] == 1 and st["tasks_released_on_device"] == K * F
        O = orc.lib()
        for k in range(K):
            d = C.c_void_p(ctx.l.pb2_dc_data_of(dc, k, 0))
            od = orc.OrcData(); O.orc_data_create(C.byref(od), 3, 0)
            req = C.c_int()
            assert O.orc_gpu_stage_in(C.byref(od), 2, 0, 0x0C, 0b100, C.byref(req)) == 0
            O.orc_gpu_stage_in_complete(C.byref(od), 2, 0x0C); O.orc_gpu_task_complete(C.byref(od), 2, 0x0C, 0)
            for _ in range(F):
                assert O.orc_gpu_stage_in(C.byref(od), 2, 2, 0x04, 0b100, C.byref(req)) == -1
                O.orc_gpu_task_complete(C.byref(od), 2, 0x04, 0)
            for dev in (0, 2):
                s = ctx.copy_state(d, dev)
                assert (s["coherency"], s["readers"], s["version"]) == (od.copy[dev].coherency_state, od.copy[dev].readers, od.copy[dev].version), (k, dev, s)
            assert ctx.l.pb2_data_owner_device(d) == od.owner_device == 2
        clean, owned = C.c_int(), C.c_int()
        ctx.l.pb2_device_lru_sizes(ctx.devices[0], C.byref(clean), C.byref(owned))
        assert (clean.value, owned.value) == (0, K)                         # written, not pushed out: dirty LRU
        t, dvs = ctx.trace(tp)
        assert len(t) == K * (1 + F) and np.all(dvs == 2)


def test_get_best_device_placement_dry_run():
    """get_best_device_check.jdf:68-83: task(m,n) runs on GPU (n*nt + m) % ngpu (testing_get_best_device.c:158)."""
    nt, ngpu, mb = 5, 4, 8
    with R.Context(cuda_devices=(0,) * ngpu, dry_run=True) as ctx:
        host = np.zeros(nt * nt * mb * mb, np.float64)
        dc = ctx.block_cyclic(8, mb, mb, nt * mb, nt * mb, mat=host)
        info = np.full(nt * (nt + 1) // 2 + 1, -1, np.int32)
        tp = C.c_void_p(ctx.l.pb2_ptg_get_best_device_new(ctx.h, dc, info.ctypes.data_as(C.c_void_p)))
        ctx.wait()
        idx = 0
        for m in range(nt):
            for n in range(m + 1):
                assert info[idx] == 2 + (n * nt + m) % ngpu, (m, n)
                idx += 1
        assert sum(ctx.stats(d)["executed_tasks"] for d in ctx.devices) == nt * (nt + 1) // 2


def test_select_best_device_load_balance_matches_oracle():
    """No affinity: least ETA with the skew rule; cross-check a random load pattern with the oracle."""
    with R.Context(cuda_devices=(0, 0, 0), dry_run=True) as ctx:
        tp = C.c_void_p(ctx.l.pb2_dtd_taskpool_new(ctx.h))
        ops = np.array([R.INPUT], np.int32)
        tc = C.c_void_p(ctx.l.pb2_dtd_create_task_class(tp, b"r", 1, ops.ctypes.data_as(C.c_void_p)))
        ctx.l.pb2_dtd_task_class_add_chore(tp, tc, R.DEV_CUDA, L.BODY_NOP, None)
        n = 30
        for _ in range(n):
            tile = C.c_void_p(ctx.l.pb2_dtd_tile_new(tp, 64))
            arr = (C.c_void_p * 1)(tile)
            ctx.l.pb2_dtd_insert_task_with_task_class(tp, tc, 0, R.DEV_CUDA, arr, None, None, 0.0)
        ctx.wait()
        _, dev = ctx.trace(tp)
        est = ctx.stats(ctx.devices[0])["time_estimate_default"]
        # replay with the oracle: every task adds its estimate to the chosen device until the window retires
        loads = [0, 0, 0, 0, 0]
        chosen = []
        for _ in range(n):
            devs = [(0, 0, 0, 0, 1), (0, 1, 0, 0, 1)] + [(1, 0, 1, loads[2 + g], est) for g in range(3)]
            arr = (orc.SelDev * 5)(*[orc.SelDev(*d) for d in devs])
            a = [np.array(x, np.int32) for x in ([0x04], [1], [-1], [0])]
            c = orc.lib().orc_select_best_device(arr, 5, 1, *[x.ctypes.data_as(C.c_void_p) for x in a], 20, 0)
            chosen.append(c); loads[c] += est
        assert sorted(dev.tolist()) == sorted(chosen)


def test_tiles_of_4_gib_or_more_are_refused_not_truncated():
    """pb2_tile_t::bytes is 32 bits: a datum whose span does not fit makes the window builder fail loudly instead of
    describing a truncated tile to the device."""
    with R.Context(cuda_devices=(0,), dry_run=True,
                   mca={"device_cuda_memory_number_of_blocks": 16, "device_cuda_memory_block_size": 1 << 30}) as ctx:
        dc = ctx.block_cyclic(4, 1 << 30, 1, 1 << 30, 1)              # one tile of exactly 4 GiB
        C.c_void_p(ctx.l.pb2_ptg_ex05_broadcast_new(ctx.h, dc, 1, 4))
        with pytest.raises(L.Pb2Error) as ei:
            ctx.wait()
        assert ei.value.rc == L.PB2_ERR_VALUE_OUT_OF_BOUNDS


def test_cpu_only_chain_is_config1():
    """BASELINE config 1: Ex02_Chain, 1000 tasks, CPU-only: task k sees k, final value 999."""
    with R.Context(cuda_devices=(), dry_run=True) as ctx:
        tp = C.c_void_p(ctx.l.pb2_ptg_ex02_chain_new(ctx.h, 999))
        ctx.l.pb2_taskpool_set_device_types(tp, R.DEV_CPU)
        ctx.wait()
        t, d = ctx.trace(tp)
        assert np.array_equal(t, np.arange(1000)) and np.all(d == 0)
        info = ctx.task_info(tp)
        assert np.array_equal(info["seen_version"][:, 0], np.arange(1000))


def test_large_batches_are_cut_into_pipelined_windows():
    """A big batch of ready GPU tasks is cut into `device_engine_pipeline` windows of whole dependency closures (two in
    flight at a time on a device): same host-visible result as one window -- every tile staged once, every in-closure
    edge released on the device, every task executed once."""
    K, NB, tb = 64, 14, 256
    F = NB // 2 + 1
    host = np.zeros(K * tb // 4, np.int32)
    with R.Context(cuda_devices=(0,), dry_run=True, mca={"device_engine_pipeline": 4, "device_engine_pipeline_min_roots": 16}) as ctx:
        dc = ctx.block_cyclic(4, tb // 4, 1, K * tb // 4, 1, mat=host)
        tp = C.c_void_p(ctx.l.pb2_ptg_ex05_broadcast_new(ctx.h, dc, K, NB))
        ctx.wait()
        st = ctx.stats(ctx.devices[0])
        assert st["windows_launched"] == 4
        assert st["executed_tasks"] == K * (1 + F) and st["tasks_released_on_device"] == K * F
        assert st["data_in_from_device"][0] == K * tb
        ctx.l.pb2_taskpool_free(tp)


def test_pipelined_windows_never_share_a_tile():
    """DTD GEMM: all tasks of a row share A(i,k), all of a column B(k,j).  A window that would need a tile held by the
    window in flight defers those tasks; the pool still completes with every task executed exactly once."""
    NT, T = 6, 32
    with R.Context(cuda_devices=(0,), dry_run=True, mca={"device_engine_pipeline": 3, "device_engine_pipeline_min_roots": 4}) as ctx:
        mats = []
        dcs = []
        for _ in range(3):
            m = np.zeros(NT * NT * T * T, np.uint16)
            mats.append(m)
            dcs.append(ctx.block_cyclic(2, T, T, NT * T, NT * T, mat=m))
        secs = C.c_double(0)
        keep = C.c_void_p()
        rc = ctx.l.pb2_app_dtd_simple_gemm(ctx.h, dcs[0], dcs[1], dcs[2], R.DEV_CUDA, C.byref(secs), C.byref(keep))
        assert rc == 0
        st = ctx.stats(ctx.devices[0])
        assert st["executed_tasks"] == NT ** 3
        assert st["windows_launched"] >= 3
        ctx.l.pb2_taskpool_free(keep)


def test_user_submit_tasks_get_their_own_lane():
    """PB2_BODY_USER chores (user `submit` functions) never enter an engine window: FILL -> user -> CHECK on one tile is
    three windows (engine, host-driven stream lane, engine), in dependency order, versions 0 -> 1 -> 2."""
    with R.Context(cuda_devices=(0,), dry_run=True) as ctx:
        cb = R.GPU_SUBMIT(lambda d, g, s: 0)
        tp = C.c_void_p(ctx.l.pb2_dtd_taskpool_new(ctx.h))
        op = np.array([R.INOUT], np.int32)
        nc = lambda: C.c_void_p(ctx.l.pb2_dtd_create_task_class(tp, b"k", 1, op.ctypes.data_as(C.c_void_p)))
        fill, user, chk = nc(), nc(), nc()
        assert ctx.l.pb2_dtd_task_class_add_chore(tp, fill, R.DEV_CUDA, L.BODY_FILL_I32, None) == 0
        assert ctx.l.pb2_dtd_task_class_add_submit(tp, user, cb) == 0
        assert ctx.l.pb2_dtd_task_class_add_submit(tp, user, R.GPU_SUBMIT()) != 0          # NULL function is refused
        assert ctx.l.pb2_dtd_task_class_add_chore(tp, chk, R.DEV_CUDA, L.BODY_CHECK_I32, None) == 0
        p = lambda *v: np.array(v, np.int32).ctypes.data_as(C.c_void_p)
        for _ in range(3):
            arr = (C.c_void_p * 1)(C.c_void_p(ctx.l.pb2_dtd_tile_new(tp, 4096)))
            ctx.l.pb2_dtd_insert_task_with_task_class(tp, fill, 0, R.DEV_CUDA, arr, p(R.OUTPUT), p(7, 0, 0), 0.0)
            ctx.l.pb2_dtd_insert_task_with_task_class(tp, user, 0, R.DEV_CUDA, arr, p(R.INOUT), p(1, 0, 0), 0.0)
            ctx.l.pb2_dtd_insert_task_with_task_class(tp, chk, 0, R.DEV_CUDA, arr, p(R.INPUT), p(0x01010101, 0, 0), 0.0)
        ctx.wait()
        st = ctx.stats(ctx.devices[0])
        assert st["executed_tasks"] == 9 and st["windows_launched"] == 3
        t, dev = ctx.trace(tp)
        pos = {int(task): i for i, task in enumerate(t)}
        for tile in range(3):
            assert pos[3 * tile] < pos[3 * tile + 1] < pos[3 * tile + 2]


def test_statistics_table_reports_what_the_devices_did():
    """parsec_devices_print_statistics (device.c:499-590): per device kernels, required vs moved bytes, evictions."""
    K, NB, tb = 6, 4, 256
    host = np.zeros(K * tb // 4, np.int32)
    with R.Context(cuda_devices=(0,), dry_run=True) as ctx:
        dc = ctx.block_cyclic(4, tb // 4, 1, K * tb // 4, 1, mat=host)
        C.c_void_p(ctx.l.pb2_ptg_ex05_broadcast_new(ctx.h, dc, K, NB))
        ctx.wait()
        table = ctx.statistics_table()
    rows = [l for l in table.splitlines() if l.strip().startswith("2 |")]
    assert len(rows) == 1 and "cuda(0)" in rows[0]
    cols = [c.strip() for c in rows[0].split("|")]
    F = NB // 2 + 1
    assert int(cols[2]) == K * (1 + F)                               # kernels
    assert cols[4] == "6.00 KB" and cols[5].startswith("1.50 KB") and "25.00" in cols[5]   # each tile H2D once of (1+F) uses
    assert int(cols[10]) == 1 and int(cols[11]) == K * F
    assert table.splitlines()[-1].strip().startswith("all")


@pytest.mark.parametrize("seed", range(16))
def test_random_dtd_pools_under_pipelining_and_memory_pressure(seed):
    """Random DTD pools (GPU bodies, CPU bodies, user-submit bodies on shared tiles) on two dry-run GPUs with a device
    heap far smaller than the working set and windows cut into small pipelined pieces: every task completes exactly
    once, in an order that respects every RAW / WAW / WAR hazard of the insertion order, and the host copy of every
    tile ends at version == number of writers after the final flush."""
    rng = np.random.default_rng(seed)
    ntiles, ntasks, tb = 10, 120, 1024
    mca = {"device_cuda_memory_number_of_blocks": 6, "device_cuda_memory_block_size": tb,
           "device_engine_pipeline": 3, "device_engine_pipeline_min_roots": 2}
    cpu_hook = R.CPU_HOOK(lambda t, p, ip, fp: 0)
    with R.Context(cuda_devices=(0, 1), dry_run=True, mca=mca) as ctx:
        tp = C.c_void_p(ctx.l.pb2_dtd_taskpool_new(ctx.h))
        keep = [R.GPU_SUBMIT(lambda d, g, s: 0)]
        classes = {}
        for nf in (1, 2):
            ops = np.array([R.INOUT] * nf, np.int32)
            for kind in ("gpu", "user"):
                tc = C.c_void_p(ctx.l.pb2_dtd_create_task_class(tp, b"k", nf, ops.ctypes.data_as(C.c_void_p)))
                if kind == "gpu":
                    assert ctx.l.pb2_dtd_task_class_add_chore(tp, tc, R.DEV_CUDA, L.BODY_NOP, None) == 0
                else:
                    assert ctx.l.pb2_dtd_task_class_add_submit(tp, tc, keep[0]) == 0
                classes[(nf, kind)] = tc
            # CPU-only classes: GPU -> CPU hand-over without pushout, CPU writes between GPU uses
            tc = C.c_void_p(ctx.l.pb2_dtd_create_task_class(tp, b"c", nf, ops.ctypes.data_as(C.c_void_p)))
            assert ctx.l.pb2_dtd_task_class_add_chore(tp, tc, R.DEV_CPU, 0, C.cast(cpu_hook, C.c_void_p)) == 0
            classes[(nf, "cpu")] = tc
        tiles = [C.c_void_p(ctx.l.pb2_dtd_tile_new(tp, tb)) for _ in range(ntiles)]
        uses = []                                              # (task, tile, writes)
        kinds, expect_seen, uses_of = [], {}, []
        uses_reads = lambda op: op != R.OUTPUT
        writers = np.zeros(ntiles, np.int64)
        for t in range(ntasks):
            nf = int(rng.integers(1, 3))
            u = rng.random()
            kind = "user" if u < 0.12 else ("cpu" if u < 0.32 else "gpu")
            kinds.append(kind)
            sel = rng.choice(ntiles, nf, replace=False)
            ops = [int(rng.choice([R.INPUT, R.INOUT, R.OUTPUT])) for _ in range(nf)]
            uses_of.append(ops)
            arr = (C.c_void_p * nf)(*[tiles[int(i)] for i in sel])
            tid = ctx.l.pb2_dtd_insert_task_with_task_class(tp, classes[(nf, kind)], int(rng.integers(0, 4)),
                                                            R.DEV_CPU if kind == "cpu" else R.DEV_CUDA, arr,
                                                            np.array(ops, np.int32).ctypes.data_as(C.c_void_p), None, 0.0)
            assert tid == t
            for f, (i, op) in enumerate(zip(sel, ops)):
                uses.append((t, int(i), op != R.INPUT))
                expect_seen[(t, f)] = int(writers[int(i)])      # version a reader must see: writers inserted before it
                writers[int(i)] += op != R.INPUT
            if t % 37 == 36:
                ctx.wait()                                      # several rounds: later inserts chain behind completed tasks
        for tl in tiles:
            ctx.l.pb2_dtd_data_flush(tp, tl)
        ctx.wait()
        order, dev = ctx.trace(tp)
        assert sorted(order.tolist()) == list(range(ntasks))
        pos = np.empty(ntasks, np.int64)
        pos[order] = np.arange(ntasks)
        last_writer, readers = {}, {}
        for t, tile, w in uses:                                # hazards of the sequential insertion order
            if tile in last_writer:
                assert pos[last_writer[tile]] < pos[t], (seed, "RAW/WAW", last_writer[tile], t, tile)
            if w:
                for r in readers.get(tile, []):
                    if r != t:
                        assert pos[r] < pos[t], (seed, "WAR", r, t, tile)
                last_writer[tile] = t; readers[tile] = []
            else:
                readers.setdefault(tile, []).append(t)
        # a CPU body reads the newest version wherever it was produced (GPU in-place writes leave no pushout behind)
        info = ctx.task_info(tp)
        for (t, f), v in expect_seen.items():
            if kinds[t] == "cpu" and uses_reads(uses_of[t][f]):
                assert info["seen_version"][t, f] == v, (seed, t, f, kinds[t], int(info["seen_version"][t, f]), v)
        st = [ctx.stats(d) for d in ctx.devices] + [ctx.stats(ctx.l.pb2_mca_device_get(ctx.h, 0))]
        assert sum(s["executed_tasks"] for s in st) == ntasks
        assert sum(s["windows_launched"] for s in st) > 3
        for i, tl in enumerate(tiles):                         # flushed home: the host copy carries the last version
            hc = ctx.copy_state(C.c_void_p(ctx.l.pb2_dtd_tile_data(tl)), 0)
            assert hc["version"] == writers[i], (seed, i, hc, int(writers[i]))


def test_product_coherency_protocol_equals_oracle_on_random_sequences():
    """pb2_data_start/end_transfer_ownership_to_copy (the product's own statement of parsec/data.c:313-458) against the
    oracle, which tests/test_oracle.py pins to the reference's own build of data.c: same return value, owner, states,
    readers and versions after every step of 300 random access sequences."""
    import random
    from oracle import orc

    class Copy(C.Structure):            # include/pb2_parsec.h: struct pb2_data_copy_s (public part)
        _fields_ = [("device_index", C.c_int8), ("flags", C.c_uint8), ("coherency_state", C.c_uint8),
                    ("data_transfer_status", C.c_uint8), ("readers", C.c_int32), ("version", C.c_uint32)]

    l = R.lib()
    O = orc.lib()
    NDEV = 6
    Rd, Wr = 0x04, 0x08
    buf = np.zeros(16, np.int32)
    for seed in range(300):
        rnd = random.Random(1000 + seed)
        pd = C.c_void_p(l.pb2_data_create(None, seed, buf.ctypes.data_as(C.c_void_p), buf.nbytes))
        od = orc.OrcData()
        O.orc_data_create(C.byref(od), NDEV, 0)
        for step in range(14):
            dev = rnd.randrange(1, NDEV)
            acc = rnd.choice([Rd, Wr, Rd | Wr, Rd, Rd | Wr])
            if not od.copy[dev].present:
                assert l.pb2_data_copy_attach(pd, dev)
                od.copy[dev].present = 1
            if od.copy[dev].coherency_state == 0x1 and od.owner_device != dev:
                continue                                    # two owners: a sane program never asks for this
            r_p = l.pb2_data_start_transfer_ownership_to_copy(None, pd, dev, acc)
            r_o = O.orc_data_start_transfer_ownership(C.byref(od), dev, acc)
            assert r_p == r_o, (seed, step, dev, acc)
            l.pb2_data_end_transfer_ownership_to_copy(pd, dev, acc)
            O.orc_data_end_transfer_ownership(C.byref(od), dev, acc)
            pc = C.cast(l.pb2_data_get_copy(pd, dev), C.POINTER(Copy)).contents
            if acc & Wr:
                src = r_o if r_o >= 0 else dev
                pc.version = od.copy[dev].version = od.copy[src].version + 1
            elif r_o >= 0:
                pc.version = od.copy[dev].version = od.copy[r_o].version
            assert l.pb2_data_owner_device(pd) == od.owner_device, (seed, step)
            for d in range(NDEV):
                cp = l.pb2_data_get_copy(pd, d)
                assert bool(cp) == bool(od.copy[d].present)
                if cp:
                    c = C.cast(cp, C.POINTER(Copy)).contents
                    assert (c.coherency_state, c.readers, c.version) == (od.copy[d].coherency_state, od.copy[d].readers, od.copy[d].version), (seed, step, d)
