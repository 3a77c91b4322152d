"""Host logic of the streaming ring (include/pb2_stream.h) on dry-run streams: no CUDA is touched, tasks retire in
dependency order without running their bodies.  The GPU tests in test_stream_gpu.py run the same drivers for real."""
import ctypes as C

import numpy as np
import pytest

from oracle import orc, orc_dags as dags
from parsec_b200 import _lib as L
from parsec_b200.stream import Stream, STREAM_SYMBOLS, lib, run_dag


def test_symbols_exported():
    l = lib()
    for n in STREAM_SYMBOLS:
        assert hasattr(l, n), n


def test_create_without_engine_needs_dry_run():
    with pytest.raises(L.Pb2Error):
        Stream(None)                       # a real stream without an engine is refused, loudly


def _tiles(dag):
    t = np.zeros(dag.ntiles, L.TILE_DTYPE)
    t["bytes"] = dag.tile_bytes
    t["state"] = L.TILE_VALID
    return t


@pytest.mark.parametrize("mode", ["lookahead", "host"])
@pytest.mark.parametrize("maker", [lambda: dags.ex05_broadcast(16, 6, 64), lambda: dags.ex02_chain(50),
                                   lambda: dags.ep(8, 5), lambda: dags.rtt_chain(12, 3, 256)])
def test_dry_run_order_is_a_linear_extension_and_versions_match_oracle(mode, maker):
    dag = maker()
    tiles = _tiles(dag)
    with Stream(None, dry_run=1, cmd_slots=1024, max_tiles=max(dag.ntiles, 1)) as s:
        out = run_dag(s, dag, tiles, mode=mode)
        st = s.stats()
    n = len(dag.tasks)
    order = out["retire_order"]
    assert sorted(order.tolist()) == list(range(n))
    pos = np.empty(n, np.int64)
    pos[order] = np.arange(n)
    for u in range(n):
        t = dag.tasks[u]
        for e in dag.succ[t["succ_begin"]:t["succ_begin"] + t["succ_count"]]:
            assert pos[u] < pos[int(e) & 0x07FFFFFF]
    spec = np.zeros(dag.ntiles, orc.TILE_DTYPE)
    spec["bytes"] = dag.tile_bytes
    spec["state"] = L.TILE_VALID
    spec["src_ptr"] = np.arange(dag.ntiles, dtype=np.uint64) * np.uint64(max(dag.tile_bytes, 1))
    ref = orc.run_window(dag.tasks, dag.succ, spec, dag.ready, np.zeros(max(dag.ntiles * dag.tile_bytes, 4) // 4 + 1, np.int32))
    assert ref["rc"] == 0
    assert np.array_equal(out["seen_version"], ref["seen_version"])
    assert st["submitted"] == n and st["retired"] == n
    if mode == "lookahead":
        assert st["edges"] == sum(int(t["succ_count"]) for t in dag.tasks)


def test_ring_full_is_reported_not_dropped():
    with Stream(None, dry_run=1, cmd_slots=1024, max_tiles=4) as s:
        row = np.zeros(1, L.TASK_DTYPE)
        row["tile"] = -1
        got = [s.submit(row, cookie=i, allow_full=True) for i in range(1100)]
        assert got.count(None) == 1100 - 1024 and None not in got[:1024]
        assert s.inflight() == 1024
        recs = s.poll(4096)
        assert [r[0] for r in recs] == list(range(1024))          # FIFO among ready tasks
        assert s.submit(row, cookie=7, allow_full=True) is not None


def test_bad_arguments():
    with Stream(None, dry_run=1, cmd_slots=1024, max_tiles=4) as s:
        row = np.zeros(1, L.TASK_DTYPE)
        row["tile"] = -1
        row["body"] = L.BODY_GEMM_BF16
        with pytest.raises(L.Pb2Error):
            s.submit(row)                  # tensor-core bodies run in windows, never in the streaming kernel
        row["body"] = 0
        row["tile"][0, 0] = 9
        row["nb_flows"] = 1
        with pytest.raises(L.Pb2Error):
            s.submit(row)                  # tile id outside the table
        with pytest.raises(L.Pb2Error):
            s.add_edge(1, 2)               # tickets that are not in flight


def _two_sided(stream, n):
    """One thread submits n empty tasks (retrying when tickets or ring space run out), another polls at the same time:
    every cookie comes back exactly once (the threading contract of include/pb2_stream.h)."""
    import threading
    from parsec_b200 import _lib as L
    task = np.zeros(1, L.TASK_DTYPE)
    task["tile"][:] = -1
    task["body"] = L.BODY_NOP
    seen, errors = [], []

    def submitter():
        try:
            i = 0
            while i < n:
                if stream.submit(task, cookie=i + 1, allow_full=True) is None:
                    continue
                i += 1
                if (i & 63) == 0:
                    stream.kick()
            stream.kick()
        except Exception as exc:                       # pragma: no cover
            errors.append(exc)

    th = threading.Thread(target=submitter)
    th.start()
    spins = 0
    while len(seen) < n and spins < 50_000_000 and not errors:
        got = stream.poll()
        seen.extend(r[0] for r in got)
        spins += 1
    th.join()
    assert not errors, errors
    assert len(seen) == n and sorted(seen) == list(range(1, n + 1))
    assert stream.inflight() == 0


def test_submit_and_poll_from_two_threads_dry_run():
    from parsec_b200.stream import Stream
    with Stream(None, dry_run=1, cmd_slots=1024) as s:
        _two_sided(s, 20000)
