"""Splitting one window over the GPUs of a box (pb2_partition_*, host logic; runs without a GPU).

The per-rank windows + remote edges are replayed as ONE merged DAG by the sequential oracle (remote edges become
ordinary edges, every rank's slab is a host buffer, a PEER descriptor's src_ptr points into the producer's buffer
and is read when the consumer runs -- the lazy pull the device does).  Results must equal the unsplit window's, in
FIFO, LIFO and random ready orders: a missing write-after-read edge shows up as a wrong value in some order.
Reference behaviour being mirrored: remote_dep.c:451 (which ranks get an activation), remote_dep_mpi.c:1860
(release of the local successors on the receiver), insert_function.c:2603 / overlap_strategies.c:139-352 (WAR).
"""
import numpy as np
import pytest

from oracle import orc
from oracle import orc_dags as D
from parsec_b200 import _lib as L
from parsec_b200 import multigpu as M


def merged_run(part, world, policy=0, seed=1):
    sizes = [part.sizes(r) for r in range(world)]
    slabs = [np.zeros(max(int(z["slab_bytes"]), 8), np.uint8) for z in sizes]
    base = [s.ctypes.data for s in slabs]
    parts = [part.get(r, base) for r in range(world)]
    toff = np.cumsum([0] + [len(p["tasks"]) for p in parts])
    doff = np.cumsum([0] + [len(p["tiles"]) for p in parts])
    tasks, succ, tiles, ready, gid = [], [], [], [], []
    nsucc = 0
    for r, p in enumerate(parts):
        t = p["tasks"].copy()
        tl = t["tile"]
        tl[tl >= 0] += doff[r]
        t["tile"] = tl
        rs_cnt = np.diff(p["rs_begin"])
        sb = np.zeros(len(t), np.int32)
        for l in range(len(t)):                       # local entries, then remote entries (as ordinary edges)
            loc = p["succ"][t["succ_begin"][l]: t["succ_begin"][l] + t["succ_count"][l]]
            loc = ((loc >> 27) << 27) | ((loc & 0x7FFFFFF) + toff[r])
            rem = []
            for e in range(p["rs_begin"][l], p["rs_begin"][l + 1]):
                rem.append(int(p["rs_target"][e]) + toff[p["rs_rank"][e]])
            sb[l] = nsucc
            succ.extend(int(x) for x in loc)
            succ.extend(rem)
            nsucc += len(loc) + len(rem)
        t["succ_begin"], t["succ_count"] = sb, t["succ_count"] + rs_cnt
        tasks.append(t)
        tiles.append(p["tiles"])
        ready.extend(int(x) + toff[r] for x in p["ready"])
        gid.append(p["global_id"])
    tasks = np.concatenate(tasks)
    res = orc.run_window_raw(tasks, np.array(succ, np.uint32), np.concatenate(tiles), np.array(ready, np.int32), policy, seed)
    return res, np.concatenate(gid), parts, slabs


def check_split(g, world, tile_data_check=True, split_input=None):
    tasks, succ, tiles, ready, task_rank, tile_rank = g
    glob = orc.run_window(tasks, succ, tiles, ready)
    assert glob["rc"] == 0
    part = M.Partition(*(split_input or g), nranks=world)
    assert sum(part.sizes(r)["ntasks"] for r in range(world)) == len(tasks)
    for policy, seed in [(0, 1), (1, 1), (2, 7), (2, 12345), (2, 99)]:
        res, gid, parts, slabs = merged_run(part, world, policy, seed)
        assert res["rc"] == 0, (policy, seed)
        assert res["stats"]["tasks_retired"] == len(tasks)
        assert np.array_equal(res["result"], glob["result"][gid]), (policy, seed)
        assert np.array_equal(res["seen_version"], glob["seen_version"][gid]), (policy, seed)   # same version per flow
        if tile_data_check:
            # the rank that ran the last writer of a tile holds its final version
            last = {}
            for t in glob["retire_order"]:
                for f in range(tasks["nb_flows"][t]):
                    if tasks["tile"][t, f] >= 0 and tasks["access"][t, f] & L.ACCESS_WRITE:
                        last[int(tasks["tile"][t, f])] = int(task_rank[t])
            for tile, r in last.items():
                p = parts[r]
                s = int(np.nonzero(p["slot_tile"] == tile)[0][0])
                o, b = int(p["slot_offset"][s]), int(tiles["bytes"][tile])
                assert np.array_equal(slabs[r][o:o + b], glob["device"][tile]), (policy, seed, tile)
    return part


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_ex05_split_matches_unsplit(world):
    g = M.ex05_global(8 * world, 14, world, 256)
    part = check_split(g, world)
    F = 8
    for r in range(world):
        z = part.sizes(r)
        assert z["ntasks"] == 8 * (1 + F)                       # 8 broadcasts + the receivers whose loc lands here
        # every TaskRecv(k, n) with (k + n) % world != k % world is a remote edge of TaskBcast(k)'s rank
        remote_per_bcast = sum(1 for n in range(0, 15, 2) if n % world != 0)
        assert z["nremote"] == 8 * remote_per_bcast


def test_ex05_matches_oracle_builder():
    """the product-side builder and the oracle's Ex05 builder describe the same window"""
    d = D.ex05_broadcast(12, 14, 256)
    t, succ, tiles, ready, _, _ = M.ex05_global(12, 14, 1, 256)
    a = orc.run_window(t, succ, tiles, ready)
    b = orc.run_window(d.tasks, d.succ, d.tiles_spec() if hasattr(d, "tiles_spec") else tiles, d.ready)
    assert a["rc"] == 0 and b["rc"] == 0
    assert sorted(a["result"].tolist()) == sorted(b["result"].tolist())


@pytest.mark.parametrize("world,n", [(2, 10), (3, 17), (4, 64)])
def test_rtt_ring(world, n):
    g = M.rtt_global(n, world, 512)
    part = check_split(g, world)
    # a chain needs no extra ordering: one remote edge per hop, nothing else
    assert sum(part.sizes(r)["nremote"] for r in range(world)) == n - 1
    assert sum(part.sizes(r)["nsucc"] for r in range(world)) == 0


def random_dtd(ntasks, ntiles, world, seed, tile_bytes=64):
    rng = np.random.default_rng(seed)
    t = np.zeros(ntasks, L.TASK_DTYPE)
    t["tile"][:] = -1
    kind = rng.integers(0, 4, ntasks)
    a = rng.integers(0, ntiles, ntasks)
    b = (a + 1 + rng.integers(0, ntiles - 1, ntasks)) % ntiles
    op = np.zeros((ntasks, 4), np.int32)
    ft = np.full((ntasks, 4), -1, np.int32)
    for i in range(ntasks):
        if kind[i] == 0:
            t["body"][i], t["nb_flows"][i] = L.BODY_INCR_I32, 1
            t["iparam"][i, 0] = int(rng.integers(1, 100))
            t["tile"][i, 0], t["access"][i, 0], op[i, 0], ft[i, 0] = a[i], L.ACCESS_RW, 3, a[i]
        elif kind[i] == 1:
            t["body"][i], t["nb_flows"][i] = L.BODY_SCALE_I32, 1
            t["iparam"][i, 0] = 3
            t["tile"][i, 0], t["access"][i, 0], op[i, 0], ft[i, 0] = a[i], L.ACCESS_RW, 3, a[i]
        elif kind[i] == 2:
            t["body"][i], t["nb_flows"][i] = L.BODY_COPY, 2
            t["tile"][i, 0], t["access"][i, 0], op[i, 0], ft[i, 0] = a[i], L.ACCESS_READ, 1, a[i]
            t["tile"][i, 1], t["access"][i, 1], op[i, 1], ft[i, 1] = b[i], L.ACCESS_WRITE, 2, b[i]
        else:
            t["body"][i], t["nb_flows"][i] = L.BODY_CHECK_I32, 1
            t["tile"][i, 0], t["access"][i, 0], op[i, 0], ft[i, 0] = a[i], L.ACCESS_READ, 1, a[i]
    src, dst, fl, dep = orc.dtd_build(t["nb_flows"].astype(np.int32), ft, op, ntiles)
    begin, count, succ = D._csr_from_edges(ntasks, src, dst, fl)
    t["succ_begin"], t["succ_count"], t["dep_goal"] = begin, count, dep
    tiles = np.zeros(ntiles, L.TILE_DTYPE)
    tiles["bytes"], tiles["state"] = tile_bytes, L.TILE_VALID
    ready = np.nonzero(dep == 0)[0].astype(np.int32)
    task_rank = rng.integers(0, world, ntasks).astype(np.int32)
    tile_rank = rng.integers(0, world, ntiles).astype(np.int32)
    return t, succ, tiles, ready, task_rank, tile_rank


@pytest.mark.parametrize("seed", range(12))
def test_random_dtd_dag_any_owner_map(seed):
    world = 2 + seed % 3
    g = random_dtd(150, 5, world, seed)
    check_split(g, world)


def drop_cross_rank_control_edges(g):
    """remove every edge between two ranks that carries no data (the WAR edges DTD adds); same-rank edges stay"""
    tasks, succ, tiles, ready, task_rank, tile_rank = g
    src, dst, fl = [], [], []
    for p in range(len(tasks)):
        wr = {int(tasks["tile"][p, f]) for f in range(tasks["nb_flows"][p]) if tasks["access"][p, f] & L.ACCESS_WRITE}
        for e in range(tasks["succ_begin"][p], tasks["succ_begin"][p] + tasks["succ_count"][p]):
            s_, f_ = int(succ[e]) & 0x7FFFFFF, int(succ[e]) >> 27
            data = int(tasks["tile"][s_, f_]) in wr
            if data or task_rank[p] == task_rank[s_]:
                src.append(p); dst.append(s_); fl.append(f_)
    t = tasks.copy()
    begin, count, succ2 = D._csr_from_edges(len(t), np.array(src), np.array(dst), np.array(fl))
    t["succ_begin"], t["succ_count"] = begin, count
    dep = np.zeros(len(t), np.int32)
    np.add.at(dep, np.array(dst), 1)
    t["dep_goal"] = dep
    return t, succ2, tiles, np.nonzero(dep == 0)[0].astype(np.int32), task_rank, tile_rank


@pytest.mark.parametrize("seed", range(12))
def test_partitioner_restores_cross_rank_ordering(seed):
    """a DAG that only carries data edges between ranks (what a PTG describes): the partitioner has to hold back
    every overwrite of a slot that a remote reader still pulls from"""
    world = 2 + seed % 3
    g = random_dtd(150, 4, world, 100 + seed)
    g2 = drop_cross_rank_control_edges(g)
    assert len(g2[1]) < len(g[1])
    check_split(g, world, split_input=g2)


@pytest.mark.parametrize("world,P,Q", [(2, 1, 2), (4, 2, 2), (8, 2, 4)])
def test_cholesky_shape_split(world, P, Q):
    """BASELINE configs[4] shape (reduced): GEMM-class bodies, 2D block-cyclic owners; the split replays to the same
    tiles (sparse +-1 data, exact) and the same per-flow versions as the unsplit window."""
    from parsec_b200.bf16 import f32_to_bf16_bits
    NT, nb = 3, 32
    g = M.cholesky_global(NT, nb, P, Q)
    tasks, succ, tiles, ready, task_rank, tile_rank = g
    assert len(tasks) == NT + 2 * (NT * (NT - 1) // 2) + NT * (NT - 1) * (NT - 2) // 6
    rng = np.random.default_rng(3)
    vals = (rng.random((len(tiles), nb, nb)) < 1 / 48) * rng.choice([-1.0, 1.0], (len(tiles), nb, nb))
    bits = f32_to_bf16_bits(vals.astype(np.float32)).reshape(len(tiles), -1)
    spec = np.zeros(len(tiles), orc.TILE_DTYPE)
    spec["bytes"], spec["state"] = nb * nb * 2, orc.TILE_INVALID
    spec["src_ptr"] = np.arange(len(tiles), dtype=np.uint64) * np.uint64(nb * nb * 2)
    glob = orc.run_window(tasks, succ, spec, ready, bits.copy().reshape(-1))
    assert glob["rc"] == 0
    part = M.Partition(*g, nranks=world)
    sizes = [part.sizes(r) for r in range(world)]
    slabs = [np.zeros(max(int(z["slab_bytes"]), 8), np.uint8) for z in sizes]
    parts = [part.get(r, [s.ctypes.data for s in slabs]) for r in range(world)]
    for r, p in enumerate(parts):                                  # initial contents into the home slots
        for sl, tile in enumerate(p["slot_tile"]):
            if tile_rank[tile] == r:
                o = int(p["slot_offset"][sl])
                slabs[r][o:o + nb * nb * 2] = bits[tile].view(np.uint8)
    # merged replay (same construction as merged_run, on the pre-filled slabs)
    toff = np.cumsum([0] + [len(p["tasks"]) for p in parts]); doff = np.cumsum([0] + [len(p["tiles"]) for p in parts])
    mt, ms, mr = [], [], []
    for r, p in enumerate(parts):
        t = p["tasks"].copy(); tl = t["tile"]; tl[tl >= 0] += doff[r]; t["tile"] = tl
        sb = np.zeros(len(t), np.int32)
        for l in range(len(t)):
            loc = p["succ"][t["succ_begin"][l]: t["succ_begin"][l] + t["succ_count"][l]]
            sb[l] = len(ms)
            ms.extend(int(((x >> 27) << 27) | ((x & 0x7FFFFFF) + toff[r])) for x in loc)
            ms.extend(int(p["rs_target"][e]) + int(toff[p["rs_rank"][e]]) for e in range(p["rs_begin"][l], p["rs_begin"][l + 1]))
        t["succ_count"] = t["succ_count"] + np.diff(p["rs_begin"]); t["succ_begin"] = sb
        mt.append(t); mr.extend(int(x) + int(toff[r]) for x in p["ready"])
    gid = np.concatenate([p["global_id"] for p in parts])
    for policy, seed in [(0, 1), (1, 1), (2, 5)]:
        for r, p in enumerate(parts):
            for sl, tile in enumerate(p["slot_tile"]):
                o = int(p["slot_offset"][sl])
                slabs[r][o:o + nb * nb * 2] = bits[tile].view(np.uint8) if tile_rank[tile] == r else 0
        res = orc.run_window_raw(np.concatenate(mt), np.array(ms, np.uint32), np.concatenate([p["tiles"] for p in parts]),
                                 np.array(mr, np.int32), policy, seed)
        assert res["rc"] == 0 and res["stats"]["tasks_retired"] == len(tasks)
        assert np.array_equal(res["seen_version"], glob["seen_version"][gid])
        for tile in range(len(tiles)):
            r = int(tile_rank[tile])                              # the RW chain of a tile lives on its home rank
            p = parts[r]; sl = int(np.nonzero(p["slot_tile"] == tile)[0][0]); o = int(p["slot_offset"][sl])
            assert np.array_equal(slabs[r][o:o + nb * nb * 2], glob["device"][tile]), (policy, tile)


def test_remote_edge_targets_are_consistent():
    """every rank's dep_goal == local in-edges + remote in-edges addressed to it by the other ranks"""
    world = 4
    g = random_dtd(300, 7, world, 1234)
    part = M.Partition(*g, nranks=world)
    parts = [part.get(r, [0x10000000 * (i + 1) for i in range(world)]) for r in range(world)]
    for r, p in enumerate(parts):
        indeg = np.zeros(len(p["tasks"]), np.int64)
        np.add.at(indeg, (p["succ"] & 0x7FFFFFF).astype(np.int64), 1)
        for q, o in enumerate(parts):
            m = o["rs_rank"] == r
            assert q != r or not m.any()
            np.add.at(indeg, o["rs_target"][m].astype(np.int64), 1)
        assert np.array_equal(indeg, p["tasks"]["dep_goal"].astype(np.int64))
        assert not (p["tasks"]["flags"] & L.TASK_DEPS_MASK).any()
        assert np.array_equal(np.sort(p["ready"]), np.nonzero(p["tasks"]["dep_goal"] == 0)[0])


def test_partition_rejects_bad_input():
    g = list(M.ex05_global(4, 14, 2, 256))
    g[4] = g[4].copy(); g[4][0] = 5
    with pytest.raises(L.Pb2Error):
        M.Partition(*g, nranks=2)
    g = list(M.ex05_global(4, 14, 2, 256))
    g[0] = g[0].copy(); g[0]["dep_goal"][10] = 3                 # mask that no edge satisfies
    with pytest.raises(L.Pb2Error):
        M.Partition(*g, nranks=2)


def test_producer_side_push_lists():
    """pb2_partition_set_push: a version a rank reads FIRST in a slot is written there by its producer.  Ex05 on 4 ranks:
    every (tile, remote reader rank) pair is one push of the broadcast task; the rtt ring reuses one slot per rank, so only
    the first hop into each rank can be pushed, the later ones are pulled."""
    from parsec_b200 import multigpu as M
    from parsec_b200 import _lib as L
    world, K = 4, 32
    g = M.ex05_global(K * world, 14, world, 4096)
    part = M.Partition(*g, nranks=world)
    part.set_push(True)
    base = np.arange(world, dtype=np.uint64) * np.uint64(1 << 40)
    tasks, succ, tiles, ready, task_rank, tile_rank = g
    total = 0
    parts = [part.get(r, base) for r in range(world)]
    for r in range(world):
        ps_begin, push = part.get_push(r, base)
        p = parts[r]
        assert ps_begin[0] == 0 and ps_begin[-1] == len(push) and np.all(np.diff(ps_begin) >= 0)
        total += len(push)
        for l in range(len(p["tasks"])):
            for e in push[ps_begin[l]:ps_begin[l + 1]]:
                gid = p["global_id"][l]
                assert tasks["body"][gid] == L.BODY_FILL_I32                    # only the broadcasts write
                dst = parts[e["rank"]]["tiles"][e["desc"]]
                assert dst["src_kind"] == 2 and dst["state"] == L.TILE_INVALID   # PB2_SRC_PUSH, filled by the producer
                assert int(dst["dev_ptr"]) == int(e["dst"]) and dst["bytes"] == e["bytes"] == 4096
                assert int(dst["src_ptr"]) == int(p["tiles"][e["src_tile"]]["dev_ptr"])   # the slot a pull would have read
    # n even: (k + n) % 4 is k % 4 or (k + 2) % 4 -- one remote reader rank per tile
    assert total == K * world
    npush_kinds = sum(int((parts[r]["tiles"]["src_kind"] == 2).sum()) for r in range(world))
    assert npush_kinds == total
    # pushes off: the same descriptors are pulls
    part.set_push(False)
    assert all(int((part.get(r, base)["tiles"]["src_kind"] == 2).sum()) == 0 for r in range(world))
    assert all(len(part.get_push(r, base)[1]) == 0 for r in range(world))
    # rtt: one slot per rank, reused at every lap
    g = M.rtt_global(16, world, 4096)
    part = M.Partition(*g, nranks=world)
    part.set_push(True)
    n = sum(len(part.get_push(r, base)[1]) for r in range(world))
    assert n == world - 1          # hops 1, 2, 3 enter a fresh slot; rank 0's slot holds the initial tile
