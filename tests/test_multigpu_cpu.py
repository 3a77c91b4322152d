"""N > 1 host logic on CPU: world_size-2/4 gloo processes shard Ex05 by the block-cyclic owner map, run their
windows through the ORACLE (no GPU here), exchange tiles with the same batched send/recv plan the NCCL path
uses, and every receiver must observe its k (Ex05_Broadcast.jdf:53-57)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, K, NB, tb, q):
    import torch
    import torch.distributed as dist
    from oracle import orc
    from parsec_b200 import multigpu as mg
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = mg.ex05_shard(K, NB, world, rank, tb)
    slab = torch.full((K * tb // 4,), -1, dtype=torch.int32)
    recv = [torch.full((K * tb // 4,), -2, dtype=torch.int32) for _ in sh["recv_from"]]

    def run(phase, bufs, valid):
        n = K * len(bufs)
        if len(phase["tasks"]) == 0:
            return 0, []
        host = np.concatenate([b.numpy() for b in bufs]) if bufs else np.zeros(0, np.int32)
        t = np.zeros(n, orc.TILE_DTYPE)
        t["bytes"], t["state"] = tb, orc.TILE_VALID if valid else orc.TILE_INVALID
        t["src_ptr"] = np.arange(n, dtype=np.uint64) * np.uint64(tb)
        out = orc.run_window(phase["tasks"], phase["succ"], t, phase["ready"], host)
        assert out["rc"] == 0 and out["stats"]["body_errors"] == 0, (rank, out["stats"])
        seen = [(int(phase["tasks"]["locals"][i, 0]), int(phase["tasks"]["locals"][i, 1]), int(out["result"][i] & 0xFFFFFFFF))
                for i in range(len(phase["tasks"])) if phase["tasks"]["class_id"][i] == 1]
        return out, seen

    outA, seenA = run(sh["phase_a"], [slab], False)
    for l in range(K):                                         # the oracle's "device" tiles are what gets sent
        slab[l * tb // 4:(l + 1) * tb // 4] = torch.from_numpy(outA["device"][l].view(np.int32).copy())
    mg.exchange(dist, slab, recv, sh["send_to"], sh["recv_from"])
    outB, seenB = run(sh["phase_b"], recv, False)    # receive buffers are the source the oracle stages from
    q.put((rank, sh["tasks_per_rank"], seenA + seenB, sh["send_to"], sh["recv_from"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_ex05_sharded_over_ranks_gloo(world):
    K, NB, tb = 6, 14, 64
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, K, NB, tb, q)) for r in range(world)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    F = NB // 2 + 1
    allseen = []
    for rank, ntasks, seen, send_to, recv_from in res:
        assert ntasks == K * (1 + F)                            # weak scaling: every rank owns the same share
        for k, n, val in seen:
            assert val == k                                     # every TaskRecv(k, n) observes k
            assert (k + n) % world == rank                      # and runs on the owner of mydata(k + n)
        allseen += [(k, n) for k, n, _ in seen]
        assert sorted(send_to) == sorted((rank + o) % world for o in {n % world for n in range(0, NB + 1, 2)} - {0})
    nodes = K * world
    assert sorted(allseen) == sorted((k, n) for k in range(nodes) for n in range(0, NB + 1, 2))


def test_owner_map_matches_collection():
    """The shard's owner rule is the collection's rank_of (1 x N grid)."""
    import ctypes as C
    from parsec_b200 import multigpu as mg
    from parsec_b200 import runtime as R
    with R.Context(cuda_devices=(), dry_run=True) as ctx:
        for world in (2, 4, 8):
            dc = ctx.block_cyclic(4, 4, 1, 4 * 32, 1, P=1, Q=1)
            dcq = ctx.block_cyclic(4, 1, 4, 1, 4 * 32, P=1, Q=world)
            for k in range(32):
                assert ctx.l.pb2_dc_rank_of(dcq, 0, k) == mg.owner_1xN(k, world)
            ctx.l.pb2_data_collection_free(dc); ctx.l.pb2_data_collection_free(dcq)


def _direct_worker(rank, world, port, q):
    """Host protocol of the direct path, one process per rank, no GPU: every rank partitions the same global window,
    publishes (fake) slab addresses and its window's release table like SharedRun does with IPC handles, translates
    its remote edges, and the ranks cross-check that what the peers will release adds up to every task's goal."""
    import torch.distributed as dist
    from parsec_b200 import multigpu as mg
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, Q = {2: (1, 2), 4: (2, 2)}[world]
    g = mg.cholesky_global(5, 64, P, Q)
    part = mg.Partition(*g, nranks=world)
    bases = [None] * world
    dist.all_gather_object(bases, 0x10000000000 * (rank + 1))          # stands for the IPC-mapped slab of each rank
    p = part.get(rank, bases)
    # what pb2_window_task_entries returns for an HBM window without wide tasks: entry = task id, one ring entry
    tables = [None] * world
    dist.all_gather_object(tables, np.arange(len(p["tasks"]), dtype=np.int32) + 0)
    tgt = mg.translate_remote_targets(p, tables)
    sent = [np.bincount(tgt[p["rs_rank"] == r].astype(np.int64), minlength=len(tables[r])) for r in range(world)]
    allsent = [None] * world
    dist.all_gather_object(allsent, sent)
    indeg = np.zeros(len(p["tasks"]), np.int64)
    np.add.at(indeg, (p["succ"] & 0x7FFFFFF).astype(np.int64), 1)
    for r in range(world):
        indeg += allsent[r][rank]
    peers_pull_from = sorted({int(a) // 0x10000000000 - 1 for a in p["tiles"]["src_ptr"][p["tiles"]["src_kind"] == 1]})
    q.put((rank, bool(np.array_equal(indeg, p["tasks"]["dep_goal"])), int(len(p["rs_rank"])), peers_pull_from,
           int(len(p["tasks"])), int((p["tiles"]["src_kind"] == 1).sum())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_direct_path_host_protocol_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_direct_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)                                   # goals == local + remote in-edges on every rank
    assert sum(r[2] for r in res) > 0 and sum(r[5] for r in res) > 0
    assert sum(r[4] for r in res) == 5 + 2 * 10 + 10                 # POTRF + TRSM + SYRK + GEMM of NT = 5
    for r in res:
        assert r[0] not in r[3]                                     # a rank never pulls from itself
