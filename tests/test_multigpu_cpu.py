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
