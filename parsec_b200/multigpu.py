"""Sharding a task pool over the GPUs of one box: one process per GPU, 2D block-cyclic owner map, NCCL over
NVLink for the dependency edges that cross GPUs.

The reference spreads a PTG over processes with the collection's rank_of (two_dim_rectangle_cyclic.c:258-286):
a task runs where its affinity datum lives; an edge whose endpoints have different owners is a remote
dependency (remote_dep.c): ACTIVATE message, then the tile itself (remote_dep_mpi.c:1681, :2120).  Here the
same split is computed on the host of every rank, the local parts run as device windows, and all the tiles a
window produced for remote successors travel in one NCCL exchange (send/recv pairs over NVLink) before the
window that consumes them -- the ACTIVATE/GET/PUT hand-shake of a whole dependency frontier batched into one
collective step, no per-edge host round trip.

Host logic only (numpy + torch.distributed plumbing); the kernels are the engine's.
"""
import numpy as np

from . import _lib as L


def owner_1xN(k, world):
    """rank_of for a 1 x world grid, kp = kq = 1, ip = jq = 0: tile column k lives on rank k % world."""
    return k % world


def ex05_shard(K_local, NB, world, rank, tile_bytes):
    """Split Ex05_Broadcast (examples/Ex05_Broadcast.jdf:24-58) with nodes = K_local * world over `world` ranks.

    TaskBcast(k) runs on owner(mydata(k)); TaskRecv(k, n) on owner(mydata(k + n)) (loc = k + n, :45-47).
    Returns a dict with
      phase_a : window arrays (tasks, succ, ready, tile ids) of TaskBcast(k) + the receivers that are local
      phase_b : window arrays of the receivers whose tile arrives from another rank
      send_to : ranks that need every tile this rank broadcasts (n even => offsets n % world)
      recv_from : ranks whose tiles this rank's phase-b receivers read, in receive-buffer order
    Local tile l of this rank is global group k = l * world + rank.  Receive buffer j holds the K_local tiles of
    rank recv_from[j], in that rank's local order.
    """
    F = NB // 2 + 1
    offsets = sorted({n % world for n in range(0, NB + 1, 2)})          # (k + n) % world - k % world
    remote_off = [o for o in offsets if o != 0]
    send_to = [(rank + o) % world for o in remote_off]
    recv_from = [(rank - o) % world for o in remote_off]
    ns = np.arange(0, NB + 1, 2)
    local_n = ns[(ns % world) == 0]                                     # receivers of my own tiles that stay here
    # ---- phase A: my TaskBcast + local receivers
    nA = K_local * (1 + len(local_n))
    a = np.zeros(nA, dtype=L.TASK_DTYPE)
    a["tile"][:] = -1
    l = np.arange(K_local, dtype=np.int32)
    kglob = l * world + rank
    b = a[:K_local]
    b["body"], b["nb_flows"], b["flags"] = L.BODY_FILL_I32, 1, L.TASK_DEPS_MASK
    b["tile"][:, 0], b["access"][:, 0], b["iparam"][:, 0], b["locals"][:, 0] = l, L.ACCESS_RW, kglob, kglob
    r = a[K_local:]
    ll = np.repeat(l, len(local_n))
    r["body"], r["nb_flows"], r["flags"], r["class_id"], r["dep_goal"] = L.BODY_CHECK_I32, 1, L.TASK_DEPS_MASK, 1, 0x1
    r["tile"][:, 0], r["access"][:, 0] = ll, L.ACCESS_READ
    r["iparam"][:, 0] = ll * world + rank
    r["locals"][:, 0] = ll * world + rank
    r["locals"][:, 1] = np.tile(local_n, K_local)
    a["succ_begin"][:K_local] = np.arange(K_local) * len(local_n)
    a["succ_count"][:K_local] = len(local_n)
    succ_a = (K_local + np.arange(K_local * len(local_n))).astype(np.uint32)
    # ---- phase B: receivers of tiles that arrive from recv_from[j]
    per_src = []
    for j, (src, off) in enumerate(zip(recv_from, remote_off)):
        n_src = ns[(ns % world) == off]                                 # the n's of rank src's tiles that land here
        nb_ = K_local * len(n_src)
        t = np.zeros(nb_, dtype=L.TASK_DTYPE)
        t["tile"][:] = -1
        ls = np.repeat(l, len(n_src))
        t["body"], t["nb_flows"], t["flags"], t["class_id"], t["dep_goal"] = L.BODY_CHECK_I32, 1, L.TASK_DEPS_MASK, 1, 0
        t["tile"][:, 0], t["access"][:, 0] = j * K_local + ls, L.ACCESS_READ
        t["iparam"][:, 0] = ls * world + src
        t["locals"][:, 0] = ls * world + src
        t["locals"][:, 1] = np.tile(n_src, K_local)
        per_src.append(t)
    bt = np.concatenate(per_src) if per_src else np.zeros(0, L.TASK_DTYPE)
    return {
        "phase_a": dict(tasks=a, succ=succ_a, ready=np.arange(K_local, dtype=np.int32), ntiles=K_local),
        "phase_b": dict(tasks=bt, succ=np.zeros(0, np.uint32), ready=np.arange(len(bt), dtype=np.int32), ntiles=K_local * len(recv_from)),
        "send_to": send_to, "recv_from": recv_from, "F": F, "tile_bytes": tile_bytes,
        "tasks_per_rank": nA + len(bt),
    }


def exchange(dist, send_buf, recv_bufs, send_to, recv_from):
    """One batched send/recv of the whole slab per (destination, source) pair: the PUT of every remote edge."""
    if not send_to:
        return
    ops = []
    for dst in send_to:
        ops.append(dist.P2POp(dist.isend, send_buf, dst))
    for buf, src in zip(recv_bufs, recv_from):
        ops.append(dist.P2POp(dist.irecv, buf, src))
    for w in dist.batch_isend_irecv(ops):
        w.wait()


def ex05_multi_gpu_step_factory(ctx, dev, dc, K, NB, tile_bytes, rank, world, local_rank):
    """Device windows + NCCL exchange for one rank; everything is enqueued on torch's current CUDA stream."""
    import torch
    import torch.distributed as dist
    from .engine import Engine

    sh = ex05_shard(K, NB, world, rank, tile_bytes)
    eng = Engine(local_rank)
    eng.use_stream(torch.cuda.current_stream().cuda_stream)
    slab = torch.zeros(K * tile_bytes // 4, dtype=torch.int32, device="cuda")
    recv = [torch.empty(K * tile_bytes // 4, dtype=torch.int32, device="cuda") for _ in sh["recv_from"]]

    def tiles_for(bufs):
        n = K * len(bufs)
        t = np.zeros(n, L.TILE_DTYPE)
        for j, b in enumerate(bufs):
            t["dev_ptr"][j * K:(j + 1) * K] = b.data_ptr() + np.arange(K, dtype=np.uint64) * np.uint64(tile_bytes)
        t["bytes"], t["state"] = tile_bytes, L.TILE_VALID
        return t

    pa, pb = sh["phase_a"], sh["phase_b"]
    wa = eng.window(0, pa["tasks"], pa["succ"], tiles_for([slab]), pa["ready"])
    wb = eng.window(0, pb["tasks"], pb["succ"], tiles_for(recv), pb["ready"]) if len(pb["tasks"]) else None
    keep = (eng, slab, recv, wa, wb)

    def step():
        wa.launch()
        exchange(dist, slab, recv, sh["send_to"], sh["recv_from"])
        if wb is not None:
            wb.launch()

    def finish():
        torch.cuda.synchronize()
        st = wa.wait()
        assert st["body_errors"] == 0 and st["tasks_retired"] == len(pa["tasks"])
        if wb is not None:
            st = wb.wait()
            assert st["body_errors"] == 0 and st["tasks_retired"] == len(pb["tasks"]), st
        return keep

    return step, finish, (4 if wb is not None else 2)
