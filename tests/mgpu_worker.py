"""One rank of tests/test_multigpu_gpu.py (run under torch.distributed.run)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def cholesky(case, rank, world, local, torch, dist, orc, M, L, Engine):
    """Cholesky-shaped DAG (BASELINE configs[4] shape, reduced) with tensor-core GEMM bodies on a P x Q grid of GPUs.
    (a) NT=3, sparse +-1 data: every partial sum is an integer below 2^24, so fp32 accumulation is exact in any order
        and the per-task bf16 rounding is deterministic: tiles must be bit-identical to the oracle's.
    (b) NT=8, zero data: the values are trivial, the check is the dependency order -- every task retired once and
        every flow saw the tile version the sequential oracle shows it."""
    from parsec_b200.bf16 import f32_to_bf16_bits, bf16_bits_to_f32
    mode = 2 if case.endswith("unfused") else 0
    P, Q = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}[world]
    ok, d2d, nremote, big = True, 0, 0, 0.0
    for NT, nb, dens in [(3, 128, 1.0 / 192), (8, 64, 0.0)]:
        g = M.cholesky_global(NT, nb, P, Q)
        tasks, succ, tiles, ready, task_rank, tile_rank = g
        rng = np.random.default_rng(1805)
        ntiles = len(tiles)
        vals = (rng.random((ntiles, nb, nb)) < dens) * rng.choice([-1.0, 1.0], (ntiles, nb, nb))
        bits = f32_to_bf16_bits(vals.astype(np.float32)).reshape(ntiles, -1)
        host = bits.copy().reshape(-1)
        spec = np.zeros(ntiles, orc.TILE_DTYPE)
        spec["bytes"], spec["state"] = nb * nb * 2, orc.TILE_INVALID
        spec["src_ptr"] = np.arange(ntiles, dtype=np.uint64) * np.uint64(nb * nb * 2)
        glob = orc.run_window(tasks, succ, spec, ready, host.view(np.uint16))
        assert glob["rc"] == 0
        big = max(big, max(float(np.abs(bf16_bits_to_f32(d.view(np.uint16))).max()) for d in glob["device"]))
        part = M.Partition(*g, nranks=world)
        eng = Engine(local, timeout_ms=20000, gemm_mode=mode)
        eng.use_stream(M.work_stream(torch))
        run = M.SharedRun(eng, part, rank, world, dist, torch, kind=1)
        gid = run.p["global_id"]
        last = {}
        for t in glob["retire_order"]:
            for f in range(tasks["nb_flows"][t]):
                if tasks["tile"][t, f] >= 0 and tasks["access"][t, f] & L.ACCESS_WRITE:
                    last[int(tasks["tile"][t, f])] = int(task_rank[t])
        for it in range(2):
            run.load_home_tiles(tile_rank, bits.view(np.uint8).reshape(ntiles, -1))
            run.step()
            torch.cuda.synchronize()
            st = run.wait()
            res = run.w.results()
            if st["tasks_retired"] != len(gid):
                print("rank", rank, "NT", NT, "retired", st["tasks_retired"], "of", len(gid), flush=True)
                ok = False
            sv_ok = np.array_equal(res["seen_version"][:, :3], glob["seen_version"][gid][:, :3])
            if not sv_ok:
                bad = np.nonzero((res["seen_version"][:, :3] != glob["seen_version"][gid][:, :3]).any(axis=1))[0][:4]
                print("rank", rank, "NT", NT, "seen_version differs at global tasks", gid[bad], res["seen_version"][bad], glob["seen_version"][gid][bad], flush=True)
                ok = False
            for tile, r in last.items():
                if r != rank:
                    continue
                got = run.read_tile(tile, nb * nb * 2)
                want = glob["device"][tile]
                if not np.array_equal(got, want[: len(got)]):
                    print("rank", rank, "NT", NT, "tile", tile, "differs:", int((got != want[: len(got)]).sum()), "bytes", flush=True)
                    ok = False
        d2d += int(st["bytes_d2d"]); nremote += len(run.p["rs_rank"])
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    tot = torch.tensor([d2d, nremote], device="cuda")
    dist.all_reduce(tot)
    return {"ok": bool(flag.item()), "world": world, "case": case, "bytes_d2d": int(tot[0].item()),
            "remote_edges": int(tot[1].item()), "max_abs": big, "exact_range": big < 2.0 ** 23}


def run_case(case, rank, world, local, torch, dist):
    """One parity case on an initialised process group (tests run it through main(), bench.py calls it in-run).
    Returns the verdict (identical on every rank)."""
    from oracle import orc
    from parsec_b200 import multigpu as M
    from parsec_b200 import _lib as L
    from parsec_b200.engine import Engine
    g = None
    if case == "ex05":
        g = M.ex05_global(64 * world, 14, world, 65536)
    elif case == "rtt":
        g = M.rtt_global(200, world, 1 << 20)
    elif case == "random_dtd":
        import test_partition as T
        # the oracle runs the fully ordered (DTD) window; the partitioner gets it WITHOUT the cross-rank control edges
        full = T.random_dtd(400, 6, world, 4242, tile_bytes=4096)
        g = T.drop_cross_rank_control_edges(full)
    if case.startswith("cholesky"):
        return cholesky(case, rank, world, local, torch, dist, orc, M, L, Engine)
    tasks, succ, tiles, ready, task_rank, tile_rank = g
    glob = orc.run_window(*(full if case == "random_dtd" else g)[:4])
    assert glob["rc"] == 0
    part = M.Partition(*g, nranks=world)
    eng = Engine(local, timeout_ms=20000)
    eng.use_stream(M.work_stream(torch))
    run = M.SharedRun(eng, part, rank, world, dist, torch)
    ok = True
    d2d = 0
    last = {}
    zeros = np.zeros(run.slab_bytes, np.uint8)
    for it in range(3):                       # the window is re-armed and re-run: reset + barrier protocol
        eng.h2d(run.slab, zeros)
        eng.synchronize()
        dist.barrier()
        run.step()
        torch.cuda.synchronize()
        st = run.wait()
        res = run.w.results()
        gid = run.p["global_id"]
        if st["tasks_retired"] != len(gid):
            print("rank", rank, "iteration", it, "retired", st["tasks_retired"], "of", len(gid), flush=True)
            ok = False
        if it == 0:
            # CHECK bodies count elements != iparam[0]: the split run must count what the unsplit oracle run counts
            chk = tasks["body"][gid] == L.BODY_CHECK_I32
            want = int((glob["result"][gid][chk] >> np.uint64(32)).sum())
            if st["body_errors"] != want:
                print("rank", rank, "body_errors", st["body_errors"], "oracle", want, flush=True)
                ok = False
            same = res["result"] == glob["result"][gid]
            if not same.all():
                bad = np.nonzero(~same)[0][:5]
                print("rank", rank, "mismatch at global tasks", gid[bad], res["result"][bad], glob["result"][gid][bad], flush=True)
            ok = ok and bool(same.all())
            # final versions: the rank that ran the last writer of a tile holds the oracle's final bytes
            slab = np.zeros(run.slab_bytes, np.uint8)
            eng.d2h(slab, run.slab)
            eng.synchronize()
            for t in glob["retire_order"]:
                for f in range(tasks["nb_flows"][t]):
                    if tasks["tile"][t, f] >= 0 and tasks["access"][t, f] & L.ACCESS_WRITE:
                        last[int(tasks["tile"][t, f])] = int(task_rank[t])
            for tile, r in last.items():
                if r != rank:
                    continue
                sl = int(np.nonzero(run.p["slot_tile"] == tile)[0][0])
                o, b = int(run.p["slot_offset"][sl]), int(tiles["bytes"][tile])
                if not np.array_equal(slab[o:o + b], glob["device"][tile]):
                    print("rank", rank, "final bytes of tile", tile, "differ", slab[o:o + 16], glob["device"][tile][:16], flush=True)
                    ok = False
        d2d += int(st["bytes_d2d"])
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    tot = torch.tensor([d2d, len(run.p["rs_rank"])], device="cuda")
    dist.all_reduce(tot)
    return {"ok": bool(flag.item()), "world": world, "case": case, "bytes_d2d": int(tot[0].item()), "remote_edges": int(tot[1].item())}


def main():
    import torch
    import torch.distributed as dist
    case = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = run_case(case, rank, world, local, torch, dist)
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
