"""BASELINE configs[3]: rtt.jdf ring, 1024x1024 fp32 tiles (4 MiB), NT hops over WS GPUs, FRAGS chains.
Run under torch.distributed.run.  Prints tile GB/s = (NT-1)*FRAGS*tile_bytes / t (rtt_main.c:239-241, decimal GB)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from parsec_b200 import multigpu as M
    from parsec_b200.engine import Engine

    nt = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    frags_list = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["1", "16"])]
    tile = int(sys.argv[3]) if len(sys.argv) > 3 else 4 << 20
    part_bytes = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    steps = 5
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for frags in frags_list:
        g = M.rtt_global(nt, world, tile, frags)
        part = M.Partition(*g, nranks=world, part_bytes=part_bytes)
        eng = Engine(local, timeout_ms=20000, part_bytes=part_bytes)
        eng.use_stream(M.work_stream(torch))
        run = M.SharedRun(eng, part, rank, world, dist, torch)
        for _ in range(2):
            run.step()
        torch.cuda.synchronize(); run.wait()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            run.step()
        e1.record()
        torch.cuda.synchronize()
        st = run.wait()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        # parity: a run starts from rank 0's slot (left at k0 + 1 by its last task of the previous run) and adds NT;
        # the final version lives on the rank that ran the last hop
        ok = True
        if (nt - 1) % world == rank:
            slab = np.zeros(run.slab_bytes // 4, np.int32)
            eng.d2h(slab, run.slab); eng.synchronize()
            k0 = ((nt - 1) // world) * world
            ok = bool(np.all(slab[: frags * tile // 4] == (1 + steps) * (k0 + 1) + nt))
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            t = float(ms.item()) / 1e3
            print(json.dumps({"config": "rtt ring", "n_gpus": world, "NT": nt, "FRAGS": frags, "tile_bytes": tile,
                              "ms_per_run": t * 1e3, "hops_per_s": (nt - 1) * frags / t,
                              "tile_GBs": (nt - 1) * frags * tile / t / 1e9, "us_per_hop": t / (nt - 1) * 1e6,
                              "parity_ok": bool(flag.item()), "retired": st["tasks_retired"], "d2d_bytes": st["bytes_d2d"]}), flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
