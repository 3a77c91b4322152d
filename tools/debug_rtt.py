"""debug: bench.secondary_rtt's check at N ranks with a small chain, printing what each rank holds"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from parsec_b200 import multigpu as M
from parsec_b200.engine import Engine
for (nt, tile, frags, pb, steps) in [(16, 1 << 20, 1, 32768, 3), (1024, 4 << 20, 1, 32768, 3), (1024, 4 << 20, 16, 32768, 3), (1024, 4 << 20, 1, 0, 3)]:
    g = M.rtt_global(nt, world, tile, frags)
    part = M.Partition(*g, nranks=world, part_bytes=pb) if pb else M.Partition(*g, nranks=world)
    eng = Engine(local, timeout_ms=20000, part_bytes=pb) if pb else Engine(local, timeout_ms=20000)
    eng.use_stream(M.work_stream(torch))
    run = M.SharedRun(eng, part, rank, world, dist, torch)
    vals = []
    for r in range(2 + steps):
        run.step()
        torch.cuda.synchronize(); st = run.wait()
        dist.barrier(); torch.cuda.synchronize()
        slab = np.zeros(run.slab_bytes // 4, np.int32)
        eng.d2h(slab, run.slab); eng.synchronize()
        t0 = slab[: frags * tile // 4]
        vals.append((int(t0.min()), int(t0.max()), st["tasks_retired"]))
        dist.barrier()
    k0 = ((nt - 1) // world) * world
    print("rank", rank, "nt", nt, "frags", frags, "pb", pb, "slab_bytes", run.slab_bytes, "slots", run.p["slot_offset"][:4], "per-run (min,max,retired)", vals,
          "expected at last rank after R runs: (R-1)*%d+%d" % (k0 + 1, nt), "last rank", (nt - 1) % world, flush=True)
    del run
    eng.close()
dist.destroy_process_group()
