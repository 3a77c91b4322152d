"""BASELINE configs[4]: tile-Cholesky-shaped DAG, 1024x1024 bf16 tiles, N = NT*1024, P x Q grid of GPUs, tensor-core
GEMM bodies, cross-GPU edges released by the device.  Run under torch.distributed.run.
Prints tasks/s and TFLOP/s (GEMM-class tasks x 2*nb^3)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from parsec_b200 import multigpu as M
    from parsec_b200.bf16 import f32_to_bf16_bits
    from parsec_b200.engine import Engine

    NT = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    steps = 5
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    P, Q = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}[world]
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = M.cholesky_global(NT, nb, P, Q)
    tasks, succ, tiles, ready, task_rank, tile_rank = g
    part = M.Partition(*g, nranks=world)
    eng = Engine(local, timeout_ms=30000, gemm_mode=mode)
    eng.use_stream(M.work_stream(torch))
    run = M.SharedRun(eng, part, rank, world, dist, torch, kind=1)
    one = f32_to_bf16_bits(np.random.default_rng(7).uniform(-0.01, 0.01, nb * nb).astype(np.float32)).view(np.uint8)
    run.load_home_tiles(tile_rank, {int(t): one for t in range(len(tiles))})
    for _ in range(2):
        run.step()
    torch.cuda.synchronize(); st = run.wait()
    assert st["tasks_retired"] == len(run.p["global_id"]), st
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
    tot = torch.tensor([float(st["bytes_d2d"]), float(len(run.p["rs_rank"]))], device="cuda", dtype=torch.float64)
    dist.all_reduce(tot)
    if rank == 0:
        t = float(ms.item()) / 1e3
        ngemm = int((tasks["body"] == 16).sum())
        print(json.dumps({"config": "cholesky-shaped DAG", "n_gpus": world, "grid": [P, Q], "NT": NT, "tile": [nb, nb, "bf16"],
                          "tasks": len(tasks), "ms_per_run": t * 1e3, "tasks_per_s": len(tasks) / t,
                          "tflops": ngemm * 2.0 * nb ** 3 / t / 1e12, "tflops_per_gpu": ngemm * 2.0 * nb ** 3 / t / 1e12 / world,
                          "gemm_mode": mode, "d2d_bytes_per_run": tot[0].item(), "remote_edges": int(tot[1].item())}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
