#!/usr/bin/env python
"""bench.py -- BASELINE.json metric on its config: tasks/s (+ tile GB/s) of a 2D-block-cyclic tile DAG on B200.

Workload at N=1 = BASELINE configs[1]: Ex05_Broadcast dataflow, 256x256 fp32 tiles (262 144 B), K = 4096
broadcast groups, fan-out F = 8 (NB = 14): 36 864 tasks, 9.66 GB of algorithmic tile traffic per step.
A "step" is one complete pass of that DAG through the device engine.

  value      whole-job tasks/s with the tiles already resident in HBM; timed on the device with CUDA events
             around the window re-arm + persistent-kernel launch (K steps back to back on the engine stream).
  e2e        the same metric through the reference-shaped host API (PTG front end -> device module ->
             kernel_scheduler) with HOST buffers: every step builds the task pool, stages every tile in from
             pinned host memory inside the kernel, and writes every tile back to host memory
             (parsec_device_flush_lru); bytes are counted from the buffers that move.
  roofline   HBM bound: algorithmic bytes of the window kernel / its CUDA-event duration vs MEASURED_PEAKS.json.
  cpu_baseline / --impl reference
             the oracle's multi-threaded CPU port of the reference scheduler path (oracle/orc_cpu_sched.c; the
             reference runtime itself is not buildable here) on all host cores, same DAG, same tiles
             in host memory.
N > 1 (torchrun, one rank per GPU): weak scaling, every rank owns K groups of a K*N-group collection laid out
by the 2D block-cyclic map (1 x N grid); TaskRecv(k, n) lives on the owner of mydata(k+n), so tiles whose
receivers are remote cross NVLink through one NCCL all-to-all per step between two device windows.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TILE = 256 * 256 * 4
K_GROUPS = 4096
NB = 14
F = NB // 2 + 1


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference_arm(steps, warmup, sample_groups):
    """The reference's CPU path (oracle port) on all host cores; a step = the full DAG on `sample_groups` groups."""
    from oracle import orc, orc_dags as dags
    cores = os.cpu_count() or 1
    dag = dags.ex05_broadcast(sample_groups, NB, TILE)
    host = np.zeros(sample_groups * TILE // 4, np.int32)
    tiles = np.zeros(sample_groups, orc.TILE_DTYPE)
    tiles["bytes"] = TILE
    tiles["state"] = orc.TILE_VALID
    tiles["dev_ptr"] = host.ctypes.data + np.arange(sample_groups, dtype=np.uint64) * np.uint64(TILE)
    times = []
    for it in range(warmup + steps):
        secs, _, errs = orc.cpu_sched_run(dag.tasks, dag.succ, tiles, dag.ready, cores)
        assert secs > 0 and errs == 0
        if it >= warmup:
            times.append(secs)
    total = sum(times)
    return dag.ntasks * len(times) / total, cores, total / len(times), dag.ntasks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--groups", type=int, default=K_GROUPS, help="broadcast groups per GPU (config value: 4096)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample-groups", type=int, default=1024)
    ap.add_argument("--mgpu", default="direct", choices=["direct", "exchange"],
                    help="N>1: device-released cross-GPU edges (default) or two windows + one NCCL exchange")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K = args.groups
    ntasks = K * (1 + F)
    algo_bytes = K * (1 + F) * TILE
    cfg = {"workload": "Ex05_Broadcast dataflow (BASELINE configs[1]), 256x256 fp32 tiles, K=%d groups/GPU, fan-out %d" % (K, F),
           "tile_bytes": TILE, "groups_per_gpu": K, "tasks_per_gpu_step": ntasks, "distribution": "two_dim_block_cyclic 1x%d" % world,
           "l2": "inputs larger than L2: %.2f GiB of tiles per GPU vs 126 MB L2, FIFO ready order" % (K * TILE / 2 ** 30)}

    if args.impl == "reference":
        if rank != 0:
            return
        v, cores, sec, nt = cpu_reference_arm(args.steps, args.warmup, args.cpu_sample_groups)
        sample = "%d of %d groups per step (%d tasks), all tiles in host memory" % (args.cpu_sample_groups, K, nt)
        print(json.dumps({"impl": "reference", "metric": "tasks/s", "value": v, "unit": "tasks/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic", "config": cfg,
                          "cpu_baseline": {"value": v, "unit": "tasks/s", "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": v, "unit": "tasks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "tile_gbs": v * TILE / 1e9}))
        return

    import torch
    import torch.distributed as dist
    from parsec_b200 import _lib as L
    from parsec_b200 import runtime as R
    from parsec_b200.engine import Window
    from parsec_b200.multigpu import ex05_multi_gpu_step_factory, ex05_direct_step_factory, work_stream

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; parsec_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    host = np.zeros(K * TILE // 4, np.int32)
    ctx = R.Context(nb_cores=os.cpu_count() or 1, cuda_devices=(local_rank,))
    dev = ctx.devices[0]
    dc = ctx.block_cyclic(4, TILE // 4, 1, K * TILE // 4, 1, mat=host)
    assert ctx.l.pb2_dc_register_memory(dc, dev) == 0            # twoDBC_memory_register: pin the collection once

    # ---------------------------------------------------------------- e2e through the host API, host buffers
    # One step = the application hands a freshly written collection (host memory) to a new task pool:
    #   host_write_all  : the host copies are the newest version (what CPU producer tasks would leave behind)
    #   ptg_new + wait  : PTG front end -> kernel_scheduler -> 4 pipelined windows (two in flight); every tile comes
    #                     from pinned host memory (H2D, K * 262144 B: strided runs through the copy engine, anything
    #                     else staged by the worker CTAs), bodies run, successors are released on-device
    #   task_info       : the per-task results (what TaskRecv "prints") are read back to the host (D2H)
    tsplit = {"new": 0.0, "wait": 0.0, "read": 0.0}

    def e2e_step():
        t0 = time.perf_counter()
        assert ctx.l.pb2_dc_host_write_all(dc) == 0
        tp = C.c_void_p(ctx.l.pb2_ptg_ex05_broadcast_new(ctx.h, dc, K, NB))
        t1 = time.perf_counter()
        ctx.wait()
        t2 = time.perf_counter()
        info = ctx.task_info(tp)
        recv = info["class_id"] == 1
        ok = bool(np.all(info["result"][recv] == info["locals"][recv, 0].astype(np.uint64)))   # observed k, 0 mismatches
        ctx.l.pb2_taskpool_free(tp)
        t3 = time.perf_counter()
        tsplit["new"] += t1 - t0; tsplit["wait"] += t2 - t1; tsplit["read"] += t3 - t2
        return ok

    if world == 1:
        for _ in range(2):
            assert e2e_step()
        for k in tsplit: tsplit[k] = 0.0
        h2d0 = ctx.stats(dev)["data_in_from_device"][0]
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            assert e2e_step()
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / args.e2e_steps
        h2d_tiles = (ctx.stats(dev)["data_in_from_device"][0] - h2d0) // args.e2e_steps
        assert h2d_tiles == K * TILE                                      # every tile came from host memory, once
        h2d_step = int(h2d_tiles) + ntasks * 64 + K * F * 4 + K * 32 + K * 4   # tiles (in-kernel) + descriptors
        d2h_step = ntasks * (4 + 16 + 8)                                  # retire log + flow versions + results
        e2e = {"value": ntasks / e2e_s, "unit": "tasks/s", "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step,
               "ms_per_step": e2e_s * 1e3, "tile_gbs": algo_bytes / e2e_s / 1e9,
               "ms_split": {k: v / args.e2e_steps * 1e3 for k, v in tsplit.items()}}
    else:
        # N > 1: the split window is built once and re-armed every step; per step every rank's tiles are staged in
        # from its pinned host buffer inside the kernel (H2D), cross-GPU edges are released by the device, and the
        # per-task results + retire log are read back (D2H).
        from parsec_b200.engine import Engine
        eng = Engine(local_rank)
        eng.use_stream(work_stream(torch))
        alias = eng.host_register(host)
        estep, efinish, _, _ = ex05_direct_step_factory(K, NB, TILE, rank, world, local_rank, eng=eng, host_tiles=alias)

        def e2e_step_n():
            host[::TILE // 4] += 1                                        # the application rewrites its tiles
            estep()
            run = efinish()
            res = run.w.results()
            tl = run.p["tasks"]
            recv = tl["class_id"] == 1
            assert run.w.stats["bytes_h2d"] == K * TILE
            return bool(np.all(res["result"][recv] == tl["locals"][recv, 0].astype(np.uint64)))

        for _ in range(2):
            assert e2e_step_n()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            assert e2e_step_n()
        torch.cuda.synchronize()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
        e2e = {"value": world * ntasks / e2e_s, "unit": "tasks/s", "h2d_bytes_per_step": world * K * TILE,
               "d2h_bytes_per_step": world * ntasks * (4 + 16 + 8), "ms_per_step": e2e_s * 1e3,
               "tile_gbs": world * algo_bytes / e2e_s / 1e9,
               "note": "split window built once and re-armed per step; tiles staged from pinned host memory in-kernel"}
        eng.synchronize()

    # ---------------------------------------------------------------- device-resident value
    if world == 1:
        tp = C.c_void_p(ctx.l.pb2_ptg_ex05_broadcast_new(ctx.h, dc, K, NB))
        win = ctx.export_window(tp, dev)                          # the window the module builds for this pool
        assert len(win["tasks"]) == ntasks
        from parsec_b200.engine import Engine
        eng = Engine(local_rank)
        slab = eng.malloc(K * TILE)
        tiles = win["tiles"].copy()
        order = np.argsort(tiles["src_ptr"])
        tiles["dev_ptr"][order] = slab + np.arange(K, dtype=np.uint64) * np.uint64(TILE)
        tiles["state"] = L.TILE_VALID                             # inputs already resident in HBM
        w = eng.window(0, win["tasks"], win["succ"], tiles, win["ready"])
        step = lambda: w.launch()
        finish = lambda: w.wait()
        launches_per_step = 2
    elif args.mgpu == "direct":
        step, finish, _, nt_rank = ex05_direct_step_factory(K, NB, TILE, rank, world, local_rank)
        launches_per_step = 2                                     # re-arm + persistent kernel (the NCCL barrier kernel is not ours)
        assert nt_rank == ntasks
        cfg["multi_gpu"] = "one window per GPU; cross-GPU edges released by the producer's CTA (system-scope atomics over NVLink), tiles pulled by the consumer; NCCL only as the per-step barrier"
    else:
        step, finish, launches_per_step = ex05_multi_gpu_step_factory(ctx, dev, dc, K, NB, TILE, rank, world, local_rank)
        cfg["multi_gpu"] = "two windows per GPU + one batched NCCL send/recv of the frontier"

    for _ in range(args.warmup):
        step()
    finish()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    kernel_ms = 0.0
    ev0.record()
    for _ in range(args.steps):
        step()
        if world == 1:
            kernel_ms += w.wait()["kernel_ms"] + w.stats["reset_ms"]
    ev1.record()
    finish()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if world == 1:
        # the engine runs on its own stream: the CUDA events recorded there by the library are the device time
        dev_ms = kernel_ms
        st = w.stats
        assert st["body_errors"] == 0 and st["tasks_retired"] == ntasks
        only_kernel_ms = st["kernel_ms"]
    else:
        dev_ms = ev0.elapsed_time(ev1)                             # engine work is enqueued on torch's current stream
    ms_per_step = max_over_ranks(dev_ms / args.steps)
    value = world * ntasks / (ms_per_step / 1e3)

    out = {"metric": "tasks/s", "value": value, "unit": "tasks/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32",
           "data": "synthetic", "config": cfg, "tile_gbs": world * algo_bytes / (ms_per_step / 1e3) / 1e9,
           "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
           "engine": {"hbm_worker": "64 threads x 24 per SM", "gemm_worker": "CTA pair, cta_group::2, fused k-chains"}}
    if world == 1:
        peak, how = peaks()
        ach = algo_bytes / (only_kernel_ms / 1e3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_ex05_traffic.json")   # from the committed ncu --set full capture
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        out["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                           "kernel": "pb2_engine_hbm_kernel", "kernel_ms": only_kernel_ms, "algorithmic_bytes": algo_bytes,
                           "peak_source": how,
                           "note": "algorithmic bytes = K*(1+F)*262144; frac can exceed 1 when successor reads hit the 126 MB L2"}
        v, cores, sec, nt = cpu_reference_arm(3, 1, args.cpu_sample_groups)
        out["cpu_baseline"] = {"value": v, "unit": "tasks/s", "cores": cores, "kind": "port",
                               "sample": "%d of %d groups (%d tasks) x 3 runs, tiles in host memory, oracle/orc_cpu_sched.c" % (args.cpu_sample_groups, K, nt)}
    out["e2e"] = e2e
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
