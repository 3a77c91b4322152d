import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from parsec_b200 import _lib as L, runtime as R
from parsec_b200.engine import Engine
from oracle import orc_dags as dags
K, NB, TILE = 4096, 14, 256*256*4
def timeit(tag, eng, tasks, succ, tiles, ready):
    w = eng.window(0, tasks, succ, tiles, ready)
    for _ in range(3): st = w.run()
    print(tag, "kernel_ms", [round(w.run()["kernel_ms"],3) for _ in range(4)])
    w.close()
mode = sys.argv[1]
if mode == "A":   # oracle builder + small engine only
    eng = Engine(0); slab = eng.malloc(K*TILE)
    dag = dags.ex05_broadcast(K, NB, TILE)
    tiles = np.zeros(K, L.TILE_DTYPE); tiles["dev_ptr"] = slab + np.arange(K, dtype=np.uint64)*np.uint64(TILE); tiles["bytes"]=TILE; tiles["state"]=2
    timeit("A builder/no-ctx", eng, dag.tasks, dag.succ, tiles, dag.ready)
if mode in ("B", "C"):   # module window; B: big ctx slab present, C: tiny ctx slab
    host = np.zeros(K*TILE//4, np.int32)
    mca = {} if mode == "B" else {"device_cuda_memory_number_of_blocks": 16}
    ctx = R.Context(cuda_devices=(0,), mca=mca)
    dc = ctx.block_cyclic(4, TILE//4, 1, K*TILE//4, 1, mat=host)
    tp = C.c_void_p(ctx.l.pb2_ptg_ex05_broadcast_new(ctx.h, dc, K, NB))
    win = ctx.export_window(tp, ctx.devices[0])
    eng = Engine(0); slab = eng.malloc(K*TILE)
    tiles = win["tiles"].copy(); order = np.argsort(tiles["src_ptr"])
    tiles["dev_ptr"][order] = slab + np.arange(K, dtype=np.uint64)*np.uint64(TILE); tiles["state"] = 2
    timeit(mode+" module window", eng, win["tasks"], win["succ"], tiles, win["ready"])
    dag = dags.ex05_broadcast(K, NB, TILE)
    t2 = np.zeros(K, L.TILE_DTYPE); t2["dev_ptr"] = slab + np.arange(K, dtype=np.uint64)*np.uint64(TILE); t2["bytes"]=TILE; t2["state"]=2
    timeit(mode+" builder window, ctx alive", eng, dag.tasks, dag.succ, t2, dag.ready)
    d = win["tasks"]; print("first tasks", d[:2], "recv", d[K:K+2]); print("succ", win["succ"][:10], "ready", win["ready"][:5])
