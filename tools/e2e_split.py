"""Where the e2e step goes (host API, N=1): run with PB2_TIMING=1 to get the runtime's own split."""
import os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parsec_b200 import runtime as R

K, NB, TILE = 4096, 14, 262144
host = np.zeros(K * TILE // 4, np.int32)
ctx = R.Context(nb_cores=os.cpu_count() or 1, cuda_devices=(0,))
dev = ctx.devices[0]
dc = ctx.block_cyclic(4, TILE // 4, 1, K * TILE // 4, 1, mat=host)
assert ctx.l.pb2_dc_register_memory(dc, dev) == 0
ts = {"write": 0, "new": 0, "wait": 0, "read": 0, "free": 0}
N = 5
for it in range(N + 2):
    t0 = time.perf_counter()
    ctx.l.pb2_dc_host_write_all(dc)
    ta = time.perf_counter()
    tp = C.c_void_p(ctx.l.pb2_ptg_ex05_broadcast_new(ctx.h, dc, K, NB))
    t1 = time.perf_counter()
    ctx.wait()
    t2 = time.perf_counter()
    info = ctx.task_info(tp)
    t3 = time.perf_counter()
    ctx.l.pb2_taskpool_free(tp)
    t4 = time.perf_counter()
    if it >= 2:
        ts["write"] += ta - t0; ts["new"] += t1 - ta; ts["wait"] += t2 - t1; ts["read"] += t3 - t2; ts["free"] += t4 - t3
print({k: round(v / N * 1e3, 3) for k, v in ts.items()})
