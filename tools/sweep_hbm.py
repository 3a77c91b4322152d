"""Sweep engine launch parameters on the device-resident Ex05 window (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from parsec_b200 import _lib as L
from parsec_b200.engine import Engine
from oracle import orc_dags as dags
K, NB, TILE = 4096, 14, 256*256*4
dag = dags.ex05_broadcast(K, NB, TILE)
algo = K * 9 * TILE
cfgs = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [(4,256,0)]
for wps, thr, pol in cfgs:
    with Engine(0, workers_per_sm=wps, threads=thr, queue_policy=pol) as eng:
        slab = eng.malloc(K*TILE)
        tiles = np.zeros(K, L.TILE_DTYPE); tiles["dev_ptr"] = slab + np.arange(K, dtype=np.uint64)*np.uint64(TILE); tiles["bytes"]=TILE; tiles["state"]=2
        w = eng.window(0, dag.tasks, dag.succ, tiles, dag.ready)
        for _ in range(3): w.run()
        ms = min(w.run()["kernel_ms"] for _ in range(5))
        print(f"wps={wps} thr={thr} policy={pol} workers={eng.info()['nworkers']} kernel_ms={ms:.3f} tasks/s={dag.ntasks/ms*1e3:.3e} algoGB/s={algo/ms/1e6:.0f}", flush=True)
        w.close()
