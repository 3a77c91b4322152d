"""Small driver for ncu: the device-resident Ex05 window, engine only (no 95% slab), 5 launches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from parsec_b200 import _lib as L
from parsec_b200.engine import Engine
from oracle import orc_dags as dags
K, NB, TILE = 4096, 14, 256*256*4
dag = dags.ex05_broadcast(K, NB, TILE)
with Engine(0) as eng:
    slab = eng.malloc(K*TILE)
    tiles = np.zeros(K, L.TILE_DTYPE); tiles["dev_ptr"] = slab + np.arange(K, dtype=np.uint64)*np.uint64(TILE); tiles["bytes"]=TILE; tiles["state"]=2
    w = eng.window(0, dag.tasks, dag.succ, tiles, dag.ready)
    for _ in range(5):
        st = w.run()
    print("kernel_ms", st["kernel_ms"], "errors", st["body_errors"])
    w.close()
