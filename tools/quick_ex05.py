"""Quick device-resident timing of the Ex05 window (development aid, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from parsec_b200 import _lib as L
from oracle import orc_dags as dags
from parsec_b200.engine import Engine

K = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
wps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
thr = int(sys.argv[3]) if len(sys.argv) > 3 else 256
tb = 256 * 256 * 4
with Engine(0, workers_per_sm=wps, threads=thr) as e:
    print(e.info())
    dag = dags.ex05_broadcast(K, 14, tb)
    slab = e.malloc(K * tb)
    host = np.zeros(K * tb // 4, np.int32)
    alias = e.host_register(host)
    for valid in (True, False):
        tiles = np.zeros(K, L.TILE_DTYPE)
        tiles["dev_ptr"] = slab + np.arange(K, dtype=np.uint64) * np.uint64(tb)
        tiles["src_ptr"] = alias + np.arange(K, dtype=np.uint64) * np.uint64(tb)
        tiles["bytes"] = tb
        tiles["state"] = L.TILE_VALID if valid else L.TILE_INVALID
        w = e.window(0, dag.tasks, dag.succ, tiles, dag.ready)
        for it in range(4):
            t0 = time.perf_counter()
            st = w.run()
            t1 = time.perf_counter()
            algo = K * (1 + dag.meta["F"]) * tb
            print(f"valid={valid} it={it} kernel_ms={st['kernel_ms']:.3f} reset_ms={st['reset_ms']:.3f} wall_ms={(t1-t0)*1e3:.3f} "
                  f"tasks/s={dag.ntasks/st['kernel_ms']*1e3:.3e} algoGB/s={algo/st['kernel_ms']/1e6:.1f} "
                  f"h2d={st['bytes_h2d']} errs={st['body_errors']}")
        w.close()
    dag = dags.ep(4096, 64)
    w = e.window(0, dag.tasks, dag.succ, np.zeros(0, L.TILE_DTYPE), dag.ready)
    for it in range(3):
        st = w.run()
        print(f"ep 4096x64: kernel_ms={st['kernel_ms']:.3f} tasks/s={dag.ntasks/st['kernel_ms']*1e3:.3e}")
