"""Quick device-resident timing of the DTD GEMM window (development aid, not the bench)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from parsec_b200 import _lib as L
from oracle import orc_dags as dags
from parsec_b200.engine import Engine

NT = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 512
tb = T * T * 2
with Engine(0) as e:
    dag = dags.dtd_gemm(NT, T)
    dag.tasks["access"][:, 2] &= ~np.uint8(L.FLOW_PUSHOUT)
    slab = e.malloc(dag.ntiles * tb)
    init = np.zeros(dag.ntiles * tb // 2, np.uint16)
    init[:] = 0x3C00  # small bf16 values
    e.h2d(slab, init)
    tiles = np.zeros(dag.ntiles, L.TILE_DTYPE)
    tiles["dev_ptr"] = slab + np.arange(dag.ntiles, dtype=np.uint64) * np.uint64(tb)
    tiles["bytes"] = tb
    tiles["state"] = L.TILE_VALID
    w = e.window(1, dag.tasks, dag.succ, tiles, dag.ready)
    flops = 2.0 * (NT * T) ** 3
    for it in range(4):
        st = w.run()
        print(f"NT={NT} T={T} it={it} kernel_ms={st['kernel_ms']:.3f} tasks/s={dag.ntasks/st['kernel_ms']*1e3:.3e} TFLOP/s={flops/st['kernel_ms']/1e9:.1f}")
    w.close()
