"""Small driver for ncu: the device-resident DTD GEMM window (config 3: NT=32, 512^2 bf16), default kernel mode."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from parsec_b200 import _lib as L
from oracle import orc_dags as dags
from parsec_b200.engine import Engine
NT, T = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 512
tb = T * T * 2
with Engine(0) as e:
    dag = dags.dtd_gemm(NT, T)
    dag.tasks["access"][:, 2] &= ~np.uint8(L.FLOW_PUSHOUT)
    slab = e.malloc(dag.ntiles * tb)
    # random operands (uniform in [-1, 1), bf16): constant data toggles few bits and clocks higher than real data
    from parsec_b200.bf16 import f32_to_bf16_bits
    rng = np.random.default_rng(2026)
    e.h2d(slab, f32_to_bf16_bits(rng.uniform(-1.0, 1.0, dag.ntiles * tb // 2).astype(np.float32)))
    tiles = np.zeros(dag.ntiles, L.TILE_DTYPE)
    tiles["dev_ptr"] = slab + np.arange(dag.ntiles, dtype=np.uint64) * np.uint64(tb)
    tiles["bytes"], tiles["state"] = tb, L.TILE_VALID
    w = e.window(1, dag.tasks, dag.succ, tiles, dag.ready)
    for _ in range(4):
        st = w.run()
    print("kernel_ms", st["kernel_ms"], "TFLOP/s", 2.0 * (NT * T) ** 3 / st["kernel_ms"] / 1e9)
    w.close()
