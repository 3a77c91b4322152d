"""GPU tests of the tensor-core GEMM window (BASELINE config 3 shape, reduced NT)."""
import numpy as np
import pytest

from parsec_b200 import _lib as L
from oracle import orc_dags as dags
from parsec_b200.bf16 import bf16_bits_to_f32, f32_to_bf16_bits, round_to_bf16

pytestmark = pytest.mark.gpu


def gemm_chain_reference(A, B, C, NT, T):
    """Per-task semantics: C(i,j) <- bf16( f32(C(i,j)) + A(i,k) @ B(k,j)^T [fp32 accumulate] ), k = 0..NT-1.
    A[i,k] is [T][T] row-major (M x K), B[k,j] is stored [N][K].  Returns final C and the per-element
    running max magnitude (for the tolerance)."""
    out = np.empty_like(C)
    mag = np.empty_like(C)
    for i in range(NT):
        for j in range(NT):
            c = C[i, j].copy()
            m = np.abs(c)
            for k in range(NT):
                c = round_to_bf16(c + A[i, k].astype(np.float32) @ B[k, j].astype(np.float32).T)
                m = np.maximum(m, np.abs(c))
            out[i, j], mag[i, j] = c, m
    return out, mag


@pytest.mark.parametrize("NT,T", [(1, 128), (2, 256), (3, 512), (2, 320)])
def test_dtd_gemm_chain(engine, NT, T):
    rng = np.random.default_rng(1789 + NT * 1000 + T)
    def rnd(shape):
        return round_to_bf16(rng.uniform(-0.5, 0.5, shape).astype(np.float32))
    A, B, C = rnd((NT, NT, T, T)), rnd((NT, NT, T, T)), rnd((NT, NT, T, T))
    dag = dags.dtd_gemm(NT, T)
    tb = T * T * 2
    host = np.concatenate([f32_to_bf16_bits(A).ravel(), f32_to_bf16_bits(B).ravel(), f32_to_bf16_bits(C).ravel()])
    assert host.nbytes == dag.ntiles * tb
    slab = engine.malloc(dag.ntiles * tb)
    alias = engine.host_register(host)
    tiles = np.zeros(dag.ntiles, L.TILE_DTYPE)
    tiles["dev_ptr"] = slab + np.arange(dag.ntiles, dtype=np.uint64) * np.uint64(tb)
    tiles["src_ptr"] = alias + np.arange(dag.ntiles, dtype=np.uint64) * np.uint64(tb)
    tiles["bytes"] = tb
    w = engine.window(1, dag.tasks, dag.succ, tiles, dag.ready)
    st = w.run()
    res = w.results()
    w.close()
    assert st["tasks_retired"] == NT ** 3
    assert all(v == 0 for v in dags.check_execution(dag, res).values())
    assert st["bytes_h2d"] == dag.ntiles * tb and st["bytes_d2h"] == NT * NT * tb
    # every C(i,j) chain ran in k order: task (i,j,k) saw version k of C
    assert np.array_equal(res["seen_version"][:, 2], np.tile(np.arange(NT), NT * NT))
    got = bf16_bits_to_f32(host[2 * NT * NT * T * T:]).reshape(NT, NT, T, T)   # pushed out on the last k
    ref, mag = gemm_chain_reference(A, B, C, NT, T)
    # tolerance: 2 bf16 ulps (2^-7 relative) of the largest magnitude the element took along its chain;
    # fp32 accumulation order inside the tensor core differs from numpy's, which can flip a rounding.
    tol = 2.0 ** -7 * np.maximum(mag, 1.0)
    bad = np.abs(got - ref) > tol
    assert not bad.any(), f"{bad.sum()} / {bad.size} elements out of tolerance, max err {np.abs(got - ref).max()}"
    engine.host_unregister(host)
    engine.free(slab)
