"""Two GPUs, one process per GPU: windows split by pb2_partition_*, cross-GPU edges released by the device.

Launched by the test as `python -m torch.distributed.run --nproc-per-node 2` on 127.0.0.1; skipped on a box with
one GPU.  Parity: every rank's per-task results equal the sequential oracle's run of the UNSPLIT window.
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _run(world, case, extra=()):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29611 + world), os.path.join(ROOT, "tests", "mgpu_worker.py"), case, *extra]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stdout[-2000:]
    out = json.loads(lines[-1])
    out["log"] = [l for l in r.stdout.splitlines() if l.startswith("rank")][:20]
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["ex05", "rtt", "random_dtd", "cholesky", "cholesky_unfused"])
def test_direct_path_two_gpus(case):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    out = _run(2, case)
    assert out["ok"] and out["world"] == 2, out
    assert out.get("exact_range", True), out
    if case != "ex05":            # NB = 14, n even: with two ranks every TaskRecv(k, n) lives on TaskBcast(k)'s rank
        assert out["remote_edges"] > 0 and out["bytes_d2d"] > 0, out


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["ex05", "rtt", "random_dtd", "cholesky"])
def test_direct_path_four_gpus(case):
    if _ngpus() < 4:
        pytest.skip("needs 4 GPUs")
    out = _run(4, case)
    assert out["ok"] and out["world"] == 4, out
    assert out["remote_edges"] > 0 and out["bytes_d2d"] > 0, out
