"""Golden vectors: tests/golden/reference_runtime_cpu.json holds what the UNMODIFIED reference runtime computes for the test
applications with their CPU incarnations (generated in the build container by tests/golden/make_golden.py: counts, error
totals and the FNV-1a checksum of the final host data).  The CPU test re-derives them where the reference runtime is
built; the GPU tests run the same command lines through the b200 device component and require the same final data,
bit for bit."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_runtime_cpu.json")))
CPU_FLAG = {"ex05_b200": ["-m", "cpu"], "ex02_b200": ["-m", "cpu"], "stage_b200": ["-m", "cpu"], "dtd_b200": ["-C"], "batch_b200": ["-m", "cpu"]}
GPU_FLAG = {"ex05_b200": ["-m", "gpu"], "ex02_b200": ["-m", "gpu"], "stage_b200": ["-m", "gpu"], "dtd_b200": [], "batch_b200": ["-m", "gpu"]}


def run(app, args, env):
    exe = os.path.join(BIN, app)
    if not os.path.exists(exe):
        pytest.fail(f"{exe} is missing: run __graft_entry__.build() where /root/reference is mounted")
    e = dict(os.environ)
    for k in ("PARSEC_MCA_device_b200_enabled", "PARSEC_MCA_device_b200_dry_run", "PARSEC_MCA_device_cuda_enabled"):
        e.pop(k, None)
    e.update(env)
    p = subprocess.run([exe] + list(args), env=e, cwd="/tmp", capture_output=True, text=True, timeout=300)
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert lines, (p.returncode, p.stderr[-1500:])
    return p.returncode, json.loads(lines[-1])


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_reference_runtime_reproduces_the_golden_vectors(name):
    g = GOLDEN[name]
    rc, d = run(g["app"], CPU_FLAG[g["app"]] + g["args"], {"PARSEC_MCA_device_cuda_enabled": "0"})
    assert rc == g["rc"]
    for f, v in g["fields"].items():
        assert d[f] == v, (name, f, d[f], v)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_b200_component_matches_the_golden_vectors(name):
    g = GOLDEN[name]
    rc, d = run(g["app"], GPU_FLAG[g["app"]] + g["args"], {"PARSEC_MCA_device_b200_enabled": "1"})
    assert rc == g["rc"], d
    assert d["b200_modules"] == 1 and d["executed_on_gpu"] > 0        # the tasks really went through the component
    for f, v in g["fields"].items():
        assert d[f] == v, (name, f, d[f], v)
