"""Several task pools at once through the b200 component on the GPU (tests/parsec/ex05_main.c -P n): the final host data of
every collection must equal, bit for bit, what the reference runtime computes with its CPU incarnations for the same
command line.  Runs last (file name): it is the youngest GPU test of the suite."""
import pytest

from test_mca_component import run, CPU_ENV


@pytest.mark.gpu
def test_component_gpu_three_taskpools_at_once_match_the_cpu_run():
    K, P, rep = 256, 3, 2
    args = ["-K", K, "-t", 65536, "-c", 8, "-w", "-P", P]
    rc, c, _ = run("ex05_b200", args + ["-m", "cpu", "-r", 1], CPU_ENV)
    assert rc == 0 and c["errors"] == 0 and c["pools"] == P
    rc, d, err = run("ex05_b200", args + ["-m", "gpu", "-r", rep],
                     {"PARSEC_MCA_device_b200_enabled": "1", "PARSEC_MCA_device_b200_nvtx": "1"})
    assert rc == 0 and d["errors"] == 0 and d["b200"]["check_mismatches"] == 0, err[-1000:]
    assert d["executed_on_gpu"] == K * 9 * P * rep and d["b200"]["tasks_engine"] == K * 9 * P * rep
    assert d["checksum"] == c["checksum"]
