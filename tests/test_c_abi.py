"""The C ABI used from plain C: tests/c/*.c are compiled with gcc -std=c99 -pedantic against include/*.h and
libparsec_b200.so, the way a PaRSEC component or application written in C would use them."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "parsec_b200")


def _build(tmp_path, name):
    exe = str(tmp_path / name)
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c", name + ".c"), "-o", exe, "-L" + LIBDIR, "-lparsec_b200", "-Wl,-rpath," + LIBDIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_headers_are_c99(tmp_path):
    src = tmp_path / "hdr.c"
    src.write_text('#include "pb2_engine.h"\n#include "pb2_parsec.h"\nint main(void) { return 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                        "-fsyntax-only", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_ex05_from_c_dry_run(tmp_path):
    exe = _build(tmp_path, "ex05_broadcast")
    r = subprocess.run([exe, "24", "14", "1024", "1"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_ex05_from_c_on_the_gpu(tmp_path):
    exe = _build(tmp_path, "ex05_broadcast")
    r = subprocess.run([exe, "256", "14", "262144", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr
