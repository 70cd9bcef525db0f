"""The drop-in boundary from plain C: tests/c_abi/scan_smoke.c (no Python, no torch in the caller) is compiled with gcc
against include/ss2d_b200.h + libss2d_b200.so, run on the GPU, and checks forward and all gradients of both kernel families
against the C oracle it links in."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi", "scan_smoke.c")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _compile(out):
    lib = os.path.join(ROOT, "focalnet_b200", "lib")
    cmd = ["gcc", "-O2", "-std=c11", SRC, os.path.join(ROOT, "oracle", "ss2d_oracle.c"), "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(CUDA, "include"), "-L" + lib, "-lss2d_b200", "-L" + os.path.join(CUDA, "lib64"), "-lcudart", "-lm", "-fopenmp",
           "-Wl,-rpath," + lib, "-Wl,-rpath," + os.path.join(CUDA, "lib64"), "-o", out]
    return subprocess.run(cmd, capture_output=True, text=True)


def test_c_caller_compiles_and_links_against_the_header(tmp_path):
    """CPU: the header is valid C11 and every symbol the C caller uses resolves in libss2d_b200.so."""
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    r = _compile(str(tmp_path / "scan_smoke"))
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_c_caller_matches_the_oracle_on_the_gpu(tmp_path):
    exe = str(tmp_path / "scan_smoke")
    r = _compile(exe)
    assert r.returncode == 0, r.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "C ABI smoke: OK" in run.stdout and "family 1" in run.stdout and "family 2" in run.stdout
