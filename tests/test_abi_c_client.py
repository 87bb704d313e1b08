"""The drop-in boundary without Python in the way: tests/abi_client/client.c is a plain C program (gcc, cudart) that
drives libcnrma_b200.so through include/cnrma_b200.h with cudaMalloc'd buffers and checks Stage A against the C oracle
bit for bit.  On a box without a GPU it must still build, link, load the library and fail cleanly."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _build(tmp_path):
    import cnrma_b200
    import oracle
    cnrma_b200.build()
    oracle.build()
    exe = str(tmp_path / "abi_client")
    libdir = os.path.join(ROOT, "cn-rma_b200")
    odir = os.path.join(ROOT, "oracle")
    cmd = [shutil.which("gcc") or "gcc", "-std=c99", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           os.path.join(ROOT, "tests", "abi_client", "client.c"), "-o", exe,
           "-L", libdir, "-lcnrma_b200", "-L", odir, "-lcnrma_oracle", "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-lm",
           "-Wl,-rpath," + libdir, "-Wl,-rpath," + odir, "-Wl,-rpath," + os.path.join(CUDA, "lib64")]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_c_client_builds_and_fails_cleanly_without_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "abi version 2" in out.stdout
    if not torch.cuda.is_available():
        assert "no usable device" in out.stdout


@pytest.mark.gpu
def test_c_client_stage_a_matches_oracle(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches 0" in out.stdout
