"""The C ABI used from plain C (no Python, no torch in the client): compile tests/c/abi_smoke.c with gcc against
include/mmr_b200.h + libmmr_b200.so + libcudart and run it."""
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "multimodal-rag-for-image-text-search_b200")


def test_plain_c_client(tmp_path):
    gcc = shutil.which("gcc")
    cuda = "/usr/local/cuda"
    if gcc is None or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime.h")):
        pytest.skip("gcc / CUDA headers not available")
    exe = str(tmp_path / "abi_smoke")
    cmd = [gcc, "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", exe, "-L", PKG_DIR, "-lmmr_b200",
           "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm", f"-Wl,-rpath,{PKG_DIR}", f"-Wl,-rpath,{cuda}/lib64"]
    build = subprocess.run(cmd, capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "abi_smoke: ok" in run.stdout
