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


def test_integration_md_ctypes_stub_runs():
    """The binding printed in INTEGRATION.md (examples/ctypes_stub.py is the same code) against the oracle."""
    import importlib.util

    import numpy as np
    import torch

    from oracle import flat_search as ofs
    from tests import util

    spec = importlib.util.spec_from_file_location("ctypes_stub", os.path.join(ROOT, "examples", "ctypes_stub.py"))
    stub = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(stub)
    rows = util.unit_rows(30_000, 512, seed=3)
    dev_rows = torch.from_numpy(rows).cuda().to(torch.bfloat16)
    seg = [0, 10_000, 30_000]
    table = stub.ResidentTable(dev_rows, seg, [f"c{i}" for i in range(30_000)], ['{"i": %d}' % i for i in range(30_000)],
                               {"alice": 0, "bob": 1})
    q = util.queries(1, 512)[0]
    hits = table.search("bob", q.tolist(), 12)
    d, ids = ofs.flat_search(rows, q, 12, lo=10_000, hi=30_000)
    assert len(hits) == 12 and hits[0]["chunk_id"] == f"c{ids[0]}"
    assert abs(hits[0]["score"] - (1.0 - float(d[0]))) < util.TOL_BF16
    assert hits[0]["meta"] == {"i": int(ids[0])}
    assert table.search("nobody", q.tolist(), 12) == []
    assert len(table.search("alice", q.tolist(), 0)) == 1
