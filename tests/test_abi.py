"""The C-ABI library loads here (no GPU) and exports every symbol include/mmr_b200.h declares; argument
errors are reported through status codes + mmr_last_error; device work fails loudly without a GPU."""
import ctypes as C
import importlib
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "multimodal-rag-for-image-text-search_b200"
native = importlib.import_module(PKG + "._native")


def _declared():
    text = open(os.path.join(ROOT, "include", "mmr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mmr_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.exists(native.LIB_PATH), "run __graft_entry__.build() first"
    assert os.path.dirname(native.LIB_PATH).startswith(ROOT)


def test_exports_every_declared_symbol():
    handle = C.CDLL(native.LIB_PATH)
    declared = _declared()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/mmr_b200.h but not exported"
    bound = {s[0] for s in native.SYMBOLS}
    assert bound == set(declared), f"binding and header disagree: {bound ^ set(declared)}"
    assert handle.mmr_abi_version() == native.ABI_VERSION


def test_argument_errors_have_messages():
    lib = native.lib()
    out = C.c_void_p()
    assert lib.mmr_index_create(0, 512, 99, 10, None, None, 0, 0, C.byref(out)) == 1  # MMR_ERR_INVALID
    assert b"dtype" in lib.mmr_last_error()
    assert lib.mmr_index_create(0, 500, 0, 0, None, None, 0, 0, C.byref(out)) == 3    # MMR_ERR_UNSUPPORTED
    assert b"384" in lib.mmr_last_error()
    assert lib.mmr_search(None, None, None, 1, 10, None, None, None, 0, None) == 1
    assert lib.mmr_merge_topk(None, None, 0, 1, 10, None, None, None) == 1
    assert lib.mmr_fuse(None, None, 0, None, None, 0, 0, 4, 0.25, None, None, None, None, None, None) == 1
    assert lib.mmr_search_workspace_bytes(None, 1, 10) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    lib = native.lib()
    out = C.c_void_p()
    rc = lib.mmr_index_create(0, 512, 0, 0, None, None, 0, 0, C.byref(out))
    assert rc == 2, "without a GPU the library must fail with MMR_ERR_CUDA, not fall back"   # MMR_ERR_CUDA
    pkg = importlib.import_module(PKG)
    with pytest.raises(pkg.NativeError):
        pkg.B200Store()
    with pytest.raises(pkg.NativeError):
        pkg.ResidentIndex(torch.zeros(8, 512, dtype=torch.bfloat16))
