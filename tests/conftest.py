"""pytest configuration: marker registration and import paths.

``-m "not gpu"`` runs here (no GPU): oracle vs golden vectors, host logic, C-ABI symbol checks,
world_size-2 gloo tests.  ``-m gpu`` runs on a B200: the parity tests proper, through the C-ABI.
"""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def pytest_sessionstart(session):
    """The C-ABI library is a build artefact (git-ignored): compile it in-tree when it is missing or older than its
    sources, so a fresh checkout can run the suite without a separate build step.  nvcc cross-compiles sm_100a
    without a GPU; on a box without nvcc the prebuilt .so that travelled with the repo is used as is."""
    import shutil

    builder = importlib.import_module("multimodal-rag-for-image-text-search_b200.build")
    if builder.needs_build() and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        builder.build(verbose=False)


@pytest.fixture(scope="session")
def pkg():
    """The product package (its directory name has hyphens, so it is imported by string)."""
    return importlib.import_module("multimodal-rag-for-image-text-search_b200")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
