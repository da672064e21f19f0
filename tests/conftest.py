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


@pytest.fixture(scope="session")
def pkg():
    """The product package (its directory name has hyphens, so it is imported by string)."""
    return importlib.import_module("multimodal-rag-for-image-text-search_b200")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
