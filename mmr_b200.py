"""Importable alias for the package whose directory name has hyphens."""
import importlib
import sys

_pkg = importlib.import_module("multimodal-rag-for-image-text-search_b200")
sys.modules[__name__] = _pkg
