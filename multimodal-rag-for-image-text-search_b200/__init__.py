"""B200-native exact-scan hot path for Multimodal-RAG (retrieve_text / retrieve_images).

The directory name carries hyphens (it is fixed by the build contract), so import it with
``importlib.import_module("multimodal-rag-for-image-text-search_b200")`` or through the ``mmr_b200`` alias
module at the repository root.
"""
from . import _native
from ._native import NativeError
from .build import build as build_native
from .index import MultiIndex, ResidentIndex, fuse, fuse_f64, merge_topk
from .sharded import ShardedIndex, shard_bounds
from .store import B200Store, VectorRow, make_arrow_table
from .settings import RetrievalSettings, load_retrieval_settings
from .batcher import MicroBatcher

__all__ = [
    "ResidentIndex", "MultiIndex", "B200Store", "VectorRow", "RetrievalSettings", "load_retrieval_settings", "NativeError",
    "merge_topk", "fuse", "fuse_f64", "ShardedIndex", "shard_bounds", "make_arrow_table", "build_native", "MicroBatcher",
]
