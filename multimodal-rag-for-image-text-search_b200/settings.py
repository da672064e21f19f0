"""Retrieval knobs read by the hot path, same names / env vars / defaults as the reference
(config.py:42-50 RetrievalDefaults, app/settings.py:95-103 RetrievalSettings, :198-205 env loader)."""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Mapping, Optional

_TRUE = {"1", "true", "yes", "on", "y", "t"}
_FALSE = {"0", "false", "no", "off", "n", "f"}


def _env_bool(env: Mapping[str, str], key: str, default: bool) -> bool:
    raw = env.get(key)
    if raw is None:
        return default
    low = raw.strip().lower()
    if low in _TRUE:
        return True
    if low in _FALSE:
        return False
    return default


def _env_num(env: Mapping[str, str], key: str, default, cast):
    raw = env.get(key)
    if raw is None:
        return default
    try:
        return cast(raw)
    except (TypeError, ValueError):
        return default


@dataclass(frozen=True)
class RetrievalSettings:
    use_rerank: bool = True        # RERANK_ENABLED
    index_topk_text: int = 50      # INDEX_TOPK_TEXT
    index_topk_image: int = 12     # INDEX_TOPK_IMG
    rerank_topk: int = 8           # RERANK_TOPK
    final_n: int = 4               # FINAL_N
    confidence_tau: float = 0.25   # CONFIDENCE_TAU


def load_retrieval_settings(env: Optional[Mapping[str, str]] = None) -> RetrievalSettings:
    env = os.environ if env is None else env
    d = RetrievalSettings()
    return RetrievalSettings(
        use_rerank=_env_bool(env, "RERANK_ENABLED", d.use_rerank),
        index_topk_text=_env_num(env, "INDEX_TOPK_TEXT", d.index_topk_text, int),
        index_topk_image=_env_num(env, "INDEX_TOPK_IMG", d.index_topk_image, int),
        rerank_topk=_env_num(env, "RERANK_TOPK", d.rerank_topk, int),
        final_n=_env_num(env, "FINAL_N", d.final_n, int),
        confidence_tau=_env_num(env, "CONFIDENCE_TAU", d.confidence_tau, float),
    )
