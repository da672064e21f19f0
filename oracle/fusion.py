"""CPU restatement of the text+image score fusion and the confidence gate.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  PINNED against outputs of the
reference's own function bodies (tests/golden/fusion_golden.json, made by
oracle/make_golden.py).

  * ``z_scores``        <- _z_scores         app/ml/retrieve.py:186-195
  * ``fuse_results``    <- _fuse_results     app/ml/retrieve.py:158-183
  * ``rerank_text``     <- _rerank_text      app/ml/retrieve.py:132-155 (model supplied by caller)
  * ``confidence_low``  <- _confidence_low   app/ml/generate.py:56-60
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Sequence

import numpy as np


def z_scores(values: Sequence[Optional[float]]) -> List[float]:
    """retrieve.py:186-195 -- mean/std (population, ddof=0) in float32; z in float64; None -> 0.0."""
    numeric = [v for v in values if v is not None]
    if not numeric:
        return []
    arr = np.array(numeric, dtype=np.float32)
    mean = float(arr.mean())
    std = float(arr.std())
    if std == 0:
        return [0.0 for _ in values]
    return [float((v - mean) / std) if v is not None else 0.0 for v in values]


def fuse_results(
    text_results: List[Dict[str, Any]],
    image_results: List[Dict[str, Any]],
    final_n: int = 4,
) -> List[Dict[str, Any]]:
    """retrieve.py:158-183.

    Keeps the reference's positional quirk: ``text_rerank_z`` is the z-score list of the
    *compacted* rerank scores and is indexed with the text item's position (:162,:173).
    """
    items: List[Dict[str, Any]] = []
    text_cos = [it["score"] for it in text_results]
    text_rr = [it.get("rerank_score") for it in text_results if "rerank_score" in it]
    image_cos = [it["score"] for it in image_results]

    text_cos_z = z_scores(text_cos)
    text_rr_z = z_scores(text_rr) if text_rr else []
    image_cos_z = z_scores(image_cos)

    for idx, it in enumerate(text_results):
        z_vals: List[float] = []
        if text_cos_z:
            z_vals.append(text_cos_z[idx])
        if text_rr_z and idx < len(text_rr_z):
            z_vals.append(text_rr_z[idx])
        combined = float(np.mean(z_vals)) if z_vals else it["score"]
        items.append({**it, "combined_score": combined})

    for idx, it in enumerate(image_results):
        z_val = image_cos_z[idx] if image_cos_z else it["score"]
        items.append({**it, "combined_score": float(z_val)})

    items.sort(key=lambda e: e["combined_score"], reverse=True)  # stable
    return items[:final_n]


def rerank_text(
    query: str,
    results: List[Dict[str, Any]],
    predict: Optional[Callable[[List[Any]], Sequence[float]]],
    use_rerank: bool = True,
    rerank_topk: int = 8,
) -> List[Dict[str, Any]]:
    """retrieve.py:132-155 with the cross-encoder passed in as ``predict`` (None = unavailable)."""
    if not results or not use_rerank:
        return results
    if not predict:
        return results
    top = results[:rerank_topk]
    if not top:
        return results
    pairs = [(query, it["text"]) for it in top if it.get("text")]
    if not pairs:
        return results
    scores = predict(pairs)
    for it, s in zip(top, scores):
        it["rerank_score"] = float(s)
    reranked = top + results[len(top):]
    reranked.sort(key=lambda it: it.get("rerank_score", it["score"]), reverse=True)
    return reranked


def confidence_low(items: List[Dict[str, Any]], tau: float = 0.25) -> bool:
    """generate.py:56-60."""
    if not items:
        return True
    top = max(it.get("combined_score", it.get("score", 0.0)) for it in items)
    return top < tau
