"""Host mirror of the reference's retrieval orchestration (app/ml/retrieve.py) and confidence gate
(app/ml/generate.py:56-60), with the scan behind `_LANCEDB_STORE` served by B200Store.

Same seams as the reference so its tests read the same (tests/test_retrieve.py:62-67 monkeypatches exactly
these module attributes): `_LANCEDB_STORE`, `_METADATA_STORE`, `embed_text_batch`, `embed_query_for_images`,
`_get_cross_encoder`, `get_index_version`, `settings`.

  retrieve_text   <- retrieve.py:41-69     retrieve_images <- :72-100     retrieve <- :103-117
  _get_embeddings <- :120-129              _rerank_text    <- :132-155
  _fuse_results   <- :158-183              _z_scores       <- :186-195
  _prepare_metadata <- :198-206            _confidence_low <- generate.py:56-60

With RERANK_ENABLED=false the fusion + gate can also run on the device for a whole batch of requests
(`retrieve_batch_device`, kernel K5); with rerank on, the cross-encoder sits between scan and fusion
(SURVEY 8a/a13), so the host functions below finish the job with identical arithmetic.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import cache as _cache
from .settings import load_retrieval_settings

settings = SimpleNamespace(retrieval=load_retrieval_settings())

# injected by the application (or by tests): the vector store, the chunk metadata store, the encoders
_LANCEDB_STORE: Any = None
_METADATA_STORE: Any = None
_CROSS_ENCODER: Any = None
embed_text_batch: Optional[Callable[[List[str]], np.ndarray]] = None
embed_query_for_images: Optional[Callable[[str], np.ndarray]] = None
# optional device-side models (encoders.TextQueryEncoder / ImageQueryEncoder / DeviceCrossEncoder): when set,
# retrieve_batch_device embeds the whole micro-batch on the GPU and the embeddings never visit the host
_DEVICE_TEXT_ENCODER: Any = None
_DEVICE_IMAGE_ENCODER: Any = None
_DEVICE_CROSS_ENCODER: Any = None

TEXT_DIM, IMAGE_DIM = 384, 512


def configure(store=None, metadata=None, text_encoder=None, image_query_encoder=None, cross_encoder=None,
              retrieval_settings=None, device_text_encoder=None, device_image_encoder=None,
              device_cross_encoder=None) -> None:
    """Wire the module-level seams in one call.  A device encoder also serves the matching host seam (it is callable with
    the reference's signature), so `retrieve()` and `retrieve_batch_device()` use the same model."""
    global _LANCEDB_STORE, _METADATA_STORE, embed_text_batch, embed_query_for_images, _CROSS_ENCODER
    global _DEVICE_TEXT_ENCODER, _DEVICE_IMAGE_ENCODER, _DEVICE_CROSS_ENCODER
    if device_text_encoder is not None:
        _DEVICE_TEXT_ENCODER = device_text_encoder
        text_encoder = text_encoder or device_text_encoder
    if device_image_encoder is not None:
        _DEVICE_IMAGE_ENCODER = device_image_encoder
        image_query_encoder = image_query_encoder or device_image_encoder
    if device_cross_encoder is not None:
        _DEVICE_CROSS_ENCODER = device_cross_encoder
        cross_encoder = cross_encoder or device_cross_encoder
    if store is not None:
        _LANCEDB_STORE = store
    if metadata is not None:
        _METADATA_STORE = metadata
    if text_encoder is not None:
        embed_text_batch = text_encoder
    if image_query_encoder is not None:
        embed_query_for_images = image_query_encoder
    if cross_encoder is not None:
        _CROSS_ENCODER = cross_encoder
    if retrieval_settings is not None:
        settings.retrieval = retrieval_settings


def get_index_version(user_id: str) -> int:
    """app/ml/index_build.py:40-43; served by the store's version table."""
    getter = getattr(_LANCEDB_STORE, "get_index_version", None)
    return int(getter(user_id)) if getter else 0


def _get_cross_encoder():
    """retrieve.py:29-38: a falsy value means "no reranker available" and rerank is skipped."""
    return _CROSS_ENCODER or False


def _get_embeddings(query: str) -> Tuple[np.ndarray, np.ndarray]:
    hit = _cache.get_query_embeddings(query)
    if hit:
        return hit
    text = embed_text_batch([query])
    image = embed_query_for_images(query)
    first = text[0] if text.size else np.zeros(TEXT_DIM, dtype=np.float32)
    _cache.set_query_embeddings(query, first, image)
    return first, image


def _prepare_metadata(chunk) -> Dict[str, Any]:
    meta = dict(getattr(chunk, "meta", None) or {})
    for key, attr in (("doc_id", "document_id"), ("modality", "modality"), ("page_no", "page_no"),
                      ("start_ts", "start_ts"), ("end_ts", "end_ts"), ("file_path", "file_path")):
        meta.setdefault(key, getattr(chunk, attr, None))
    return meta


def _join(raw: List[Dict[str, Any]], modality: str, need_text: bool) -> List[Dict[str, Any]]:
    """Per-hit metadata join and drop rules (retrieve.py:55-67 / 86-98).

    The reference issues one SELECT per hit (N+1); a metadata store that offers
    `get_chunks(ids) -> {id: chunk}` is asked once per result list instead."""
    out = []
    bulk = getattr(_METADATA_STORE, "get_chunks", None)
    found = bulk([hit["chunk_id"] for hit in raw]) if (bulk and raw) else None
    for hit in raw:
        chunk = found.get(hit["chunk_id"]) if found is not None else _METADATA_STORE.get_chunk(hit["chunk_id"])
        if not chunk or (need_text and not chunk.text):
            continue
        out.append({
            "chunk_id": chunk.id,
            "modality": modality,
            "score": float(hit["score"]),
            "metadata": _prepare_metadata(chunk),
            "text": chunk.text if need_text else None,
        })
    return out


def retrieve_text(user_id: str, query: str, top_k: Optional[int] = None) -> List[Dict[str, Any]]:
    top_k = top_k or settings.retrieval.index_topk_text
    version = get_index_version(user_id)
    key = f"text::{query}"
    hit = _cache.get_retrieval_results(user_id, key, version)
    if hit is not None:
        return hit
    vec, _ = _get_embeddings(query)
    if vec.size == 0:
        return []
    # (the reference passes vec.tolist(), retrieve.py:53; the store takes any float sequence, and handing it the ndarray
    #  skips a list round trip of ~40 us per request)
    results = _join(_LANCEDB_STORE.search_text(user_id, vec, top_k), "text", need_text=True)
    _cache.set_retrieval_results(user_id, key, version, results)
    return results


def retrieve_images(user_id: str, query: str, top_k: Optional[int] = None) -> List[Dict[str, Any]]:
    top_k = top_k or settings.retrieval.index_topk_image
    version = get_index_version(user_id)
    key = f"image::{query}"
    hit = _cache.get_retrieval_results(user_id, key, version)
    if hit is not None:
        return hit
    _, vec = _get_embeddings(query)
    if vec.size == 0:
        return []
    results = _join(_LANCEDB_STORE.search_image(user_id, vec, top_k), "image", need_text=False)
    _cache.set_retrieval_results(user_id, key, version, results)
    return results


def retrieve(user_id: str, query: str) -> List[Dict[str, Any]]:
    version = get_index_version(user_id)
    key = _cache.normalize_query(query)
    hit = _cache.get_retrieval_results(user_id, key, version)
    if hit is not None:
        return hit
    text = retrieve_text(user_id, query)
    images = retrieve_images(user_id, query)
    fused = _fuse_results(_rerank_text(query, text), images)
    _cache.set_retrieval_results(user_id, key, version, fused)
    return fused


def _embed_batch(queries: Sequence[str]):
    """Query embeddings of a micro-batch: on the device when device encoders are configured (the embeddings are born in
    HBM and feed the scan directly), else through the host seams of the reference (cached, retrieve.py:120-129)."""
    if _DEVICE_TEXT_ENCODER is not None and _DEVICE_IMAGE_ENCODER is not None:
        return _DEVICE_TEXT_ENCODER.encode_device(list(queries)), _DEVICE_IMAGE_ENCODER.encode_device(list(queries))
    vecs = [_get_embeddings(q) for q in queries]
    return np.stack([v[0] for v in vecs]), np.stack([v[1] for v in vecs])


def retrieve_batch_device(user_ids: Sequence[str], queries: Sequence[str]) -> List[Tuple[List[Dict[str, Any]], bool]]:
    """Micro-batched `retrieve` + `_confidence_low` for B concurrent requests with the fusion and the gate on the
    device.  One launch chain serves the whole batch: (query encoders,) text scan, image scan, (cross-encoder,) K5.
    Only the FINAL_N winners -- plus, with rerank on, the RERANK_TOPK candidates whose text the cross-encoder needs --
    are joined with the metadata store (<= 12 lookups per request in ONE batched statement instead of 62).

      * RERANK_ENABLED=false (or no reranker): mmr_fuse on the scans' f32 cosines.
      * rerank on with a device cross-encoder (configure(device_cross_encoder=...)): the (query, passage) pairs of ALL
        requests share one forward pass, its logits stay on the device and mmr_fuse_f64 reproduces _rerank_text's
        re-ordering + the rerank z-scores + fusion + gate (reference app/ml/retrieve.py:132-183, generate.py:56-60).

    The device fuses every hit the scan returns; the reference fuses the hits that survive the metadata join
    (retrieve.py:57-58,88-89).  Both agree whenever each indexed chunk exists with non-empty text -- which the write
    path guarantees (index_build.py:52-55 skips empty text).  If a candidate or winner nevertheless fails the join, that
    request is redone through the host path so the reference's semantics are kept."""
    cfg = settings.retrieval
    rerank_on = bool(cfg.use_rerank and _get_cross_encoder())
    if rerank_on and _DEVICE_CROSS_ENCODER is None:
        raise RuntimeError("rerank is on but no device cross-encoder is configured: use retrieve(), or "
                           "configure(device_cross_encoder=...), or RERANK_ENABLED=false")
    fused_search = getattr(_LANCEDB_STORE, "fused_search_batch", None)
    if fused_search is None:
        raise RuntimeError("the configured store has no device fusion (needs B200Store)")
    text_vecs, image_vecs = _embed_batch(queries)
    if rerank_on:
        batch = _LANCEDB_STORE.fused_search_batch_rerank(
            user_ids, queries, text_vecs, image_vecs, cfg.index_topk_text, cfg.index_topk_image, cfg.rerank_topk, cfg.final_n,
            cfg.confidence_tau, _DEVICE_CROSS_ENCODER, _METADATA_STORE)
    else:
        batch = fused_search(user_ids, text_vecs, image_vecs, cfg.index_topk_text, cfg.index_topk_image, cfg.final_n,
                             cfg.confidence_tau)
    bulk = getattr(_METADATA_STORE, "get_chunks", None)
    out: List[Tuple[List[Dict[str, Any]], bool]] = []
    for user_id, query, res in zip(user_ids, queries, batch):
        if res is None:                               # the device path could not serve this request (join failure)
            host = retrieve(user_id, query)
            out.append((host, _confidence_low(host)))
            continue
        items, low = res
        found = bulk([it["chunk_id"] for it in items]) if (bulk and items) else None
        joined, complete = [], True
        for it in items:
            chunk = found.get(it["chunk_id"]) if found is not None else _METADATA_STORE.get_chunk(it["chunk_id"])
            if not chunk or (it["modality"] == "text" and not chunk.text):
                complete = False
                break
            row = {"chunk_id": chunk.id, "modality": it["modality"], "score": it["score"],
                   "metadata": _prepare_metadata(chunk), "text": chunk.text if it["modality"] == "text" else None}
            if "rerank_score" in it:
                row["rerank_score"] = it["rerank_score"]
            row["combined_score"] = it["combined_score"]
            joined.append(row)
        if not complete:
            host = retrieve(user_id, query)
            out.append((host, _confidence_low(host)))
        else:
            out.append((joined, low))
    return out


def _rerank_text(query: str, results: List[Dict[str, Any]]) -> List[Dict[str, Any]]:
    cfg = settings.retrieval
    if not results or not cfg.use_rerank:
        return results
    model = _get_cross_encoder()
    if not model:
        return results
    head = results[: cfg.rerank_topk]
    pairs = [(query, item["text"]) for item in head if item.get("text")]
    if not head or not pairs:
        return results
    for item, logit in zip(head, model.predict(pairs)):
        item["rerank_score"] = float(logit)
    merged = head + results[len(head):]
    merged.sort(key=lambda item: item.get("rerank_score", item["score"]), reverse=True)
    return merged


def _z_scores(values: Sequence[Optional[float]]) -> List[float]:
    """Population z-scores with float32 moments and float64 quotient; None -> 0.0 (retrieve.py:186-195)."""
    present = [v for v in values if v is not None]
    if not present:
        return []
    moments = np.asarray(present, dtype=np.float32)
    mu, sigma = float(moments.mean()), float(moments.std())
    if sigma == 0:
        return [0.0] * len(values)
    return [0.0 if v is None else float((v - mu) / sigma) for v in values]


def _fuse_results(text_results: List[Dict[str, Any]], image_results: List[Dict[str, Any]]) -> List[Dict[str, Any]]:
    """Per-modality z-scores -> combined score -> stable descending order -> FINAL_N (retrieve.py:158-183).

    The rerank z-list is built from the items that HAVE a rerank score and is then indexed by the text
    item's position -- the reference's behaviour, preserved."""
    z_cos = _z_scores([it["score"] for it in text_results])
    rerank_vals = [it["rerank_score"] for it in text_results if "rerank_score" in it]
    z_rr = _z_scores(rerank_vals) if rerank_vals else []
    z_img = _z_scores([it["score"] for it in image_results])

    fused: List[Dict[str, Any]] = []
    for pos, it in enumerate(text_results):
        parts = ([z_cos[pos]] if z_cos else []) + ([z_rr[pos]] if pos < len(z_rr) else [])
        combined = float(np.mean(parts)) if parts else it["score"]
        fused.append(dict(it, combined_score=combined))
    for pos, it in enumerate(image_results):
        fused.append(dict(it, combined_score=float(z_img[pos]) if z_img else it["score"]))
    fused.sort(key=lambda it: it["combined_score"], reverse=True)
    return fused[: settings.retrieval.final_n]


def _confidence_low(items: List[Dict[str, Any]]) -> bool:
    """generate.py:56-60: abstain when the best (combined) score is under CONFIDENCE_TAU."""
    if not items:
        return True
    best = max(it.get("combined_score", it.get("score", 0.0)) for it in items)
    return best < settings.retrieval.confidence_tau


__all__ = ["retrieve_text", "retrieve_images", "retrieve", "retrieve_batch_device", "configure"]
