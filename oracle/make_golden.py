#!/usr/bin/env python
"""Generate tests/golden/*.json by EXECUTING the reference's own function bodies.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden.py

The reference modules cannot be imported here (``app.ml.retrieve`` constructs a LanceDBStore
at import, and lancedb / sentence_transformers / llama_index are absent), so the numpy-only
functions on the hot path are lifted out of the source with ``ast`` and exec'd unchanged in a
namespace that provides just ``np``, ``json``, typing names and a ``settings`` stub.  Only the
OUTPUTS are committed; no reference source is copied into this repository.

Functions executed (reference file:line):
  app/ml/retrieve.py:132-155  _rerank_text      (cross-encoder stubbed like tests/test_retrieve.py:33-35)
  app/ml/retrieve.py:158-183  _fuse_results
  app/ml/retrieve.py:186-195  _z_scores
  app/ml/generate.py:56-60    _confidence_low
  app/storage/lancedb_store.py:63-69    LanceDBStore._normalize
  app/storage/lancedb_store.py:71-85    LanceDBStore._prepare_rows
  app/storage/lancedb_store.py:125-139  LanceDBStore._format_results
  app/storage/lancedb_store.py:141-144  LanceDBStore._where_clause
"""
from __future__ import annotations

import ast
import copy
import json
import os
import sys
import types
from dataclasses import dataclass
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

REF = os.environ.get("MMR_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _lift(path: str, names: Sequence[str], cls: Optional[str] = None) -> Dict[str, Any]:
    """Compile the named top-level (or ``cls``-level) functions of ``path`` and return them."""
    with open(os.path.join(REF, path)) as fh:
        tree = ast.parse(fh.read())
    body = tree.body
    if cls is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    picked = [n for n in body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(picked) == len(names), (path, names, [n.name for n in picked])
    for fn in picked:
        fn.decorator_list = []  # @staticmethod -> plain function
    mod = ast.Module(body=picked, type_ignores=[])
    return mod


def _settings(final_n=4, use_rerank=True, rerank_topk=8, tau=0.25):
    return types.SimpleNamespace(
        retrieval=types.SimpleNamespace(
            final_n=final_n, use_rerank=use_rerank, rerank_topk=rerank_topk, confidence_tau=tau
        )
    )


class _LinspaceCrossEncoder:
    """tests/test_retrieve.py:33-35."""

    def predict(self, pairs):
        return np.linspace(0.1, 0.9, len(pairs))


class _SeededCrossEncoder:
    def __init__(self, seed):
        self._rng = np.random.default_rng(seed)

    def predict(self, pairs):
        return self._rng.normal(0.0, 3.0, size=len(pairs)).astype(np.float32)


def _namespace(**extra) -> Dict[str, Any]:
    ns: Dict[str, Any] = {
        "np": np, "json": json, "Any": Any, "Dict": Dict, "Iterable": Iterable, "List": List,
        "Optional": Optional, "Sequence": Sequence, "Tuple": Tuple,
    }
    ns.update(extra)
    return ns


def _load_retrieve(settings, cross_encoder):
    ns = _namespace(settings=settings, _get_cross_encoder=lambda: cross_encoder)
    code = compile(_lift("app/ml/retrieve.py", ["_rerank_text", "_fuse_results", "_z_scores"]),
                   "ref:app/ml/retrieve.py", "exec")
    exec(code, ns)
    return ns


def _load_generate(settings):
    ns = _namespace(settings=settings)
    exec(compile(_lift("app/ml/generate.py", ["_confidence_low"]), "ref:app/ml/generate.py", "exec"), ns)
    return ns


def _load_store():
    ns = _namespace()
    exec(compile(_lift("app/storage/lancedb_store.py",
                       ["_normalize", "_format_results", "_where_clause"], cls="LanceDBStore"),
                 "ref:app/storage/lancedb_store.py", "exec"), ns)
    # _prepare_rows refers to LanceDBStore._normalize: give it a holder class
    ns["LanceDBStore"] = types.SimpleNamespace(_normalize=ns["_normalize"])
    exec(compile(_lift("app/storage/lancedb_store.py", ["_prepare_rows"], cls="LanceDBStore"),
                 "ref:app/storage/lancedb_store.py", "exec"), ns)
    return ns


@dataclass
class _Row:
    chunk_id: str
    user_id: str
    document_id: str
    modality: str
    embedding: Sequence[float]
    meta: Dict[str, Any]


def _text_items(rng, n, with_text=True):
    scores = np.sort(rng.uniform(-0.2, 0.95, size=n).astype(np.float32))[::-1]
    return [
        {"chunk_id": f"t{i}", "modality": "text", "score": float(s), "metadata": {"doc_id": f"d{i % 3}"},
         "text": (f"text {i}" if with_text or i % 2 == 0 else "")}
        for i, s in enumerate(scores)
    ]


def _image_items(rng, n):
    scores = np.sort(rng.uniform(0.05, 0.45, size=n).astype(np.float32))[::-1]
    return [
        {"chunk_id": f"i{i}", "modality": "image", "score": float(s), "metadata": {"file_path": f"/tmp/{i}.jpg"},
         "text": None}
        for i, s in enumerate(scores)
    ]


def make_fusion() -> Dict[str, Any]:
    out: Dict[str, Any] = {"z_scores": [], "fuse": [], "rerank_fuse": [], "confidence": []}
    rng = np.random.default_rng(20260118)

    ns = _load_retrieve(_settings(), False)
    z_cases: List[List[Optional[float]]] = [
        [], [0.7], [0.6, 0.8], [0.5, None, 0.7], [None], [None, None], [0.3, 0.3, 0.3],
        [1.0, 0.0], [0.1, 0.2, 0.3, 0.4], [-1.5, 2.5, 0.0], [1e-8, 2e-8], [1e6, 1e6 + 1],
        [0.123456789, 0.123456789, 0.987654321],
    ]
    for n in (2, 3, 5, 8, 12, 50, 64):
        z_cases.append([float(v) for v in rng.uniform(-1, 1, size=n).astype(np.float32)])
        z_cases.append([float(v) for v in rng.uniform(-1, 1, size=n)])  # f64-valued python floats
    withnone = [float(v) for v in rng.uniform(0, 1, size=9)]
    withnone[2] = None
    withnone[7] = None
    z_cases.append(withnone)
    for vals in z_cases:
        out["z_scores"].append({"values": vals, "expect": ns["_z_scores"](copy.deepcopy(vals))})

    # _fuse_results without rerank scores, several final_n
    fuse_inputs = [
        # tests/test_retrieve.py:46-52 shapes, minus the cross encoder
        ([{"chunk_id": "t1", "modality": "text", "score": 0.8, "metadata": {}, "text": "alpha"},
          {"chunk_id": "t2", "modality": "text", "score": 0.6, "metadata": {}, "text": "beta"}],
         [{"chunk_id": "i1", "modality": "image", "score": 0.7, "metadata": {"file_path": "/tmp/img.jpg"}, "text": None}]),
        ([], []),
        (_text_items(rng, 1), []),
        ([], _image_items(rng, 1)),
        ([], _image_items(rng, 12)),
        (_text_items(rng, 50), []),
        (_text_items(rng, 50), _image_items(rng, 12)),
        (_text_items(rng, 10), _image_items(rng, 10)),
        (_text_items(rng, 3), _image_items(rng, 2)),
        # all-equal scores -> std == 0 -> every z == 0 -> stable order decides
        ([{"chunk_id": f"t{i}", "modality": "text", "score": 0.5, "metadata": {}, "text": "x"} for i in range(4)],
         [{"chunk_id": f"i{i}", "modality": "image", "score": 0.25, "metadata": {}, "text": None} for i in range(3)]),
    ]
    for final_n in (4, 1, 10, 100):
        ns = _load_retrieve(_settings(final_n=final_n), False)
        for text, image in fuse_inputs:
            got = ns["_fuse_results"](copy.deepcopy(text), copy.deepcopy(image))
            out["fuse"].append({"final_n": final_n, "text": text, "image": image, "expect": got})

    # _rerank_text -> _fuse_results (the retrieve() tail, retrieve.py:113-114)
    rr_inputs = [
        (fuse_inputs[0][0], fuse_inputs[0][1], "linspace", 0),
        (_text_items(rng, 50), _image_items(rng, 12), "linspace", 0),
        (_text_items(rng, 50), _image_items(rng, 12), "seeded", 11),
        (_text_items(rng, 5), _image_items(rng, 12), "seeded", 12),
        (_text_items(rng, 9, with_text=False), _image_items(rng, 4), "seeded", 13),
        (_text_items(rng, 8), [], "seeded", 14),
    ]
    for text, image, kind, seed in rr_inputs:
        for rerank_topk, final_n in ((8, 4), (3, 4), (8, 20)):
            ce = _LinspaceCrossEncoder() if kind == "linspace" else _SeededCrossEncoder(seed)
            ns = _load_retrieve(_settings(final_n=final_n, rerank_topk=rerank_topk), ce)
            t = copy.deepcopy(text)
            reranked = ns["_rerank_text"]("example query", t)
            fused = ns["_fuse_results"](reranked, copy.deepcopy(image))
            # what the encoder returned, so the oracle can replay it without the stub class
            ce2 = _LinspaceCrossEncoder() if kind == "linspace" else _SeededCrossEncoder(seed)
            n_pairs = len([it for it in text[:rerank_topk] if it.get("text")])
            replay = [float(v) for v in ce2.predict([None] * n_pairs)] if n_pairs else []
            out["rerank_fuse"].append({
                "final_n": final_n, "rerank_topk": rerank_topk, "text": text, "image": image,
                "predict": replay, "reranked": reranked, "expect": fused,
            })

    for tau in (0.25, 0.0, 0.9):
        ns = _load_generate(_settings(tau=tau))
        cases = [
            [],
            [{"modality": "text", "score": 0.1, "metadata": {"doc_id": "doc1"}, "text": "sample"}],  # test_generate.py:19-21
            [{"modality": "text", "score": 1.0, "combined_score": 1.0, "metadata": {"doc_id": "doc1"}, "text": "fact"}],
            [{"score": 0.9, "combined_score": 0.0}],
            [{"combined_score": 0.25}],
            [{"combined_score": 0.24999999}, {"score": 0.3}],
            [{"metadata": {}}],
            [{"combined_score": -1.2}, {"combined_score": 0.26}],
        ]
        for items in cases:
            out["confidence"].append({"tau": tau, "items": items, "expect": bool(ns["_confidence_low"](copy.deepcopy(items)))})
    return out


def make_store() -> Dict[str, Any]:
    out: Dict[str, Any] = {"normalize": [], "format_results": [], "where_clause": [], "prepare_rows": []}
    rng = np.random.default_rng(7)
    ns = _load_store()

    vecs: List[List[float]] = [
        [0.0] * 8, [1.0, 0.0, 0.0], [3, 4], [1e-30, 1e-30], [1e20, 1e20, 1e20], [-2.0, 0.5, 0.25, 8.0],
    ]
    for d in (4, 384, 512):
        vecs.append([float(v) for v in rng.normal(size=d)])
        vecs.append([float(v) for v in rng.normal(size=d).astype(np.float32)])
    for v in vecs:
        out["normalize"].append({"vector": v, "expect": ns["_normalize"](list(v))})

    fr_cases = [
        [],
        [{"chunk_id": "a", "meta": "{}", "_distance": float(np.float32(0.25))}],
        [{"chunk_id": "a", "meta": None, "_distance": float(np.float32(0.3))},
         {"chunk_id": "b", "meta": "{\"k\": 1}", "_distance": float(np.float32(0.1))},
         {"chunk_id": "c", "_distance": float(np.float32(0.3))},          # tie with a: stable order keeps a before c
         {"chunk_id": "d", "meta": "{\"x\": [1, 2]}"}],                    # missing _distance -> 0.0 -> score 1.0
    ]
    d = np.sort(rng.uniform(0, 2, size=50).astype(np.float32))
    fr_cases.append([{"chunk_id": f"c{i}", "meta": json.dumps({"i": i}), "_distance": float(x)} for i, x in enumerate(d)])
    for rows in fr_cases:
        out["format_results"].append({"rows": rows, "expect": ns["_format_results"](copy.deepcopy(rows))})

    for col, val in (("user_id", "alice"), ("user_id", "o'brien"), ("chunk_id", "a''b"), ("user_id", 42), ("user_id", "")):
        out["where_clause"].append({"column": col, "value": val, "expect": ns["_where_clause"](col, val)})

    rows = [
        _Row("c1", "u1", "d1", "text", [3.0, 4.0], {"a": 1}),
        _Row("c2", "u2", "d1", "image", [0.0, 0.0], None),
        _Row("c3", "u1", "d2", "text", [float(v) for v in rng.normal(size=16)], {}),
    ]
    out["prepare_rows"].append({
        "rows": [r.__dict__ for r in copy.deepcopy(rows)],
        "expect": ns["_prepare_rows"](copy.deepcopy(rows)),
    })
    return out


def main() -> None:
    if not os.path.isdir(REF):
        sys.exit(f"{REF} not present: golden vectors can only be regenerated in the build container")
    os.makedirs(OUT, exist_ok=True)
    for name, fn in (("fusion_golden.json", make_fusion), ("store_golden.json", make_store)):
        data = fn()
        data["_generator"] = "oracle/make_golden.py (executes reference function bodies lifted with ast)"
        data["_numpy"] = np.__version__
        with open(os.path.join(OUT, name), "w") as fh:
            json.dump(data, fh, sort_keys=True, separators=(",", ":"))
        print(name, {k: len(v) for k, v in data.items() if isinstance(v, list)})


if __name__ == "__main__":
    main()
