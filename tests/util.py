"""Shared helpers for the parity tests: seeded synthetic tables and the tolerance-aware checker."""
from __future__ import annotations

import numpy as np

from oracle import flat_search as ofs

TOL_BF16 = 1e-3   # north star: |score - oracle| <= 1e-3 for bf16 storage
TOL_F32 = 1e-5    # and <= 1e-5 for fp32 storage
TOL_STRICT = 2e-6  # same stored values, different fp32 summation order only


def unit_rows(n: int, d: int, seed: int, cone: float = 0.0) -> np.ndarray:
    """Index A (isotropic, cone=0) / Index B (CLIP-like cone: normalize(mu*c + g))."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    if cone:
        mu = np.random.default_rng(12345).standard_normal(d).astype(np.float32)
        mu /= np.linalg.norm(mu)
        x += np.float32(cone * np.sqrt(d)) * mu
    return ofs.normalize_rows(x)


def queries(b: int, d: int, seed: int = 0xC0FFEE, cone: float = 0.0) -> np.ndarray:
    return unit_rows(b, d, seed, cone)


def oracle_scores(rows_f32: np.ndarray, q: np.ndarray) -> np.ndarray:
    """fp32 cosine similarity of one query against all rows = 1 - oracle distance (as f32)."""
    qn = np.asarray(ofs.normalize(q), dtype=np.float32)
    return (np.float32(1.0) - ofs.cosine_distances(rows_f32, qn, unit_rows=True)).astype(np.float32)


def check_topk(gpu_scores, gpu_rows, full_scores: np.ndarray, k: int, tol: float, lo: int = 0, hi=None, what=""):
    """Tolerance-aware parity of one query's result against the oracle's full score vector over [lo, hi).

    * result is sorted by score desc (row asc among equal scores), ids unique and inside [lo, hi)
    * every returned score is within `tol` of the oracle score of that row
    * every oracle row scoring more than tol above the oracle's k-th score is present
    * no returned row scores more than tol below the oracle's k-th score
    """
    hi = full_scores.shape[0] + lo if hi is None else hi
    n = hi - lo
    kk = min(k, n)
    gpu_scores = np.asarray(gpu_scores)
    gpu_rows = np.asarray(gpu_rows)
    valid = gpu_rows >= 0
    assert valid.sum() == kk, f"{what}: expected {kk} hits, got {valid.sum()}"
    assert valid[:kk].all(), f"{what}: hits must be a prefix"
    assert np.isneginf(gpu_scores[kk:]).all(), f"{what}: padding scores must be -inf"
    r = gpu_rows[:kk]
    s = gpu_scores[:kk]
    assert ((r >= lo) & (r < hi)).all(), f"{what}: row outside its segment"
    assert len(set(r.tolist())) == kk, f"{what}: duplicate rows"
    for j in range(kk - 1):
        assert s[j] > s[j + 1] or (s[j] == s[j + 1] and r[j] < r[j + 1]), f"{what}: order broken at {j}"
    ref = full_scores[r - lo]
    err = np.abs(s.astype(np.float64) - ref.astype(np.float64)).max() if kk else 0.0
    assert err <= tol, f"{what}: score error {err} > {tol}"
    if kk:
        kth = np.partition(full_scores, n - kk)[n - kk]
        must = np.nonzero(full_scores > kth + tol)[0] + lo
        missing = set(must.tolist()) - set(r.tolist())
        assert not missing, f"{what}: missing clear winners {sorted(missing)[:5]}"
        assert (ref >= kth - tol).all(), f"{what}: returned a row below the k-th score by more than tol"
    return err
