"""CPU oracle for the exact-scan hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs
(``cpu_baseline`` / ``--impl reference``) may import this package.  The product
package (``multimodal-rag-for-image-text-search_b200``) never imports it and has
no CPU fallback: it raises when the CUDA extension is missing.

PARITY STATUS
  * fusion / z-score / tau-gate / normalise / format_results / where-clause:
    PINNED.  ``oracle/make_golden.py`` executes the reference's own function
    bodies (extracted with ``ast`` from /root/reference, numpy-only) and the
    outputs are committed under ``tests/golden/``; ``tests/test_oracle_golden.py``
    checks this restatement against them bit-for-bit.
  * flat cosine scan + top-k (``flat_search``): PARITY UNPINNED.  The arithmetic
    lives in the un-vendored, un-pinned third-party ``lancedb`` -> ``lance`` Rust
    crate (reference requirements.txt:11, no version, no lock file), which is not
    installed here and cannot be fetched (no network), and no reference test
    exercises it (tests/test_retrieve.py:13-22 swaps in a DummyStore).  The
    restatement follows Lance's published flat-KNN definition of the cosine
    metric, d = 1 - x.q/(|x||q|) in float32, k smallest d, anchored on the
    reference call sites app/storage/lancedb_store.py:103-123.
"""
