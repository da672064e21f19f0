"""ResidentIndex -- one HBM-resident row matrix (bf16/fp16/fp32) + tenant segments, searched through the
C ABI.  PyTorch is only the plumbing here: it owns the device memory and the stream.

Reference anchors: the matrix is the `embedding` column of one LanceDB collection
(app/storage/lancedb_store.py:33-44), the segments are the `user_id == '...'` predicate (:107,118,141-144),
`search` is the scan + top-k of search_text / search_image (:103-123).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as N

_DTYPES = {
    "bf16": (N.MMR_BF16, torch.bfloat16),
    "f16": (N.MMR_F16, torch.float16),
    "f32": (N.MMR_F32, torch.float32),
}


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class ResidentIndex:
    """Rows live in `self.rows` ([n, dim] on `device`); `seg_offsets` ([T+1]) delimit tenants."""

    def __init__(self, rows: torch.Tensor, seg_offsets: Optional[Sequence[int]] = None, row_base: int = 0):
        if not rows.is_cuda:
            raise N.NativeError("ResidentIndex needs CUDA rows: this package has no CPU path")
        if rows.dim() != 2 or not rows.is_contiguous():
            raise ValueError("rows must be a contiguous [n, dim] tensor")
        kinds = {v[1]: (k, v[0]) for k, v in _DTYPES.items()}
        if rows.dtype not in kinds:
            raise ValueError(f"unsupported row dtype {rows.dtype}")
        self.dtype_name, self._dtype_code = kinds[rows.dtype]
        self.rows = rows
        self.device = rows.device
        self.n_rows, self.dim = int(rows.shape[0]), int(rows.shape[1])
        self.row_base = int(row_base)
        self.seg_offsets = None if seg_offsets is None else np.ascontiguousarray(seg_offsets, dtype=np.int64)
        self._handle = C.c_void_p()
        self._ws: Optional[torch.Tensor] = None
        self._one: dict = {}     # search_host_one: per-k cached output arrays + their addresses
        self._ws_key = (0, 0, 0)
        lib = N.lib()
        nseg = 0 if self.seg_offsets is None else len(self.seg_offsets) - 1
        segp = None if self.seg_offsets is None else self.seg_offsets.ctypes.data
        N.check(lib.mmr_index_create(self.device.index or 0, self.dim, self._dtype_code, self.n_rows,
                                     rows.data_ptr() if self.n_rows else None, segp, nseg, self.row_base,
                                     C.byref(self._handle)))

    # -- construction helpers -------------------------------------------------------------------
    @classmethod
    def from_f32(cls, rows_f32, seg_offsets=None, dtype: str = "bf16", device="cuda:0", normalize: bool = False,
                 row_base: int = 0) -> "ResidentIndex":
        """L1 loader: fp32 rows (host numpy / host or device torch) -> resident rows of `dtype`."""
        code, tdt = _DTYPES[dtype]
        device = torch.device(device)
        lib = N.lib()
        if isinstance(rows_f32, np.ndarray):
            src = np.ascontiguousarray(rows_f32, dtype=np.float32)
            n, d = src.shape
            dst = torch.empty((n, d), dtype=tdt, device=device)
            with torch.cuda.device(device):
                N.check(lib.mmr_load_rows_f32_host(device.index or 0, src.ctypes.data, _ptr(dst) if n else None, code,
                                                   n, d, int(normalize), _stream_ptr(device)))
        else:
            src = rows_f32.to(device=device, dtype=torch.float32).contiguous()
            n, d = src.shape
            dst = torch.empty((n, d), dtype=tdt, device=device)
            with torch.cuda.device(device):
                N.check(lib.mmr_convert_rows_f32(_ptr(src) if n else None, _ptr(dst) if n else None, code, n, d,
                                                 int(normalize), _stream_ptr(device)))
        return cls(dst, seg_offsets, row_base)

    @staticmethod
    def alloc_rows(n: int, dim: int, dtype: str, device) -> torch.Tensor:
        """Uninitialised resident buffer [n, dim] of the storage type `dtype`."""
        return torch.empty((int(n), int(dim)), dtype=_DTYPES[dtype][1], device=torch.device(device))

    @staticmethod
    def load_rows_into(buf: torch.Tensor, src_f32: np.ndarray, dst_rows: Optional[np.ndarray] = None,
                       dst_offset: int = 0, normalize: bool = False) -> None:
        """L1 loader into an existing resident buffer: host fp32 rows -> `buf`'s storage type.

        dst_rows given : source row i lands in buf[dst_rows[i]] (negative = skip) -- a host block in insertion order goes
                         straight to its tenant-sorted place, nothing is gathered on the host;
        dst_rows None  : the rows land contiguously at buf[dst_offset : dst_offset + n].
        Streams through pinned staging (mmr_load_rows_f32_host_scatter); returns when the rows are resident."""
        kinds = {v[1]: v[0] for v in _DTYPES.values()}
        src = np.ascontiguousarray(src_f32, dtype=np.float32)
        n, d = src.shape
        if n == 0:
            return
        if d != buf.shape[1]:
            raise ValueError(f"row length {d} != resident row length {buf.shape[1]}")
        dev = buf.device
        esize = buf.element_size()
        if dst_rows is not None:
            m = np.ascontiguousarray(dst_rows, dtype=np.int64)
            if m.shape != (n,) or (m.size and int(m.max()) >= buf.shape[0]):
                raise ValueError("dst_rows must hold one in-range resident row per source row")
            dst_ptr, mp = buf.data_ptr(), m.ctypes.data
        else:
            if dst_offset < 0 or dst_offset + n > buf.shape[0]:
                raise ValueError("rows do not fit the resident buffer")
            dst_ptr, mp = buf.data_ptr() + dst_offset * d * esize, None
        with torch.cuda.device(dev):
            N.check(N.lib().mmr_load_rows_f32_host_scatter(dev.index or 0, src.ctypes.data, dst_ptr, kinds[buf.dtype], n, d,
                                                           int(normalize), mp, _stream_ptr(dev)))

    def update(self, rows: torch.Tensor, n_rows: Optional[int] = None, seg_offsets=None) -> None:
        """Re-point the handle at grown / rewritten rows (same dim and dtype) after an upsert; the first n_rows rows
        of `rows` are live (`rows` may be a larger capacity buffer)."""
        if rows.dtype != self.rows.dtype or rows.shape[1] != self.dim or not rows.is_contiguous():
            raise ValueError("update needs a contiguous tensor of the same dtype and dim")
        n = int(rows.shape[0]) if n_rows is None else int(n_rows)
        self.seg_offsets = None if seg_offsets is None else np.ascontiguousarray(seg_offsets, dtype=np.int64)
        nseg = 0 if self.seg_offsets is None else len(self.seg_offsets) - 1
        segp = None if self.seg_offsets is None else self.seg_offsets.ctypes.data
        N.check(N.lib().mmr_index_update(self._handle, n, rows.data_ptr() if n else None, segp, nseg))
        self.rows, self.n_rows = rows, n

    def search_ranges(self, queries: torch.Tensor, k: int, ranges) -> Tuple[torch.Tensor, torch.Tensor]:
        """Like `search`, but query b scans the explicit row ranges `ranges[b]` = [(lo, hi), ...]."""
        if queries.dim() == 1:
            queries = queries[None, :]
        b = int(queries.shape[0])
        k = max(int(k), 1)
        off = np.zeros(b + 1, dtype=np.int32)
        flat = []
        for i, rq in enumerate(ranges):
            flat.extend(rq)
            off[i + 1] = len(flat)
        arr = np.ascontiguousarray(flat, dtype=np.int64).reshape(-1, 2) if flat else np.zeros((1, 2), np.int64)
        scores = torch.empty((b, k), dtype=torch.float32, device=self.device)
        rows = torch.empty((b, k), dtype=torch.int64, device=self.device)
        ws = self._workspace(b, k, n_ranges=len(flat))
        with torch.cuda.device(self.device):
            N.check(N.lib().mmr_search_ranges(self._handle, queries.data_ptr(), b, k, off.ctypes.data, arr.ctypes.data,
                                              scores.data_ptr(), rows.data_ptr(), ws.data_ptr(), ws.numel(),
                                              _stream_ptr(self.device)))
        return scores, rows

    def close(self) -> None:
        if self._handle:
            N.lib().mmr_index_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- search ---------------------------------------------------------------------------------
    def _workspace(self, b: int, k: int, n_ranges: int = 0) -> torch.Tensor:
        key = (b, k, n_ranges)
        if key != self._ws_key or self._ws is None:
            lib = N.lib()
            need = max(lib.mmr_search_workspace_bytes(self._handle, b, k),
                       lib.mmr_search_ranges_workspace_bytes(self._handle, b, k, n_ranges) if n_ranges else 0)
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.zeros(int(need), dtype=torch.uint8, device=self.device)
            self._ws_key = key
        return self._ws

    def set_query_precision(self, mode: str) -> None:
        """"auto": batches of >= 3 queries on one row range use the tensor-core kernels (16-bit queries);
        "f32": every query is scored in fp32 (K1 passes), so a request's result never depends on its batch;
        "rescore": tensor-core candidates re-scored in fp32 with K1's arithmetic and proven exact -- the same
        batch-independent, bit-identical results at tensor-core throughput (serving default of B200Store)."""
        N.check(N.lib().mmr_index_set_query_precision(
            self._handle, {"auto": N.MMR_QP_AUTO, "f32": N.MMR_QP_F32, "rescore": N.MMR_QP_RESCORE}[mode]))

    def search(self, queries: torch.Tensor, k: int, segments: Optional[Sequence[int]] = None,
               out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Device-resident search: queries [B, dim] f32 on the index device -> (scores [B,k] f32, rows [B,k] i64)."""
        if queries.dim() == 1:
            queries = queries[None, :]
        if not queries.is_cuda or queries.dtype != torch.float32 or not queries.is_contiguous():
            raise ValueError("queries must be a contiguous float32 CUDA tensor")
        if queries.shape[1] != self.dim:
            raise ValueError(f"query dim {queries.shape[1]} != index dim {self.dim}")
        b = int(queries.shape[0])
        k = max(int(k), 1)
        if out is None:
            scores = torch.empty((b, k), dtype=torch.float32, device=self.device)
            rows = torch.empty((b, k), dtype=torch.int64, device=self.device)
        else:
            scores, rows = out
        seg_arr = None
        if segments is not None:
            seg_arr = np.ascontiguousarray(segments, dtype=np.int32)
            if seg_arr.shape != (b,):
                raise ValueError("segments must have one entry per query")
        ws = self._workspace(b, k)
        args = (self._handle, queries.data_ptr(), None if seg_arr is None else seg_arr.ctypes.data, b, k, scores.data_ptr(),
                rows.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(self.device))
        if torch.cuda.current_device() == (self.device.index or 0):   # (entering torch's device context costs ~10 us)
            N.check(N.lib().mmr_search(*args))
        else:
            with torch.cuda.device(self.device):
                N.check(N.lib().mmr_search(*args))
        return scores, rows

    def debug_umma_scores(self, queries: torch.Tensor, row_begin: int, row_end: int) -> torch.Tensor:
        """Validation hook: raw tensor-core (K2) scores [B, row_end-row_begin] without the top-k epilogue."""
        b = int(queries.shape[0])
        n = int(row_end - row_begin)
        out = torch.full((b, n), float("nan"), dtype=torch.float32, device=self.device)
        ws = self._workspace(max(b, 8), 10)
        with torch.cuda.device(self.device):
            N.check(N.lib().mmr_debug_umma_scores(self._handle, queries.contiguous().data_ptr(), b, int(row_begin),
                                                  int(row_end), out.data_ptr(), n, ws.data_ptr(), ws.numel(),
                                                  _stream_ptr(self.device)))
        return out

    def search_host(self, queries: np.ndarray, k: int, segments: Optional[Sequence[int]] = None):
        """Host-buffer search (H2D + scan + D2H + sync inside the C call) -> numpy (scores, rows)."""
        q = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.float32)
        if q.shape[1] != self.dim:
            raise ValueError(f"query dim {q.shape[1]} != index dim {self.dim}")
        b = q.shape[0]
        k = max(int(k), 1)
        scores = np.empty((b, k), dtype=np.float32)
        rows = np.empty((b, k), dtype=np.int64)
        seg_arr = None if segments is None else np.ascontiguousarray(segments, dtype=np.int32)
        # (the C call selects the index's device itself; entering torch's device context costs ~10 us per request, so it
        #  is only done when another device is current)
        if torch.cuda.current_device() == (self.device.index or 0):
            N.check(N.lib().mmr_search_host(self._handle, q.ctypes.data, None if seg_arr is None else seg_arr.ctypes.data,
                                            b, k, scores.ctypes.data, rows.ctypes.data, _stream_ptr(self.device)))
        else:
            with torch.cuda.device(self.device):
                N.check(N.lib().mmr_search_host(self._handle, q.ctypes.data, None if seg_arr is None else seg_arr.ctypes.data,
                                                b, k, scores.ctypes.data, rows.ctypes.data, _stream_ptr(self.device)))
        return scores, rows


    def search_host_one(self, q: np.ndarray, k: int, segment: int = 0):
        """The single-request form of search_host for the serving store: `q` is ONE query, float32, C-contiguous, [1, dim] or
        [dim].  The output arrays and the segment word are cached per k and REUSED by the next call (the caller holds the
        collection lock and converts them before it releases it), so a request allocates nothing and builds no ctypes
        views: ~5 us less host work per call than search_host."""
        k = max(int(k), 1)
        slot = self._one.get(k)
        if slot is None:
            scores = np.empty((1, k), dtype=np.float32)
            rows = np.empty((1, k), dtype=np.int64)
            seg = np.zeros(1, dtype=np.int32)
            slot = self._one[k] = (scores, rows, seg, scores.ctypes.data, rows.ctypes.data, seg.ctypes.data)
        scores, rows, seg, p_s, p_r, p_seg = slot
        if q.dtype != np.float32 or not q.flags.c_contiguous or q.size != self.dim:
            raise ValueError(f"search_host_one needs one contiguous float32 query of dim {self.dim}")
        seg[0] = segment
        fn = N.lib().mmr_search_host
        if torch.cuda.current_device() == (self.device.index or 0):
            N.check(fn(self._handle, q.ctypes.data, p_seg, 1, k, p_s, p_r, _stream_ptr(self.device)))
        else:
            with torch.cuda.device(self.device):
                N.check(fn(self._handle, q.ctypes.data, p_seg, 1, k, p_s, p_r, _stream_ptr(self.device)))
        return scores, rows


class MultiIndex:
    """G row-range shards of ONE table on the G GPUs of a box, searched from ONE process (mmr_multi_*): one scan launch
    per device from per-device launcher threads, results pushed into the collector's buffer over NVLink peer mappings,
    merged result + completion flag in a mapped host mailbox.  `shards[g]` must have row_base = its first global row."""

    def __init__(self, shards: Sequence[ResidentIndex]) -> None:
        self.shards = list(shards)
        self._handle = C.c_void_p()
        arr = (C.c_void_p * len(self.shards))(*[s._handle for s in self.shards])
        N.check(N.lib().mmr_multi_create(arr, len(self.shards), C.byref(self._handle)))
        self.dim = self.shards[0].dim
        self.n_rows = max(s.row_base + s.n_rows for s in self.shards)

    def search_host(self, queries: np.ndarray, k: int, ranges=None):
        """queries [B, dim] f32 (host) -> numpy (scores [B,k], rows [B,k] global ids).  ranges[b] = [(lo, hi), ...]
        global row ranges; None = the whole table for every query."""
        q = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.float32)
        if q.shape[1] != self.dim:
            raise ValueError(f"query dim {q.shape[1]} != index dim {self.dim}")
        b = q.shape[0]
        k = max(int(k), 1)
        if ranges is None:
            ranges = [[(0, self.n_rows)]] * b
        off = np.zeros(b + 1, dtype=np.int32)
        flat = []
        for i, rq in enumerate(ranges):
            flat.extend(rq)
            off[i + 1] = len(flat)
        arr = np.ascontiguousarray(flat, dtype=np.int64).reshape(-1, 2) if flat else np.zeros((1, 2), np.int64)
        scores = np.empty((b, k), dtype=np.float32)
        rows = np.empty((b, k), dtype=np.int64)
        with torch.cuda.device(self.shards[0].device):   # the call leaves the collector's device current: restore ours
            N.check(N.lib().mmr_multi_search_host(self._handle, q.ctypes.data, b, k, off.ctypes.data, arr.ctypes.data,
                                                  scores.ctypes.data, rows.ctypes.data))
        return scores, rows

    def close(self) -> None:
        if self._handle:
            N.lib().mmr_multi_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def merge_topk(scores: torch.Tensor, rows: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """K4: [G, B, k] shard-local results -> [B, k]."""
    g, b, k = scores.shape
    out_s = torch.empty((b, k), dtype=torch.float32, device=scores.device)
    out_r = torch.empty((b, k), dtype=torch.int64, device=scores.device)
    with torch.cuda.device(scores.device):
        N.check(N.lib().mmr_merge_topk(scores.contiguous().data_ptr(), rows.contiguous().data_ptr(), g, b, k,
                                       out_s.data_ptr(), out_r.data_ptr(), _stream_ptr(scores.device)))
    return out_s, out_r


def fuse(text: Optional[Tuple[torch.Tensor, torch.Tensor]], image: Optional[Tuple[torch.Tensor, torch.Tensor]],
         final_n: int, tau: float):
    """K5: device fusion + gate for the rerank-off path.  Returns dict of device tensors."""
    ref = text if text is not None else image
    if ref is None:
        raise ValueError("need at least one modality")
    dev = ref[0].device
    b = int(ref[0].shape[0])
    kt = 0 if text is None else int(text[0].shape[1])
    ki = 0 if image is None else int(image[0].shape[1])
    comb = torch.empty((b, final_n), dtype=torch.float64, device=dev)
    score = torch.empty((b, final_n), dtype=torch.float64, device=dev)
    rows = torch.empty((b, final_n), dtype=torch.int64, device=dev)
    mod = torch.empty((b, final_n), dtype=torch.int8, device=dev)
    low = torch.empty((b,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().mmr_fuse(_ptr(text[0].contiguous()) if kt else None, _ptr(text[1].contiguous()) if kt else None, kt,
                                 _ptr(image[0].contiguous()) if ki else None, _ptr(image[1].contiguous()) if ki else None, ki,
                                 b, int(final_n), float(tau), comb.data_ptr(), score.data_ptr(), rows.data_ptr(),
                                 mod.data_ptr(), low.data_ptr(), _stream_ptr(dev)))
    return {"combined": comb, "score": score, "rows": rows, "modality": mod, "low_conf": low}


def fuse_f64(text_scores, text_count, img_scores, img_count, final_n: int, tau: float, rerank=None, rerank_count=None):
    """K5, complete form (mmr_fuse_f64): float64 scores [B, kt] / [B, ki] + counts (+ optional cross-encoder logits for
    the first rerank_count[b] text items) -> {"combined" [B, final_n] f64, "index" [B, final_n] i32, "low_conf" [B] u8}."""
    ref = text_scores if text_scores is not None else img_scores
    dev = ref.device
    b = int(ref.shape[0])
    kt = 0 if text_scores is None else int(text_scores.shape[1])
    ki = 0 if img_scores is None else int(img_scores.shape[1])
    comb = torch.empty((b, final_n), dtype=torch.float64, device=dev)
    index = torch.empty((b, final_n), dtype=torch.int32, device=dev)
    low = torch.empty((b,), dtype=torch.uint8, device=dev)

    def p(t, dt):
        return None if t is None else t.to(device=dev, dtype=dt).contiguous()

    ts, tc, rr, rc = p(text_scores, torch.float64), p(text_count, torch.int32), p(rerank, torch.float64), p(rerank_count, torch.int32)
    is_, ic = p(img_scores, torch.float64), p(img_count, torch.int32)
    with torch.cuda.device(dev):
        N.check(N.lib().mmr_fuse_f64(_ptr(ts), _ptr(tc), _ptr(rr), _ptr(rc), _ptr(is_), _ptr(ic), kt, ki, b, int(final_n),
                                     float(tau), comb.data_ptr(), index.data_ptr(), low.data_ptr(), _stream_ptr(dev)))
    return {"combined": comb, "index": index, "low_conf": low}
