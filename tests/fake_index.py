"""numpy stand-in for ResidentIndex (test infrastructure, like the reference's DummyStore): lets the host logic of the
store -- which rows a tenant owns after any sequence of upserts, durable log, cross-process refresh -- run without a GPU."""
import importlib

import numpy as np
import torch

from oracle import flat_search as ofs

PKG = "multimodal-rag-for-image-text-search_b200"


class FakeIndex:
    """numpy stand-in for ResidentIndex: exact fp32 scan, NaN rows never returned, order (score desc, row asc)."""

    def __init__(self, rows, seg_offsets=None, row_base=0):
        self.rows, self.n_rows, self.dim = rows, int(rows.shape[0]), int(rows.shape[1])
        self.seg_offsets = None if seg_offsets is None else np.asarray(seg_offsets, np.int64)
        self.device = rows.device

    @classmethod
    def from_f32(cls, rows_f32, seg_offsets=None, dtype="bf16", device="cpu", normalize=False, row_base=0):
        return cls(torch.from_numpy(np.ascontiguousarray(rows_f32, dtype=np.float32)).clone(), seg_offsets, row_base)

    @staticmethod
    def alloc_rows(n, dim, dtype, device):
        return torch.full((int(n), int(dim)), float("nan"), dtype=torch.float32)

    @staticmethod
    def load_rows_into(buf, src_f32, dst_rows=None, dst_offset=0, normalize=False):
        src = torch.from_numpy(np.array(src_f32, dtype=np.float32, copy=True))
        if dst_rows is None:
            buf[dst_offset:dst_offset + src.shape[0]] = src
        else:
            keep = np.nonzero(np.asarray(dst_rows) >= 0)[0]
            buf[torch.from_numpy(np.asarray(dst_rows)[keep])] = src[torch.from_numpy(keep)]

    def update(self, rows, n_rows=None, seg_offsets=None):
        self.rows, self.n_rows = rows, int(rows.shape[0] if n_rows is None else n_rows)
        self.seg_offsets = None if seg_offsets is None else np.asarray(seg_offsets, np.int64)

    def close(self):
        pass

    def set_query_precision(self, mode):
        pass

    def _scan(self, q, k, ranges):
        mat = self.rows[: self.n_rows].numpy()
        qn = np.asarray(ofs.normalize(q), np.float32)
        idx = np.concatenate([np.arange(lo, hi) for lo, hi in ranges] + [np.zeros(0, np.int64)]).astype(np.int64)
        s = mat[idx] @ qn
        keep = ~np.isnan(s)
        idx, s = idx[keep], s[keep]
        order = np.lexsort((idx, -s.astype(np.float64)))[:k]
        out_s, out_r = np.full(k, -np.inf, np.float32), np.full(k, -1, np.int64)
        out_s[: len(order)], out_r[: len(order)] = s[order], idx[order]
        return out_s, out_r

    def search_host(self, q, k, segments=None):
        q = np.atleast_2d(q)
        res = [self._scan(q[b], k, [(int(self.seg_offsets[segments[b]]), int(self.seg_offsets[segments[b] + 1]))])
               for b in range(q.shape[0])]
        return np.stack([r[0] for r in res]), np.stack([r[1] for r in res])

    def search_ranges(self, q, k, ranges):
        q = q.numpy()
        res = [self._scan(q[b], k, ranges[b]) for b in range(q.shape[0])]
        return torch.from_numpy(np.stack([r[0] for r in res])), torch.from_numpy(np.stack([r[1] for r in res]))



def make_cpu_store(db_path=None):
    """A B200Store whose resident side is the FakeIndex (the CUDA checks of __init__ are skipped)."""
    store_mod = importlib.import_module(PKG + ".store")
    store_mod.ResidentIndex = FakeIndex
    st = store_mod.B200Store.__new__(store_mod.B200Store)
    st._init_state(db_path, torch.device("cpu"), "f32")
    return st
