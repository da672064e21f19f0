"""Row-range sharding of one table over the GPUs of a box.

One process per GPU (torch.distributed).  Every rank holds rows [bounds[r], bounds[r+1]) of the tenant-sorted
table as a ResidentIndex with row_base = bounds[r]; queries are replicated; each rank scans its shard, the tiny
per-shard results ([B, k] scores + rows, 12 B per hit) are all-gathered in ONE collective on a packed wire
buffer, and every rank runs the K4 merge kernel on the gathered buffer in place.  The answer is bit-identical
for every world size (same total order: score desc, row asc).

The scan itself runs only on CUDA (no CPU fallback); the bounds / wire / gather logic here is backend-agnostic
and is exercised with the gloo backend on CPU in tests/test_sharded_gloo.py.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, world: int, align: int = 1) -> List[int]:
    """Contiguous, balanced row ranges; interior bounds rounded to `align` rows."""
    bounds = [0]
    for r in range(1, world):
        b = int(round(r * n_rows / world))
        b = min(n_rows, (b + align - 1) // align * align)
        bounds.append(max(b, bounds[-1]))
    bounds.append(n_rows)
    return bounds


def split_segments(seg_offsets: Sequence[int], lo: int, hi: int) -> np.ndarray:
    """Tenant segment table of the shard [lo, hi): global offsets clipped to the shard and made shard-local.
    A tenant that straddles a shard boundary keeps its id on both sides (each GPU scans its slice)."""
    seg = np.asarray(seg_offsets, dtype=np.int64)
    return np.clip(seg, lo, hi) - lo


class Wire:
    """Packed per-rank result buffer: [scores f32 B*k | pad to 8 | rows i64 B*k] as one uint8 tensor."""

    def __init__(self, b: int, k: int, device) -> None:
        self.b, self.k = b, k
        self.score_bytes = (b * k * 4 + 7) // 8 * 8
        self.nbytes = self.score_bytes + b * k * 8
        self.buf = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
        self.scores = self.buf[: b * k * 4].view(torch.float32).view(b, k)
        self.rows = self.buf[self.score_bytes:].view(torch.int64).view(b, k)


class GatheredWire:
    """world x Wire, gathered in place; views for the strided K4 merge."""

    def __init__(self, wire: Wire, world: int) -> None:
        self.world, self.wire = world, wire
        self.buf = torch.zeros((world, wire.nbytes), dtype=torch.uint8, device=wire.buf.device)

    def views(self) -> Tuple[torch.Tensor, torch.Tensor]:
        w = self.wire
        scores = self.buf[:, : w.b * w.k * 4].view(torch.float32)      # [G, B*k], row stride nbytes/4
        rows = self.buf[:, w.score_bytes:].view(torch.int64)            # [G, B*k], row stride nbytes/8
        return scores, rows


def gather_wire(wire: Wire, gathered: GatheredWire, group=None) -> None:
    """The one exchange step of the path: all-gather of the packed per-shard top-k."""
    dist.all_gather_into_tensor(gathered.buf.view(-1), wire.buf, group=group)


class PeerExchange:
    """Symmetric buffers for the fused exchange (mmr_search_exchange): one buffer per rank, mapped into every
    process with torch's symmetric memory (CUDA VMM handles over the process-group store; NVLink peer access).
    `ptrs[g]` = rank g's buffer as addressed from this process."""

    def __init__(self, world: int, rank: int, b: int, k: int, device, group=None) -> None:
        import torch.distributed._symmetric_memory as symm
        from . import _native as N

        self.b, self.k = b, k
        nbytes = int(N.lib().mmr_exchange_buffer_bytes(world, b, k))
        self.buf = symm.empty(nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        grp = group if group is not None else dist.group.WORLD
        try:
            self.handle = symm.rendezvous(self.buf, grp)
        except Exception:
            symm.enable_symm_mem_for_group(grp.group_name)
            self.handle = symm.rendezvous(self.buf, grp.group_name)
        self.ptrs = np.ascontiguousarray([int(p) for p in self.handle.buffer_ptrs], dtype=np.uint64)
        assert len(self.ptrs) == world and int(self.ptrs[rank]) == self.buf.data_ptr()
        self.seq = 0
        torch.cuda.synchronize(device)
        dist.barrier(group=group)          # every buffer is zeroed before anybody pushes into it


class ShardedIndex:
    """This rank's shard + the exchange.  `local` is a ResidentIndex built with row_base = bounds[rank].

    exchange = "fused": scan kernels store their result straight into every peer's symmetric buffer over NVLink and
                        a wait+merge kernel finishes (no NCCL on the data path);
               "nccl" : scan -> one packed all-gather -> K4 merge (correctness baseline, any backend);
               "auto" : fused when the symmetric-memory rendezvous works, else nccl.
    """

    def __init__(self, local, group=None, exchange: str = "auto") -> None:
        self.local = local
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.exchange = exchange
        self._wire: Optional[Wire] = None
        self._gathered: Optional[GatheredWire] = None
        self._out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
        self._peer: Optional[PeerExchange] = None
        self._ws: Optional[torch.Tensor] = None

    def _peer_exchange(self, b: int, k: int) -> Optional[PeerExchange]:
        if self.exchange == "nccl" or self.world == 1:
            return None
        if self._peer is None or (self._peer.b, self._peer.k) != (b, k):
            try:
                self._peer = PeerExchange(self.world, self.rank, b, k, self.local.device, self.group)
            except Exception:
                if self.exchange == "fused":
                    raise
                self.exchange = "nccl"     # symmetric memory unavailable: every rank fails the same way
                return None
        return self._peer

    def _search_fused(self, peer: PeerExchange, queries: torch.Tensor, b: int, k: int, segments, out):
        from . import _native as N
        from .index import _stream_ptr

        lib = N.lib()
        ix = self.local
        need = int(lib.mmr_search_exchange_workspace_bytes(ix._handle, b, k))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.zeros(need, dtype=torch.uint8, device=ix.device)
        seg_arr = None if segments is None else np.ascontiguousarray(segments, dtype=np.int32)
        peer.seq += 1
        with torch.cuda.device(ix.device):
            N.check(lib.mmr_search_exchange(ix._handle, queries.data_ptr(), None if seg_arr is None else seg_arr.ctypes.data,
                                            b, k, peer.ptrs.ctypes.data, self.world, self.rank, peer.seq,
                                            out[0].data_ptr(), out[1].data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                                            _stream_ptr(ix.device)))
        return out

    def search_host(self, queries: np.ndarray, k: int, segments=None) -> Tuple[np.ndarray, np.ndarray]:
        """Host query in, host result out (identical on every rank) through mmr_search_exchange_host: for B <= 2 the query
        rides in the scan kernel's parameters and the merged result lands in a mapped mailbox -- no copies, no stream
        synchronisation.  Needs the fused exchange (symmetric memory)."""
        from . import _native as N
        from .index import _stream_ptr

        q = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.float32)
        b, k = int(q.shape[0]), max(int(k), 1)
        peer = self._peer_exchange(b, k)
        if peer is None:
            out = self.search(torch.from_numpy(q).to(self.local.device), k, segments)
            return out[0].cpu().numpy(), out[1].cpu().numpy()
        ix = self.local
        scores = np.empty((b, k), dtype=np.float32)
        rows = np.empty((b, k), dtype=np.int64)
        seg_arr = None if segments is None else np.ascontiguousarray(segments, dtype=np.int32)
        peer.seq += 1
        if True:   # (the C call selects the device itself)
            N.check(N.lib().mmr_search_exchange_host(ix._handle, q.ctypes.data, None if seg_arr is None else seg_arr.ctypes.data,
                                                     b, k, peer.ptrs.ctypes.data, self.world, self.rank, peer.seq,
                                                     scores.ctypes.data, rows.ctypes.data, _stream_ptr(ix.device)))
        return scores, rows

    def _buffers(self, b: int, k: int):
        if self._wire is None or (self._wire.b, self._wire.k) != (b, k):
            dev = self.local.device
            self._wire = Wire(b, k, dev)
            self._gathered = GatheredWire(self._wire, self.world)
            self._out = (torch.empty((b, k), dtype=torch.float32, device=dev),
                         torch.empty((b, k), dtype=torch.int64, device=dev))
        return self._wire, self._gathered, self._out

    def search(self, queries: torch.Tensor, k: int, segments=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Replicated queries [B, dim] -> global top-k (identical on every rank)."""
        from . import _native as N
        from .index import _stream_ptr

        if queries.dim() == 1:
            queries = queries[None, :]
        b = int(queries.shape[0])
        k = max(int(k), 1)
        wire, gathered, out = self._buffers(b, k)
        peer = self._peer_exchange(b, k)
        if peer is not None:
            return self._search_fused(peer, queries, b, k, segments, out)
        self.local.search(queries, k, segments, out=(wire.scores, wire.rows))
        if self.world == 1:
            return wire.scores, wire.rows
        gather_wire(wire, gathered, self.group)
        scores, rows = gathered.views()
        dev = self.local.device
        with torch.cuda.device(dev):
            N.check(N.lib().mmr_merge_topk_strided(scores.data_ptr(), rows.data_ptr(), wire.nbytes // 4, wire.nbytes // 8,
                                                   self.world, b, k, out[0].data_ptr(), out[1].data_ptr(),
                                                   _stream_ptr(dev)))
        return out
