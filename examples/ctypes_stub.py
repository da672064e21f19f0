"""The reference-side ctypes binding shown in INTEGRATION.md, kept as a runnable file (tests/test_gpu_c_abi.py runs it).
Point MMR_LIB at libmmr_b200.so if it is not next to the package."""
import os
_HERE = os.path.dirname(os.path.abspath(__file__))
_DEFAULT_LIB = os.path.join(_HERE, "..", "multimodal-rag-for-image-text-search_b200", "libmmr_b200.so")
import ctypes as C, json, numpy as np, torch

_lib = C.CDLL(os.environ.get("MMR_LIB", _DEFAULT_LIB))
_lib.mmr_last_error.restype = C.c_char_p
_lib.mmr_index_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64,
                                  C.POINTER(C.c_void_p)]
_lib.mmr_search_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]

def _check(rc):
    if rc != 0:
        raise RuntimeError(_lib.mmr_last_error().decode())

class ResidentTable:
    """rows: torch.bfloat16 [n, dim] on cuda, tenant-sorted; seg: int64 [T+1] host offsets; chunk_ids/metas: host lists."""
    def __init__(self, rows, seg, chunk_ids, metas, tenant_of):
        self.rows, self.seg = rows, np.ascontiguousarray(seg, np.int64)
        self.chunk_ids, self.metas, self.tenant_of = chunk_ids, metas, tenant_of
        self.h = C.c_void_p()
        _check(_lib.mmr_index_create(rows.device.index, rows.shape[1], 0, rows.shape[0], rows.data_ptr(),
                                     self.seg.ctypes.data, len(self.seg) - 1, 0, C.byref(self.h)))

    def search(self, user_id, query_vec, top_k):                      # == LanceDBStore.search_text / search_image
        t = self.tenant_of.get(str(user_id))
        if t is None:
            return []
        k = max(int(top_k), 1)
        q = np.ascontiguousarray(query_vec, np.float32)[None, :]
        seg = np.array([t], np.int32)
        scores, ids = np.empty((1, k), np.float32), np.empty((1, k), np.int64)
        _check(_lib.mmr_search_host(self.h, q.ctypes.data, seg.ctypes.data, 1, k, scores.ctypes.data, ids.ctypes.data,
                                    torch.cuda.current_stream().cuda_stream))
        out = []
        for s, r in zip(scores[0], ids[0]):
            if r < 0:
                break
            distance = float(np.float32(1.0) - s)                      # Lance's _distance
            out.append({"chunk_id": self.chunk_ids[r], "score": 1.0 - distance,          # _format_results :130-131
                        "meta": json.loads(self.metas[r] or "{}")})
        return out
