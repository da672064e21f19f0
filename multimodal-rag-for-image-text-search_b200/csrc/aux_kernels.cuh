// Small kernels either side of the scan: the resident-index loader (L1), the cross-shard merge (K4)
// and the fusion / confidence-gate epilogue (K5).
#pragma once
#include "common.cuh"
#include "topk.cuh"

namespace mmr {

// ------------------------------------------------------------------------------------------------
// L1 -- fp32 rows (the reference's list<float32> embedding column, app/storage/lancedb_store.py:33-44)
// -> resident bf16 / fp16 / fp32 rows.  One warp per row.  normalize != 0 re-applies
// LanceDBStore._normalize (:63-69, f32: x / ||x|| unless ||x|| <= 0) before the narrowing cast, which is
// what _prepare_rows (:71-85) does on the write path.
// ------------------------------------------------------------------------------------------------
template <typename E>
__device__ __forceinline__ E cast_elem(float x);
template <>
__device__ __forceinline__ __nv_bfloat16 cast_elem<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }
template <>
__device__ __forceinline__ __half cast_elem<__half>(float x) { return __float2half_rn(x); }
template <>
__device__ __forceinline__ float cast_elem<float>(float x) { return x; }

// dst_row != nullptr: source row i lands in resident row dst_row[i] (the tenant-sorted position the store computed);
// a negative entry skips the row (deleted / superseded rows of a host block are streamed past, never gathered on the host).
template <typename E>
__global__ void convert_rows_kernel(const float* __restrict__ src, E* __restrict__ dst, int64_t n_rows, int dim,
                                    int normalize, const int64_t* __restrict__ dst_row = nullptr) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = int64_t(gridDim.x) * (blockDim.x >> 5);
  for (int64_t row = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n_rows; row += warps) {
    const int64_t drow = dst_row ? dst_row[row] : row;
    if (drow < 0) continue;
    const float* s = src + row * dim;
    E* d = dst + drow * dim;
    float inv = 1.f;
    bool scale = false;
    if (normalize) {
      float ss = 0.f;
      for (int i = lane; i < dim; i += 32) {
        const float x = s[i];
        ss = fmaf(x, x, ss);
      }
      ss = warp_allreduce_sum(ss);
      const float nrm = sqrtf(ss);
      scale = nrm > 0.f;
      inv = nrm;
    }
    for (int i = lane; i < dim; i += 32) {
      const float x = s[i];
      d[i] = cast_elem<E>(scale ? x / inv : x);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K4 -- merge G shard-local top-k lists into one.  Inputs are what each shard's scan wrote:
// scores [G, B, k] f32 and rows [G, B, k] i64 (global ordinals, -1 = empty), shard g starting at element
// g * score_stride / g * row_stride (so the gathered wire buffers can be read in place).  One warp per query.
// Same total order as everywhere else: (score desc, row asc); result is independent of G.
// ------------------------------------------------------------------------------------------------
template <int KPL>
__global__ void merge_shards_kernel(const float* __restrict__ scores, const int64_t* __restrict__ rows,
                                    int64_t score_stride, int64_t row_stride, int G, int B, int k,
                                    float* __restrict__ out_scores, int64_t* __restrict__ out_rows) {
  const int lane = threadIdx.x & 31;
  const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (qi >= B) return;
  WarpTopK<KPL> m;
  m.clear();
  uint64_t thr = 0ull;
  const int per = k;  // entries per shard for this query
  const int total = G * per;
  for (int i0 = 0; i0 < total; i0 += 32) {
    const int i = i0 + lane;
    bool valid = i < total;
    uint64_t key = 0ull;
    if (valid) {
      const int g = i / per, j = i % per;
      const size_t o = size_t(qi) * k + j;
      const int64_t r = rows[size_t(g) * row_stride + o];
      valid = r >= 0;
      if (valid) key = make_key(scores[size_t(g) * score_stride + o], uint32_t(r));
    }
    thr = m.offer_batch(key, valid, thr, k, lane);
  }
#pragma unroll
  for (int j = 0; j < KPL; ++j) {
    const int pos = j * 32 + lane;
    if (pos < k) {
      const uint64_t key = m.key[j];
      out_scores[size_t(qi) * k + pos] = key ? key_score(key) : -INFINITY;
      out_rows[size_t(qi) * k + pos] = key ? int64_t(key_row(key)) : int64_t(-1);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fused exchange over peer memory (replaces the NCCL all-gather of the tiny per-shard results).
// Every rank owns one symmetric buffer, identical layout everywhere:
//   [0, 256)                       flags  u32 [2 parities][MMR_XCHG_MAX_PEERS]   (sequence numbers, start at 0)
//   [256, ...)                     slots  [2 parities][G source ranks][wire_bytes], wire = scores f32 [B*k] | rows i64 [B*k]
// push_wire_kernel   : CTA g copies this rank's wire into peer g's slot [parity][rank] with peer-mapped stores, fences at
//                      system scope and releases peer g's flag [parity][rank] = seq.  (K1 does this itself in its last CTA.)
// merge_wait_kernel  : one warp per query; lane g acquires flag [parity][g] >= seq (bounded spin), then the warp merges the
//                      G x k candidates out of the local slots (written by the peers) -- K4 fused with the wait.
// Two parities are enough: a rank cannot start search seq+2 before every peer finished merging search seq.
// ------------------------------------------------------------------------------------------------
constexpr int MMR_XCHG_MAX_PEERS = 16;
constexpr int MMR_XCHG_HEADER = 256;

struct PeerPtrs {
  uint64_t slot[MMR_XCHG_MAX_PEERS];  // this rank's slot inside peer g's buffer
  uint64_t flag[MMR_XCHG_MAX_PEERS];  // this rank's flag inside peer g's buffer
};

__global__ void push_wire_kernel(const uint8_t* __restrict__ wire, uint32_t wire_bytes, const PeerPtrs peers, uint32_t seq) {
  const int g = blockIdx.x;
  const uint64_t* peer_slot = peers.slot;
  const uint64_t* peer_flag = peers.flag;
  const uint4* src = reinterpret_cast<const uint4*>(wire);
  uint4* dst = reinterpret_cast<uint4*>(peer_slot[g]);
  const uint32_t n16 = wire_bytes / 16;
  for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
  for (uint32_t i = n16 * 16 + threadIdx.x; i < wire_bytes; i += blockDim.x)
    reinterpret_cast<uint8_t*>(peer_slot[g])[i] = wire[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flag[g]), "r"(seq) : "memory");
}

__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int KPL>
__device__ __forceinline__ void merge_wait_body(const uint8_t* __restrict__ xbuf, int parity, uint32_t seq,
                                                uint32_t wire_bytes, uint32_t score_bytes, int G, int B, int k,
                                                float* __restrict__ out_scores, int64_t* __restrict__ out_rows,
                                                uint64_t timeout_ns, int qi, int lane);

template <int KPL>
__global__ void merge_wait_kernel(const uint8_t* __restrict__ xbuf, int parity, uint32_t seq, uint32_t wire_bytes,
                                  uint32_t score_bytes, int G, int B, int k, float* __restrict__ out_scores,
                                  int64_t* __restrict__ out_rows, uint64_t timeout_ns, uint32_t* done_flag,
                                  uint32_t done_seq) {
  const int lane = threadIdx.x & 31;
  const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  // let the next search's scan start while we wait for the peers (it uses the other slot parity)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (done_flag != nullptr) {
    // single-block form (B <= 4, host-buffer calls): out_* are mapped host memory; once every warp has written its
    // result the block releases the completion flag at system scope and the host stops spinning
    merge_wait_body<KPL>(xbuf, parity, seq, wire_bytes, score_bytes, G, B, k, out_scores, out_rows, timeout_ns, qi, lane);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(done_flag), "r"(done_seq) : "memory");
    return;
  }
  merge_wait_body<KPL>(xbuf, parity, seq, wire_bytes, score_bytes, G, B, k, out_scores, out_rows, timeout_ns, qi, lane);
}

template <int KPL>
__device__ __forceinline__ void merge_wait_body(const uint8_t* __restrict__ xbuf, int parity, uint32_t seq,
                                                uint32_t wire_bytes, uint32_t score_bytes, int G, int B, int k,
                                                float* __restrict__ out_scores, int64_t* __restrict__ out_rows,
                                                uint64_t timeout_ns, int qi, int lane) {
  if (qi >= B) return;
  const uint32_t* flags = reinterpret_cast<const uint32_t*>(xbuf) + parity * MMR_XCHG_MAX_PEERS;
  bool ok = true;
  if (lane < G) {
    const uint64_t t0 = globaltimer_ns();
    while (ld_acquire_sys_u32(flags + lane) < seq) {
      if (globaltimer_ns() - t0 > timeout_ns) {  // a peer died: do not hang the GPU, report instead
        ok = false;
        break;
      }
    }
  }
  ok = __all_sync(0xffffffffu, ok);
  if (!ok) {
    if (lane == 0) {
      out_rows[size_t(qi) * k] = -2;  // MMR exchange timeout marker
      out_scores[size_t(qi) * k] = -INFINITY;
    }
    return;
  }
  const uint8_t* slots = xbuf + MMR_XCHG_HEADER + size_t(parity) * G * wire_bytes;
  WarpTopK<KPL> m;
  m.clear();
  uint64_t thr = 0ull;
  const int total = G * k;
  for (int i0 = 0; i0 < total; i0 += 32) {
    const int i = i0 + lane;
    bool valid = i < total;
    uint64_t key = 0ull;
    if (valid) {
      const int g = i / k, j = i % k;
      const uint8_t* slot = slots + size_t(g) * wire_bytes;
      const size_t o = size_t(qi) * k + j;
      const long long r = __ldcg(reinterpret_cast<const long long*>(slot + score_bytes) + o);  // L2: written by peers
      valid = r >= 0;
      if (valid) key = make_key(__ldcg(reinterpret_cast<const float*>(slot) + o), uint32_t(r));
    }
    thr = m.offer_batch(key, valid, thr, k, lane);
  }
#pragma unroll
  for (int j = 0; j < KPL; ++j) {
    const int pos = j * 32 + lane;
    if (pos < k) {
      const uint64_t key = m.key[j];
      out_scores[size_t(qi) * k + pos] = key ? key_score(key) : -INFINITY;
      out_rows[size_t(qi) * k + pos] = key ? int64_t(key_row(key)) : int64_t(-1);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K5 -- text+image score fusion and the CONFIDENCE_TAU gate for the rerank-off path.
// Follows _z_scores / _fuse_results (reference app/ml/retrieve.py:186-195, 158-183) and _confidence_low
// (app/ml/generate.py:56-60) operation by operation so the result is bit-identical to the reference:
//   v_i      = 1.0 - double(f32(1) - cos_i)          (_format_results, lancedb_store.py:130-131)
//   arr      = float32(v)                             (np.array(numeric, dtype=np.float32))
//   mean,std = numpy float32 mean / population std:   pairwise sum with 8 accumulators for n >= 8,
//              plain left-to-right for n < 8, divide by n in f32, sqrt in f32
//   z_i      = (v_i - double(mean)) / double(std)     in float64; all 0.0 when std == 0
//   combined = z_i  (np.mean of a one-element list), items = text then image, stable sort descending,
//   cut to final_n; low_conf = no items or max(combined) < tau.
// One thread per query (k <= 64 per modality; the work is ~200 flops).  No FMA contraction anywhere.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float np_pairwise_sum_f32(const float* a, int n) {
  if (n < 8) {
    float res = 0.f;
    for (int i = 0; i < n; ++i) res = __fadd_rn(res, a[i]);
    return res;
  }
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  int i = 8;
  const int lim = n - (n % 8);
  for (; i < lim; i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
  }
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __fadd_rn(res, a[i]);
  return res;
}

constexpr int FUSE_MAXK = 64;

// z-scores of n cosine scores -> z[0..n) (float64) ; v (the python-float score) also returned
__device__ __forceinline__ void np_z_scores(const float* cos, int n, double* v, double* z) {
  float arr[FUSE_MAXK], sq[FUSE_MAXK];
  for (int i = 0; i < n; ++i) {
    const float d = __fsub_rn(1.0f, cos[i]);
    v[i] = __dsub_rn(1.0, double(d));
    arr[i] = __double2float_rn(v[i]);
  }
  const float fn = float(n);
  const float mean = __fdiv_rn(np_pairwise_sum_f32(arr, n), fn);
  for (int i = 0; i < n; ++i) {
    const float x = __fsub_rn(arr[i], mean);
    sq[i] = __fmul_rn(x, x);
  }
  const float var = __fdiv_rn(np_pairwise_sum_f32(sq, n), fn);
  const float sd = __fsqrt_rn(var);
  if (sd == 0.f) {
    for (int i = 0; i < n; ++i) z[i] = 0.0;
  } else {
    const double dm = double(mean), ds = double(sd);
    for (int i = 0; i < n; ++i) z[i] = __ddiv_rn(__dsub_rn(v[i], dm), ds);
  }
}

// z-scores of n python-float scores v (float64) -> z; same arithmetic as np_z_scores without the cosine step
__device__ __forceinline__ void np_z_scores_f64(const double* v, int n, double* z) {
  float arr[FUSE_MAXK], sq[FUSE_MAXK];
  for (int i = 0; i < n; ++i) arr[i] = __double2float_rn(v[i]);
  const float fn = float(n);
  const float mean = __fdiv_rn(np_pairwise_sum_f32(arr, n), fn);
  for (int i = 0; i < n; ++i) {
    const float x = __fsub_rn(arr[i], mean);
    sq[i] = __fmul_rn(x, x);
  }
  const float sd = __fsqrt_rn(__fdiv_rn(np_pairwise_sum_f32(sq, n), fn));
  if (sd == 0.f) {
    for (int i = 0; i < n; ++i) z[i] = 0.0;
  } else {
    const double dm = double(mean), ds = double(sd);
    for (int i = 0; i < n; ++i) z[i] = __ddiv_rn(__dsub_rn(v[i], dm), ds);
  }
}

// stable descending order of key[0..n) -> idx (insertion sort: n <= 128, one thread per request)
__device__ __forceinline__ void stable_desc_order(const double* key, int n, int* idx) {
  for (int i = 0; i < n; ++i) {
    int j = i;
    while (j > 0 && key[idx[j - 1]] < key[i]) {
      idx[j] = idx[j - 1];
      --j;
    }
    idx[j] = i;
  }
}

// K5 (complete form): _rerank_text's reordering + _fuse_results + _confidence_low for one request per thread, from the
// float64 scores the reference sees (reference app/ml/retrieve.py:132-195, app/ml/generate.py:56-60).
//   text_scores [B, kt] in scan order, text_count [B]; rerank [B, kt]: the cross-encoder logits, assigned -- exactly
//   as `zip(top_candidates, scores)` does -- to the FIRST rr_count[b] text items (rr_count NULL / 0 = no rerank);
//   img_scores [B, ki], img_count [B].  out_index: text item j (scan position) -> j, image item j -> kt + j, -1 = none.
__global__ void fuse_full_kernel(const double* __restrict__ text_scores, const int32_t* __restrict__ text_count,
                                 const double* __restrict__ rerank, const int32_t* __restrict__ rr_count,
                                 const double* __restrict__ img_scores, const int32_t* __restrict__ img_count, int kt,
                                 int ki, int B, int final_n, double tau, double* __restrict__ out_combined,
                                 int32_t* __restrict__ out_index, uint8_t* __restrict__ out_low_conf) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int nt = text_count ? min(text_count[b], kt) : 0;
  const int ni = img_count ? min(img_count[b], ki) : 0;
  const int nr = (rr_count && rerank) ? min(rr_count[b], nt) : 0;
  double key[FUSE_MAXK], v[FUSE_MAXK], z[FUSE_MAXK], rr[FUSE_MAXK], zr[FUSE_MAXK];
  double comb[2 * FUSE_MAXK];
  int order[FUSE_MAXK], src[2 * FUSE_MAXK], fin[2 * FUSE_MAXK];
  // _rerank_text: sort by rerank_score where present, else score (stable, descending); untouched without logits
  for (int j = 0; j < nt; ++j) {
    key[j] = j < nr ? rerank[size_t(b) * kt + j] : text_scores[size_t(b) * kt + j];
    order[j] = j;
  }
  if (nr > 0) stable_desc_order(key, nt, order);
  // _fuse_results
  int nrr = 0;
  for (int j = 0; j < nt; ++j) {
    v[j] = text_scores[size_t(b) * kt + order[j]];
    if (order[j] < nr) rr[nrr++] = rerank[size_t(b) * kt + order[j]];
  }
  if (nt) np_z_scores_f64(v, nt, z);
  if (nrr) np_z_scores_f64(rr, nrr, zr);
  for (int j = 0; j < nt; ++j) {
    // np.mean of [z_cos] or [z_cos, z_rerank]: pairwise sum starting at 0.0, then a true divide by the count
    comb[j] = j < nrr ? __ddiv_rn(__dadd_rn(__dadd_rn(0.0, z[j]), zr[j]), 2.0) : __ddiv_rn(__dadd_rn(0.0, z[j]), 1.0);
    src[j] = order[j];
  }
  if (ni) {
    for (int j = 0; j < ni; ++j) v[j] = img_scores[size_t(b) * ki + j];
    np_z_scores_f64(v, ni, z);
    for (int j = 0; j < ni; ++j) {
      comb[nt + j] = z[j];
      src[nt + j] = kt + j;
    }
  }
  const int n = nt + ni;
  stable_desc_order(comb, n, fin);
  for (int o = 0; o < final_n; ++o) {
    const size_t oo = size_t(b) * final_n + o;
    out_combined[oo] = o < n ? comb[fin[o]] : 0.0;
    out_index[oo] = o < n ? src[fin[o]] : -1;
  }
  out_low_conf[b] = (n == 0 || final_n <= 0) ? 1 : (comb[fin[0]] < tau ? 1 : 0);
}

__global__ void fuse_kernel(const float* __restrict__ text_scores, const int64_t* __restrict__ text_rows, int kt,
                            const float* __restrict__ img_scores, const int64_t* __restrict__ img_rows, int ki, int B,
                            int final_n, double tau, double* __restrict__ out_combined,
                            double* __restrict__ out_score, int64_t* __restrict__ out_rows,
                            int8_t* __restrict__ out_modality, uint8_t* __restrict__ out_low_conf) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= B) return;
  double v[2 * FUSE_MAXK], z[2 * FUSE_MAXK];
  // valid hits are a prefix (the scan pads with row = -1 at the end)
  int nt = 0, ni = 0;
  if (text_rows) {
    while (nt < kt && text_rows[size_t(qi) * kt + nt] >= 0) ++nt;
  }
  if (img_rows) {
    while (ni < ki && img_rows[size_t(qi) * ki + ni] >= 0) ++ni;
  }
  if (nt) np_z_scores(text_scores + size_t(qi) * kt, nt, v, z);
  if (ni) np_z_scores(img_scores + size_t(qi) * ki, ni, v + nt, z + nt);
  const int n = nt + ni;
  // stable selection of the final_n largest combined scores (n <= 128, final_n small)
  unsigned long long taken_lo = 0ull, taken_hi = 0ull;
  double best0 = 0.0;
  for (int o = 0; o < final_n; ++o) {
    int bi = -1;
    double bz = 0.0;
    for (int i = 0; i < n; ++i) {
      const bool taken = i < 64 ? ((taken_lo >> i) & 1ull) : ((taken_hi >> (i - 64)) & 1ull);
      if (taken) continue;
      if (bi < 0 || z[i] > bz) {  // strict > keeps the earliest among equals = stable descending sort
        bi = i;
        bz = z[i];
      }
    }
    const size_t oo = size_t(qi) * final_n + o;
    if (bi < 0) {
      out_combined[oo] = 0.0;
      out_score[oo] = 0.0;
      out_rows[oo] = -1;
      out_modality[oo] = -1;
      continue;
    }
    if (bi < 64) taken_lo |= 1ull << bi; else taken_hi |= 1ull << (bi - 64);
    if (o == 0) best0 = bz;
    out_combined[oo] = bz;
    out_score[oo] = v[bi];
    const bool is_text = bi < nt;
    out_rows[oo] = is_text ? text_rows[size_t(qi) * kt + bi] : img_rows[size_t(qi) * ki + (bi - nt)];
    out_modality[oo] = is_text ? 0 : 1;
  }
  // _confidence_low looks at the returned items only (generate.py:56-60)
  out_low_conf[qi] = (n == 0 || final_n <= 0) ? 1 : (best0 < tau ? 1 : 0);
}

}  // namespace mmr
