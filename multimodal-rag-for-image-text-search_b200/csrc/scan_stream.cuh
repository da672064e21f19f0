// K1 -- HBM-streaming exact scan for small query groups (1..4 queries sharing one row range, 8 on fp32 rows).
//
// Replaces the Lance flat KNN behind LanceDBStore.search_text / search_image
// (reference app/storage/lancedb_store.py:103-123): cos(q, x_n) for every row of the tenant's row
// range, keep the best k under (score desc, row asc).
//
// Shape of the kernel (bandwidth-bound: algorithmic bytes = rows * D * sizeof(elem), read once):
//   * persistent grid, one CTA per SM, NW consumer warps per CTA;
//   * every warp owns a private ring of S stages in shared memory.  One stage = R consecutive index rows
//     (R * D * sizeof(elem) contiguous bytes), filled by ONE 1-D bulk async copy (cp.async.bulk, the TMA
//     engine without a descriptor) that signals an mbarrier.  NW * S stages are in flight, independent of register
//     pressure (tuned geometry below: 8 warps x 2 stages x 4 rows = 64 KB per SM for one query; deeper / wider for groups);
//   * chunk c of the row range goes to global warp (c mod total_warps): at any instant the chip reads
//     one contiguous window of HBM;
//   * per stage a warp computes R x NQ dot products: lane l holds 8/16-byte column slices of each row and
//     the matching slice of every (normalised, fp32) query in registers; partial sums are combined with
//     a transposing butterfly (V values -> V-1 shuffles, result spread one value per lane group);
//   * scores never leave registers: a warp-resident sorted top-k list (topk.cuh) takes the rare score
//     that beats the current k-th key;
//   * warp lists -> CTA list (shared memory) -> per-CTA partial in global memory; the last CTA to finish
//     (atomic ticket) merges all partials and writes the final [NQ, k] (score, row) -- one launch.
#pragma once
#include "common.cuh"
#include "topk.cuh"

namespace mmr {

// Ring geometry, tuned on B200 (profiles/r01_k1_summary.md, "bytes in flight"): 8 warps x 2 stages x 4 rows.
// Fewer bytes in flight per SM stream FASTER here: 64 KB/SM reached 7.4-7.5 TB/s where the first version's 192 KB/SM
// (3 stages x 8 rows) reached 6.8-6.9 TB/s on the same box; 2-row stages lose again (per-stage overheads).
#ifndef MMR_K1_NW
#define MMR_K1_NW 8
#endif
#ifndef MMR_K1_STAGES
#define MMR_K1_STAGES 2
#endif
#ifndef MMR_K1_ROWS
#define MMR_K1_ROWS 4
#endif
// Groups of 2 / 4+ queries spend longer on a stage (2x / 4x the FMAs per row byte), and with two stages that compute time
// ADDS to the copy latency instead of hiding under it (a warp's period is max(S*Tc, L + Tc) per S stages): deeper rings
// for them.  The single-query ring stays at two stages (more bytes in flight made it slower, see above).
#ifndef MMR_K1_STAGES_NQ2
#define MMR_K1_STAGES_NQ2 3   // measured: K1 B = 2 on 10M rows 1.427 -> 1.394 ms, K6 2-query items 6.69 -> 6.82 TB/s
#endif
#ifndef MMR_K1_STAGES_NQ4
#define MMR_K1_STAGES_NQ4 MMR_K1_STAGES
#endif
// Four-query groups on 16-bit rows are ISSUE-paced, not latency-paced (ncu, profiles/r02_k6_summary.md: 57 % issue-active
// with two warps per scheduler, top stall "wait" = fixed-latency FFMA dependencies; ring depth changes nothing): they get
// three warps per scheduler.
#ifndef MMR_K1_NW_NQ4
#define MMR_K1_NW_NQ4 12
#endif
constexpr int K1_NW = MMR_K1_NW;   // consumer warps per CTA (1- and 2-query groups, fp32 rows)
constexpr int MMR_MAX_PEERS = 16;

constexpr int MMR_MAX_K_DEVICE = 64;  // == MMR_MAX_K (include/mmr_b200.h)
constexpr int K1_ITEM_NQ = 4;        // queries per varlen work item (one pass over the item's rows serves all of them)
constexpr int K1_INLINE_FLOATS = 1024;  // queries that ride in the kernel parameters (2 x 512 floats)

struct ScanItem {  // varlen mode: one contiguous row range scanned for up to K1_ITEM_NQ queries that share it
  uint32_t row_begin;
  uint32_t row_end;
  int32_t query[K1_ITEM_NQ];  // query ordinals; entries >= nq are ignored
  int32_t nq;
  int32_t out;                // which list group of `partial` this item writes (items are launched sorted by size)
};

struct QuerySlot {  // varlen mode: where query b's partial lists are: items [item0, item0 + n_items), list `slot` of each
  int32_t item0;
  int32_t n_items;
  int32_t slot;
  int32_t pad;
};

struct StreamParams {
  const void* rows;      // [n_rows, D] row-major, elem type E
  const float* queries;  // [B, D] fp32, not necessarily unit norm (re-normalised here, lancedb_store.py:104);
                         // nullptr = the queries are in `qinline` (host-buffer calls: no H2D copy at all)
  int32_t q_first;       // first query of this group
  int32_t nq;            // live queries in this group (1..NQ)
  int32_t k;
  uint32_t row_begin, row_end;  // uniform mode: the shared row range
  uint64_t* partial;            // uniform: [grid, NQ, k]; varlen: [n_items, NQ, k]
  unsigned int* ticket;         // two words, zero on entry, left zero on exit (uniform: arrival ticket; varlen: claim + done)
  float* out_scores;            // [B, k]  (uniform mode only; may be mapped host memory)
  int64_t* out_rows;            // [B, k]
  int64_t row_base;             // shard base added to the int64 row ids written out
  const ScanItem* items;        // varlen mode (nullptr = uniform)
  int32_t n_items;
  int32_t item_nq;              // varlen mode: lists per item in `partial` (K1_ITEM_NQ, whatever NQ this launch runs with)
  // Fused exchange (row-range shards over NVLink): the last CTA also stores the final [B, k] result straight into
  // every peer's symmetric buffer (this rank's slot) and then releases that peer's flag with `seq`.
  int32_t n_peers;              // 0 = no exchange
  uint32_t seq;
  uint32_t wire_score_bytes;    // offset of the int64 rows inside one wire slot
  uint64_t peer_slot[MMR_MAX_PEERS];  // address of this rank's slot inside peer g's buffer (peer-mapped)
  uint64_t peer_flag[MMR_MAX_PEERS];  // address of this rank's flag inside peer g's buffer
  // Completion mailbox (host-buffer calls): after the results are written, done_flag (mapped host memory) is released
  // with done_seq at system scope, so the host can spin on it instead of synchronising the stream.
  uint32_t* done_flag;
  uint32_t done_seq;
  float qinline[K1_INLINE_FLOATS];
};

template <typename E>
struct ElemTraits;
template <>
struct ElemTraits<__nv_bfloat16> {
  static constexpr int BYTES = 2;
  // 32-bit word holds two bf16: low half = element 2i, high half = element 2i+1
  static __device__ __forceinline__ void unpack(uint32_t w, float& a, float& b) {
    a = __uint_as_float(w << 16);
    b = __uint_as_float(w & 0xFFFF0000u);
  }
};
template <>
struct ElemTraits<__half> {
  static constexpr int BYTES = 2;
  static __device__ __forceinline__ void unpack(uint32_t w, float& a, float& b) {
    const __half2 h = *reinterpret_cast<const __half2*>(&w);
    const float2 f = __half22float2(h);
    a = f.x;
    b = f.y;
  }
};
template <>
struct ElemTraits<float> {
  static constexpr int BYTES = 4;
};

// V partial sums per lane -> full sums, value i ending up (replicated) in lanes [i*32/V, (i+1)*32/V).
// Each halving step exchanges half of the live values with the partner lane, so V values cost V-1
// shuffles instead of 5V.
template <int N, int OFF>
__device__ __forceinline__ void transpose_reduce_step(float* v, int lane) {
  if constexpr (OFF >= 1) {
    if constexpr (N > 1) {
      constexpr int H = N / 2;
      const bool upper = (lane & OFF) != 0;
#pragma unroll
      for (int i = 0; i < H; ++i) {
        const float send = upper ? v[i] : v[i + H];
        const float keep = upper ? v[i + H] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
      }
      transpose_reduce_step<H, OFF / 2>(v, lane);
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], OFF);
      transpose_reduce_step<1, OFF / 2>(v, lane);
    }
  }
}
template <int V>
__device__ __forceinline__ float transpose_reduce(float (&v)[V], int lane) {
  static_assert(V >= 1 && V <= 32 && (V & (V - 1)) == 0, "V must be a power of two <= 32");
  transpose_reduce_step<V, 16>(v, lane);
  return v[0];
}

template <typename E, int D, int NQ, int KPL>
struct StreamCfg {
  static constexpr int EB = ElemTraits<E>::BYTES;
  static constexpr int ROW_BYTES = D * EB;
  static constexpr int R = MMR_K1_ROWS;                              // rows per stage
  static constexpr int V = R * NQ;                                   // dot products per stage per warp
  static constexpr int STAGE_BYTES = R * ROW_BYTES;
  static constexpr int S_WANT = NQ >= 4 ? MMR_K1_STAGES_NQ4 : NQ >= 2 ? MMR_K1_STAGES_NQ2 : MMR_K1_STAGES;
  static constexpr int NW = (NQ >= 4 && EB == 2) ? MMR_K1_NW_NQ4 : K1_NW;   // consumer warps per CTA
  static constexpr int THREADS = NW * 32;
  static constexpr int S_FIT = (160 * 1024) / (NW * STAGE_BYTES);   // fp32 rows: 8 KB stages, the ring stays <= 160 KB
  static constexpr int S = S_WANT < S_FIT ? S_WANT : (S_FIT < 2 ? 2 : S_FIT);   // stages per warp
  static constexpr int VECB = (ROW_BYTES % 512 == 0) ? 16 : 8;       // bytes per lane per vector load
  static constexpr int NV = ROW_BYTES / (32 * VECB);                 // vector loads per lane per row
  static constexpr int RPV = VECB / 4;                               // 32-bit registers per vector
  static constexpr int EPR = 4 / EB;                                 // elements per register
  static constexpr int EPL = D / 32;                                 // elements per lane per row
  static constexpr int KSLOTS = 32 * KPL;
  static constexpr int RING_BYTES = NW * S * STAGE_BYTES;
  static constexpr int SMEM_BYTES = RING_BYTES + NW * S * 8 + 64;
  static_assert(ROW_BYTES % (32 * 8) == 0, "D * sizeof(elem) must be a multiple of 256 bytes");
  static_assert(S >= 2, "ring too shallow");
  static_assert(V <= 32, "too many dot products per stage");
  static_assert(NW * NQ * KSLOTS * 8 <= RING_BYTES, "merge scratch must fit in the ring");
};

template <typename E, int D, int NQ, int KPL>
__global__ void __launch_bounds__((StreamCfg<E, D, NQ, KPL>::THREADS), 1) scan_stream_kernel(const __grid_constant__ StreamParams p) {
  using C = StreamCfg<E, D, NQ, KPL>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::RING_BYTES);
  __shared__ int s_last;

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int k = p.k;

  // Programmatic dependent launch: the next search on this stream may start its CTAs as ours retire (its scan only
  // reads the index and its own queries); it waits (griddepcontrol.wait below) before it touches the shared workspace.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (p.items != nullptr) asm volatile("griddepcontrol.wait;" ::: "memory");  // varlen writes partials as it goes

  // ---- this warp's ring ----
  const uint32_t ring0 = smem_u32(smem) + uint32_t(warp) * (C::S * C::STAGE_BYTES);
  const uint32_t bar0 = smem_u32(bars) + uint32_t(warp) * (C::S * 8);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < C::S; ++s) mbar_init(bar0 + s * 8, 1);
    fence_mbar_init();
  }
  __syncwarp();

  // ---- queries: normalise (f32, q / ||q||) and keep this lane's column slices in registers ----
  float q[NQ][C::EPL];
#pragma unroll
  for (int qi = 0; qi < NQ; ++qi) {
    const bool live = qi < p.nq;
    const float* qbase = p.queries ? p.queries : p.qinline;
    const float* qsrc = qbase + size_t(p.items ? 0 : (p.q_first + (live ? qi : 0))) * D;
    float ss = 0.f;
#pragma unroll
    for (int t = 0; t < C::NV; ++t)
#pragma unroll
      for (int j = 0; j < C::RPV * C::EPR; ++j) {
        const int e = (t * 32 + lane) * (C::RPV * C::EPR) + j;
        const float x = (live && !p.items) ? qsrc[e] : 0.f;
        q[qi][t * C::RPV * C::EPR + j] = x;
        ss += x * x;
      }
    ss = warp_allreduce_sum(ss);
    const float nrm = sqrtf(ss);
    if (nrm > 0.f) {
#pragma unroll
      for (int e = 0; e < C::EPL; ++e) q[qi][e] = q[qi][e] / nrm;
    }
  }

  WarpTopK<KPL> list[NQ];
  uint64_t thr[NQ];
#pragma unroll
  for (int qi = 0; qi < NQ; ++qi) {
    list[qi].clear();
    thr[qi] = 0ull;
  }

  const int gwarp = blockIdx.x * C::NW + warp;
  const int twarps = gridDim.x * C::NW;
  const uint8_t* rows8 = reinterpret_cast<const uint8_t*>(p.rows);

  int stage = 0;          // next ring stage this warp consumes (persists across row ranges)
  uint32_t phase = 0;     // mbarrier parity of that stage
  int nq_live = p.nq;     // live queries of the current job (varlen: per item)

  // One "job" = a row range scanned with chunk index c = first, first+stride, ...
  auto scan_range = [&](uint32_t row_begin, uint32_t row_end, int first, int stride) {
    const uint32_t nrows = row_end > row_begin ? row_end - row_begin : 0u;
    const int nchunks = int((nrows + C::R - 1) / C::R);
    auto issue = [&](int c, int s) {
      // lane 0 only
      const uint32_t r0 = row_begin + uint32_t(c) * C::R;
      const uint32_t nr = min(uint32_t(C::R), row_end - r0);
      const uint32_t bytes = nr * C::ROW_BYTES;
      const uint32_t bar = bar0 + s * 8;
      mbar_arrive_expect_tx(bar, bytes);
      bulk_g2s(ring0 + s * C::STAGE_BYTES, rows8 + size_t(r0) * C::ROW_BYTES, bytes, bar);
    };
    // prologue: fill the ring
    if (lane == 0) {
      int sj = stage;
#pragma unroll
      for (int j = 0; j < C::S; ++j) {
        const int c = first + j * stride;
        if (c < nchunks) issue(c, sj);
        sj = (sj + 1 == C::S) ? 0 : sj + 1;
      }
    }
    for (int c = first; c < nchunks; c += stride) {
      const int s = stage;
      mbar_wait(bar0 + s * 8, phase);

      const uint32_t base = ring0 + s * C::STAGE_BYTES + lane * C::VECB;
      float acc[C::V];
#pragma unroll
      for (int i = 0; i < C::V; ++i) acc[i] = 0.f;
#pragma unroll
      for (int r = 0; r < C::R; ++r) {
#pragma unroll
        for (int t = 0; t < C::NV; ++t) {
          uint32_t w[C::RPV];
          const uint32_t addr = base + r * C::ROW_BYTES + t * (32 * C::VECB);
          if constexpr (C::VECB == 16) {
            lds_v4(addr, reinterpret_cast<uint32_t(&)[4]>(w));
          } else {
            lds_v2(addr, reinterpret_cast<uint32_t(&)[2]>(w));
          }
#pragma unroll
          for (int j = 0; j < C::RPV; ++j) {
            if constexpr (C::EB == 2) {
              float a, b;
              ElemTraits<E>::unpack(w[j], a, b);
#pragma unroll
              for (int qi = 0; qi < NQ; ++qi) {
                acc[r * NQ + qi] = fmaf(a, q[qi][(t * C::RPV + j) * 2 + 0], acc[r * NQ + qi]);
                acc[r * NQ + qi] = fmaf(b, q[qi][(t * C::RPV + j) * 2 + 1], acc[r * NQ + qi]);
              }
            } else {
              const float a = __uint_as_float(w[j]);
#pragma unroll
              for (int qi = 0; qi < NQ; ++qi)
                acc[r * NQ + qi] = fmaf(a, q[qi][t * C::RPV + j], acc[r * NQ + qi]);
            }
          }
        }
      }
      // the stage has been read into registers: refill it (generic reads -> async-proxy write)
      __syncwarp();
      if (lane == 0) {
        const int cn = c + C::S * stride;
        if (cn < nchunks) {
          fence_proxy_async();
          issue(cn, s);
        }
      }

      const float score = transpose_reduce<C::V>(acc, lane);
      constexpr int LPV = 32 / C::V;  // lanes per value
      const int vi = lane / LPV;      // value index = r * NQ + qi
      const int r = vi / NQ;
      const int myq = vi % NQ;
      const uint32_t row = row_begin + uint32_t(c) * C::R + uint32_t(r);
      // NaN scores (tombstoned rows are overwritten with NaN by the store) never enter a list
      const bool owner = (lane % LPV == 0) && row < row_end && score == score;
      const uint64_t key = make_key(score, row);
#pragma unroll
      for (int qi = 0; qi < NQ; ++qi) {
        if (qi < nq_live) thr[qi] = list[qi].offer(key, owner && myq == qi, thr[qi], k, lane);
      }
      if (++stage == C::S) {
        stage = 0;
        phase ^= 1u;
      }
    }
  };

  if (p.items == nullptr) {
    // ------------------------------ uniform mode ------------------------------
    scan_range(p.row_begin, p.row_end, gwarp, twarps);

    // the previous search on this stream must be completely done before we touch partials / ticket / outputs
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // warp lists -> shared memory (ring is idle now: every issued copy has been consumed)
    __syncthreads();
    uint64_t* wk = reinterpret_cast<uint64_t*>(smem);  // [NW][NQ][KSLOTS]
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
#pragma unroll
      for (int j = 0; j < KPL; ++j) wk[(warp * NQ + qi) * C::KSLOTS + j * 32 + lane] = list[qi].key[j];
    }
    __syncthreads();
    // CTA-level merge: warp qi merges the NW lists of query qi and writes the CTA partial
    if (warp < p.nq) {
      const int qi = warp;
      WarpTopK<KPL> m;
      m.clear();
      uint64_t t = 0ull;
      // all heads first (position-major over the NW warp lists), so most later candidates fail the ballot
      t = m.template merge_batched<4>(
          [&](int i) -> uint64_t { return wk[((i % C::NW) * NQ + qi) * C::KSLOTS + i / C::NW]; }, C::NW * k, t, k, lane);
      m.store(p.partial + (size_t(blockIdx.x) * NQ + qi) * k, k, lane);
    }
    // grid-level merge by the last CTA to arrive
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int prev = atomicAdd(p.ticket, 1u);
      s_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int nparts = gridDim.x;
    for (int qi = 0; qi < p.nq; ++qi) {
      WarpTopK<KPL> m;
      m.clear();
      uint64_t t = 0ull;
      {
        // this warp's share of the per-CTA lists: parts warp, warp + NW, ...; candidates are visited position-major
        // (all heads first) so the threshold rises early, and fetched 8 per lane per round trip through L2
        const int nmine = (nparts - warp + C::NW - 1) / C::NW;
        const uint64_t* base = p.partial;
        t = m.template merge_batched<8>(
            [&](int i) -> uint64_t {
              const int pos = i / nmine, part = warp + (i % nmine) * C::NW;
              return __ldcg(reinterpret_cast<const unsigned long long*>(base) + (size_t(part) * NQ + qi) * k + pos);
            },
            nmine * k, t, k, lane);
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < KPL; ++j) wk[warp * C::KSLOTS + j * 32 + lane] = m.key[j];
      __syncthreads();
      if (warp == 0) {
        WarpTopK<KPL> f;
        f.clear();
        uint64_t tf = 0ull;
        tf = f.template merge_batched<4>(
            [&](int i) -> uint64_t { return wk[(i % C::NW) * C::KSLOTS + i / C::NW]; }, C::NW * k, tf, k, lane);
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
          const int pos = j * 32 + lane;
          if (pos < k) {
            const uint64_t key = f.key[j];
            const size_t o = size_t(p.q_first + qi) * k + pos;
            const float sc = key ? key_score(key) : -INFINITY;
            const int64_t rw = key ? int64_t(key_row(key)) + p.row_base : int64_t(-1);
            p.out_scores[o] = sc;
            p.out_rows[o] = rw;
            for (int g = 0; g < p.n_peers; ++g) {  // peer-mapped stores over NVLink (or local for g == rank)
              reinterpret_cast<float*>(p.peer_slot[g])[o] = sc;
              reinterpret_cast<int64_t*>(p.peer_slot[g] + p.wire_score_bytes)[o] = rw;
            }
          }
        }
      }
    }
    if (threadIdx.x == 0) *p.ticket = 0u;  // ready for the next launch (before the flags: a pipelined successor
                                           // may pass its dependency wait as soon as the exchange completes)
    if (p.n_peers > 0 || p.done_flag != nullptr) {
      __threadfence_system();
      __syncthreads();
      if (int(threadIdx.x) < p.n_peers)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.peer_flag[threadIdx.x]), "r"(p.seq) : "memory");
      if (threadIdx.x == 0 && p.done_flag != nullptr)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.done_flag), "r"(p.done_seq) : "memory");
    }
  } else {
    // ------------------------------ varlen mode ------------------------------
    // An item is one piece of one row range scanned for up to NQ queries that share it (the queries of one tenant in this
    // batch, grouped by the host planner: the rows are read once for the group).  The warps of the CTA interleave the
    // item's chunks exactly as in uniform mode, so the CTA streams ONE contiguous window of HBM (per-warp items made every
    // warp stream its own distant region).  Items are claimed DYNAMICALLY: the host sorts them by size, largest first, CTA c
    // starts with item c and then takes the next unclaimed one (atomic counter, claimed one item ahead so the latency hides
    // under the scan) -- longest-processing-time-first, so the launch ends on the smallest pieces instead of on whichever CTA
    // a static round-robin happened to overload.  Each item's lists are merged in shared memory and written to its own slot
    // (`ScanItem::out`) of `partial`; merge_items_kernel reduces them per query.
    uint64_t* wk = reinterpret_cast<uint64_t*>(smem);  // [NW][NQ][KSLOTS], aliases the (idle) ring between items
    __shared__ int s_claim[2];
    int it = blockIdx.x;
    for (int round = 0; it < p.n_items; ++round) {
      if (threadIdx.x == 0) s_claim[round & 1] = int(gridDim.x + atomicAdd(p.ticket, 1u));
      const ScanItem item = p.items[it];
      nq_live = min(item.nq, NQ);
#pragma unroll
      for (int qi = 0; qi < NQ; ++qi) {
        // load + normalise this item's queries (dead slots read query 0 of the item and are never offered)
        const float* qsrc = p.queries + size_t(item.query[qi < nq_live ? qi : 0]) * D;
        float ss = 0.f;
#pragma unroll
        for (int t = 0; t < C::NV; ++t)
#pragma unroll
          for (int j = 0; j < C::RPV * C::EPR; ++j) {
            const float x = qsrc[(t * 32 + lane) * (C::RPV * C::EPR) + j];
            q[qi][t * C::RPV * C::EPR + j] = x;
            ss += x * x;
          }
        ss = warp_allreduce_sum(ss);
        const float nrm = sqrtf(ss);
        if (nrm > 0.f) {
#pragma unroll
          for (int e = 0; e < C::EPL; ++e) q[qi][e] = q[qi][e] / nrm;
        }
        list[qi].clear();
        thr[qi] = 0ull;
      }
      scan_range(item.row_begin, item.row_end, warp, C::NW);
      __syncthreads();   // every warp has consumed the copies it issued: the ring is free to hold the lists
#pragma unroll
      for (int qi = 0; qi < NQ; ++qi) {
#pragma unroll
        for (int j = 0; j < KPL; ++j) wk[(warp * NQ + qi) * C::KSLOTS + j * 32 + lane] = list[qi].key[j];
      }
      __syncthreads();
      if (warp < nq_live) {
        const int qi = warp;
        WarpTopK<KPL> m;
        m.clear();
        uint64_t t = 0ull;
        t = m.template merge_batched<4>(
            [&](int i) -> uint64_t { return wk[((i % C::NW) * NQ + qi) * C::KSLOTS + i / C::NW]; }, C::NW * k, t, k, lane);
        m.store(p.partial + (size_t(item.out) * p.item_nq + qi) * k, k, lane);
      }
      __syncthreads();   // the lists have been read: the next item's copies may overwrite them
      it = s_claim[round & 1];   // written before this item's barriers; rewritten two rounds (>= one barrier) from now
    }
    // every CTA has made its last claim before it arrives here: the last one to arrive re-arms both counters
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(p.ticket + 1, 1u) == gridDim.x - 1) {
        p.ticket[0] = 0u;
        p.ticket[1] = 0u;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Exact rescoring of tensor-core candidates (query precision policy MMR_QP_RESCORE).
// K2 scores rows against the 16-bit rounding of the query; a serving store must not let a request's answer depend on
// whether it was batched onto the tensor cores.  So K2 only NOMINATES: it returns kc > k candidates per query, this
// kernel re-scores them with the fp32 query using EXACTLY K1's arithmetic (same lane <-> element mapping, same fmaf
// order, same butterfly; the functions below are the ones K1's inner loop is written from), keeps the best k, and
// proves the answer equal to K1's:  every row K2 did not nominate has a 16-bit score <= c_min (the kc-th candidate's),
// hence an fp32 score <= c_min + eps with eps >= |s32 - s16| for any row (eps = 1.01 ||q - q16|| + 2e-5, Cauchy-
// Schwarz on unit rows).  If the k-th rescored score is above that, no outsider can belong to the top-k.  Otherwise
// (rare: it needs the candidates' scores packed inside ~1e-3) the query is flagged and the host reruns it on K1.
// ------------------------------------------------------------------------------------------------
template <typename E, int D>
__device__ __forceinline__ void k1_load_unit_query(const float* __restrict__ qsrc, int lane, float (&q)[D / 32]) {
  using C = StreamCfg<E, D, 1, 1>;
  float ss = 0.f;
#pragma unroll
  for (int t = 0; t < C::NV; ++t)
#pragma unroll
    for (int j = 0; j < C::RPV * C::EPR; ++j) {
      const float x = qsrc[(t * 32 + lane) * (C::RPV * C::EPR) + j];
      q[t * C::RPV * C::EPR + j] = x;
      ss += x * x;
    }
  ss = warp_allreduce_sum(ss);
  const float nrm = sqrtf(ss);
  if (nrm > 0.f) {
#pragma unroll
    for (int e = 0; e < C::EPL; ++e) q[e] = q[e] / nrm;
  }
}

template <typename E, int D>
__device__ __forceinline__ float k1_row_score(const uint8_t* __restrict__ row, const float (&q)[D / 32], int lane) {
  using C = StreamCfg<E, D, 1, 1>;
  float acc = 0.f;
#pragma unroll
  for (int t = 0; t < C::NV; ++t) {
    uint32_t w[C::RPV];
    const uint8_t* src = row + lane * C::VECB + t * (32 * C::VECB);
    if constexpr (C::VECB == 16) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
      w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else {
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(src));
      w[0] = v.x; w[1] = v.y;
    }
#pragma unroll
    for (int j = 0; j < C::RPV; ++j) {
      if constexpr (C::EB == 2) {
        float a, b;
        ElemTraits<E>::unpack(w[j], a, b);
        acc = fmaf(a, q[(t * C::RPV + j) * 2 + 0], acc);
        acc = fmaf(b, q[(t * C::RPV + j) * 2 + 1], acc);
      } else {
        acc = fmaf(__uint_as_float(w[j]), q[t * C::RPV + j], acc);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);   // K1's butterfly: offsets 16, 8, 4, 2, 1
  return acc;
}

// one CTA (8 warps) per query: candidates [B, kc] (16-bit-query scores + GLOBAL row ids as written by K2's merge) ->
// exact top-k [B, k] + flag[b] (1 = proven exact, 0 = rerun on K1)
template <typename E, int D, int KPL>
__global__ void __launch_bounds__(256) rescore_kernel(const void* __restrict__ rows, const float* __restrict__ queries,
                                                      const float* __restrict__ cand_scores, const int64_t* __restrict__ cand_rows,
                                                      const float* __restrict__ qerr, int kc, int k, int64_t row_base,
                                                      float* __restrict__ out_scores, int64_t* __restrict__ out_rows,
                                                      uint8_t* __restrict__ exact_flag) {
  __shared__ uint64_t keys[MMR_MAX_K_DEVICE];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float q[D / 32];
  k1_load_unit_query<E, D>(queries + size_t(b) * D, lane, q);
  const uint8_t* rows8 = reinterpret_cast<const uint8_t*>(rows);
  for (int c = warp; c < kc; c += 8) {
    const int64_t r = cand_rows[size_t(b) * kc + c];
    uint64_t key = 0ull;
    if (r >= 0) {
      const uint32_t local = uint32_t(r - row_base);
      const float s = k1_row_score<E, D>(rows8 + size_t(local) * (D * sizeof(E)), q, lane);
      key = make_key(s, local);
    }
    if (lane == 0) keys[c] = key;
  }
  __syncthreads();
  if (warp == 0) {
    WarpTopK<KPL> m;
    m.clear();
    uint64_t thr = 0ull;
    thr = m.template merge_batched<2>([&](int i) -> uint64_t { return keys[i]; }, kc, thr, k, lane);
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      const int pos = j * 32 + lane;
      if (pos < k) {
        const uint64_t key = m.key[j];
        out_scores[size_t(b) * k + pos] = key ? key_score(key) : -INFINITY;
        out_rows[size_t(b) * k + pos] = key ? int64_t(key_row(key)) + row_base : int64_t(-1);
      }
    }
    if (lane == 0) {
      // candidates are sorted best-first by K2's merge: the last VALID one carries c_min; fewer than kc valid candidates
      // means K2 nominated every row of the range
      const bool all_rows = cand_rows[size_t(b) * kc + kc - 1] < 0;
      const float c_min = cand_scores[size_t(b) * kc + kc - 1];
      const float eps = 1.01f * qerr[b] + 2e-5f;
      const float kth32 = thr ? key_score(thr) : -INFINITY;      // k-th best rescored score (thr = k-th key, 0 if < k hits)
      exact_flag[b] = (all_rows || (thr != 0ull && kth32 > c_min + eps)) ? 1 : 0;
    }
  }
}

// Varlen tail: one warp per query merges the partial lists of that query (list `slot` of the items
// [item0, item0 + n_items) of its group; partial is [n_items_total][nq_per_item][k]).
template <int KPL>
__global__ void merge_items_kernel(const uint64_t* __restrict__ partial, const QuerySlot* __restrict__ slots, int nq,
                                   int nq_per_item, int k, float* __restrict__ out_scores,
                                   int64_t* __restrict__ out_rows, int64_t row_base) {
  const int lane = threadIdx.x & 31;
  const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (qi >= nq) return;
  WarpTopK<KPL> m;
  m.clear();
  uint64_t t = 0ull;
  const QuerySlot sl = slots[qi];
  const uint64_t* base = partial + (size_t(sl.item0) * nq_per_item + sl.slot) * k;
  const size_t item_stride = size_t(nq_per_item) * k;
  // position-major over the items' lists (all heads first), 4 loads in flight per lane
  t = m.template merge_batched<4>(
      [&](int i) -> uint64_t {
        const int item = i % sl.n_items, pos = i / sl.n_items;
        return __ldcg(reinterpret_cast<const unsigned long long*>(base) + size_t(item) * item_stride + pos);
      },
      sl.n_items * k, t, k, lane);
#pragma unroll
  for (int j = 0; j < KPL; ++j) {
    const int pos = j * 32 + lane;
    if (pos < k) {
      const uint64_t key = m.key[j];
      out_scores[size_t(qi) * k + pos] = key ? key_score(key) : -INFINITY;
      out_rows[size_t(qi) * k + pos] = key ? int64_t(key_row(key)) + row_base : int64_t(-1);
    }
  }
}

}  // namespace mmr
