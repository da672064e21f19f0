// Warp-resident exact top-k: a descending-sorted list of u64 keys spread over the 32 lanes of a warp
// (KPL keys per lane -> capacity 32*KPL), maintained with shuffles only.  Inserts are rare once the
// threshold (the k-th key) has warmed up, so the scan's common path is one compare per score.
#pragma once
#include "common.cuh"

namespace mmr {

template <int KPL>
struct WarpTopK {
  // position p (0 = best) lives in lane p % 32, slot p / 32
  uint64_t key[KPL];

  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int j = 0; j < KPL; ++j) key[j] = 0ull;
  }

  // k-th best key (threshold); k in [1, 32*KPL].  Uniform across the warp.
  __device__ __forceinline__ uint64_t kth(int k) const {
    const int p = k - 1;
    uint64_t v = key[0];
#pragma unroll
    for (int j = 1; j < KPL; ++j)
      if ((p >> 5) == j) v = key[j];
    return shfl_u64(v, p & 31);
  }

  // Insert a warp-uniform candidate `c` (distinct from every key already held).  Keys ranking below
  // it shift down one position; the last one falls off the end.
  __device__ __forceinline__ void insert(uint64_t c, int lane) {
    uint64_t carry = ~0ull;  // "element before position 0": never outranked by c
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      const uint64_t mine = key[j];
      uint64_t up = shfl_up_u64(mine, 1);
      if (lane == 0) up = carry;
      carry = shfl_u64(mine, 31);
      key[j] = (c > mine) ? ((c > up) ? up : c) : mine;
    }
  }

  // Offer up to 32 candidates (one per lane, `valid` lanes only) against threshold `thr`; returns the
  // updated threshold.  All lanes must call.
  __device__ __forceinline__ uint64_t offer(uint64_t cand, bool valid, uint64_t thr, int k, int lane) {
    unsigned m = __ballot_sync(0xffffffffu, valid && cand > thr);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const uint64_t c = shfl_u64(cand, src);
      if (c > thr) {
        insert(c, lane);
        thr = kth(k);
      }
    }
    return thr;
  }

  // ---- batch merge: 32 candidates at once through a bitonic network --------------------------------------------
  // `offer` costs one dependent shuffle chain per winning candidate, which is right for the scan (winners are rare)
  // and wrong for the merges, where most of the first candidates win: merging the per-warp / per-CTA lists that way
  // was ~13 us of serial tail per search (benchmarks/fixed_cost.py).  Here the 32 candidates (one per lane, 0 = none)
  // are sorted with a 15-step bitonic network and merged into the sorted list with 6 (KPL = 1) or 17 (KPL = 2) more
  // compare-exchange steps, whatever the number of winners.  Keys are distinct (or 0), so the result is the same
  // sorted list insertion would give.
  static __device__ __forceinline__ uint64_t cmpx(uint64_t v, int j, bool keep_max) {
    const uint64_t o = shfl_u64_xor(v, j);
    return keep_max ? (v > o ? v : o) : (v < o ? v : o);
  }
  // sort 32 lane-values descending (lane 0 = largest)
  static __device__ __forceinline__ uint64_t sort32_desc(uint64_t v, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
        const bool desc = (lane & k) == 0 || k == 32;   // final pass: the whole warp descending
        const bool lower = (lane & j) == 0;
        v = cmpx(v, j, lower == desc);
      }
    }
    return v;
  }
  // v is bitonic over the 32 lanes -> sorted descending
  static __device__ __forceinline__ uint64_t bitonic_merge32_desc(uint64_t v, int lane) {
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) v = cmpx(v, j, (lane & j) == 0);
    return v;
  }
  // Merge 32 candidates (one per lane; 0 = none) into the list.  All lanes must call.
  __device__ __forceinline__ void merge32(uint64_t cand, int lane) {
    const uint64_t c = sort32_desc(cand, lane);
    const uint64_t crev = shfl_u64(c, 31 - lane);
    if constexpr (KPL == 1) {
      const uint64_t m = key[0] > crev ? key[0] : crev;       // top 32 of the 64, as a bitonic sequence
      key[0] = bitonic_merge32_desc(m, lane);
    } else {
      static_assert(KPL <= 2, "merge32 is written for lists of at most 64 keys");
      // lower half (positions 32..63) vs the candidates: keep the best 32 of those 64, sorted
      uint64_t lo = key[1] > crev ? key[1] : crev;
      lo = bitonic_merge32_desc(lo, lane);
      // now two sorted runs of 32: key[0] and lo -> sorted 64
      const uint64_t lrev = shfl_u64(lo, 31 - lane);
      const uint64_t hi = key[0] > lrev ? key[0] : lrev;      // the 32 largest (bitonic)
      const uint64_t rest = key[0] > lrev ? lrev : key[0];    // the 32 smallest (bitonic)
      key[0] = bitonic_merge32_desc(hi, lane);
      key[1] = bitonic_merge32_desc(rest, lane);
    }
  }

  // Merge `count` keys from memory, element i at base[i * stride].  GLOBAL = true reads through L2
  // only (ld.global.cg): the keys were written by other CTAs of the same launch.
  template <bool GLOBAL = false>
  __device__ __forceinline__ uint64_t merge_from(const uint64_t* base, int count, int stride, uint64_t thr, int k,
                                                 int lane) {
    for (int i0 = 0; i0 < count; i0 += 32) {
      const int i = i0 + lane;
      const bool valid = i < count;
      uint64_t c = 0ull;
      if (valid) {
        if constexpr (GLOBAL) c = __ldcg(reinterpret_cast<const unsigned long long*>(base) + size_t(i) * stride);
        else c = base[size_t(i) * stride];
      }
      thr = offer_batch(c, valid && c != 0ull, thr, k, lane);
    }
    return thr;
  }

  // One batch of up to 32 candidates against the threshold: nothing to do when none beats it (the common case late in a
  // merge), one insertion when exactly one does, the bitonic merge otherwise.
  __device__ __forceinline__ uint64_t offer_batch(uint64_t cand, bool valid, uint64_t thr, int k, int lane) {
    const bool win = valid && cand > thr;
    const unsigned m = __ballot_sync(0xffffffffu, win);
    if (m == 0u) return thr;
    if ((m & (m - 1u)) == 0u) {
      insert(shfl_u64(cand, __ffs(m) - 1), lane);
      return kth(k);
    }
    merge32(win ? cand : 0ull, lane);
    return kth(k);
  }

  // Merge `count` candidates fetched through `key_at(i)` (returns 0 for "no key").  Loads are issued U at a time per
  // lane before any of them is consumed, so a merge over memory costs count / (32 U) round trips, not count / 32.
  template <int U, typename F>
  __device__ __forceinline__ uint64_t merge_batched(F key_at, int count, uint64_t thr, int k, int lane) {
    for (int i0 = 0; i0 < count; i0 += 32 * U) {
      uint64_t c[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * 32 + lane;
        c[u] = i < count ? key_at(i) : 0ull;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (i0 + u * 32 < count) thr = offer_batch(c[u], c[u] != 0ull, thr, k, lane);
      }
    }
    return thr;
  }

  // Write the first k positions to dst[0..k).
  __device__ __forceinline__ void store(uint64_t* dst, int k, int lane) const {
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      const int p = j * 32 + lane;
      if (p < k) dst[p] = key[j];
    }
  }
};

}  // namespace mmr
