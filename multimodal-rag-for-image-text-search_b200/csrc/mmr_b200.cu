// libmmr_b200.so -- C ABI (include/mmr_b200.h) over the sm_100a scan kernels.
// Host side: argument checks, planning (which kernel, grid, workspace carve-up), launches.
#include "../../include/mmr_b200.h"

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "aux_kernels.cuh"
#include "common.cuh"
#include "scan_stream.cuh"
#ifdef MMR_WITH_UMMA
#include "scan_umma2.cuh"
#endif

using namespace mmr;

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static thread_local int g_last_kernel = 0;
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CUDA_TRY(expr)                                                                                \
  do {                                                                                                \
    cudaError_t e_ = (expr);                                                                          \
    if (e_ != cudaSuccess) return fail(MMR_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                                       __FILE__, __LINE__);                                           \
  } while (0)

// ------------------------------------------------------------------------------------------------ index
struct mmr_index {
  int device = 0;
  int dim = 0;
  int dtype = MMR_BF16;
  int64_t n_rows = 0;
  const void* rows = nullptr;
  int64_t row_base = 0;
  std::vector<int64_t> seg;  // [n_segments + 1]
  int sm_count = 0;
  // small staging for mmr_search_host
  float* h_q = nullptr;
  float* d_q = nullptr;
  float* d_scores = nullptr;
  int64_t* d_rows = nullptr;
  float* h_scores = nullptr;
  int64_t* h_rows = nullptr;
  void* d_ws = nullptr;
  size_t ws_bytes = 0;
  int cap_b = 0, cap_k = 0;
#ifdef MMR_WITH_UMMA
  mutable UmmaIndexState umma;
  mutable Umma2IndexState umma2;
#endif
};

static int elem_bytes(int dtype) { return dtype == MMR_F32 ? 4 : 2; }

static int set_segments(mmr_index* ix, const int64_t* seg, int32_t nseg) {
  ix->seg.clear();
  if (seg == nullptr || nseg <= 0) {
    ix->seg = {0, ix->n_rows};
    return MMR_OK;
  }
  ix->seg.assign(seg, seg + nseg + 1);
  if (ix->seg.front() < 0 || ix->seg.back() > ix->n_rows) return fail(MMR_ERR_INVALID, "segment offsets out of range");
  for (int i = 0; i < nseg; ++i)
    if (ix->seg[i] > ix->seg[i + 1]) return fail(MMR_ERR_INVALID, "segment offsets must be ascending");
  return MMR_OK;
}

extern "C" int mmr_abi_version(void) { return MMR_ABI_VERSION; }
extern "C" const char* mmr_last_error(void) { return g_err.c_str(); }
extern "C" int64_t mmr_launch_count(void) { return g_launches.load(); }
extern "C" int mmr_last_kernel(void) { return g_last_kernel; }

extern "C" int mmr_device_sm_count(int device, int* out_sms) {
  if (!out_sms) return fail(MMR_ERR_INVALID, "out_sms is NULL");
  CUDA_TRY(cudaDeviceGetAttribute(out_sms, cudaDevAttrMultiProcessorCount, device));
  return MMR_OK;
}

extern "C" int mmr_index_create(int device, int dim, int dtype, int64_t n_rows, const void* rows_dev,
                                const int64_t* seg_offsets_host, int32_t n_segments, int64_t row_base,
                                mmr_index** out) {
  if (!out) return fail(MMR_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (dtype != MMR_BF16 && dtype != MMR_F32 && dtype != MMR_F16) return fail(MMR_ERR_INVALID, "unknown dtype %d", dtype);
  if (dim <= 0 || n_rows < 0) return fail(MMR_ERR_INVALID, "bad shape [%lld, %d]", (long long)n_rows, dim);
  if (n_rows >= (int64_t(1) << 32)) return fail(MMR_ERR_UNSUPPORTED, "at most 2^32-1 rows per resident index");
  if (n_rows > 0 && rows_dev == nullptr) return fail(MMR_ERR_INVALID, "rows_dev is NULL");
  if ((reinterpret_cast<uintptr_t>(rows_dev) & 15) != 0) return fail(MMR_ERR_INVALID, "rows_dev must be 16-byte aligned");
  if (dim != 384 && dim != 512)
    return fail(MMR_ERR_UNSUPPORTED, "dim %d: kernels are built for 384 (MiniLM) and 512 (CLIP)", dim);
  int major = 0, minor = 0, sms = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  CUDA_TRY(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  if (major != 10) return fail(MMR_ERR_CUDA, "device %d is sm_%d%d; this library is sm_100a only", device, major, minor);
  mmr_index* ix = new mmr_index();
  ix->device = device;
  ix->dim = dim;
  ix->dtype = dtype;
  ix->n_rows = n_rows;
  ix->rows = rows_dev;
  ix->row_base = row_base;
  ix->sm_count = sms;
  int rc = set_segments(ix, seg_offsets_host, n_segments);
  if (rc != MMR_OK) {
    delete ix;
    return rc;
  }
  *out = ix;
  return MMR_OK;
}

extern "C" int mmr_index_update(mmr_index* ix, int64_t n_rows, const void* rows_dev, const int64_t* seg, int32_t nseg) {
  if (!ix) return fail(MMR_ERR_INVALID, "index is NULL");
  if (n_rows < 0 || n_rows >= (int64_t(1) << 32)) return fail(MMR_ERR_INVALID, "bad n_rows");
  if (n_rows > 0 && rows_dev == nullptr) return fail(MMR_ERR_INVALID, "rows_dev is NULL");
  if ((reinterpret_cast<uintptr_t>(rows_dev) & 15) != 0) return fail(MMR_ERR_INVALID, "rows_dev must be 16-byte aligned");
  ix->n_rows = n_rows;
  ix->rows = rows_dev;
#ifdef MMR_WITH_UMMA
  ix->umma.valid = false;
  ix->umma2.valid = false;
#endif
  return set_segments(ix, seg, nseg);
}

static void free_staging(mmr_index* ix) {
  if (ix->h_q) cudaFreeHost(ix->h_q);
  if (ix->h_scores) cudaFreeHost(ix->h_scores);
  if (ix->h_rows) cudaFreeHost(ix->h_rows);
  if (ix->d_q) cudaFree(ix->d_q);
  if (ix->d_scores) cudaFree(ix->d_scores);
  if (ix->d_rows) cudaFree(ix->d_rows);
  if (ix->d_ws) cudaFree(ix->d_ws);
  ix->h_q = ix->d_q = ix->d_scores = ix->h_scores = nullptr;
  ix->h_rows = ix->d_rows = nullptr;
  ix->d_ws = nullptr;
  ix->cap_b = ix->cap_k = 0;
}

extern "C" int mmr_index_destroy(mmr_index* ix) {
  if (!ix) return MMR_OK;
  cudaSetDevice(ix->device);
  free_staging(ix);
  delete ix;
  return MMR_OK;
}

// ------------------------------------------------------------------------------------------------ loader
template <typename E>
static int launch_convert(const float* src, void* dst, int64_t n, int dim, int normalize, cudaStream_t st,
                          const int64_t* dst_row = nullptr) {
  if (n == 0) return MMR_OK;
  const int threads = 256;
  const int64_t blocks = std::min<int64_t>((n + 7) / 8, 148 * 16);
  convert_rows_kernel<E><<<(unsigned)blocks, threads, 0, st>>>(src, reinterpret_cast<E*>(dst), n, dim, normalize, dst_row);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return MMR_OK;
}

static int convert_dispatch(const float* src_dev, void* dst_dev, int dtype, int64_t n_rows, int dim, int normalize,
                            cudaStream_t st, const int64_t* dst_row_dev) {
  switch (dtype) {
    case MMR_BF16: return launch_convert<__nv_bfloat16>(src_dev, dst_dev, n_rows, dim, normalize, st, dst_row_dev);
    case MMR_F16: return launch_convert<__half>(src_dev, dst_dev, n_rows, dim, normalize, st, dst_row_dev);
    case MMR_F32: return launch_convert<float>(src_dev, dst_dev, n_rows, dim, normalize, st, dst_row_dev);
  }
  return fail(MMR_ERR_INVALID, "unknown dtype %d", dtype);
}

extern "C" int mmr_convert_rows_f32(const float* src_dev, void* dst_dev, int dtype, int64_t n_rows, int dim,
                                    int normalize, void* stream) {
  if (n_rows < 0 || dim <= 0) return fail(MMR_ERR_INVALID, "bad shape");
  if (n_rows > 0 && (!src_dev || !dst_dev)) return fail(MMR_ERR_INVALID, "NULL buffer");
  return convert_dispatch(src_dev, dst_dev, dtype, n_rows, dim, normalize, static_cast<cudaStream_t>(stream), nullptr);
}

// Pageable host memory -> pinned staging with a few threads (one memcpy thread tops out near 10 GB/s, well under the
// PCIe gen5 link), double-buffered against the H2D copy + convert kernel of the previous chunk.
static void parallel_memcpy(void* dst, const void* src, size_t bytes, int threads) {
  if (threads <= 1 || bytes < (size_t(4) << 20)) {
    memcpy(dst, src, bytes);
    return;
  }
  std::vector<std::thread> pool;
  const size_t per = (bytes / threads + 4095) / 4096 * 4096;
  for (int t = 0; t < threads; ++t) {
    const size_t o = size_t(t) * per;
    if (o >= bytes) break;
    const size_t len = std::min(per, bytes - o);
    pool.emplace_back([=]() { memcpy(static_cast<uint8_t*>(dst) + o, static_cast<const uint8_t*>(src) + o, len); });
  }
  for (auto& th : pool) th.join();
}

extern "C" int mmr_load_rows_f32_host_scatter(int device, const float* src_host, void* dst_dev, int dtype, int64_t n_rows,
                                              int dim, int normalize, const int64_t* dst_row_host, void* stream) {
  if (n_rows < 0 || dim <= 0) return fail(MMR_ERR_INVALID, "bad shape");
  if (n_rows == 0) return MMR_OK;
  if (!src_host || !dst_dev) return fail(MMR_ERR_INVALID, "NULL buffer");
  if (dtype != MMR_BF16 && dtype != MMR_F32 && dtype != MMR_F16) return fail(MMR_ERR_INVALID, "unknown dtype %d", dtype);
  CUDA_TRY(cudaSetDevice(device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t chunk_rows = std::max<int64_t>(1, (int64_t(64) << 20) / (int64_t(dim) * 4));  // 64 MiB chunks
  const size_t chunk_bytes = size_t(chunk_rows) * dim * 4;
  const size_t map_bytes = dst_row_host ? size_t(chunk_rows) * 8 : 0;
  uint8_t* h[2] = {nullptr, nullptr};
  uint8_t* d[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  int rc = MMR_OK;
  auto cleanup = [&]() {
    for (int i = 0; i < 2; ++i) {
      if (h[i]) cudaFreeHost(h[i]);
      if (d[i]) cudaFree(d[i]);
      if (done[i]) cudaEventDestroy(done[i]);
    }
  };
  for (int i = 0; i < 2; ++i) {
    if (cudaMallocHost(&h[i], chunk_bytes + map_bytes) != cudaSuccess ||
        cudaMalloc(&d[i], chunk_bytes + map_bytes) != cudaSuccess ||
        cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) != cudaSuccess) {
      cleanup();
      return fail(MMR_ERR_CUDA, "staging allocation of %zu bytes failed", chunk_bytes + map_bytes);
    }
  }
  const int threads = int(std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency() / 2)));
  int slot = 0;
  for (int64_t r0 = 0; r0 < n_rows && rc == MMR_OK; r0 += chunk_rows, slot ^= 1) {
    const int64_t nr = std::min(chunk_rows, n_rows - r0);
    const size_t bytes = size_t(nr) * dim * 4;
    cudaEventSynchronize(done[slot]);  // staging slot free again
    parallel_memcpy(h[slot], src_host + r0 * dim, bytes, threads);
    if (dst_row_host) memcpy(h[slot] + chunk_bytes, dst_row_host + r0, size_t(nr) * 8);
    // one copy when the row map rides along (it sits right behind the chunk in both staging buffers)
    const size_t copy_bytes = dst_row_host ? chunk_bytes + size_t(nr) * 8 : bytes;
    if (cudaMemcpyAsync(d[slot], h[slot], copy_bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) {
      rc = fail(MMR_ERR_CUDA, "H2D copy failed");
      break;
    }
    const int64_t* map_dev = dst_row_host ? reinterpret_cast<const int64_t*>(d[slot] + chunk_bytes) : nullptr;
    const int eb = elem_bytes(dtype);
    void* dst = dst_row_host ? dst_dev : static_cast<void*>(static_cast<uint8_t*>(dst_dev) + size_t(r0) * dim * eb);
    rc = convert_dispatch(reinterpret_cast<const float*>(d[slot]), dst, dtype, nr, dim, normalize, st, map_dev);
    cudaEventRecord(done[slot], st);
  }
  cudaStreamSynchronize(st);
  cleanup();
  return rc;
}

extern "C" int mmr_load_rows_f32_host(int device, const float* src_host, void* dst_dev, int dtype, int64_t n_rows,
                                      int dim, int normalize, void* stream) {
  return mmr_load_rows_f32_host_scatter(device, src_host, dst_dev, dtype, n_rows, dim, normalize, nullptr, stream);
}

// Host-side helper of the columnar store: 64-bit hashes of n strings held Arrow-style (one byte buffer + n+1 int32
// offsets).  The store keeps (hash, row) sorted to find the rows an upsert replaces (delete-by-chunk_id,
// lancedb_store.py:91-92) without a Python dict over millions of ids.
extern "C" int mmr_hash_strings(const uint8_t* data, const int32_t* offsets, int64_t n, uint64_t* out) {
  if (n < 0 || (n > 0 && (!offsets || !out))) return fail(MMR_ERR_INVALID, "bad arguments");
  for (int64_t i = 0; i < n; ++i) {
    const uint8_t* p = data + offsets[i];
    const int32_t len = offsets[i + 1] - offsets[i];
    uint64_t h = 0xcbf29ce484222325ull ^ (uint64_t(uint32_t(len)) * 0x9E3779B97F4A7C15ull);
    int32_t j = 0;
    for (; j + 8 <= len; j += 8) {
      uint64_t w;
      memcpy(&w, p + j, 8);
      h = (h ^ w) * 0x100000001b3ull;
      h ^= h >> 29;
    }
    uint64_t tail = 0;
    if (j < len) memcpy(&tail, p + j, size_t(len - j));
    h = (h ^ tail) * 0x100000001b3ull;
    h ^= h >> 32;  // final avalanche (murmur3 fmix64)
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33;
    h *= 0xc4ceb9fe1a85ec53ull;
    h ^= h >> 33;
    out[i] = h;
  }
  return MMR_OK;
}

// ------------------------------------------------------------------------------------------------ K1 launch
typedef void (*stream_kernel_t)(const StreamParams);

template <typename E, int D, int NQ, int KPL>
static int launch_stream_t(const StreamParams& p, int grid, cudaStream_t st) {
  using C = StreamCfg<E, D, NQ, KPL>;
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    CUDA_TRY(cudaFuncSetAttribute(scan_stream_kernel<E, D, NQ, KPL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C::SMEM_BYTES));
    attr_set[dev] = true;
  }
  // MMR_PDL=1 (opt-in): launch with programmatic stream serialization so consecutive searches on one stream overlap
  // tail and head.  Contract: the queries of a search must not be produced by the kernel launched immediately before
  // it on the same stream (the scan starts before that kernel's memory is guaranteed visible).
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(K1_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  const char* pdl = getenv("MMR_PDL");
  attr[0].val.programmaticStreamSerializationAllowed = (p.items == nullptr && pdl && pdl[0] == '1') ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, scan_stream_kernel<E, D, NQ, KPL>, p));
  g_launches++;
  return MMR_OK;
}

template <typename E, int D>
static int launch_stream_ed(const StreamParams& p, int nq_pad, int kpl, int grid, cudaStream_t st) {
#define MMR_CASE(NQ_, KPL_) \
  if (nq_pad == NQ_ && kpl == KPL_) return launch_stream_t<E, D, NQ_, KPL_>(p, grid, st);
  MMR_CASE(1, 1) MMR_CASE(1, 2) MMR_CASE(2, 1) MMR_CASE(2, 2) MMR_CASE(4, 1) MMR_CASE(4, 2)
  if constexpr (sizeof(E) == 4) {  // fp32 rows stream half as many rows per byte: 8 queries per pass stay HBM-bound
    MMR_CASE(8, 1) MMR_CASE(8, 2)
  }
#undef MMR_CASE
  return fail(MMR_ERR_UNSUPPORTED, "no stream kernel for nq=%d kpl=%d", nq_pad, kpl);
}

static int launch_stream(const mmr_index* ix, const StreamParams& p, int nq_pad, int kpl, int grid, cudaStream_t st) {
  const int key = ix->dtype * 1000 + ix->dim;
  switch (key) {
    case MMR_BF16 * 1000 + 512: return launch_stream_ed<__nv_bfloat16, 512>(p, nq_pad, kpl, grid, st);
    case MMR_BF16 * 1000 + 384: return launch_stream_ed<__nv_bfloat16, 384>(p, nq_pad, kpl, grid, st);
    case MMR_F16 * 1000 + 512: return launch_stream_ed<__half, 512>(p, nq_pad, kpl, grid, st);
    case MMR_F16 * 1000 + 384: return launch_stream_ed<__half, 384>(p, nq_pad, kpl, grid, st);
    case MMR_F32 * 1000 + 512: return launch_stream_ed<float, 512>(p, nq_pad, kpl, grid, st);
    case MMR_F32 * 1000 + 384: return launch_stream_ed<float, 384>(p, nq_pad, kpl, grid, st);
  }
  return fail(MMR_ERR_UNSUPPORTED, "no stream kernel for dtype %d dim %d", ix->dtype, ix->dim);
}

static int rows_per_stage(int) { return MMR_K1_ROWS; }

// ------------------------------------------------------------------------------------------------ planning
// Workspace layout (bytes):
//   [0, 256)                          control block (ticket counter at +0)
//   [256, 256 + PART)                 partial top-k keys  (uniform: grid*4*k u64; varlen: n_items*k u64)
//   then (varlen only)                ScanItem[n_items], int32 item_off[B+1]
static constexpr size_t WS_CTRL = 256;
static constexpr int VARLEN_ITEMS_PER_WARP = 8;

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int max_varlen_items(const mmr_index* ix, int B) { return ix->sm_count * K1_NW * VARLEN_ITEMS_PER_WARP + 2 * B + 64; }

extern "C" size_t mmr_search_workspace_bytes(const mmr_index* ix, int32_t B, int32_t k) {
  if (!ix || B <= 0 || k <= 0) return 0;
  const size_t kk = size_t(std::min<int32_t>(k, MMR_MAX_K));
  const size_t uniform = size_t(ix->sm_count) * 8 * kk * 8;
  const size_t items = size_t(max_varlen_items(ix, B));
  const size_t varlen = align_up(items * kk * 8, 256) + align_up(items * sizeof(ScanItem), 256) +
                        align_up(size_t(B + 1) * 4, 256);
  size_t total = WS_CTRL + align_up(std::max(uniform, varlen), 256);
#ifdef MMR_WITH_UMMA
  total += umma_workspace_bytes(ix->sm_count, ix->dim, B, int(kk));
#endif
  return total;
}

struct ExchangeInfo {  // fused push by K1's last CTA (only when one launch covers all B queries)
  int n_peers = 0;
  uint32_t seq = 0;
  uint32_t wire_score_bytes = 0;
  uint64_t slot[MMR_MAX_PEERS] = {};
  uint64_t flag[MMR_MAX_PEERS] = {};
};

static int search_uniform_stream(const mmr_index* ix, const float* q, int B, int k, uint32_t r0, uint32_t r1,
                                 float* out_s, int64_t* out_r, uint8_t* ws, cudaStream_t st,
                                 const ExchangeInfo* xi = nullptr) {
  const int kpl = k <= 32 ? 1 : 2;
  const int R = rows_per_stage(ix->dtype);
  const int64_t nchunks = (int64_t(r1) - r0 + R - 1) / R;
  int grid = int(std::min<int64_t>(ix->sm_count, std::max<int64_t>(1, (nchunks + K1_NW - 1) / K1_NW)));
  const int group = ix->dtype == MMR_F32 ? 8 : 4;  // queries per pass
  for (int q0 = 0; q0 < B; q0 += group) {
    const int nq = std::min(group, B - q0);
    const int nq_pad = nq <= 2 ? nq : (nq <= 4 ? 4 : 8);
    StreamParams p{};
    p.rows = ix->rows;
    p.queries = q;
    p.q_first = q0;
    p.nq = nq;
    p.k = k;
    p.row_begin = r0;
    p.row_end = r1;
    p.partial = reinterpret_cast<uint64_t*>(ws + WS_CTRL);
    p.ticket = reinterpret_cast<unsigned int*>(ws);
    p.out_scores = out_s;
    p.out_rows = out_r;
    p.row_base = ix->row_base;
    p.items = nullptr;
    p.n_items = 0;
    if (xi && B <= group) {
      p.n_peers = xi->n_peers;
      p.seq = xi->seq;
      p.wire_score_bytes = xi->wire_score_bytes;
      for (int g = 0; g < xi->n_peers; ++g) {
        p.peer_slot[g] = xi->slot[g];
        p.peer_flag[g] = xi->flag[g];
      }
    }
    int rc = launch_stream(ix, p, nq_pad, kpl, grid, st);
    if (rc != MMR_OK) return rc;
  }
  g_last_kernel = 1;
  return MMR_OK;
}

// `ranges[b]` = the row ranges query b scans (a tenant is one base segment plus any appended delta segments).
static int search_varlen_stream(const mmr_index* ix, const float* q,
                                const std::vector<std::vector<std::pair<uint32_t, uint32_t>>>& ranges, int k, float* out_s,
                                int64_t* out_r, uint8_t* ws, size_t ws_bytes, cudaStream_t st) {
  const int B = int(ranges.size());
  const int kpl = k <= 32 ? 1 : 2;
  const int R = rows_per_stage(ix->dtype);
  int64_t total_rows = 0;
  for (auto& rq : ranges)
    for (auto& r : rq) total_rows += int64_t(r.second) - r.first;
  const int twarps = ix->sm_count * K1_NW;
  const int64_t target = int64_t(twarps) * VARLEN_ITEMS_PER_WARP;
  int64_t item_rows = std::max<int64_t>(64, (total_rows + target - 1) / target);
  item_rows = (item_rows + R - 1) / R * R;
  std::vector<ScanItem> items;
  std::vector<int32_t> off(B + 1, 0);
  for (int b = 0; b < B; ++b) {
    off[b] = int32_t(items.size());
    for (auto& r : ranges[b]) {
      for (int64_t s = r.first; s < int64_t(r.second); s += item_rows) {
        ScanItem it;
        it.row_begin = uint32_t(s);
        it.row_end = uint32_t(std::min<int64_t>(s + item_rows, r.second));
        it.query = b;
        it.pad = 0;
        items.push_back(it);
      }
    }
  }
  off[B] = int32_t(items.size());
  const int n_items = int(items.size());
  if (n_items > max_varlen_items(ix, B)) return fail(MMR_ERR_WORKSPACE, "varlen plan produced %d items (too many row ranges)", n_items);
  const size_t part_bytes = align_up(size_t(std::max(n_items, 1)) * k * 8, 256);
  const size_t item_bytes = align_up(size_t(std::max(n_items, 1)) * sizeof(ScanItem), 256);
  const size_t off_bytes = align_up(size_t(B + 1) * 4, 256);
  if (WS_CTRL + part_bytes + item_bytes + off_bytes > ws_bytes) return fail(MMR_ERR_WORKSPACE, "workspace too small");
  uint64_t* d_part = reinterpret_cast<uint64_t*>(ws + WS_CTRL);
  ScanItem* d_items = reinterpret_cast<ScanItem*>(ws + WS_CTRL + part_bytes);
  int32_t* d_off = reinterpret_cast<int32_t*>(ws + WS_CTRL + part_bytes + item_bytes);
  if (n_items > 0) CUDA_TRY(cudaMemcpyAsync(d_items, items.data(), size_t(n_items) * sizeof(ScanItem), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_off, off.data(), size_t(B + 1) * 4, cudaMemcpyHostToDevice, st));
  // pageable sources: the copies above are staged before returning, the vectors may die after this call
  if (n_items > 0) {
    StreamParams p{};
    p.rows = ix->rows;
    p.queries = q;
    p.q_first = 0;
    p.nq = 1;
    p.k = k;
    p.partial = d_part;
    p.ticket = reinterpret_cast<unsigned int*>(ws);
    p.row_base = ix->row_base;
    p.items = d_items;
    p.n_items = n_items;
    const int grid = int(std::min<int64_t>(ix->sm_count, (n_items + K1_NW - 1) / K1_NW));
    int rc = launch_stream(ix, p, 1, kpl, grid, st);
    if (rc != MMR_OK) return rc;
  }
  const int wpb = 4;
  if (kpl == 1)
    merge_items_kernel<1><<<(B + wpb - 1) / wpb, wpb * 32, 0, st>>>(d_part, d_off, B, k, out_s, out_r, ix->row_base);
  else
    merge_items_kernel<2><<<(B + wpb - 1) / wpb, wpb * 32, 0, st>>>(d_part, d_off, B, k, out_s, out_r, ix->row_base);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  g_last_kernel = 3;
  return MMR_OK;
}

extern "C" int mmr_search(const mmr_index* ix, const float* queries_dev, const int32_t* query_seg_host, int32_t B,
                          int32_t k, float* out_scores_dev, int64_t* out_rows_dev, void* workspace_dev,
                          size_t workspace_bytes, void* stream) {
  if (!ix) return fail(MMR_ERR_INVALID, "index is NULL");
  if (B <= 0) return fail(MMR_ERR_INVALID, "B must be >= 1");
  if (k < 1 || k > MMR_MAX_K) return fail(MMR_ERR_INVALID, "k must be in [1, %d]", MMR_MAX_K);
  if (!queries_dev || !out_scores_dev || !out_rows_dev || !workspace_dev) return fail(MMR_ERR_INVALID, "NULL buffer");
  if (workspace_bytes < mmr_search_workspace_bytes(ix, B, k)) return fail(MMR_ERR_WORKSPACE, "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace_dev);
  const int nseg = int(ix->seg.size()) - 1;

  // row range of every query
  std::vector<std::pair<uint32_t, uint32_t>> ranges(B);
  bool uniform = true;
  for (int b = 0; b < B; ++b) {
    const int s = query_seg_host ? query_seg_host[b] : -1;
    if (s < -1 || s >= nseg) return fail(MMR_ERR_INVALID, "query %d: segment %d out of range [0, %d)", b, s, nseg);
    ranges[b] = s < 0 ? std::make_pair(uint32_t(0), uint32_t(ix->n_rows))
                      : std::make_pair(uint32_t(ix->seg[s]), uint32_t(ix->seg[s + 1]));
    if (ranges[b] != ranges[0]) uniform = false;
  }
  if (uniform) {
#ifdef MMR_WITH_UMMA
    if (umma_preferred(ix->dtype, ix->dim, B, k, int64_t(ranges[0].second) - ranges[0].first)) {
      int rc = umma_search(ix->umma, ix->umma2, ix->rows, ix->n_rows, ix->dim, ix->dtype, ix->sm_count, queries_dev, B, k, ranges[0].first,
                           ranges[0].second, ix->row_base, out_scores_dev, out_rows_dev,
                           ws + mmr_search_workspace_bytes(ix, B, k) - umma_workspace_bytes(ix->sm_count, ix->dim, B, k),
                           st, g_err);
      if (rc == MMR_OK) {
        g_launches += umma_launches_per_search();
        g_last_kernel = 2;
      }
      return rc;
    }
#endif
    return search_uniform_stream(ix, queries_dev, B, k, ranges[0].first, ranges[0].second, out_scores_dev,
                                 out_rows_dev, ws, st);
  }
  std::vector<std::vector<std::pair<uint32_t, uint32_t>>> per_query(B);
  for (int b = 0; b < B; ++b) per_query[b].push_back(ranges[b]);
  return search_varlen_stream(ix, queries_dev, per_query, k, out_scores_dev, out_rows_dev, ws, workspace_bytes, st);
}

// Explicit row ranges per query: query b scans ranges[range_off[b] .. range_off[b+1]) (pairs of [begin, end) row
// ordinals).  Used by stores that append delta segments between compactions.
extern "C" int mmr_search_ranges(const mmr_index* ix, const float* queries_dev, int32_t B, int32_t k,
                                 const int32_t* range_off_host, const int64_t* ranges_host, float* out_scores_dev,
                                 int64_t* out_rows_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
  if (!ix) return fail(MMR_ERR_INVALID, "index is NULL");
  if (B <= 0) return fail(MMR_ERR_INVALID, "B must be >= 1");
  if (k < 1 || k > MMR_MAX_K) return fail(MMR_ERR_INVALID, "k must be in [1, %d]", MMR_MAX_K);
  if (!queries_dev || !range_off_host || !out_scores_dev || !out_rows_dev || !workspace_dev)
    return fail(MMR_ERR_INVALID, "NULL buffer");
  if (workspace_bytes < mmr_search_workspace_bytes(ix, B, k)) return fail(MMR_ERR_WORKSPACE, "workspace too small");
  std::vector<std::vector<std::pair<uint32_t, uint32_t>>> per_query(B);
  bool single_shared = true;
  for (int b = 0; b < B; ++b) {
    if (range_off_host[b + 1] < range_off_host[b]) return fail(MMR_ERR_INVALID, "range offsets must be ascending");
    for (int32_t r = range_off_host[b]; r < range_off_host[b + 1]; ++r) {
      const int64_t lo = ranges_host[2 * r], hi = ranges_host[2 * r + 1];
      if (lo < 0 || hi < lo || hi > ix->n_rows) return fail(MMR_ERR_INVALID, "query %d: bad row range [%lld, %lld)", b, (long long)lo, (long long)hi);
      if (hi > lo) per_query[b].push_back({uint32_t(lo), uint32_t(hi)});
    }
    if (per_query[b].size() != 1 || per_query[b] != per_query[0]) single_shared = false;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace_dev);
  if (single_shared) {  // everyone scans the same single range: the uniform kernels apply
    const uint32_t r0 = per_query[0][0].first, r1 = per_query[0][0].second;
#ifdef MMR_WITH_UMMA
    if (umma_preferred(ix->dtype, ix->dim, B, k, int64_t(r1) - r0)) {
      int rc = umma_search(ix->umma, ix->umma2, ix->rows, ix->n_rows, ix->dim, ix->dtype, ix->sm_count, queries_dev, B, k, r0, r1,
                           ix->row_base, out_scores_dev, out_rows_dev,
                           ws + mmr_search_workspace_bytes(ix, B, k) - umma_workspace_bytes(ix->sm_count, ix->dim, B, k),
                           st, g_err);
      if (rc == MMR_OK) {
        g_launches += umma_launches_per_search();
        g_last_kernel = 2;
      }
      return rc;
    }
#endif
    return search_uniform_stream(ix, queries_dev, B, k, r0, r1, out_scores_dev, out_rows_dev, ws, st);
  }
  return search_varlen_stream(ix, queries_dev, per_query, k, out_scores_dev, out_rows_dev, ws, workspace_bytes, st);
}

// Result staging is ONE buffer on each side ([scores f32 B*k | pad | rows i64 B*k]) so a search costs one D2H copy.
static size_t staging_rows_off(int cb, int ck) { return align_up(size_t(cb) * ck * 4, 16); }

static int ensure_staging(mmr_index* ix, int B, int k) {
  if (B <= ix->cap_b && k <= ix->cap_k) return MMR_OK;
  free_staging(ix);
  const int cb = std::max(B, 8), ck = std::max(k, 16);
  const size_t res_bytes = staging_rows_off(cb, ck) + size_t(cb) * ck * 8;
  CUDA_TRY(cudaMallocHost(&ix->h_q, size_t(cb) * ix->dim * 4));
  CUDA_TRY(cudaMallocHost(&ix->h_scores, res_bytes));
  CUDA_TRY(cudaMalloc(&ix->d_q, size_t(cb) * ix->dim * 4));
  CUDA_TRY(cudaMalloc(&ix->d_scores, res_bytes));
  ix->h_rows = nullptr;  // views into the packed buffers are computed per call (they depend on B and k)
  ix->d_rows = nullptr;
  ix->ws_bytes = mmr_search_workspace_bytes(ix, cb, ck);
  CUDA_TRY(cudaMalloc(&ix->d_ws, ix->ws_bytes));
  CUDA_TRY(cudaMemset(ix->d_ws, 0, ix->ws_bytes));
  ix->cap_b = cb;
  ix->cap_k = ck;
  return MMR_OK;
}

extern "C" int mmr_search_host(mmr_index* ix, const float* queries_host, const int32_t* query_seg_host, int32_t B,
                               int32_t k, float* out_scores_host, int64_t* out_rows_host, void* stream) {
  if (!ix) return fail(MMR_ERR_INVALID, "index is NULL");
  if (B <= 0) return fail(MMR_ERR_INVALID, "B must be >= 1");
  if (k < 1 || k > MMR_MAX_K) return fail(MMR_ERR_INVALID, "k must be in [1, %d]", MMR_MAX_K);
  if (!queries_host || !out_scores_host || !out_rows_host) return fail(MMR_ERR_INVALID, "NULL buffer");
  CUDA_TRY(cudaSetDevice(ix->device));
  int rc = ensure_staging(ix, B, k);
  if (rc != MMR_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t rows_off = staging_rows_off(B, k);
  const size_t res_bytes = rows_off + size_t(B) * k * 8;
  uint8_t* d_res = reinterpret_cast<uint8_t*>(ix->d_scores);
  uint8_t* h_res = reinterpret_cast<uint8_t*>(ix->h_scores);
  memcpy(ix->h_q, queries_host, size_t(B) * ix->dim * 4);
  CUDA_TRY(cudaMemcpyAsync(ix->d_q, ix->h_q, size_t(B) * ix->dim * 4, cudaMemcpyHostToDevice, st));
  rc = mmr_search(ix, ix->d_q, query_seg_host, B, k, reinterpret_cast<float*>(d_res),
                  reinterpret_cast<int64_t*>(d_res + rows_off), ix->d_ws, ix->ws_bytes, st);
  if (rc != MMR_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(h_res, d_res, res_bytes, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  memcpy(out_scores_host, h_res, size_t(B) * k * 4);
  memcpy(out_rows_host, h_res + rows_off, size_t(B) * k * 8);
  return MMR_OK;
}

// ------------------------------------------------------------------------------------------------ fused exchange
static uint32_t xchg_score_bytes(int B, int k) { return uint32_t((size_t(B) * k * 4 + 7) / 8 * 8); }
static uint32_t xchg_wire_bytes(int B, int k) { return uint32_t((xchg_score_bytes(B, k) + size_t(B) * k * 8 + 15) / 16 * 16); }

extern "C" size_t mmr_exchange_buffer_bytes(int32_t G, int32_t B, int32_t k) {
  if (G <= 0 || G > MMR_XCHG_MAX_PEERS || B <= 0 || k <= 0) return 0;
  return size_t(MMR_XCHG_HEADER) + size_t(2) * G * xchg_wire_bytes(B, k);
}

extern "C" size_t mmr_search_exchange_workspace_bytes(const mmr_index* ix, int32_t B, int32_t k) {
  const size_t base = mmr_search_workspace_bytes(ix, B, k);
  return base ? base + align_up(xchg_wire_bytes(B, std::min<int32_t>(k, MMR_MAX_K)), 256) : 0;
}

extern "C" int mmr_search_exchange(const mmr_index* ix, const float* queries_dev, const int32_t* query_seg_host,
                                   int32_t B, int32_t k, const uint64_t* peer_bufs_host, int32_t G, int32_t rank,
                                   uint32_t seq, float* out_scores_dev, int64_t* out_rows_dev, void* workspace_dev,
                                   size_t workspace_bytes, void* stream) {
  if (!ix) return fail(MMR_ERR_INVALID, "index is NULL");
  if (!peer_bufs_host || G < 1 || G > MMR_XCHG_MAX_PEERS || rank < 0 || rank >= G)
    return fail(MMR_ERR_INVALID, "bad peer table (G=%d, rank=%d)", G, rank);
  if (seq == 0) return fail(MMR_ERR_INVALID, "seq must start at 1 and grow by 1 per search");
  if (B <= 0 || k < 1 || k > MMR_MAX_K) return fail(MMR_ERR_INVALID, "bad B or k");
  if (!queries_dev || !out_scores_dev || !out_rows_dev || !workspace_dev) return fail(MMR_ERR_INVALID, "NULL buffer");
  const size_t base_ws = mmr_search_workspace_bytes(ix, B, k);
  if (workspace_bytes < mmr_search_exchange_workspace_bytes(ix, B, k)) return fail(MMR_ERR_WORKSPACE, "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace_dev);
  const uint32_t score_bytes = xchg_score_bytes(B, k), wire_bytes = xchg_wire_bytes(B, k);
  const int parity = int(seq & 1u);
  uint8_t* wire = ws + base_ws;  // this rank's own result, [scores | rows]
  float* w_scores = reinterpret_cast<float*>(wire);
  int64_t* w_rows = reinterpret_cast<int64_t*>(wire + score_bytes);
  ExchangeInfo xi;
  xi.n_peers = G;
  xi.seq = seq;
  xi.wire_score_bytes = score_bytes;
  for (int g = 0; g < G; ++g) {
    xi.slot[g] = peer_bufs_host[g] + MMR_XCHG_HEADER + (size_t(parity) * G + rank) * wire_bytes;
    xi.flag[g] = peer_bufs_host[g] + (size_t(parity) * MMR_XCHG_MAX_PEERS + rank) * 4;
  }
  // 1. the shard-local scan
  bool pushed = false;
  const int nseg = int(ix->seg.size()) - 1;
  bool uniform = true;
  int s0 = query_seg_host ? query_seg_host[0] : -1;
  for (int b = 0; b < B && uniform; ++b) uniform = (query_seg_host ? query_seg_host[b] : -1) == s0;
  if (s0 < -1 || s0 >= nseg) return fail(MMR_ERR_INVALID, "segment %d out of range", s0);
  const uint32_t r0 = s0 < 0 ? 0u : uint32_t(ix->seg[s0]);
  const uint32_t r1 = s0 < 0 ? uint32_t(ix->n_rows) : uint32_t(ix->seg[s0 + 1]);
  bool k2 = false;
#ifdef MMR_WITH_UMMA
  k2 = uniform && umma_preferred(ix->dtype, ix->dim, B, k, int64_t(r1) - r0);
#endif
  if (uniform && !k2 && B <= (ix->dtype == MMR_F32 ? 8 : 4)) {
    // K1 computes and pushes in ONE kernel: its last CTA stores the result into every peer over NVLink
    int rc = search_uniform_stream(ix, queries_dev, B, k, r0, r1, w_scores, w_rows, ws, st, &xi);
    if (rc != MMR_OK) return rc;
    pushed = true;
  } else {
    int rc = mmr_search(ix, queries_dev, query_seg_host, B, k, w_scores, w_rows, ws, base_ws, st);
    if (rc != MMR_OK) return rc;
  }
  // 2. push (when the scan kernel did not do it itself)
  if (!pushed) {
    PeerPtrs pp;
    for (int g = 0; g < MMR_XCHG_MAX_PEERS; ++g) {
      pp.slot[g] = g < G ? xi.slot[g] : 0;
      pp.flag[g] = g < G ? xi.flag[g] : 0;
    }
    push_wire_kernel<<<G, 256, 0, st>>>(wire, wire_bytes, pp, seq);
    g_launches++;
  }
  // 3. wait for every peer's slot, merge in place
  const uint8_t* local = reinterpret_cast<const uint8_t*>(peer_bufs_host[rank]);
  const int wpb = 4;
  const uint64_t timeout_ns = 5000000000ull;
  // With MMR_PDL=1 the wait+merge kernel is a programmatic dependent of the scan: it may become resident while the scan
  // still runs (it only spins on the flags, which the scan's last CTA releases at its very end) and the next search's
  // scan may in turn start behind it.  Otherwise plain stream order.
  const char* pdl = getenv("MMR_PDL");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((B + wpb - 1) / wpb);
  cfg.blockDim = dim3(wpb * 32);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (pushed && pdl && pdl[0] == '1') ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (k <= 32)
    CUDA_TRY(cudaLaunchKernelEx(&cfg, merge_wait_kernel<1>, local, parity, seq, wire_bytes, score_bytes, int(G), int(B),
                                int(k), out_scores_dev, out_rows_dev, timeout_ns));
  else
    CUDA_TRY(cudaLaunchKernelEx(&cfg, merge_wait_kernel<2>, local, parity, seq, wire_bytes, score_bytes, int(G), int(B),
                                int(k), out_scores_dev, out_rows_dev, timeout_ns));
  g_launches++;
  return MMR_OK;
}

#ifdef MMR_WITH_UMMA
// Debug / validation hook: raw K2 scores (tensor-core contraction only, no top-k) for rows [row_begin, row_end).
extern "C" int mmr_debug_umma_scores(const mmr_index* ix, const float* queries_dev, int32_t B, int64_t row_begin,
                                     int64_t row_end, float* out_scores_dev, int64_t out_ld, void* workspace_dev,
                                     size_t workspace_bytes, void* stream) {
  if (!ix || !queries_dev || !out_scores_dev || !workspace_dev) return fail(MMR_ERR_INVALID, "NULL argument");
  if (ix->dtype != MMR_BF16 && ix->dtype != MMR_F16) return fail(MMR_ERR_UNSUPPORTED, "K2 needs bf16 or fp16 rows");
  if (B <= 0 || row_begin < 0 || row_end > ix->n_rows || row_end <= row_begin || out_ld < row_end - row_begin)
    return fail(MMR_ERR_INVALID, "bad range");
  if (workspace_bytes < umma_workspace_bytes(ix->sm_count, ix->dim, B, 10)) return fail(MMR_ERR_WORKSPACE, "workspace too small");
  int rc = umma_search(ix->umma, ix->umma2, ix->rows, ix->n_rows, ix->dim, ix->dtype, ix->sm_count, queries_dev, B, 10, uint32_t(row_begin),
                       uint32_t(row_end), 0, nullptr, nullptr, static_cast<uint8_t*>(workspace_dev),
                       static_cast<cudaStream_t>(stream), g_err, out_scores_dev, out_ld);
  if (rc == MMR_OK) g_launches += 2;
  return rc;
}
#else
extern "C" int mmr_debug_umma_scores(const mmr_index*, const float*, int32_t, int64_t, int64_t, float*, int64_t, void*,
                                     size_t, void*) {
  return fail(MMR_ERR_UNSUPPORTED, "built without the tcgen05 kernel");
}
#endif

// ------------------------------------------------------------------------------------------------ K4 / K5
extern "C" int mmr_merge_topk_strided(const float* scores_dev, const int64_t* rows_dev, int64_t score_shard_stride,
                                      int64_t row_shard_stride, int32_t G, int32_t B, int32_t k, float* out_scores_dev,
                                      int64_t* out_rows_dev, void* stream) {
  if (G <= 0 || B <= 0) return fail(MMR_ERR_INVALID, "G and B must be >= 1");
  if (k < 1 || k > MMR_MAX_K) return fail(MMR_ERR_INVALID, "k must be in [1, %d]", MMR_MAX_K);
  if (!scores_dev || !rows_dev || !out_scores_dev || !out_rows_dev) return fail(MMR_ERR_INVALID, "NULL buffer");
  if (score_shard_stride < int64_t(B) * k || row_shard_stride < int64_t(B) * k)
    return fail(MMR_ERR_INVALID, "shard stride smaller than B*k");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int wpb = 4;
  const int blocks = (B + wpb - 1) / wpb;
  if (k <= 32)
    merge_shards_kernel<1><<<blocks, wpb * 32, 0, st>>>(scores_dev, rows_dev, score_shard_stride, row_shard_stride, G, B,
                                                        k, out_scores_dev, out_rows_dev);
  else
    merge_shards_kernel<2><<<blocks, wpb * 32, 0, st>>>(scores_dev, rows_dev, score_shard_stride, row_shard_stride, G, B,
                                                        k, out_scores_dev, out_rows_dev);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return MMR_OK;
}

extern "C" int mmr_merge_topk(const float* scores_dev, const int64_t* rows_dev, int32_t G, int32_t B, int32_t k,
                              float* out_scores_dev, int64_t* out_rows_dev, void* stream) {
  return mmr_merge_topk_strided(scores_dev, rows_dev, int64_t(B) * k, int64_t(B) * k, G, B, k, out_scores_dev,
                                out_rows_dev, stream);
}

extern "C" int mmr_fuse(const float* text_scores_dev, const int64_t* text_rows_dev, int32_t kt,
                        const float* img_scores_dev, const int64_t* img_rows_dev, int32_t ki, int32_t B,
                        int32_t final_n, double tau, double* out_combined_dev, double* out_score_dev,
                        int64_t* out_rows_dev, int8_t* out_modality_dev, uint8_t* out_low_conf_dev, void* stream) {
  if (B <= 0) return fail(MMR_ERR_INVALID, "B must be >= 1");
  if (kt < 0 || kt > FUSE_MAXK || ki < 0 || ki > FUSE_MAXK) return fail(MMR_ERR_INVALID, "kt, ki must be in [0, %d]", FUSE_MAXK);
  if (final_n < 1) return fail(MMR_ERR_INVALID, "final_n must be >= 1");
  if ((kt > 0 && (!text_scores_dev || !text_rows_dev)) || (ki > 0 && (!img_scores_dev || !img_rows_dev)))
    return fail(MMR_ERR_INVALID, "NULL input buffer");
  if (!out_combined_dev || !out_score_dev || !out_rows_dev || !out_modality_dev || !out_low_conf_dev)
    return fail(MMR_ERR_INVALID, "NULL output buffer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int threads = 64;
  fuse_kernel<<<(B + threads - 1) / threads, threads, 0, st>>>(
      kt > 0 ? text_scores_dev : nullptr, kt > 0 ? text_rows_dev : nullptr, kt, ki > 0 ? img_scores_dev : nullptr,
      ki > 0 ? img_rows_dev : nullptr, ki, B, final_n, tau, out_combined_dev, out_score_dev, out_rows_dev,
      out_modality_dev, out_low_conf_dev);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return MMR_OK;
}

extern "C" int mmr_fuse_f64(const double* text_scores_dev, const int32_t* text_count_dev, const double* text_rerank_dev,
                            const int32_t* rerank_count_dev, const double* img_scores_dev, const int32_t* img_count_dev,
                            int32_t kt, int32_t ki, int32_t B, int32_t final_n, double tau, double* out_combined_dev,
                            int32_t* out_index_dev, uint8_t* out_low_conf_dev, void* stream) {
  if (B <= 0) return fail(MMR_ERR_INVALID, "B must be >= 1");
  if (kt < 0 || kt > FUSE_MAXK || ki < 0 || ki > FUSE_MAXK) return fail(MMR_ERR_INVALID, "kt, ki must be in [0, %d]", FUSE_MAXK);
  if (final_n < 1) return fail(MMR_ERR_INVALID, "final_n must be >= 1");
  if ((kt > 0 && (!text_scores_dev || !text_count_dev)) || (ki > 0 && (!img_scores_dev || !img_count_dev)))
    return fail(MMR_ERR_INVALID, "NULL input buffer");
  if (!out_combined_dev || !out_index_dev || !out_low_conf_dev) return fail(MMR_ERR_INVALID, "NULL output buffer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int threads = 32;
  fuse_full_kernel<<<(B + threads - 1) / threads, threads, 0, st>>>(
      kt > 0 ? text_scores_dev : nullptr, kt > 0 ? text_count_dev : nullptr, text_rerank_dev, rerank_count_dev,
      ki > 0 ? img_scores_dev : nullptr, ki > 0 ? img_count_dev : nullptr, kt, ki, B, final_n, tau, out_combined_dev,
      out_index_dev, out_low_conf_dev);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return MMR_OK;
}
