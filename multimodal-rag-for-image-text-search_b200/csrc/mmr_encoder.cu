// libmmr_b200.so -- device-side query encoders behind the C ABI (include/mmr_b200.h, "Query encoders").
// Host side: weight store, activation arena, the launch sequence of one forward pass.
#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "abi_common.h"
#include "encoder_kernels.cuh"

using namespace mmr;
#define fail mmr_fail

namespace {

struct LayerW {
  __nv_bfloat16 *qkv_w = nullptr, *o_w = nullptr, *fc1_w = nullptr, *fc2_w = nullptr;   // [N_out, K] bf16
  float *qkv_b = nullptr, *o_b = nullptr, *fc1_b = nullptr, *fc2_b = nullptr;
  float *ln1_w = nullptr, *ln1_b = nullptr, *ln2_w = nullptr, *ln2_b = nullptr;
  CUtensorMap m_qkv, m_o, m_fc1, m_fc2;
};

bool make_map(CUtensorMap* map, const void* base, int64_t rows, int cols, int box_rows) {
  mmr_encode_tiled_fn fn = umma_encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  cuuint64_t gstr[1] = {cuuint64_t(cols) * 2};
  cuuint32_t box[2] = {64, cuuint32_t(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

struct mmr_encoder {
  int device = 0;
  mmr_encoder_config cfg{};
  int out_dim = 0;
  std::mutex mu;
  std::vector<LayerW> layers;
  float *word = nullptr, *pos = nullptr, *type_emb = nullptr, *emb_ln_w = nullptr, *emb_ln_b = nullptr;
  float *final_ln_w = nullptr, *final_ln_b = nullptr, *proj_w = nullptr;                     // CLIP
  float *pool_w = nullptr, *pool_b = nullptr, *cls_w = nullptr, *cls_b = nullptr;            // cross-encoder
  std::vector<void*> owned;
  std::map<std::string, bool> seen;
  // activation arena for up to cap_tokens tokens
  int cap_tokens = 0, cap_seqs = 0;
  float *x = nullptr, *tmp = nullptr, *qkv = nullptr;
  __nv_bfloat16 *x16 = nullptr, *ctx16 = nullptr, *h16 = nullptr;
  int32_t *d_ids = nullptr, *d_mask = nullptr, *d_types = nullptr, *h_stage = nullptr;
  CUtensorMap m_x16, m_ctx16, m_h16;
  int maps_tokens = -1;
  int maps_narrow = -1;
  int nt_x = 64, nt_ctx = 64, nt_h = 64;   // token tiles (= box rows) of the activation maps
};

static int dev_alloc(mmr_encoder* e, void** p, size_t bytes) {
  CUDA_TRY(cudaMalloc(p, std::max<size_t>(bytes, 16)));
  e->owned.push_back(*p);
  return MMR_OK;
}

extern "C" int mmr_encoder_create(int device, const mmr_encoder_config* cfg, mmr_encoder** out) {
  if (!out || !cfg) return fail(MMR_ERR_INVALID, "NULL argument");
  *out = nullptr;
  if (cfg->kind != MMR_ENC_MINILM && cfg->kind != MMR_ENC_CLIP_TEXT && cfg->kind != MMR_ENC_CROSS)
    return fail(MMR_ERR_INVALID, "unknown encoder kind %d", cfg->kind);
  const int H = cfg->hidden, I = cfg->intermediate;
  if ((H != 384 && H != 512) || H % cfg->heads != 0 || (H / cfg->heads != 32 && H / cfg->heads != 64))
    return fail(MMR_ERR_UNSUPPORTED, "hidden %d / heads %d: kernels are built for hidden 384 or 512 with head dim 32 or 64", H, cfg->heads);
  if (I % 128 != 0 || I % 64 != 0 || (3 * H) % 128 != 0) return fail(MMR_ERR_UNSUPPORTED, "intermediate size must be a multiple of 128");
  if (cfg->layers < 1 || cfg->vocab_size < 1 || cfg->max_positions < 1) return fail(MMR_ERR_INVALID, "bad config");
  int major = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) return fail(MMR_ERR_CUDA, "device %d is not sm_100: the encoders have no other backend", device);
  CUDA_TRY(cudaSetDevice(device));
  mmr_encoder* e = new mmr_encoder();
  e->device = device;
  e->cfg = *cfg;
  e->out_dim = cfg->kind == MMR_ENC_CROSS ? 1 : (cfg->kind == MMR_ENC_CLIP_TEXT ? cfg->proj_dim : H);
  e->layers.resize(cfg->layers);
  int rc = MMR_OK;
  auto A = [&](void** p, size_t bytes) { if (rc == MMR_OK) rc = dev_alloc(e, p, bytes); };
  A((void**)&e->word, size_t(cfg->vocab_size) * H * 4);
  A((void**)&e->pos, size_t(cfg->max_positions) * H * 4);
  if (cfg->kind != MMR_ENC_CLIP_TEXT) {
    A((void**)&e->type_emb, size_t(std::max(cfg->type_vocab, 1)) * H * 4);
    A((void**)&e->emb_ln_w, H * 4);
    A((void**)&e->emb_ln_b, H * 4);
  } else {
    A((void**)&e->final_ln_w, H * 4);
    A((void**)&e->final_ln_b, H * 4);
    A((void**)&e->proj_w, size_t(cfg->proj_dim) * H * 4);
  }
  if (cfg->kind == MMR_ENC_CROSS) {
    A((void**)&e->pool_w, size_t(H) * H * 4);
    A((void**)&e->pool_b, H * 4);
    A((void**)&e->cls_w, H * 4);
    A((void**)&e->cls_b, 4);
  }
  for (auto& L : e->layers) {
    A((void**)&L.qkv_w, size_t(3 * H) * H * 2);
    A((void**)&L.o_w, size_t(H) * H * 2);
    A((void**)&L.fc1_w, size_t(I) * H * 2);
    A((void**)&L.fc2_w, size_t(H) * I * 2);
    A((void**)&L.qkv_b, 3 * H * 4);
    A((void**)&L.o_b, H * 4);
    A((void**)&L.fc1_b, I * 4);
    A((void**)&L.fc2_b, H * 4);
    A((void**)&L.ln1_w, H * 4);
    A((void**)&L.ln1_b, H * 4);
    A((void**)&L.ln2_w, H * 4);
    A((void**)&L.ln2_b, H * 4);
    if (rc == MMR_OK && (!make_map(&L.m_qkv, L.qkv_w, 3 * H, H, ENC_BM) || !make_map(&L.m_o, L.o_w, H, H, ENC_BM) ||
                         !make_map(&L.m_fc1, L.fc1_w, I, H, ENC_BM) || !make_map(&L.m_fc2, L.fc2_w, H, I, ENC_BM)))
      rc = fail(MMR_ERR_CUDA, "cuTensorMapEncodeTiled failed for an encoder weight");
  }
  if (rc != MMR_OK) {
    mmr_encoder_destroy(e);
    return rc;
  }
  *out = e;
  return MMR_OK;
}

extern "C" int mmr_encoder_destroy(mmr_encoder* e) {
  if (!e) return MMR_OK;
  cudaSetDevice(e->device);
  for (void* p : e->owned) cudaFree(p);
  if (e->h_stage) cudaFreeHost(e->h_stage);
  delete e;
  return MMR_OK;
}

extern "C" int mmr_encoder_out_dim(const mmr_encoder* e) { return e ? e->out_dim : 0; }

// name -> (destination, element count, is bf16 GEMM weight)
static bool weight_slot(mmr_encoder* e, const std::string& name, void** dst, int64_t* numel, bool* bf16) {
  const int H = e->cfg.hidden, I = e->cfg.intermediate;
  *bf16 = false;
  auto set = [&](void* p, int64_t n, bool b = false) { *dst = p; *numel = n; *bf16 = b; return p != nullptr; };
  if (name == "word_emb") return set(e->word, int64_t(e->cfg.vocab_size) * H);
  if (name == "pos_emb") return set(e->pos, int64_t(e->cfg.max_positions) * H);
  if (name == "type_emb") return set(e->type_emb, int64_t(std::max(e->cfg.type_vocab, 1)) * H);
  if (name == "emb_ln_w") return set(e->emb_ln_w, H);
  if (name == "emb_ln_b") return set(e->emb_ln_b, H);
  if (name == "final_ln_w") return set(e->final_ln_w, H);
  if (name == "final_ln_b") return set(e->final_ln_b, H);
  if (name == "proj_w") return set(e->proj_w, int64_t(e->cfg.proj_dim) * H);
  if (name == "pooler_w") return set(e->pool_w, int64_t(H) * H);
  if (name == "pooler_b") return set(e->pool_b, H);
  if (name == "cls_w") return set(e->cls_w, H);
  if (name == "cls_b") return set(e->cls_b, 1);
  int l = -1;
  char field[32] = {0};
  if (sscanf(name.c_str(), "L%d.%31s", &l, field) == 2 && l >= 0 && l < int(e->layers.size())) {
    LayerW& L = e->layers[l];
    const std::string f(field);
    if (f == "qkv_w") return set(L.qkv_w, int64_t(3 * H) * H, true);
    if (f == "o_w") return set(L.o_w, int64_t(H) * H, true);
    if (f == "fc1_w") return set(L.fc1_w, int64_t(I) * H, true);
    if (f == "fc2_w") return set(L.fc2_w, int64_t(H) * I, true);
    if (f == "qkv_b") return set(L.qkv_b, 3 * H);
    if (f == "o_b") return set(L.o_b, H);
    if (f == "fc1_b") return set(L.fc1_b, I);
    if (f == "fc2_b") return set(L.fc2_b, H);
    if (f == "ln1_w") return set(L.ln1_w, H);
    if (f == "ln1_b") return set(L.ln1_b, H);
    if (f == "ln2_w") return set(L.ln2_w, H);
    if (f == "ln2_b") return set(L.ln2_b, H);
  }
  return false;
}

extern "C" int mmr_encoder_set_weight(mmr_encoder* e, const char* name, const float* data_dev, int64_t numel, void* stream) {
  if (!e || !name || !data_dev) return fail(MMR_ERR_INVALID, "NULL argument");
  void* dst = nullptr;
  int64_t want = 0;
  bool bf16 = false;
  if (!weight_slot(e, name, &dst, &want, &bf16)) return fail(MMR_ERR_INVALID, "encoder has no weight named %s", name);
  if (numel != want) return fail(MMR_ERR_INVALID, "weight %s: %lld elements, expected %lld", name, (long long)numel, (long long)want);
  CUDA_TRY(cudaSetDevice(e->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (bf16) {
    f32_to_bf16_kernel<<<std::min<int64_t>((numel + 255) / 256, 4096), 256, 0, st>>>(data_dev, static_cast<__nv_bfloat16*>(dst), numel);
    mmr_g_launches++;
    CUDA_TRY(cudaGetLastError());
  } else {
    CUDA_TRY(cudaMemcpyAsync(dst, data_dev, size_t(numel) * 4, cudaMemcpyDeviceToDevice, st));
  }
  e->seen[name] = true;
  return MMR_OK;
}

static int ensure_arena(mmr_encoder* e, int tokens, int seqs) {
  if (tokens <= e->cap_tokens && seqs <= e->cap_seqs) return MMR_OK;
  const int H = e->cfg.hidden, I = e->cfg.intermediate;
  const int cap = std::max(tokens, std::max(e->cap_tokens, 256));
  const int cs = std::max(seqs, std::max(e->cap_seqs, 8));
  // (old arena stays in `owned` until destroy: arenas only grow a few times)
  CUDA_TRY(cudaMalloc((void**)&e->x, size_t(cap) * H * 4));        e->owned.push_back(e->x);
  CUDA_TRY(cudaMalloc((void**)&e->tmp, size_t(cap) * H * 4));      e->owned.push_back(e->tmp);
  CUDA_TRY(cudaMalloc((void**)&e->qkv, size_t(cap) * 3 * H * 4));  e->owned.push_back(e->qkv);
  CUDA_TRY(cudaMalloc((void**)&e->x16, size_t(cap) * H * 2));      e->owned.push_back(e->x16);
  CUDA_TRY(cudaMalloc((void**)&e->ctx16, size_t(cap) * H * 2));    e->owned.push_back(e->ctx16);
  CUDA_TRY(cudaMalloc((void**)&e->h16, size_t(cap) * I * 2));      e->owned.push_back(e->h16);
  CUDA_TRY(cudaMalloc((void**)&e->d_ids, size_t(cap) * 3 * 4));    e->owned.push_back(e->d_ids);
  e->d_mask = e->d_ids + cap;
  e->d_types = e->d_ids + 2 * size_t(cap);
  if (e->h_stage) cudaFreeHost(e->h_stage);
  CUDA_TRY(cudaMallocHost((void**)&e->h_stage, size_t(cap) * 3 * 4));
  e->cap_tokens = cap;
  e->cap_seqs = cs;
  e->maps_tokens = -1;
  return MMR_OK;
}

static int enc_token_tile(int M) { return M <= 1024 ? 64 : (M <= 4096 ? 128 : 256); }
// Token tile of one GEMM: the M-based tile, halved while the grid (N / 128 feature tiles x token tiles) would leave SMs
// without their two CTAs -- the narrow GEMMs (out-proj and FFN2: N = hidden = 3-4 feature tiles) ran 96 CTAs on 148 SMs at
// 4096 tokens (profiles/r02_encoder_summary.md).  The tile never changes a result (accumulation runs along K only).
static int enc_token_tile_for(int M, int N) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  int nt = enc_token_tile(M);
  if (options().enc_narrow_tiles == 0) return nt;
  // 128 -> 64 until every SM has its two CTAs; 256 -> 128 only when SMs would otherwise idle (at 16384 tokens the 256-token
  // tile with 192 CTAs beats 384 CTAs of 128: profiles/r02_encoder_summary.md)
  while (nt > 64 && int64_t(N / ENC_BM) * ((M + nt - 1) / nt) < (nt == 256 ? 1 : 2) * int64_t(sms)) nt >>= 1;
  return nt;
}

template <int EPI, int NT>
static int launch_gemm_nt(const CUtensorMap& mw, const CUtensorMap& mx, int M, int N, int K, const float* bias,
                          const float* residual, float* out_f32, __nv_bfloat16* out_bf16, cudaStream_t st) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  GemmParams p{};
  p.M = M;
  p.N = N;
  p.K = K;
  const int stage_bytes = ENC_W_SLICE + NT * 128;
  p.nstages = std::max(2, std::min(std::min(ENC_MAX_STAGES, K / 64), (options().enc_gemm_smem_kb * 1024) / stage_bytes));
  p.bias = bias;
  p.residual = residual;
  p.out_f32 = out_f32;
  p.out_bf16 = out_bf16;
  const size_t smem = size_t(p.nstages) * stage_bytes + (2 * ENC_MAX_STAGES + 3) * 8 + 64;
  if (!attr_done[dev & 63]) {
    CUDA_TRY(cudaFuncSetAttribute(gemm_wt_kernel<EPI, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    attr_done[dev & 63] = true;
  }
  CUDA_TRY(launch_pdl(gemm_wt_kernel<EPI, NT>, dim3(N / ENC_BM, (M + NT - 1) / NT), dim3(ENC_THREADS), smem, st, mw, mx, p));
  mmr_g_launches++;
  return MMR_OK;
}

template <int EPI>
static int launch_gemm(const CUtensorMap& mw, const CUtensorMap& mx, int nt, int M, int N, int K, const float* bias,
                       const float* residual, float* out_f32, __nv_bfloat16* out_bf16, cudaStream_t st) {
  switch (nt) {   // = the box rows of the activation map mx
    case 64: return launch_gemm_nt<EPI, 64>(mw, mx, M, N, K, bias, residual, out_f32, out_bf16, st);
    case 128: return launch_gemm_nt<EPI, 128>(mw, mx, M, N, K, bias, residual, out_f32, out_bf16, st);
    default: return launch_gemm_nt<EPI, 256>(mw, mx, M, N, K, bias, residual, out_f32, out_bf16, st);
  }
}

// The same GEMM with the LayerNorm in front of it folded in (gemm_wt_kernel<EPI, 64, true>): X = LN(ln_src), normalised rows
// also written to ln_out (fp32) when the residual stream needs them.  Token tile 64 only (the latency-bound regime).
template <int EPI>
static int launch_gemm_ln(const CUtensorMap& mw, const CUtensorMap& any_x_map, int M, int N, int K, const float* bias,
                          float* out_f32, __nv_bfloat16* out_bf16, const float* ln_src, const float* ln_g, const float* ln_b,
                          float ln_eps, float* ln_out, cudaStream_t st) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  GemmParams p{};
  p.M = M;
  p.N = N;
  p.K = K;
  p.nstages = std::min(ENC_MAX_STAGES, K / 64);
  p.bias = bias;
  p.out_f32 = out_f32;
  p.out_bf16 = out_bf16;
  p.ln_src = ln_src;
  p.ln_g = ln_g;
  p.ln_b = ln_b;
  p.ln_out = ln_out;
  p.ln_eps = ln_eps;
  const size_t smem = size_t(p.nstages) * ENC_W_SLICE + size_t(K / 64) * (64 * 128) + (2 * ENC_MAX_STAGES + 3) * 8 + 64;
  if (!attr_done[dev & 63]) {
    CUDA_TRY(cudaFuncSetAttribute(gemm_wt_kernel<EPI, 64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    attr_done[dev & 63] = true;
  }
  CUDA_TRY(launch_pdl(gemm_wt_kernel<EPI, 64, true>, dim3(N / ENC_BM, (M + 63) / 64), dim3(ENC_THREADS), smem, st, mw, any_x_map, p));
  mmr_g_launches++;
  return MMR_OK;
}

template <int H>
static int forward_t(mmr_encoder* e, int B, int S, float* out_dev, cudaStream_t st) {
  const mmr_encoder_config& c = e->cfg;
  const int M = B * S, I = c.intermediate, DH = H / c.heads;
  const bool clip = c.kind == MMR_ENC_CLIP_TEXT;
  const int wpb = 8;
  const dim3 rows_grid((M + wpb - 1) / wpb), rows_block(wpb * 32);
  if (e->maps_tokens != M || e->maps_narrow != options().enc_narrow_tiles) {   // the activation maps depend on the live token count (TMA zero-fills rows past it)
    // box rows of an activation map = the token tile of the GEMMs that read it: x16 feeds QKV (N = 3H) and FFN1 (N = I),
    // ctx16 the out-projection and h16 FFN2 (both N = H)
    e->nt_x = enc_token_tile_for(M, std::min(3 * H, I));
    e->nt_ctx = enc_token_tile_for(M, H);
    e->nt_h = enc_token_tile_for(M, H);
    if (!make_map(&e->m_x16, e->x16, M, H, e->nt_x) || !make_map(&e->m_ctx16, e->ctx16, M, H, e->nt_ctx) ||
        !make_map(&e->m_h16, e->h16, M, I, e->nt_h))
      return fail(MMR_ERR_CUDA, "cuTensorMapEncodeTiled failed for the activations");
    e->maps_tokens = M;
    e->maps_narrow = options().enc_narrow_tiles;
  }
  static bool att_attr[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!att_attr[dev & 63]) {
    CUDA_TRY(cudaFuncSetAttribute(attention_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(attention_mma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(attention_mma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    att_attr[dev & 63] = true;
  }
  // the cross-encoder (rerank passages: 128-512 tokens) runs both attention contractions on the tensor cores, whatever the
  // sequence length (kernel choice by model, never by batch shape); MMR_ENC_ATT_MMA=0: fp32 kernel everywhere, =2: every model
  // from ATC_MIN_S tokens on (measurement)
  const int att_opt = options().enc_att_mma;
  const bool att_mma = att_opt == 2 ? S >= ATC_MIN_S : (att_opt == 1 && c.kind == MMR_ENC_CROSS);
  const size_t att_mma_smem = DH == 32 ? attention_mma_smem_bytes<32>(S) : attention_mma_smem_bytes<64>(S);
  const dim3 att_mma_grid(c.heads, B, (S + ATC_ROWS - 1) / ATC_ROWS);
  const size_t att_smem = DH == 32 ? attention_smem_bytes<32>(S) : attention_smem_bytes<64>(S);
  if (att_smem > 220 * 1024) return fail(MMR_ERR_UNSUPPORTED, "sequence length %d does not fit the attention kernel", S);

  // embeddings
  if (clip)
    CUDA_TRY(launch_pdl(embed_kernel<H>, rows_grid, rows_block, 0, st, e->d_ids, (const int32_t*)nullptr, e->word, e->pos,
                        (const float*)nullptr, (const float*)nullptr, (const float*)nullptr, c.ln_eps, M, S, e->x,
                        (__nv_bfloat16*)nullptr));
  else
    CUDA_TRY(launch_pdl(embed_kernel<H>, rows_grid, rows_block, 0, st, e->d_ids, e->d_types, e->word, e->pos, e->type_emb,
                        e->emb_ln_w, e->emb_ln_b, c.ln_eps, M, S, e->x, e->x16));
  mmr_g_launches++;
  // LayerNorms in front of a GEMM can be folded into it (MMR_ENC_FUSE_LN=1: passes of <= 16 tokens, =2: whenever the token
  // tile is 64; bit-identical activations either way).  OFF by default -- measured (profiles/r02_encoder_summary.md): 44 -> 33 /
  // 86 -> 62 launches buy nothing (MiniLM 1 x 16: 0.229 -> 0.259 ms, from 128 tokens on clearly slower): a dependent launch
  // costs no more than the LayerNorm it carried, and every CTA of the GEMM repeats its token tile's LayerNorm.
  const int fuse_opt = options().enc_fuse_ln;
  const bool fuse_ln = fuse_opt == 2 ? enc_token_tile(M) == 64 : (fuse_opt == 1 && M <= 16);
  const LayerW* prev = nullptr;   // BERT: the layer whose closing LayerNorm (ln2 over e->tmp) is still pending
  for (LayerW& L : e->layers) {
    int rc;
    if (clip) {  // pre-LN: h = LN1(x)
      if (fuse_ln) {
        rc = launch_gemm_ln<EPI_BIAS_F32>(L.m_qkv, e->m_x16, M, 3 * H, H, L.qkv_b, e->qkv, nullptr, e->x, L.ln1_w, L.ln1_b, c.ln_eps,
                                          nullptr, st);
      } else {
        CUDA_TRY(launch_pdl(layernorm_kernel<H>, rows_grid, rows_block, 0, st, e->x, L.ln1_w, L.ln1_b, c.ln_eps, M,
                            (float*)nullptr, e->x16));
        mmr_g_launches++;
        rc = launch_gemm<EPI_BIAS_F32>(L.m_qkv, e->m_x16, e->nt_x, M, 3 * H, H, L.qkv_b, nullptr, e->qkv, nullptr, st);
      }
    } else if (prev != nullptr) {   // BERT, fused: x = LN2_prev(tmp) computed by this layer's QKV GEMM
      rc = launch_gemm_ln<EPI_BIAS_F32>(L.m_qkv, e->m_x16, M, 3 * H, H, L.qkv_b, e->qkv, nullptr, e->tmp, prev->ln2_w, prev->ln2_b,
                                        c.ln_eps, e->x, st);
    } else {
      rc = launch_gemm<EPI_BIAS_F32>(L.m_qkv, e->m_x16, e->nt_x, M, 3 * H, H, L.qkv_b, nullptr, e->qkv, nullptr, st);
    }
    if (rc != MMR_OK) return rc;
    if (att_mma && DH == 32)
      CUDA_TRY(launch_pdl(attention_mma_kernel<32>, att_mma_grid, dim3(ATC_NW * 32), att_mma_smem, st, e->qkv, e->d_mask, e->ctx16, S,
                          H, clip ? 1 : 0));
    else if (att_mma)
      CUDA_TRY(launch_pdl(attention_mma_kernel<64>, att_mma_grid, dim3(ATC_NW * 32), att_mma_smem, st, e->qkv, e->d_mask, e->ctx16, S,
                          H, clip ? 1 : 0));
    else if (DH == 32)
      CUDA_TRY(launch_pdl(attention_kernel<32>, dim3(c.heads, B), dim3(ATT_NW * 32), att_smem, st, e->qkv, e->d_mask, e->ctx16, S, H,
                          clip ? 1 : 0));
    else
      CUDA_TRY(launch_pdl(attention_kernel<64>, dim3(c.heads, B), dim3(ATT_NW * 32), att_smem, st, e->qkv, e->d_mask, e->ctx16, S, H,
                          clip ? 1 : 0));
    mmr_g_launches++;
    if (clip) {
      // x = x + out_proj(ctx); h = LN2(x); x = x + fc2(quick_gelu(fc1(h)))
      rc = launch_gemm<EPI_BIAS_RES_F32>(L.m_o, e->m_ctx16, e->nt_ctx, M, H, H, L.o_b, e->x, e->x, nullptr, st);
      if (rc != MMR_OK) return rc;
      if (fuse_ln) {
        rc = launch_gemm_ln<EPI_QUICKGELU_BF16>(L.m_fc1, e->m_x16, M, I, H, L.fc1_b, nullptr, e->h16, e->x, L.ln2_w, L.ln2_b, c.ln_eps,
                                                nullptr, st);
      } else {
        CUDA_TRY(launch_pdl(layernorm_kernel<H>, rows_grid, rows_block, 0, st, e->x, L.ln2_w, L.ln2_b, c.ln_eps, M,
                            (float*)nullptr, e->x16));
        mmr_g_launches++;
        rc = launch_gemm<EPI_QUICKGELU_BF16>(L.m_fc1, e->m_x16, e->nt_x, M, I, H, L.fc1_b, nullptr, nullptr, e->h16, st);
      }
      if (rc != MMR_OK) return rc;
      rc = launch_gemm<EPI_BIAS_RES_F32>(L.m_fc2, e->m_h16, e->nt_h, M, H, I, L.fc2_b, e->x, e->x, nullptr, st);
      if (rc != MMR_OK) return rc;
    } else {
      // post-LN (BERT): x = LN1(x + out_proj(ctx)); x = LN2(x + fc2(gelu(fc1(x))))
      rc = launch_gemm<EPI_BIAS_RES_F32>(L.m_o, e->m_ctx16, e->nt_ctx, M, H, H, L.o_b, e->x, e->tmp, nullptr, st);
      if (rc != MMR_OK) return rc;
      if (fuse_ln) {
        rc = launch_gemm_ln<EPI_GELU_BF16>(L.m_fc1, e->m_x16, M, I, H, L.fc1_b, nullptr, e->h16, e->tmp, L.ln1_w, L.ln1_b, c.ln_eps,
                                           e->x, st);
      } else {
        CUDA_TRY(launch_pdl(layernorm_kernel<H>, rows_grid, rows_block, 0, st, e->tmp, L.ln1_w, L.ln1_b, c.ln_eps, M, e->x, e->x16));
        mmr_g_launches++;
        rc = launch_gemm<EPI_GELU_BF16>(L.m_fc1, e->m_x16, e->nt_x, M, I, H, L.fc1_b, nullptr, nullptr, e->h16, st);
      }
      if (rc != MMR_OK) return rc;
      rc = launch_gemm<EPI_BIAS_RES_F32>(L.m_fc2, e->m_h16, e->nt_h, M, H, I, L.fc2_b, e->x, e->tmp, nullptr, st);
      if (rc != MMR_OK) return rc;
      if (fuse_ln && &L != &e->layers.back()) {
        prev = &L;   // LN2 rides in the next layer's QKV GEMM
      } else {
        CUDA_TRY(launch_pdl(layernorm_kernel<H>, rows_grid, rows_block, 0, st, e->tmp, L.ln2_w, L.ln2_b, c.ln_eps, M, e->x, e->x16));
        mmr_g_launches++;
      }
    }
  }
  // head
  if (c.kind == MMR_ENC_MINILM)
    CUDA_TRY(launch_pdl(mean_pool_norm_kernel, dim3(B), dim3(H), 0, st, e->x, e->d_mask, S, H, out_dev));
  else if (clip)
    CUDA_TRY(launch_pdl(clip_head_kernel, dim3(B), dim3(H), size_t(H + c.proj_dim) * 4, st, e->x, e->d_ids, S, H, c.eos_token_id,
                        e->final_ln_w, e->final_ln_b, c.ln_eps, e->proj_w, c.proj_dim, out_dev));
  else
    CUDA_TRY(launch_pdl(cross_head_kernel, dim3(B), dim3(H), size_t(2 * H) * 4, st, e->x, S, H, e->pool_w, e->pool_b, e->cls_w,
                        e->cls_b, out_dev));
  mmr_g_launches++;
  return MMR_OK;
}

extern "C" int mmr_encoder_forward(mmr_encoder* e, const int32_t* input_ids_host, const int32_t* attention_mask_host,
                                   const int32_t* token_type_host, int32_t B, int32_t S, float* out_dev, void* stream) {
  if (!e || !input_ids_host || !out_dev) return fail(MMR_ERR_INVALID, "NULL argument");
  if (B < 1 || S < 1) return fail(MMR_ERR_INVALID, "B and S must be >= 1");
  if (S > e->cfg.max_positions) return fail(MMR_ERR_INVALID, "sequence length %d > max_positions %d", S, e->cfg.max_positions);
  if (S > 512) return fail(MMR_ERR_UNSUPPORTED, "sequence length %d > 512", S);
  if (e->cfg.kind == MMR_ENC_CLIP_TEXT && e->cfg.proj_dim > e->cfg.hidden) return fail(MMR_ERR_UNSUPPORTED, "projection wider than hidden");
  std::lock_guard<std::mutex> guard(e->mu);
  CUDA_TRY(cudaSetDevice(e->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int M = B * S;
  for (int i = 0; i < M; ++i)
    if (input_ids_host[i] < 0 || input_ids_host[i] >= e->cfg.vocab_size)
      return fail(MMR_ERR_INVALID, "token id %d out of range [0, %d)", input_ids_host[i], e->cfg.vocab_size);
  int rc = ensure_arena(e, M, B);
  if (rc != MMR_OK) return rc;
  // ids | mask | types ride in one pinned staging buffer, one H2D copy
  const size_t cap = size_t(e->cap_tokens);
  CUDA_TRY(cudaStreamSynchronize(st));  // the previous forward may still read the staging buffer's device copy
  memcpy(e->h_stage, input_ids_host, size_t(M) * 4);
  for (int i = 0; i < M; ++i) e->h_stage[cap + i] = attention_mask_host ? attention_mask_host[i] : 1;
  for (int i = 0; i < M; ++i) {
    const int t = token_type_host ? token_type_host[i] : 0;
    if (t < 0 || t >= std::max(e->cfg.type_vocab, 1)) return fail(MMR_ERR_INVALID, "token type %d out of range", t);
    e->h_stage[2 * cap + i] = t;
  }
  CUDA_TRY(cudaMemcpyAsync(e->d_ids, e->h_stage, cap * 3 * 4, cudaMemcpyHostToDevice, st));
  if (e->cfg.hidden == 384) return forward_t<384>(e, B, S, out_dev, st);
  return forward_t<512>(e, B, S, out_dev, st);
}
