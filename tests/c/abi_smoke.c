/*
 * Plain-C client of libmmr_b200.so: no Python, no torch.  Builds a small bf16 index with the CUDA runtime, searches it
 * through mmr_search_host and mmr_search (+ mmr_merge_topk over two row-range shards) and checks the ids against a
 * scalar CPU loop over the same bf16 values.  Compiled and run by tests/test_gpu_c_abi.py on the GPU box.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include tests/c/abi_smoke.c -o abi_smoke \
 *       -L <pkg> -lmmr_b200 -L /usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,<pkg>
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mmr_b200.h"

#define N 40000
#define D 512
#define K 10
#define B 2

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define CHECK_MMR(x) do { int rc_ = (x); if (rc_ != MMR_OK) { fprintf(stderr, "%s -> %d: %s\n", #x, rc_, mmr_last_error()); return 3; } } while (0)

static uint32_t rng_state = 12345u;
static float frand(void) {  /* xorshift, roughly uniform in [-1, 1) */
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 17; rng_state ^= rng_state << 5;
  return (float)(rng_state >> 8) / 8388608.0f - 1.0f;
}
static float bf16_to_f32(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

int main(void) {
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { fprintf(stderr, "no CUDA device\n"); return 1; }
  if (mmr_abi_version() != MMR_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 1; }

  float* rows_f32 = (float*)malloc(sizeof(float) * N * D);
  float* queries = (float*)malloc(sizeof(float) * B * D);
  for (size_t i = 0; i < (size_t)N * D; ++i) rows_f32[i] = frand();
  for (size_t i = 0; i < (size_t)B * D; ++i) queries[i] = frand();
  memcpy(rows_f32 + (size_t)31234 * D, rows_f32 + (size_t)77 * D, sizeof(float) * D); /* exact duplicate -> tie */
  memcpy(queries + D, rows_f32 + (size_t)77 * D, sizeof(float) * D);                  /* query 1 == that row  */

  /* loader: fp32 host rows -> normalised bf16 resident rows */
  void* rows_dev = NULL;
  CHECK_CUDA(cudaMalloc(&rows_dev, (size_t)N * D * 2));
  CHECK_MMR(mmr_load_rows_f32_host(0, rows_f32, rows_dev, MMR_BF16, N, D, 1, NULL));

  mmr_index* ix = NULL;
  CHECK_MMR(mmr_index_create(0, D, MMR_BF16, N, rows_dev, NULL, 0, 0, &ix));
  float scores[B * K];
  int64_t ids[B * K];
  CHECK_MMR(mmr_search_host(ix, queries, NULL, B, K, scores, ids, NULL));

  /* CPU check on the very same bf16 values */
  uint16_t* rows_bf16 = (uint16_t*)malloc((size_t)N * D * 2);
  CHECK_CUDA(cudaMemcpy(rows_bf16, rows_dev, (size_t)N * D * 2, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int b = 0; b < B; ++b) {
    double qn = 0;
    for (int d = 0; d < D; ++d) qn += (double)queries[b * D + d] * queries[b * D + d];
    qn = sqrt(qn);
    static float s[N];
    for (int r = 0; r < N; ++r) {
      double acc = 0;
      for (int d = 0; d < D; ++d) acc += (double)bf16_to_f32(rows_bf16[(size_t)r * D + d]) * (queries[b * D + d] / qn);
      s[r] = (float)acc;
    }
    for (int j = 0; j < K; ++j) {           /* selection: (score desc, row asc) */
      int best = -1;
      for (int r = 0; r < N; ++r)
        if (s[r] > -2.f && (best < 0 || s[r] > s[best])) best = r;
      if (ids[b * K + j] != best && fabsf(s[ids[b * K + j]] - s[best]) > 2e-6f) {
        fprintf(stderr, "query %d rank %d: got row %lld (%.7f), want %d (%.7f)\n", b, j, (long long)ids[b * K + j],
                scores[b * K + j], best, s[best]);
        ++bad;
      }
      if (fabsf(scores[b * K + j] - s[ids[b * K + j]]) > 2e-6f) ++bad;
      s[best] = -3.f;
    }
  }
  if (ids[K] != 77 || ids[K + 1] != 31234) { fprintf(stderr, "tie rule: got %lld, %lld\n", (long long)ids[K], (long long)ids[K + 1]); ++bad; }

  /* two row-range shards + mmr_merge_topk must reproduce the single scan bit for bit */
  mmr_index *s0 = NULL, *s1 = NULL;
  const int64_t half = N / 2;
  CHECK_MMR(mmr_index_create(0, D, MMR_BF16, half, rows_dev, NULL, 0, 0, &s0));
  CHECK_MMR(mmr_index_create(0, D, MMR_BF16, N - half, (char*)rows_dev + (size_t)half * D * 2, NULL, 0, half, &s1));
  float *q_dev, *sc_dev, *out_sc;
  int64_t *id_dev, *out_id;
  void* ws;
  size_t ws_bytes = mmr_search_workspace_bytes(s0, B, K);
  CHECK_CUDA(cudaMalloc((void**)&q_dev, sizeof(float) * B * D));
  CHECK_CUDA(cudaMalloc((void**)&sc_dev, sizeof(float) * 2 * B * K));
  CHECK_CUDA(cudaMalloc((void**)&id_dev, sizeof(int64_t) * 2 * B * K));
  CHECK_CUDA(cudaMalloc((void**)&out_sc, sizeof(float) * B * K));
  CHECK_CUDA(cudaMalloc((void**)&out_id, sizeof(int64_t) * B * K));
  CHECK_CUDA(cudaMalloc(&ws, ws_bytes));
  CHECK_CUDA(cudaMemset(ws, 0, ws_bytes));
  CHECK_CUDA(cudaMemcpy(q_dev, queries, sizeof(float) * B * D, cudaMemcpyHostToDevice));
  CHECK_MMR(mmr_search(s0, q_dev, NULL, B, K, sc_dev, id_dev, ws, ws_bytes, NULL));
  CHECK_MMR(mmr_search(s1, q_dev, NULL, B, K, sc_dev + B * K, id_dev + B * K, ws, ws_bytes, NULL));
  CHECK_MMR(mmr_merge_topk(sc_dev, id_dev, 2, B, K, out_sc, out_id, NULL));
  float m_sc[B * K];
  int64_t m_id[B * K];
  CHECK_CUDA(cudaMemcpy(m_sc, out_sc, sizeof(m_sc), cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(m_id, out_id, sizeof(m_id), cudaMemcpyDeviceToHost));
  if (memcmp(m_sc, scores, sizeof(m_sc)) != 0 || memcmp(m_id, ids, sizeof(m_id)) != 0) { fprintf(stderr, "sharded result differs\n"); ++bad; }

  /* error reporting through the ABI */
  if (mmr_search_host(ix, queries, NULL, B, 65, scores, ids, NULL) != MMR_ERR_INVALID || strlen(mmr_last_error()) == 0) ++bad;

  mmr_index_destroy(ix); mmr_index_destroy(s0); mmr_index_destroy(s1);
  printf("abi_smoke: %s (%lld kernel launches)\n", bad ? "FAILED" : "ok", (long long)mmr_launch_count());
  return bad ? 4 : 0;
}
