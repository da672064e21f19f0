// Error / accounting plumbing shared by the translation units of libmmr_b200.so.
#pragma once
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include <cuda_runtime.h>

#include "../../include/mmr_b200.h"

extern thread_local std::string mmr_g_err;       // mmr_last_error()
extern std::atomic<int64_t> mmr_g_launches;      // mmr_launch_count()
int mmr_fail(int code, const char* fmt, ...);

#define CUDA_TRY(expr)                                                                                    \
  do {                                                                                                    \
    cudaError_t e_ = (expr);                                                                              \
    if (e_ != cudaSuccess) return mmr_fail(MMR_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                                           __FILE__, __LINE__);                                           \
  } while (0)
