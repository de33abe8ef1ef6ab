// Shared host/device helpers for the msmd_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>

#include "../../include/msmd_b200.h"

namespace msmd {

void set_error(const char* fmt, ...);

#define MSMD_CHECK_CUDA(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::msmd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                        __LINE__);                                                         \
      return MSMD_ERR_CUDA;                                                                \
    }                                                                                      \
  } while (0)

#define MSMD_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      ::msmd::set_error(__VA_ARGS__);    \
      return MSMD_ERR_INVALID;           \
    }                                    \
  } while (0)

#define MSMD_CHECK_LAUNCH() MSMD_CHECK_CUDA(cudaGetLastError())

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

}  // namespace msmd
