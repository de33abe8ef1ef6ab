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

// Programmatic dependent launch (PDL): a kernel launched with launch_pdl() may start while its predecessor in
// the stream is still draining; it must call griddep_wait() before touching global memory and should call
// griddep_launch() early so that ITS successor can be scheduled.  Meant to hide launch latency + kernel
// prologues (barrier init, TMEM alloc, tensor-map prefetch).
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();   // off by default: measured 1983.5 us/step with PDL vs 1967.4 without inside the CUDA graph
                      // (graph replays already back-to-back the kernels); MSMD_PDL=1 enables it for eager use

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace msmd
