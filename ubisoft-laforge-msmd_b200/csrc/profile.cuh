// Optional per-kernel-class CUDA-event timing (bench.py's roofline figures are measured live
// with these events, on the stream the kernel is launched on).  Off by default; zero cost when off.
#pragma once
#include "common.cuh"

namespace msmd {
bool profiling_on();
// Record an event pair around a launch: {ProfileScope p("name", stream); kernel<<<...>>>;}
struct ProfileScope {
  const char* name;
  cudaStream_t st;
  cudaEvent_t e0 = nullptr;
  ProfileScope(const char* n, cudaStream_t s);
  ~ProfileScope();
};
}  // namespace msmd
