// Clip front-end of the inference driver, the two host-side numpy stages that sit
// between file I/O and the encoders:
//   msmd_audio_normalize  — inference.py:234  audio = (audio - audio.mean()) / (audio.std() + 1e-5), per clip
//   msmd_resample_linear  — inference.py:158-171  scipy interp1d(kind='linear', axis=0) of the style clip from
//                           np.linspace(0, 1, rows_in) onto np.linspace(0, 1, rows_out)
// Both are HBM-bound one-pass kernels; statistics are accumulated in double.
#include "common.cuh"
#include "profile.cuh"
#include <algorithm>

namespace msmd {
namespace {

// one CTA per clip: pass 1 sum / sum of squares (double, warp shuffles + smem), pass 2 normalise
__global__ void __launch_bounds__(1024) audio_normalize_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                               int64_t n) {
  __shared__ double s_sum[32], s_sq[32];
  __shared__ float s_mean, s_inv;
  const float* x = in + (int64_t)blockIdx.x * n;
  float* y = out + (int64_t)blockIdx.x * n;
  double a = 0.0, b = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = x[i];
    a += v;
    b += v * v;
  }
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_sum[warp] = a; s_sq[warp] = b; }
  __syncthreads();
  if (warp == 0) {
    a = lane < (int)(blockDim.x >> 5) ? s_sum[lane] : 0.0;
    b = lane < (int)(blockDim.x >> 5) ? s_sq[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
      const double mean = a / (double)n;
      const double var = fmax(b / (double)n - mean * mean, 0.0);   // numpy std: population (ddof = 0)
      s_mean = (float)mean;
      s_inv = (float)(1.0 / (sqrt(var) + 1e-5));
    }
  }
  __syncthreads();
  const float mean = s_mean, inv = s_inv;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) y[i] = (x[i] - mean) * inv;
}

// out[r, c] = in[lo, c] + frac * (in[lo + 1, c] - in[lo, c]),  position = r * (rows_in - 1) / (rows_out - 1)
__global__ void resample_linear_kernel(const float* __restrict__ in, float* __restrict__ out, int rows_in, int rows_out,
                                       int cols) {
  const int64_t total = (int64_t)rows_out * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    const double pos = rows_out > 1 ? (double)r * (double)(rows_in - 1) / (double)(rows_out - 1) : 0.0;
    int lo = (int)floor(pos);
    if (lo > rows_in - 2) lo = rows_in - 2;
    if (lo < 0) lo = 0;
    const double frac = pos - (double)lo;
    const double y0 = in[(int64_t)lo * cols + c];
    const double y1 = rows_in > 1 ? (double)in[(int64_t)(lo + 1) * cols + c] : y0;
    out[i] = (float)(y0 + frac * (y1 - y0));
  }
}

}  // namespace
}  // namespace msmd

using namespace msmd;

extern "C" int msmd_audio_normalize(const float* in, float* out, int n_clips, int64_t n_samples, void* stream) {
  MSMD_REQUIRE(n_clips >= 0 && n_samples >= 0, "msmd_audio_normalize: negative size");
  if (n_clips == 0 || n_samples == 0) return MSMD_OK;
  MSMD_REQUIRE(in && out, "msmd_audio_normalize: null pointer");
  ProfileScope prof("audio_normalize", static_cast<cudaStream_t>(stream));
  audio_normalize_kernel<<<n_clips, 1024, 0, static_cast<cudaStream_t>(stream)>>>(in, out, n_samples);
  MSMD_CHECK_LAUNCH();
  return MSMD_OK;
}

extern "C" int msmd_resample_linear(const float* in, float* out, int rows_in, int rows_out, int cols, void* stream) {
  MSMD_REQUIRE(rows_in >= 1 && rows_out >= 0 && cols >= 0, "msmd_resample_linear: need at least one input row (got %d)", rows_in);
  if (rows_out == 0 || cols == 0) return MSMD_OK;
  MSMD_REQUIRE(in && out, "msmd_resample_linear: null pointer");
  const int64_t total = (int64_t)rows_out * cols;
  const int blocks = (int)std::min<int64_t>(cdiv(total, 256), kNumSMs * 8);
  ProfileScope prof("resample_linear", static_cast<cudaStream_t>(stream));
  resample_linear_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, rows_in, rows_out, cols);
  MSMD_CHECK_LAUNCH();
  return MSMD_OK;
}
