// Non-GEMM kernels of the audio encoder: raw-audio conv layer + time-axis GroupNorm, resampling +
// LayerNorm, positional-conv packing, LayerNorm(768), flash-style attention (mma.sync bf16, online softmax).
#include "audio_kernels.cuh"
#include "profile.cuh"
#include <cmath>

namespace msmd {

namespace {
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float erf_as(float x) {  // Abramowitz-Stegun 7.1.26, |err| <= 1.5e-7
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  return copysignf(1.0f - p * t * __expf(-ax * ax), x);
}
__device__ __forceinline__ float gelu_fast(float x) { return 0.5f * x * (1.0f + erf_as(x * 0.70710678118654752f)); }

// model_common.py:110-123: reflect twice by r, then replicate one sample each side if rep
__device__ __forceinline__ int pad_src(int i, const PadSpec& ps) {
  const int len1 = ps.n + 2 * ps.r, len2 = ps.n + 4 * ps.r;
  if (ps.rep) i = min(max(i - 1, 0), len2 - 1);
  int j = i - ps.r;
  j = j < 0 ? -j : (j >= len1 ? 2 * (len1 - 1) - j : j);
  int k = j - ps.r;
  k = k < 0 ? -k : (k >= ps.n ? 2 * (ps.n - 1) - k : k);
  return k;
}
}  // namespace

PadSpec make_pad_spec(int n) {
  PadSpec ps{n, 0, 0, n};
  const int side = (int)std::ceil((320.0 * (n / 320) + 80 - n) / 2.0);
  if (side >= 0) {
    ps.r = side / 2;
    ps.rep = side % 2;
    ps.n_pad = n + 4 * ps.r + 2 * ps.rep;
  }
  return ps;
}

// ------------------------------------------------------------------------------------------- conv layer 0
constexpr int kC0Frames = 128;  // frames per block; 512 threads = one output channel each

template <bool APPLY>
__global__ void __launch_bounds__(512) conv0_kernel(const float* __restrict__ wav, PadSpec ps, int T0,
                                                    const float* __restrict__ w0, const float* __restrict__ gn_w,
                                                    const float* __restrict__ gn_b, double* __restrict__ stats,
                                                    bf16* __restrict__ out) {
  __shared__ float xs[kC0Frames * 5 + 8];
  const int n = blockIdx.y;
  const int t0 = blockIdx.x * kC0Frames;
  const int nt = min(kC0Frames, T0 - t0);
  const int c = threadIdx.x;
  for (int i = threadIdx.x; i < nt * 5 + 5; i += blockDim.x) {
    const int src = t0 * 5 + i;
    xs[i] = src < ps.n_pad ? wav[(int64_t)n * ps.n + pad_src(src, ps)] : 0.f;
  }
  float w[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) w[k] = w0[c * 10 + k];
  float mean = 0.f, rstd = 0.f, gw = 0.f, gb = 0.f;
  if (APPLY) {
    const double s = stats[((int64_t)n * 512 + c) * 2], q = stats[((int64_t)n * 512 + c) * 2 + 1];
    const double mu = s / T0;
    const double var = fmax(q / T0 - mu * mu, 0.0);
    mean = (float)mu;
    rstd = (float)(1.0 / sqrt(var + 1e-5));
    gw = gn_w[c];
    gb = gn_b[c];
  }
  __syncthreads();
  float s1 = 0.f, s2 = 0.f;
  for (int t = 0; t < nt; ++t) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 10; ++k) v = fmaf(w[k], xs[t * 5 + k], v);
    if (APPLY) {
      v = gelu_fast((v - mean) * rstd * gw + gb);
      out[((int64_t)n * T0 + t0 + t) * 512 + c] = __float2bfloat16_rn(v);
    } else {
      s1 += v;
      s2 = fmaf(v, v, s2);
    }
  }
  if (!APPLY) {
    atomicAdd(&stats[((int64_t)n * 512 + c) * 2], (double)s1);
    atomicAdd(&stats[((int64_t)n * 512 + c) * 2 + 1], (double)s2);
  }
}

int conv0_groupnorm_gelu(const float* wav, int N, PadSpec ps, int T0, const float* w0, const float* gn_w,
                         const float* gn_b, double* stats, bf16* out, cudaStream_t st) {
  MSMD_CHECK_CUDA(cudaMemsetAsync(stats, 0, (size_t)N * 512 * 2 * sizeof(double), st));
  dim3 grid(cdiv(T0, kC0Frames), N);
  conv0_kernel<false><<<grid, 512, 0, st>>>(wav, ps, T0, w0, gn_w, gn_b, stats, out);
  MSMD_CHECK_LAUNCH();
  conv0_kernel<true><<<grid, 512, 0, st>>>(wav, ps, T0, w0, gn_w, gn_b, stats, out);
  MSMD_CHECK_LAUNCH();
  return MSMD_OK;
}

// ------------------------------------------------------------------------------------------- resample + LN(512)
__device__ __forceinline__ void lerp_coords(int j, int in_len, int out_len, int& i0, int& i1, float& lam) {
  // F.interpolate(mode='linear', align_corners=False): src = (j + 0.5) * in/out - 0.5, clamped at 0
  const float scale = (float)in_len / (float)out_len;
  float src = ((float)j + 0.5f) * scale - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = (int)src;
  i0 = min(i0, in_len - 1);
  i1 = min(i0 + 1, in_len - 1);
  lam = src - (float)i0;
}

__global__ void __launch_bounds__(256) interp_ln512_kernel(const float* __restrict__ x, int N, int in_rows, int in_len,
                                                           int out_len, const float* __restrict__ g,
                                                           const float* __restrict__ b, bf16* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N * out_len) return;
  const int n = row / out_len, j = row % out_len;
  int i0, i1;
  float lam;
  lerp_coords(j, in_len, out_len, i0, i1, lam);
  const float* r0 = x + ((int64_t)n * in_rows + i0) * 512;
  const float* r1 = x + ((int64_t)n * in_rows + i1) * 512;
  float v[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 a = *reinterpret_cast<const float4*>(r0 + i * 128 + lane * 4);
    const float4 c = *reinterpret_cast<const float4*>(r1 + i * 128 + lane * 4);
    const float l0 = 1.0f - lam;
    v[4 * i] = l0 * a.x + lam * c.x; v[4 * i + 1] = l0 * a.y + lam * c.y;
    v[4 * i + 2] = l0 * a.z + lam * c.z; v[4 * i + 3] = l0 * a.w + lam * c.w;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  const float mean = wsum(s) * (1.0f / 512);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) { const float dl = v[i] - mean; q = fmaf(dl, dl, q); }
  const float rstd = rsqrtf(wsum(q) * (1.0f / 512) + 1e-5f);
  bf16* o = out + (int64_t)row * 512;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = i * 128 + lane * 4;
    const float4 gg = *reinterpret_cast<const float4*>(g + c);
    const float4 bb = *reinterpret_cast<const float4*>(b + c);
    uint2 u;
    u.x = pack2((v[4 * i] - mean) * rstd * gg.x + bb.x, (v[4 * i + 1] - mean) * rstd * gg.y + bb.y);
    u.y = pack2((v[4 * i + 2] - mean) * rstd * gg.z + bb.z, (v[4 * i + 3] - mean) * rstd * gg.w + bb.w);
    *reinterpret_cast<uint2*>(o + c) = u;
  }
}
int interp_ln512(const float* x, int N, int in_rows_per_clip, int in_len, int out_len, const float* g, const float* b,
                 bf16* out, cudaStream_t st) {
  interp_ln512_kernel<<<cdiv((int64_t)N * out_len, 8), 256, 0, st>>>(x, N, in_rows_per_clip, in_len, out_len, g, b, out);
  MSMD_CHECK_LAUNCH();
  return MSMD_OK;
}

// ------------------------------------------------------------------------------------------- positional conv glue
__global__ void pos_pack_kernel(const float* __restrict__ h0, bf16* __restrict__ xg, int N, int F) {
  const int P = F + 128;
  const int64_t n_el = (int64_t)N * 16 * P * 48;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_el; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % 48);
    const int64_t r = i / 48;
    const int tp = (int)(r % P);
    const int g = (int)((r / P) % 16);
    const int64_t n = r / ((int64_t)P * 16);
    const int t = tp - 64;
    const float v = (t >= 0 && t < F) ? h0[(n * F + t) * 768 + g * 48 + c] : 0.f;
    xg[i] = __float2bfloat16_rn(v);
  }
}
int pos_pack(const float* h0, bf16* xg, int N, int F, cudaStream_t st) {
  const int64_t n_el = (int64_t)N * 16 * (F + 128) * 48;
  pos_pack_kernel<<<(int)std::min<int64_t>(cdiv(n_el, 256), kNumSMs * 16), 256, 0, st>>>(h0, xg, N, F);
  MSMD_CHECK_LAUNCH();
  return MSMD_OK;
}

// warp per 768-wide row: lane holds columns i*128 + lane*4 + {0..3}, i = 0..5
__device__ __forceinline__ void ln768_row(float (&v)[24], const float* __restrict__ g, const float* __restrict__ b,
                                          int lane, bf16* o, float* o32) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) s += v[i];
  const float mean = wsum(s) * (1.0f / 768);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) { const float dl = v[i] - mean; q = fmaf(dl, dl, q); }
  const float rstd = rsqrtf(wsum(q) * (1.0f / 768) + 1e-5f);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int c = i * 128 + lane * 4;
    const float4 gg = *reinterpret_cast<const float4*>(g + c);
    const float4 bb = *reinterpret_cast<const float4*>(b + c);
    const float r0 = (v[4 * i] - mean) * rstd * gg.x + bb.x, r1 = (v[4 * i + 1] - mean) * rstd * gg.y + bb.y;
    const float r2 = (v[4 * i + 2] - mean) * rstd * gg.z + bb.z, r3 = (v[4 * i + 3] - mean) * rstd * gg.w + bb.w;
    uint2 u;
    u.x = pack2(r0, r1);
    u.y = pack2(r2, r3);
    *reinterpret_cast<uint2*>(o + c) = u;
    if (o32) *reinterpret_cast<float4*>(o32 + c) = make_float4(r0, r1, r2, r3);
  }
}

__global__ void __launch_bounds__(256) pos_add_ln768_kernel(const float* __restrict__ h0, const float* __restrict__ pos,
                                                            const float* __restrict__ g, const float* __restrict__ b,
                                                            bf16* __restrict__ out, int N, int F) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N * F) return;
  const int n = row / F, t = row % F;
  float v[24];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int c = i * 128 + lane * 4;
    const float4 a = *reinterpret_cast<const float4*>(h0 + (int64_t)row * 768 + c);
    const int grp = c / 48, oc = c % 48;
    const float4 p4 = *reinterpret_cast<const float4*>(pos + (((int64_t)n * 16 + grp) * F + t) * 48 + oc);
    v[4 * i] = a.x + 0.5f * p4.x * (1.0f + erff(p4.x * 0.70710678118654752f));
    v[4 * i + 1] = a.y + 0.5f * p4.y * (1.0f + erff(p4.y * 0.70710678118654752f));
    v[4 * i + 2] = a.z + 0.5f * p4.z * (1.0f + erff(p4.z * 0.70710678118654752f));
    v[4 * i + 3] = a.w + 0.5f * p4.w * (1.0f + erff(p4.w * 0.70710678118654752f));
  }
  ln768_row(v, g, b, lane, out + (int64_t)row * 768, nullptr);
}
int pos_add_ln768(const float* h0, const float* pos, const float* g, const float* b, bf16* out, int N, int F,
                  cudaStream_t st) {
  pos_add_ln768_kernel<<<cdiv((int64_t)N * F, 8), 256, 0, st>>>(h0, pos, g, b, out, N, F);
  MSMD_CHECK_LAUNCH();
  return MSMD_OK;
}

__global__ void __launch_bounds__(256) ln768_kernel(const float* __restrict__ y, const float* __restrict__ g,
                                                    const float* __restrict__ b, bf16* __restrict__ out,
                                                    float* __restrict__ out32, int M) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  float v[24];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float4 a = *reinterpret_cast<const float4*>(y + (int64_t)row * 768 + i * 128 + lane * 4);
    v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
  }
  ln768_row(v, g, b, lane, out + (int64_t)row * 768, out32 ? out32 + (int64_t)row * 768 : nullptr);
}
int ln768(const float* y, const float* g, const float* b, bf16* out, float* out_f32, int M, cudaStream_t st) {
  ln768_kernel<<<cdiv(M, 8), 256, 0, st>>>(y, g, b, out, out_f32, M);
  MSMD_CHECK_LAUNCH();
  return MSMD_OK;
}

__global__ void interp768_kernel(const float* __restrict__ hs, int N, int F, int L, bf16* __restrict__ out) {
  const int64_t n_el = (int64_t)N * L * 192;  // float4 groups
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_el; i += (int64_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % 192);
    const int64_t row = i / 192;
    const int n = (int)(row / L), j = (int)(row % L);
    int i0, i1;
    float lam;
    lerp_coords(j, F, L, i0, i1, lam);
    const float4 a = *reinterpret_cast<const float4*>(hs + ((int64_t)n * F + i0) * 768 + c4 * 4);
    const float4 c = *reinterpret_cast<const float4*>(hs + ((int64_t)n * F + i1) * 768 + c4 * 4);
    const float l0 = 1.0f - lam;
    uint2 u;
    u.x = pack2(l0 * a.x + lam * c.x, l0 * a.y + lam * c.y);
    u.y = pack2(l0 * a.z + lam * c.z, l0 * a.w + lam * c.w);
    *reinterpret_cast<uint2*>(out + row * 768 + c4 * 4) = u;
  }
}
int interp768_bf16(const float* hs, int N, int F, int L, bf16* out, cudaStream_t st) {
  const int64_t n_el = (int64_t)N * L * 192;
  interp768_kernel<<<(int)std::min<int64_t>(cdiv(n_el, 256), kNumSMs * 16), 256, 0, st>>>(hs, N, F, L, out);
  MSMD_CHECK_LAUNCH();
  return MSMD_OK;
}

// (the encoder's self-attention is flash_attn_tc.cu: tcgen05 S = Q K^T / O = P V with an online softmax)

}  // namespace msmd
