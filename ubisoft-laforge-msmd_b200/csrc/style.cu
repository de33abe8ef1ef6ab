// Style encoder (style_encoder.py:119-213, StyleEncoder_VAE2) on the tensor-core GEMM core.
//
// Layout: activations are channels-last bf16 with one zero row before and after every clip
// ([N, L+2, C]), so a k=3 / padding=1 Conv1d is a plain GEMM whose A operand is an OVERLAPPING strided
// view: row (n,t) = the 3*C contiguous elements starting at padded row t (TMA row stride = C, row length
// = 3C) - no im2col buffer.  ELU + LayerNorm (+ the single positional row pe[L]) run as
// one warp-per-row kernel that also re-creates the zero padding rows for the next conv.
#include "denoiser_kernels.cuh"
#include "gemm_tc.cuh"
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace msmd;

struct msmd_style {
  int device = 0, d_in = 67, d = 512, d_style = 256, cin_pad = 72, max_clips = 0, max_len = 0;
  bool loaded = false;
  std::vector<void*> owned;   // workspaces first (allocated by create), then the packed weights
  size_t n_workspace = (size_t)-1;   // owned[n_workspace..] are weights: released and re-packed by every load_weights
  bf16 *W1 = nullptr, *W2 = nullptr, *W3 = nullptr, *W4 = nullptr, *Wqkv = nullptr, *Wo = nullptr, *Wf1 = nullptr, *Wf2 = nullptr;
  float *b1 = nullptr, *b2 = nullptr, *b3 = nullptr, *b4 = nullptr, *bqkv = nullptr, *bo = nullptr, *bf1 = nullptr, *bf2 = nullptr;
  float *g_in1 = nullptr, *be_in1 = nullptr, *g_in2 = nullptr, *be_in2 = nullptr, *g_n1 = nullptr, *be_n1 = nullptr,
        *g_n2 = nullptr, *be_n2 = nullptr, *g_out = nullptr, *be_out = nullptr, *pe = nullptr;
  // workspaces
  bf16 *xin = nullptr, *pa = nullptr, *pb = nullptr, *xc = nullptr, *qkv = nullptr, *ctx = nullptr, *hff = nullptr;
  float* y = nullptr;
};

namespace {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// motion fp32 [N,L,cin] -> bf16 [N, L+2, cpad] with zero pad rows / channels
__global__ void style_pack_kernel(const float* __restrict__ x, bf16* __restrict__ out, int N, int L, int cin, int cpad) {
  const int64_t n_el = (int64_t)N * (L + 2) * cpad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_el; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cpad);
    const int64_t r = i / cpad;
    const int t = (int)(r % (L + 2)) - 1;
    const int64_t n = r / (L + 2);
    const float v = (t >= 0 && t < L && c < cin) ? x[(n * L + t) * cin + c] : 0.f;
    out[i] = __float2bfloat16_rn(v);
  }
}

struct RowsLnParams {
  const float* in;     // fp32 rows, row(n,t) = n*in_stride + in_off + t
  int in_stride, in_off;
  int N, L;
  int elu;             // apply ELU (alpha 1) before the LayerNorm
  const float* g; const float* b;
  const float* add;    // [512] vector added after the LayerNorm, or null
  bf16* out;           // row(n,t) = n*out_stride + out_off + t
  int out_stride, out_off;
  int zero_pad;        // also write zero rows at out_off-1 and out_off+L of every clip
};

// one warp per (n,t) row of 512 channels
__global__ void __launch_bounds__(256) rows_act_ln_kernel(RowsLnParams p) {
  constexpr int D = 512, NV = 16;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int rows_per_clip = p.L + (p.zero_pad ? 2 : 0);
  if (row >= p.N * rows_per_clip) return;
  const int n = row / rows_per_clip;
  const int t = row % rows_per_clip - (p.zero_pad ? 1 : 0);
  bf16* o = p.out + ((int64_t)n * p.out_stride + p.out_off + t) * D;
  if (t < 0 || t >= p.L) {  // padding row of the next conv
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<uint2*>(o + i * 128 + lane * 4) = make_uint2(0u, 0u);
    return;
  }
  const float* y = p.in + ((int64_t)n * p.in_stride + p.in_off + t) * D;
  float v[NV];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t4 = *reinterpret_cast<const float4*>(y + i * 128 + lane * 4);
    v[4 * i] = t4.x; v[4 * i + 1] = t4.y; v[4 * i + 2] = t4.z; v[4 * i + 3] = t4.w;
  }
  if (p.elu) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = v[i] > 0.f ? v[i] : expm1f(v[i]);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += v[i];
  const float mean = wsum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) { const float dl = v[i] - mean; q = fmaf(dl, dl, q); }
  const float rstd = rsqrtf(wsum(q) * (1.0f / D) + 1e-5f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = i * 128 + lane * 4;
    const float4 gg = *reinterpret_cast<const float4*>(p.g + c);
    const float4 bb = *reinterpret_cast<const float4*>(p.b + c);
    float4 ad = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.add) ad = *reinterpret_cast<const float4*>(p.add + c);
    const float r0 = (v[4 * i] - mean) * rstd * gg.x + bb.x + ad.x;
    const float r1 = (v[4 * i + 1] - mean) * rstd * gg.y + bb.y + ad.y;
    const float r2 = (v[4 * i + 2] - mean) * rstd * gg.z + bb.z + ad.z;
    const float r3 = (v[4 * i + 3] - mean) * rstd * gg.w + bb.w + ad.w;
    __nv_bfloat162 lo = __floats2bfloat162_rn(r0, r1), hi = __floats2bfloat162_rn(r2, r3);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(o + c) = u;
  }
}
int rows_act_ln(const RowsLnParams& p, cudaStream_t st) {
  const int rows = p.N * (p.L + (p.zero_pad ? 2 : 0));
  rows_act_ln_kernel<<<cdiv(rows, 8), 256, 0, st>>>(p);
  MSMD_CHECK_LAUNCH();
  return MSMD_OK;
}

// mean over time of conv4 output (rows n*(L+2)+t, t<L) -> mu | logvar ; style = mu + eps * exp(0.5 logvar)
__global__ void style_pool_kernel(const float* __restrict__ y, int N, int L, int d_style, const float* __restrict__ eps,
                                  float* __restrict__ style, float* __restrict__ mu, float* __restrict__ logvar) {
  const int n = blockIdx.x;
  const int D = 2 * d_style;
  for (int c = threadIdx.x; c < d_style; c += blockDim.x) {
    float sm = 0.f, sl = 0.f;
    for (int t = 0; t < L; ++t) {
      const float* r = y + ((int64_t)n * (L + 2) + t) * D;
      sm += r[c];
      sl += r[d_style + c];
    }
    sm /= (float)L;
    sl /= (float)L;
    if (mu) mu[(int64_t)n * d_style + c] = sm;
    if (logvar) logvar[(int64_t)n * d_style + c] = sl;
    if (style) style[(int64_t)n * d_style + c] = sm + (eps ? eps[(int64_t)n * d_style + c] : 0.f) * expf(0.5f * sl);
  }
}

uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

template <class Tp>
int dalloc(msmd_style* m, Tp** p, size_t n) {
  void* q = nullptr;
  MSMD_CHECK_CUDA(cudaMalloc(&q, n * sizeof(Tp)));
  m->owned.push_back(q);
  *p = static_cast<Tp*>(q);
  return MSMD_OK;
}

int gemm(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, const float* bias, const bf16* aux, int64_t ld_aux,
         void* out, int64_t ldo, int out_f32, int M, int N, int K, int act, cudaStream_t st) {
  GemmDesc d;
  d.mode = 0; d.A = A; d.W = W; d.bias = bias; d.aux = aux; d.out = out;
  d.M = M; d.N = N; d.K = K; d.lda = lda; d.ldw = ldw; d.ldo = ldo; d.ld_aux = ld_aux;
  d.out_f32 = out_f32; d.aux_f32 = 0; d.act = act; d.gelu_heavy = act;
  return gemm_tc_launch(d, st);
}

}  // namespace

extern "C" int msmd_style_create(int d_in, int d_model, int d_style, int max_clips, int max_len, int device,
                                 msmd_style** out) {
  MSMD_REQUIRE(out, "msmd_style_create: null out");
  MSMD_REQUIRE(d_in > 0 && d_in <= 72, "msmd_style_create: motion dim %d not in 1..72", d_in);
  if (d_model != 512 || 2 * d_style != 512) {
    set_error("msmd_style_create: only conv_feature_dim=512 and d_style=256 are built (got %d / %d)", d_model, d_style);
    return MSMD_ERR_UNSUPPORTED;
  }
  MSMD_REQUIRE(max_clips > 0 && max_len > 0 && max_len <= 112, "msmd_style_create: clip length %d outside 1..112 frames "
               "(single-tile attention)", max_len);
  MSMD_CHECK_CUDA(cudaSetDevice(device));
  msmd_style* m = new msmd_style();
  m->device = device; m->d_in = d_in; m->d = d_model; m->d_style = d_style; m->max_clips = max_clips; m->max_len = max_len;
  const size_t R = (size_t)max_clips * (max_len + 2), d = d_model;
  int rc = MSMD_OK;
  auto A = [&](auto** p, size_t n) { if (!rc) rc = dalloc(m, p, n); };
  A(&m->xin, R * m->cin_pad + 4 * m->cin_pad); A(&m->pa, (R + 4) * d); A(&m->pb, (R + 4) * d); A(&m->xc, R * d);
  A(&m->qkv, R * 3 * d); A(&m->ctx, R * d); A(&m->hff, R * d); A(&m->y, (R + 4) * d);
  if (rc) { msmd_style_destroy(m); return rc; }
  *out = m;
  return MSMD_OK;
}

extern "C" void msmd_style_destroy(msmd_style* m) {
  if (!m) return;
  for (void* p : m->owned) cudaFree(p);
  delete m;
}

extern "C" int msmd_style_load_weights(msmd_style* m, const char* const* names, const void* const* data,
                                       const int64_t* numel, int n) {
  MSMD_REQUIRE(m && names && data && numel, "msmd_style_load_weights: null argument");
  MSMD_CHECK_CUDA(cudaSetDevice(m->device));
  // a reload replaces the previous packed copies instead of accumulating them until destroy
  MSMD_CHECK_CUDA(cudaDeviceSynchronize());
  if (m->n_workspace == (size_t)-1) m->n_workspace = m->owned.size();
  for (size_t i = m->n_workspace; i < m->owned.size(); ++i) cudaFree(m->owned[i]);
  m->owned.resize(m->n_workspace);
  m->loaded = false;
  std::map<std::string, int> idx;
  for (int i = 0; i < n; ++i) idx[names[i]] = i;
  std::string missing;
  int rc = MSMD_OK;
  std::vector<float> h;
  auto fetch = [&](const std::string& key, size_t expect) -> bool {
    auto it = idx.find(key);
    if (it == idx.end()) { missing += key + " "; return false; }
    if ((size_t)numel[it->second] != expect) {
      set_error("msmd_style_load_weights: %s has %lld elements, expected %zu", key.c_str(), (long long)numel[it->second], expect);
      rc = MSMD_ERR_INVALID;
      return false;
    }
    h.resize(expect);
    if (cudaMemcpy(h.data(), data[it->second], expect * 4, cudaMemcpyDefault) != cudaSuccess) {
      set_error("msmd_style_load_weights: copy of %s failed", key.c_str());
      rc = MSMD_ERR_CUDA;
      return false;
    }
    return true;
  };
  auto up_bf = [&](bf16** dst, const std::vector<float>& v) {
    std::vector<uint16_t> t(v.size());
    for (size_t i = 0; i < v.size(); ++i) t[i] = f2bf(v[i]);
    if (!rc) rc = dalloc(m, dst, v.size());
    if (!rc && cudaMemcpy(*dst, t.data(), t.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) rc = MSMD_ERR_CUDA;
  };
  auto F32 = [&](const std::string& key, size_t n_, float** dst) {
    if (rc || !fetch(key, n_)) return;
    rc = dalloc(m, dst, n_);
    if (!rc && cudaMemcpy(*dst, h.data(), n_ * 4, cudaMemcpyHostToDevice) != cudaSuccess) rc = MSMD_ERR_CUDA;
  };
  auto LIN = [&](const std::string& key, size_t n_, bf16** dst) { if (!rc && fetch(key, n_)) up_bf(dst, h); };
  // Conv1d weight [Cout, Cin, 3] -> GEMM weight [Cout, 3 * Cpad] (tap-major, then channel; zero pad channels)
  auto CONV = [&](const std::string& key, int cout, int cin, int cpad, bf16** dst) {
    if (rc || !fetch(key, (size_t)cout * cin * 3)) return;
    std::vector<float> w((size_t)cout * 3 * cpad, 0.f);
    for (int o = 0; o < cout; ++o)
      for (int c = 0; c < cin; ++c)
        for (int k = 0; k < 3; ++k) w[((size_t)o * 3 + k) * cpad + c] = h[((size_t)o * cin + c) * 3 + k];
    up_bf(dst, w);
  };
  const size_t d = m->d;
  CONV("input_layers.1.weight", (int)d, m->d_in, m->cin_pad, &m->W1); F32("input_layers.1.bias", d, &m->b1);
  F32("input_layers.5.weight", d, &m->g_in1); F32("input_layers.5.bias", d, &m->be_in1);
  CONV("input_layers.7.weight", (int)d, (int)d, (int)d, &m->W2); F32("input_layers.7.bias", d, &m->b2);
  F32("input_layers.11.weight", d, &m->g_in2); F32("input_layers.11.bias", d, &m->be_in2);
  F32("PE.pe", 600 * d, &m->pe);
  LIN("encoder.self_attn.in_proj_weight", 3 * d * d, &m->Wqkv); F32("encoder.self_attn.in_proj_bias", 3 * d, &m->bqkv);
  LIN("encoder.self_attn.out_proj.weight", d * d, &m->Wo); F32("encoder.self_attn.out_proj.bias", d, &m->bo);
  LIN("encoder.linear1.weight", d * d, &m->Wf1); F32("encoder.linear1.bias", d, &m->bf1);
  LIN("encoder.linear2.weight", d * d, &m->Wf2); F32("encoder.linear2.bias", d, &m->bf2);
  F32("encoder.norm1.weight", d, &m->g_n1); F32("encoder.norm1.bias", d, &m->be_n1);
  F32("encoder.norm2.weight", d, &m->g_n2); F32("encoder.norm2.bias", d, &m->be_n2);
  CONV("output_layers.1.weight", (int)d, (int)d, (int)d, &m->W3); F32("output_layers.1.bias", d, &m->b3);
  F32("output_layers.5.weight", d, &m->g_out); F32("output_layers.5.bias", d, &m->be_out);
  CONV("output_layers.7.weight", (int)d, (int)d, (int)d, &m->W4); F32("output_layers.7.bias", d, &m->b4);
  if (rc) return rc;
  if (!missing.empty()) {
    set_error("msmd_style_load_weights: missing state_dict keys: %s", missing.c_str());
    return MSMD_ERR_INVALID;
  }
  m->loaded = true;
  return MSMD_OK;
}

extern "C" int msmd_style_encode(msmd_style* m, const float* motion, int N, int L, const float* eps, float* style_out,
                                 float* mu_out, float* logvar_out, void* stream) {
  MSMD_REQUIRE(m && motion, "msmd_style_encode: null argument");
  if (!m->loaded) { set_error("msmd_style_encode: weights not loaded"); return MSMD_ERR_STATE; }
  MSMD_REQUIRE(N > 0 && N <= m->max_clips && L > 0 && L <= m->max_len, "msmd_style_encode: N=%d L=%d outside capacity %d x %d",
               N, L, m->max_clips, m->max_len);
  MSMD_REQUIRE(L < 600, "msmd_style_encode: PositionalEncoding row %d out of range (IndexError in the reference too)", L);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int d = m->d, P = L + 2, R = N * P, Mv = R - 2, cp = m->cin_pad;
  int rc;
  style_pack_kernel<<<std::min(cdiv((int64_t)R * cp, 256), kNumSMs * 8), 256, 0, st>>>(motion, m->xin, N, L, m->d_in, cp);
  MSMD_CHECK_LAUNCH();
  RowsLnParams lp;
  lp.N = N; lp.L = L;
  // conv1 + ELU + LN -> padded
  if ((rc = gemm(m->xin, cp, m->W1, 3 * cp, m->b1, nullptr, 0, m->y, d, 1, Mv, d, 3 * cp, 0, st))) return rc;
  lp.in = m->y; lp.in_stride = P; lp.in_off = 0; lp.elu = 1; lp.g = m->g_in1; lp.b = m->be_in1; lp.add = nullptr;
  lp.out = m->pa; lp.out_stride = P; lp.out_off = 1; lp.zero_pad = 1;
  if ((rc = rows_act_ln(lp, st))) return rc;
  // conv2 + ELU + LN + pe[L] -> compact rows for the encoder layer
  if ((rc = gemm(m->pa, d, m->W2, 3 * d, m->b2, nullptr, 0, m->y, d, 1, Mv, d, 3 * d, 0, st))) return rc;
  lp.g = m->g_in2; lp.b = m->be_in2; lp.add = m->pe + (size_t)L * d; lp.out = m->xc; lp.out_stride = L; lp.out_off = 0;
  lp.zero_pad = 0;
  if ((rc = rows_act_ln(lp, st))) return rc;
  // TransformerEncoderLayer (post-LN): self-attention block
  const int M = N * L;
  if ((rc = gemm(m->xc, d, m->Wqkv, d, m->bqkv, nullptr, 0, m->qkv, 3 * d, 0, M, 3 * d, d, 0, st))) return rc;
  if ((rc = self_attn_tc_launch(m->qkv, m->ctx, N, L, 8, 0, st))) return rc;      // tcgen05 (attn_tc.cu), bf16 storage
  if ((rc = gemm(m->ctx, d, m->Wo, d, m->bo, m->xc, d, m->y, d, 1, M, d, d, 0, st))) return rc;
  lp.in = m->y; lp.in_stride = L; lp.in_off = 0; lp.elu = 0; lp.g = m->g_n1; lp.b = m->be_n1; lp.add = nullptr;
  lp.out = m->xc; lp.out_stride = L; lp.out_off = 0; lp.zero_pad = 0;
  if ((rc = rows_act_ln(lp, st))) return rc;
  // feed-forward block -> padded rows for the output convs
  if ((rc = gemm(m->xc, d, m->Wf1, d, m->bf1, nullptr, 0, m->hff, d, 0, M, d, d, 1, st))) return rc;
  if ((rc = gemm(m->hff, d, m->Wf2, d, m->bf2, m->xc, d, m->y, d, 1, M, d, d, 0, st))) return rc;
  lp.g = m->g_n2; lp.b = m->be_n2; lp.out = m->pa; lp.out_stride = P; lp.out_off = 1; lp.zero_pad = 1;
  if ((rc = rows_act_ln(lp, st))) return rc;
  // conv3 + ELU + LN -> padded ; conv4 -> fp32 ; mean over time ; reparameterise
  if ((rc = gemm(m->pa, d, m->W3, 3 * d, m->b3, nullptr, 0, m->y, d, 1, Mv, d, 3 * d, 0, st))) return rc;
  lp.in = m->y; lp.in_stride = P; lp.in_off = 0; lp.elu = 1; lp.g = m->g_out; lp.b = m->be_out; lp.out = m->pb;
  if ((rc = rows_act_ln(lp, st))) return rc;
  if ((rc = gemm(m->pb, d, m->W4, 3 * d, m->b4, nullptr, 0, m->y, d, 1, Mv, d, 3 * d, 0, st))) return rc;
  style_pool_kernel<<<N, 256, 0, st>>>(m->y, N, L, m->d_style, eps, style_out, mu_out, logvar_out);
  MSMD_CHECK_LAUNCH();
  return MSMD_OK;
}
