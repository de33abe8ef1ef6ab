// Audio encoder engine: msmd_audio_* (include/msmd_b200.h).
// Replaces utils/hubert.py:13-51 / utils/wav2vec2.py:71-119 (HF Hubert / Wav2Vec2 base forward with the
// reference's resampling) and model.py:250-264 (extract_audio_feature).
//
// Layout: channels-last bf16 activations [clip, frame, channel].  Each strided Conv1d of the feature
// extractor is a batched GEMM whose A operand is an overlapping TMA view of the previous activation
// (row t = the k*512 contiguous elements starting at frame stride*t) - no im2col buffer.  The grouped
// positional conv (k=128, 16 groups) is 16 GEMMs per clip over a group-major zero-padded copy.
#include "audio_kernels.cuh"
#include <cstdlib>
#include "gemm_tc.cuh"
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace msmd;

namespace {
constexpr int kConvK[7] = {10, 3, 3, 3, 3, 2, 2};
constexpr int kConvS[7] = {5, 2, 2, 2, 2, 2, 2};

struct EncLayer {
  bf16 *Wqkv = nullptr, *Wo = nullptr, *W1 = nullptr, *W2 = nullptr;
  float *bqkv = nullptr, *bo = nullptr, *b1 = nullptr, *b2 = nullptr, *g1 = nullptr, *be1 = nullptr, *g2 = nullptr, *be2 = nullptr;
};

uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
}  // namespace

struct msmd_audio {
  int device = 0, max_clips = 0, max_samples = 0, d_out = 512, n_layers = 12, n_heads = 12;
  bool loaded = false;
  std::vector<void*> owned;   // workspaces first (allocated by create), then the packed weights
  size_t n_workspace = (size_t)-1;   // owned[n_workspace..] are weights: released and re-packed by every load_weights
  // weights
  float *w0 = nullptr, *gn_w = nullptr, *gn_b = nullptr, *fp_g = nullptr, *fp_b = nullptr, *fp_bias = nullptr,
        *pos_bias = nullptr, *enc_g = nullptr, *enc_b = nullptr, *map_b = nullptr;
  bf16 *Wc[7] = {nullptr}, *Wfp = nullptr, *Wpos = nullptr, *Wmap = nullptr;
  std::vector<EncLayer> L;
  // workspaces
  size_t cap_act0 = 0, cap_rows = 0;
  bf16 *actA = nullptr, *actB = nullptr, *xin = nullptr, *xg = nullptr, *x = nullptr, *qkv = nullptr, *ctx = nullptr,
       *hff = nullptr, *xl = nullptr;
  float *c6 = nullptr, *h0 = nullptr, *pos = nullptr, *y = nullptr, *hs = nullptr;
  double* stats = nullptr;
};

namespace {

template <class Tp>
int dalloc(msmd_audio* m, Tp** p, size_t n) {
  void* q = nullptr;
  MSMD_CHECK_CUDA(cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(Tp)));
  m->owned.push_back(q);
  *p = static_cast<Tp*>(q);
  return MSMD_OK;
}

void conv_lengths(int n_pad, int (&T)[7]) {
  int len = n_pad;
  for (int i = 0; i < 7; ++i) {
    len = (len - kConvK[i]) / kConvS[i] + 1;
    T[i] = len;
  }
}

int gemm2d(const bf16* A, int64_t lda, const bf16* W, int64_t ldw, const float* bias, const bf16* aux, int64_t ld_aux,
           void* out, int64_t ldo, int out_f32, int M, int N, int K, int act, cudaStream_t st) {
  GemmDesc d;
  d.mode = 0; d.A = A; d.W = W; d.bias = bias; d.aux = aux; d.out = out;
  d.M = M; d.N = N; d.K = K; d.lda = lda; d.ldw = ldw; d.ldo = ldo; d.ld_aux = ld_aux;
  d.out_f32 = out_f32; d.aux_f32 = 0; d.act = act; d.gelu_heavy = act;
  return gemm_tc_launch(d, st);
}

}  // namespace

extern "C" int msmd_audio_create(int max_clips, int max_samples, int d_out, int device, msmd_audio** out) {
  MSMD_REQUIRE(out && max_clips > 0 && max_samples >= 400, "msmd_audio_create: bad sizes");
  MSMD_REQUIRE(d_out > 0 && d_out % 8 == 0, "msmd_audio_create: feature_dim %d must be a multiple of 8", d_out);
  MSMD_CHECK_CUDA(cudaSetDevice(device));
  msmd_audio* m = new msmd_audio();
  m->device = device; m->max_clips = max_clips; m->max_samples = max_samples; m->d_out = d_out;
  m->L.resize(m->n_layers);
  const PadSpec ps = make_pad_spec(max_samples);
  int T[7];
  conv_lengths(ps.n_pad + 8, T);
  const size_t N = max_clips, F = T[6] + 2, rows = N * F;
  m->cap_act0 = N * (size_t)T[0] * 512;
  m->cap_rows = rows;
  int rc = MSMD_OK;
  auto A = [&](auto** p, size_t n) { if (!rc) rc = dalloc(m, p, n); };
  A(&m->actA, m->cap_act0 + 4096); A(&m->actB, N * (size_t)T[1] * 512 + 4096); A(&m->c6, rows * 512);
  A(&m->xin, rows * 512); A(&m->h0, rows * 768); A(&m->xg, N * 16 * (F + 128) * 48 + 8192); A(&m->pos, rows * 768);
  A(&m->x, rows * 768); A(&m->qkv, rows * 2304); A(&m->ctx, rows * 768); A(&m->hff, rows * 3072);
  A(&m->y, rows * 768); A(&m->hs, rows * 768); A(&m->xl, rows * 768); A(&m->stats, N * 512 * 2);
  if (rc) { msmd_audio_destroy(m); return rc; }
  *out = m;
  return MSMD_OK;
}

extern "C" void msmd_audio_destroy(msmd_audio* m) {
  if (!m) return;
  for (void* p : m->owned) cudaFree(p);
  delete m;
}

extern "C" int msmd_audio_load_weights(msmd_audio* m, const char* const* names, const void* const* data,
                                       const int64_t* numel, int n) {
  MSMD_REQUIRE(m && names && data && numel, "msmd_audio_load_weights: null argument");
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
  auto has = [&](const std::string& k) { return idx.count(k) > 0; };
  auto fetch = [&](const std::string& key, size_t expect, std::vector<float>& dst) -> bool {
    auto it = idx.find(key);
    if (it == idx.end()) { missing += key + " "; return false; }
    if ((size_t)numel[it->second] != expect) {
      set_error("msmd_audio_load_weights: %s has %lld elements, expected %zu", key.c_str(), (long long)numel[it->second], expect);
      rc = MSMD_ERR_INVALID;
      return false;
    }
    dst.resize(expect);
    if (cudaMemcpy(dst.data(), data[it->second], expect * 4, cudaMemcpyDefault) != cudaSuccess) {
      set_error("msmd_audio_load_weights: copy of %s failed", key.c_str());
      rc = MSMD_ERR_CUDA;
      return false;
    }
    return true;
  };
  auto up32 = [&](float** dst, const std::vector<float>& v) {
    if (!rc) rc = dalloc(m, dst, v.size());
    if (!rc && cudaMemcpy(*dst, v.data(), v.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) rc = MSMD_ERR_CUDA;
  };
  auto upbf = [&](bf16** dst, const std::vector<float>& v) {
    std::vector<uint16_t> t(v.size());
    for (size_t i = 0; i < v.size(); ++i) t[i] = f2bf(v[i]);
    if (!rc) rc = dalloc(m, dst, v.size());
    if (!rc && cudaMemcpy(*dst, t.data(), t.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) rc = MSMD_ERR_CUDA;
  };
  auto F32 = [&](const std::string& key, size_t n_, float** dst) { if (!rc && fetch(key, n_, h)) up32(dst, h); };
  const std::string P = "audio_encoder.";
  const std::string fe = P + "feature_extractor.conv_layers.";
  F32(fe + "0.conv.weight", 512 * 10, &m->w0);
  F32(fe + "0.layer_norm.weight", 512, &m->gn_w);
  F32(fe + "0.layer_norm.bias", 512, &m->gn_b);
  for (int i = 1; i < 7 && !rc; ++i) {  // [512, 512, k] -> [512, k*512] tap-major
    const int k = kConvK[i];
    if (!fetch(fe + std::to_string(i) + ".conv.weight", (size_t)512 * 512 * k, h)) continue;
    std::vector<float> w((size_t)512 * k * 512);
    for (int o = 0; o < 512; ++o)
      for (int c = 0; c < 512; ++c)
        for (int kk = 0; kk < k; ++kk) w[((size_t)o * k + kk) * 512 + c] = h[((size_t)o * 512 + c) * k + kk];
    upbf(&m->Wc[i], w);
  }
  F32(P + "feature_projection.layer_norm.weight", 512, &m->fp_g);
  F32(P + "feature_projection.layer_norm.bias", 512, &m->fp_b);
  if (!rc && fetch(P + "feature_projection.projection.weight", 768 * 512, h)) upbf(&m->Wfp, h);
  F32(P + "feature_projection.projection.bias", 768, &m->fp_bias);
  {  // positional conv: fold the weight norm (dim=2): w = g[k] * v / ||v[:,:,k]||, then [16][48][k*48 + ic]
    const std::string pc = P + "encoder.pos_conv_embed.conv.";
    std::vector<float> g, v;
    const bool newn = has(pc + "parametrizations.weight.original0");
    if (!rc && fetch(pc + (newn ? "parametrizations.weight.original0" : "weight_g"), 128, g) &&
        fetch(pc + (newn ? "parametrizations.weight.original1" : "weight_v"), (size_t)768 * 48 * 128, v)) {
      std::vector<double> nrm(128, 0.0);
      for (size_t i = 0; i < v.size(); ++i) nrm[i % 128] += (double)v[i] * v[i];
      std::vector<float> w((size_t)16 * 48 * 6144);
      for (int o = 0; o < 768; ++o)
        for (int ic = 0; ic < 48; ++ic)
          for (int k = 0; k < 128; ++k) {
            const float val = (float)((double)g[k] * v[((size_t)o * 48 + ic) * 128 + k] / std::sqrt(nrm[k]));
            w[((size_t)(o / 48) * 48 + (o % 48)) * 6144 + (size_t)k * 48 + ic] = val;
          }
      upbf(&m->Wpos, w);
    }
    F32(pc + "bias", 768, &m->pos_bias);
  }
  F32(P + "encoder.layer_norm.weight", 768, &m->enc_g);
  F32(P + "encoder.layer_norm.bias", 768, &m->enc_b);
  for (int l = 0; l < m->n_layers && !rc; ++l) {
    EncLayer& w = m->L[l];
    const std::string q = P + "encoder.layers." + std::to_string(l) + ".";
    std::vector<float> wq, wk, wv, bq, bk, bv;
    if (fetch(q + "attention.q_proj.weight", 768 * 768, wq) && fetch(q + "attention.k_proj.weight", 768 * 768, wk) &&
        fetch(q + "attention.v_proj.weight", 768 * 768, wv) && fetch(q + "attention.q_proj.bias", 768, bq) &&
        fetch(q + "attention.k_proj.bias", 768, bk) && fetch(q + "attention.v_proj.bias", 768, bv)) {
      // packed q|k|v; the 1/sqrt(64) query scaling (HF *Attention: q_proj(x) * scaling) is folded in (exact: 2^-3)
      std::vector<float> W(3 * 768 * 768), B(3 * 768);
      for (size_t i = 0; i < wq.size(); ++i) { W[i] = wq[i] * 0.125f; W[768 * 768 + i] = wk[i]; W[2 * 768 * 768 + i] = wv[i]; }
      for (int i = 0; i < 768; ++i) { B[i] = bq[i] * 0.125f; B[768 + i] = bk[i]; B[1536 + i] = bv[i]; }
      upbf(&w.Wqkv, W);
      up32(&w.bqkv, B);
    }
    if (!rc && fetch(q + "attention.out_proj.weight", 768 * 768, h)) upbf(&w.Wo, h);
    F32(q + "attention.out_proj.bias", 768, &w.bo);
    F32(q + "layer_norm.weight", 768, &w.g1); F32(q + "layer_norm.bias", 768, &w.be1);
    if (!rc && fetch(q + "feed_forward.intermediate_dense.weight", 3072 * 768, h)) upbf(&w.W1, h);
    F32(q + "feed_forward.intermediate_dense.bias", 3072, &w.b1);
    if (!rc && fetch(q + "feed_forward.output_dense.weight", 768 * 3072, h)) upbf(&w.W2, h);
    F32(q + "feed_forward.output_dense.bias", 768, &w.b2);
    F32(q + "final_layer_norm.weight", 768, &w.g2); F32(q + "final_layer_norm.bias", 768, &w.be2);
  }
  if (has("audio_feature_map.weight")) {  // optional: absent when the encoder is used on its own (hubert.py forward)
    if (!rc && fetch("audio_feature_map.weight", (size_t)m->d_out * 768, h)) upbf(&m->Wmap, h);
    F32("audio_feature_map.bias", m->d_out, &m->map_b);
  }
  if (rc) return rc;
  if (!missing.empty()) {
    set_error("msmd_audio_load_weights: missing state_dict keys: %s", missing.c_str());
    return MSMD_ERR_INVALID;
  }
  m->loaded = true;
  return MSMD_OK;
}

extern "C" int msmd_audio_encode(msmd_audio* m, const float* wav, int N, int n_samples, int output_fps, int frame_num,
                                 float* hidden_out, int feat_frames, float* feat_out, void* stream) {
  MSMD_REQUIRE(m && wav, "msmd_audio_encode: null argument");
  if (!m->loaded) { set_error("msmd_audio_encode: weights not loaded"); return MSMD_ERR_STATE; }
  MSMD_REQUIRE(N > 0 && N <= m->max_clips && n_samples >= 400 && n_samples <= m->max_samples,
               "msmd_audio_encode: N=%d n=%d outside the capacity %d x %d given at create", N, n_samples, m->max_clips,
               m->max_samples);
  MSMD_REQUIRE(output_fps > 0 && frame_num > 0, "msmd_audio_encode: bad fps / frame_num");
  MSMD_REQUIRE(!feat_out || feat_frames > 0, "msmd_audio_encode: feat_frames must be > 0 with feat_out");
  MSMD_REQUIRE(!feat_out || m->Wmap, "msmd_audio_encode: audio_feature_map weights were not loaded");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const PadSpec ps = make_pad_spec(n_samples);
  int T[7];
  conv_lengths(ps.n_pad, T);
  MSMD_REQUIRE(T[6] >= 1, "msmd_audio_encode: clip too short");
  const int F = frame_num;
  MSMD_REQUIRE((size_t)N * F <= m->cap_rows && (size_t)N * T[0] * 512 <= m->cap_act0,
               "msmd_audio_encode: frame_num %d exceeds the workspace", F);
  int rc;
  // feature extractor
  if ((rc = conv0_groupnorm_gelu(wav, N, ps, T[0], m->w0, m->gn_w, m->gn_b, m->stats, m->actA, st))) return rc;
  bf16* in = m->actA;
  bf16* outb = m->actB;
  for (int i = 1; i < 7; ++i) {
    GemmDesc d;
    d.mode = 0; d.A = in; d.W = m->Wc[i]; d.bias = nullptr; d.act = 1; d.gelu_heavy = 1;
    d.M = T[i]; d.N = 512; d.K = kConvK[i] * 512;
    d.lda = (int64_t)kConvS[i] * 512; d.ldw = d.K; d.ldo = 512;
    d.batch = N; d.sA = (int64_t)T[i - 1] * 512; d.sO = (int64_t)T[i] * 512; d.wz_mod = 0;
    if (i == 6) { d.out = m->c6; d.out_f32 = 1; d.gelu_heavy = 0; } else { d.out = outb; d.out_f32 = 0; }
    if (N == 1) { d.batch = 1; }
    if ((rc = gemm_tc_launch(d, st))) return rc;
    std::swap(in, outb);
  }
  // truncate to round(frame_num * 50 / fps) frames, resample to frame_num (hubert.py:24-28), LayerNorm, projection
  const int keep = std::min(T[6], (int)std::lround((double)F * 50.0 / output_fps));
  if ((rc = interp_ln512(m->c6, N, T[6], keep, F, m->fp_g, m->fp_b, m->xin, st))) return rc;
  const int M = N * F;
  if ((rc = gemm2d(m->xin, 512, m->Wfp, 512, m->fp_bias, nullptr, 0, m->h0, 768, 1, M, 768, 512, 0, st))) return rc;
  // positional conv embedding + encoder LayerNorm
  if ((rc = pos_pack(m->h0, m->xg, N, F, st))) return rc;
  {
    GemmDesc d;
    d.mode = 0; d.A = m->xg; d.W = m->Wpos; d.bias = m->pos_bias; d.out = m->pos; d.out_f32 = 1;
    d.M = F; d.N = 48; d.K = 6144; d.lda = 48; d.ldw = 6144; d.ldo = 48;
    d.batch = N * 16; d.sA = (int64_t)(F + 128) * 48; d.sW = (int64_t)48 * 6144; d.sO = (int64_t)F * 48;
    d.wz_mod = 16; d.bias_zstride = 48;
    if ((rc = gemm_tc_launch(d, st))) return rc;
  }
  if ((rc = pos_add_ln768(m->h0, m->pos, m->enc_g, m->enc_b, m->x, N, F, st))) return rc;
  for (int l = 0; l < m->n_layers; ++l) {
    EncLayer& w = m->L[l];
    const bool last = l == m->n_layers - 1;
    if ((rc = gemm2d(m->x, 768, w.Wqkv, 768, w.bqkv, nullptr, 0, m->qkv, 2304, 0, M, 2304, 768, 0, st))) return rc;
    if ((rc = flash_attn_tc(m->qkv, m->ctx, N, F, m->n_heads, st))) return rc;
    if ((rc = gemm2d(m->ctx, 768, w.Wo, 768, w.bo, m->x, 768, m->y, 768, 1, M, 768, 768, 0, st))) return rc;
    if ((rc = ln768(m->y, w.g1, w.be1, m->x, nullptr, M, st))) return rc;
    if ((rc = gemm2d(m->x, 768, w.W1, 768, w.b1, nullptr, 0, m->hff, 3072, 0, M, 3072, 768, 1, st))) return rc;
    if ((rc = gemm2d(m->hff, 3072, w.W2, 3072, w.b2, m->x, 768, m->y, 768, 1, M, 768, 3072, 0, st))) return rc;
    if ((rc = ln768(m->y, w.g2, w.be2, m->x, last ? m->hs : nullptr, M, st))) return rc;
  }
  if (hidden_out)
    MSMD_CHECK_CUDA(cudaMemcpyAsync(hidden_out, m->hs, (size_t)M * 768 * 4, cudaMemcpyDeviceToDevice, st));
  if (feat_out) {  // model.py:259-263: resample 2L -> L, then audio_feature_map
    if ((rc = interp768_bf16(m->hs, N, F, feat_frames, m->xl, st))) return rc;
    if ((rc = gemm2d(m->xl, 768, m->Wmap, 768, m->map_b, nullptr, 0, feat_out, m->d_out, 1, N * feat_frames, m->d_out, 768,
                     0, st)))
      return rc;
  }
  return MSMD_OK;
}
