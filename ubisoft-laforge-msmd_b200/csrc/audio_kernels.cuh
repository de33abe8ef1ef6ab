// Non-GEMM kernels of the audio encoder (HF Hubert / Wav2Vec2 base as called by utils/hubert.py:13-51,
// utils/wav2vec2.py:71-119 and model.py:250-264): declarations.
#pragma once
#include "common.cuh"
#include <cuda_bf16.h>

namespace msmd {

using bf16 = __nv_bfloat16;

// pad_audio (model_common.py:110-123) as an index map: padded sample i -> source sample
struct PadSpec { int n, r, rep, n_pad; };
PadSpec make_pad_spec(int n_samples);

// conv layer 0 (1 -> 512, k=10, s=5, no bias) + GroupNorm(512 groups: per-channel over the whole time axis)
// + GELU -> bf16 channels-last [N, T0, 512].  stats: double [N,512,2] scratch (zeroed by the call).
int conv0_groupnorm_gelu(const float* wav, int N, PadSpec ps, int T0, const float* w0 /*[512,10]*/,
                         const float* gn_w, const float* gn_b, double* stats, bf16* out, cudaStream_t st);

// linear interpolation over time (F.interpolate linear, align_corners=False) of the first `in_len` frames of
// x [N, in_stride_rows, 512] fp32 to out_len frames, then LayerNorm(512) -> bf16 [N*out_len, 512]
int interp_ln512(const float* x, int N, int in_rows_per_clip, int in_len, int out_len, const float* g, const float* b,
                 bf16* out, cudaStream_t st);

// h0 fp32 [N, F, 768] -> group-major zero-padded bf16 [N, 16, F+128, 48] (pad 64 frames each side)
int pos_pack(const float* h0, bf16* xg, int N, int F, cudaStream_t st);
// x = LayerNorm768(h0 + gelu(pos)), pos fp32 [N, 16, F, 48] -> bf16 [N*F, 768]
int pos_add_ln768(const float* h0, const float* pos, const float* g, const float* b, bf16* out, int N, int F,
                  cudaStream_t st);
// LayerNorm768 of y fp32 [M,768] -> bf16 out (+ optional fp32 copy)
int ln768(const float* y, const float* g, const float* b, bf16* out, float* out_f32, int M, cudaStream_t st);

// flash-style attention, head dim 64: qkv [N*T, 3*d] bf16 (q|k|v, q pre-scaled) -> ctx [N*T, d]
int flash_attn_tc(const bf16* qkv, bf16* ctx, int N, int T, int H, cudaStream_t st);    // tcgen05 (flash_attn_tc.cu)

// hs fp32 [N, F, 768] -> linear interpolation to L frames -> bf16 [N*L, 768]
int interp768_bf16(const float* hs, int N, int F, int L, bf16* out, cudaStream_t st);

}  // namespace msmd
