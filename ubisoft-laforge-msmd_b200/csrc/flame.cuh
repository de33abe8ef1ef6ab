// FLAME decode handle shared by flame.cu (pack, pose kernel, CUDA-core path) and
// flame_tc.cu (tcgen05 path).  Replaces utils/flame.py:180-244 + utils/lbs.py:141-371.
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>

struct msmd_flame {
  int device = 0;
  int V = 0, NB = 0, NJ = 5;
  int N3 = 0;      // 3*V output columns (vertex-major, xyz inner)
  int K = 0;       // NB + 9*(NJ-1): blendshape + pose-corrective depth
  int Kpad = 0;    // K rounded up to 64 (one 128-byte fp16 k-block of the tcgen05 path)
  int N3pad = 0;   // N3 rounded up to 384 rows so every tile load is in-bounds
  int parents[8] = {-1, 0, 1, 1, 1, 0, 0, 0};
  // static device buffers
  float* basis = nullptr;     // [N3pad, Kpad] K-major: row n=(v,c): shapedirs[v,c,:] | posedirs[:,n] | 0
  // fp16 two-term split for the 3-pass tensor-core path (flame_tc.cu): s x = hi + lo with hi = fp16(s x),
  // lo = fp16(s x - hi): 22 mantissa bits.  hi and lo share ONE scale (s = kFlameScaleB for the basis, kFlameScaleA for
  // the coefficients) so that hi*hi, lo*hi and hi*lo can be summed in the same TMEM accumulator; the scales keep the
  // residuals of the small basis entries (1e-3 m) inside the fp16 normal range.  The epilogue multiplies by 1 / (sA sB).
  __half* basis_hi = nullptr;
  __half* basis_lo = nullptr;
  float* v_template = nullptr;  // [N3]
  float* weights = nullptr;     // [V, NJ]
  float* vconst = nullptr;      // [V, 8] per-vertex epilogue constants of the tensor-core path: w0..w4 | template xyz
  int weights_normalised = 0;   // every row of the skinning weights sums to 1 (+-1e-5): the blend may use joint differences
  float* Jt = nullptr;          // [NJ*3]      J_regressor @ v_template
  float* Jb = nullptr;          // [NJ*3, NB]  J_regressor @ shapedirs (joint regression folded)
  __half* Jb16 = nullptr;       // [2][16][KB] fp16 two-term split of kFlameScaleJ * Jb (rows >= NJ*3 and columns >= NB zero): the A
  int KB = 0;                   // operand of the tensor-core joint regression (flame_pose_mma_kernel); KB = NB rounded up to 16
  int* d_parents = nullptr;     // [8]
  // per-call workspaces (grown on demand)
  int64_t cap_B = 0;
  float* A = nullptr;     // [cap_B, Kpad]  betas | pose_feature | 0
  __half* A_hi = nullptr;  // same split of A
  __half* A_lo = nullptr;
  float* xf = nullptr;    // [cap_B, NJ, 12] skinning affines: G (9) | t - G j (3)
  void* tc_state = nullptr;  // tensor maps etc. owned by flame_tc.cu
};

namespace msmd {
constexpr float kFlameScaleA = 16.0f, kFlameScaleB = 256.0f, kFlameScaleJ = 1024.0f;
int flame_decode_tc(msmd_flame* fh, int64_t B, float* verts_out, cudaStream_t st);  // flame_tc.cu
void flame_tc_destroy(msmd_flame* fh);
}  // namespace msmd
