// Rotation-conversion kernels (replaces utils/rotation_conversions.py:38-569 and
// utils/lbs.py:270-301).  HBM-bound elementwise work: one rotation per thread, packed AoS
// I/O staged through shared memory so every global access is a coalesced float4.
#include "common.cuh"
#include "rot_math.cuh"

namespace msmd {

constexpr int kRotBlock = 256;

template <int KIND> struct RotTraits;
#define ROT_TRAITS(K, I, O) template <> struct RotTraits<K> { static constexpr int IN = I, OUT = O; };
ROT_TRAITS(MSMD_ROT_QUAT_TO_MATRIX, 4, 9)
ROT_TRAITS(MSMD_ROT_MATRIX_TO_QUAT, 9, 4)
ROT_TRAITS(MSMD_ROT_EULER_TO_MATRIX, 3, 9)
ROT_TRAITS(MSMD_ROT_MATRIX_TO_EULER, 9, 3)
ROT_TRAITS(MSMD_ROT_AA_TO_QUAT, 3, 4)
ROT_TRAITS(MSMD_ROT_QUAT_TO_AA, 4, 3)
ROT_TRAITS(MSMD_ROT_AA_TO_MATRIX, 3, 9)
ROT_TRAITS(MSMD_ROT_MATRIX_TO_AA, 9, 3)
ROT_TRAITS(MSMD_ROT_6D_TO_MATRIX, 6, 9)
ROT_TRAITS(MSMD_ROT_MATRIX_TO_6D, 9, 6)
ROT_TRAITS(MSMD_ROT_AA_TO_6D, 3, 6)
ROT_TRAITS(MSMD_ROT_STANDARDIZE_QUAT, 4, 4)
ROT_TRAITS(MSMD_ROT_QUAT_INVERT, 4, 4)
ROT_TRAITS(MSMD_ROT_EULER_TO_AA, 3, 3)
ROT_TRAITS(MSMD_ROT_RODRIGUES, 3, 9)

__device__ __forceinline__ Mat3 ld_mat(const float* p) {
  Mat3 m;
#pragma unroll
  for (int i = 0; i < 9; ++i) m.m[i] = p[i];
  return m;
}
__device__ __forceinline__ void st_mat(float* p, const Mat3& m) {
#pragma unroll
  for (int i = 0; i < 9; ++i) p[i] = m.m[i];
}

// CONV: Euler convention code (a*9 + b*3 + c) as a compile-time constant for the three Euler kinds (every axis test and
// matrix index folds away), 0 for the others
template <int KIND, int CONV>
__device__ __forceinline__ void rot_one(const float* i, float* o) {
  if constexpr (KIND == MSMD_ROT_QUAT_TO_MATRIX) {
    st_mat(o, quat_to_matrix(Quat{i[0], i[1], i[2], i[3]}));
  } else if constexpr (KIND == MSMD_ROT_MATRIX_TO_QUAT) {
    Quat q = matrix_to_quat(ld_mat(i));
    o[0] = q.w; o[1] = q.x; o[2] = q.y; o[3] = q.z;
  } else if constexpr (KIND == MSMD_ROT_EULER_TO_MATRIX) {
    st_mat(o, euler_to_matrix_c<CONV>(i[0], i[1], i[2]));
  } else if constexpr (KIND == MSMD_ROT_MATRIX_TO_EULER) {
    Vec3 e = matrix_to_euler_c<CONV>(ld_mat(i));
    o[0] = e.x; o[1] = e.y; o[2] = e.z;
  } else if constexpr (KIND == MSMD_ROT_AA_TO_QUAT) {
    Quat q = aa_to_quat(Vec3{i[0], i[1], i[2]});
    o[0] = q.w; o[1] = q.x; o[2] = q.y; o[3] = q.z;
  } else if constexpr (KIND == MSMD_ROT_QUAT_TO_AA) {
    Vec3 a = quat_to_aa(Quat{i[0], i[1], i[2], i[3]});
    o[0] = a.x; o[1] = a.y; o[2] = a.z;
  } else if constexpr (KIND == MSMD_ROT_AA_TO_MATRIX) {
    st_mat(o, quat_to_matrix(aa_to_quat(Vec3{i[0], i[1], i[2]})));
  } else if constexpr (KIND == MSMD_ROT_MATRIX_TO_AA) {
    Vec3 a = quat_to_aa(matrix_to_quat(ld_mat(i)));
    o[0] = a.x; o[1] = a.y; o[2] = a.z;
  } else if constexpr (KIND == MSMD_ROT_6D_TO_MATRIX) {
    st_mat(o, rot6d_to_matrix(i));
  } else if constexpr (KIND == MSMD_ROT_MATRIX_TO_6D) {
#pragma unroll
    for (int k = 0; k < 6; ++k) o[k] = i[k];
  } else if constexpr (KIND == MSMD_ROT_AA_TO_6D) {
    Mat3 m = quat_to_matrix(aa_to_quat(Vec3{i[0], i[1], i[2]}));
#pragma unroll
    for (int k = 0; k < 6; ++k) o[k] = m.m[k];
  } else if constexpr (KIND == MSMD_ROT_STANDARDIZE_QUAT) {
    const float s = (i[0] < 0.f) ? -1.f : 1.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = (s < 0.f) ? -i[k] : i[k];
  } else if constexpr (KIND == MSMD_ROT_QUAT_INVERT) {
    o[0] = i[0]; o[1] = -i[1]; o[2] = -i[2]; o[3] = -i[3];
  } else if constexpr (KIND == MSMD_ROT_EULER_TO_AA) {
    Vec3 a = quat_to_aa(matrix_to_quat(euler_to_matrix_c<CONV>(i[0], i[1], i[2])));
    o[0] = a.x; o[1] = a.y; o[2] = a.z;
  } else if constexpr (KIND == MSMD_ROT_RODRIGUES) {
    st_mat(o, rodrigues(i[0], i[1], i[2]));
  }
}

// Cooperative contiguous copy global<->shared, float4 when the block is full (base 16B aligned).
__device__ __forceinline__ void block_load(float* s, const float* g, int count, bool vec) {
  if (vec) {
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* s4 = reinterpret_cast<float4*>(s);
    for (int i = threadIdx.x; i < count / 4; i += blockDim.x) s4[i] = __ldg(g4 + i);
  } else {
    for (int i = threadIdx.x; i < count; i += blockDim.x) s[i] = g[i];
  }
}
__device__ __forceinline__ void block_store(float* g, const float* s, int count, bool vec) {
  if (vec) {
    float4* g4 = reinterpret_cast<float4*>(g);
    const float4* s4 = reinterpret_cast<const float4*>(s);
    for (int i = threadIdx.x; i < count / 4; i += blockDim.x) __stcs(g4 + i, s4[i]);
  } else {
    for (int i = threadIdx.x; i < count; i += blockDim.x) g[i] = s[i];
  }
}

template <int KIND, int CONV>
__global__ void __launch_bounds__(kRotBlock) rot_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                        int64_t n, bool aligned) {
  constexpr int IN = RotTraits<KIND>::IN, OUT = RotTraits<KIND>::OUT;
  __shared__ __align__(16) float s_in[kRotBlock * IN];
  __shared__ __align__(16) float s_out[kRotBlock * OUT];
  for (int64_t base = (int64_t)blockIdx.x * kRotBlock; base < n; base += (int64_t)gridDim.x * kRotBlock) {
    const int cnt = (int)min((int64_t)kRotBlock, n - base);
    const bool vec = aligned && cnt == kRotBlock;
    block_load(s_in, in + base * IN, cnt * IN, vec);
    __syncthreads();
    if (threadIdx.x < cnt) {
      float li[IN], lo[OUT];
#pragma unroll
      for (int k = 0; k < IN; ++k) li[k] = s_in[threadIdx.x * IN + k];
      rot_one<KIND, CONV>(li, lo);
#pragma unroll
      for (int k = 0; k < OUT; ++k) s_out[threadIdx.x * OUT + k] = lo[k];
    }
    __syncthreads();
    block_store(out + base * OUT, s_out, cnt * OUT, vec);
    __syncthreads();
  }
}

template <int OP>
__global__ void __launch_bounds__(kRotBlock) quat_binary_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                float* __restrict__ out, int64_t n) {
  constexpr int BW = (OP == MSMD_QUAT_APPLY) ? 3 : 4;
  constexpr int OW = (OP == MSMD_QUAT_APPLY) ? 3 : 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 qa = *reinterpret_cast<const float4*>(a + i * 4);
    const Quat A = {qa.x, qa.y, qa.z, qa.w};
    Quat r;
    if constexpr (OP == MSMD_QUAT_APPLY) {
      // rotation_conversions.py:396-415: (q * (0,p)) * conj(q), vector part
      const Quat P = {0.f, b[i * BW + 0], b[i * BW + 1], b[i * BW + 2]};
      const Quat inv = {A.w, -A.x, -A.y, -A.z};
      r = quat_raw_mul(quat_raw_mul(A, P), inv);
      out[i * OW + 0] = r.x; out[i * OW + 1] = r.y; out[i * OW + 2] = r.z;
    } else {
      const float4 qb = *reinterpret_cast<const float4*>(b + i * 4);
      r = quat_raw_mul(A, Quat{qb.x, qb.y, qb.z, qb.w});
      if constexpr (OP == MSMD_QUAT_MULTIPLY) {  // standardize: real part >= 0 (:326-338)
        if (r.w < 0.f) { r.w = -r.w; r.x = -r.x; r.y = -r.y; r.z = -r.z; }
      }
      *reinterpret_cast<float4*>(out + i * 4) = make_float4(r.w, r.x, r.y, r.z);
    }
  }
}

template <int KIND, int CONV>
static int launch_rot_c(const float* in, float* out, int64_t n, cudaStream_t st) {
  const int blocks = (int)std::min<int64_t>((n + kRotBlock - 1) / kRotBlock, (int64_t)kNumSMs * 16);
  const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  rot_kernel<KIND, CONV><<<blocks, kRotBlock, 0, st>>>(in, out, n, aligned);
  MSMD_CHECK_LAUNCH();
  return MSMD_OK;
}
template <int KIND>
static int launch_rot(const float* in, float* out, int64_t n, int conv, cudaStream_t st) {
  constexpr bool euler = KIND == MSMD_ROT_EULER_TO_MATRIX || KIND == MSMD_ROT_MATRIX_TO_EULER || KIND == MSMD_ROT_EULER_TO_AA;
  if constexpr (!euler) {
    return launch_rot_c<KIND, 0>(in, out, n, st);
  } else {
    switch (conv) {     // the 12 valid conventions (middle axis differs from both neighbours)
#define CV(C) case C: return launch_rot_c<KIND, C>(in, out, n, st);
      CV(3) CV(5) CV(6) CV(7) CV(10) CV(11) CV(15) CV(16) CV(19) CV(20) CV(21) CV(23)
#undef CV
      default:
        set_error("Invalid convention code %d.", conv);
        return MSMD_ERR_INVALID;
    }
  }
}

}  // namespace msmd

using namespace msmd;

extern "C" int msmd_rot_convert(int kind, const float* in, float* out, int64_t n, int convention, void* stream) {
  MSMD_REQUIRE(n >= 0, "msmd_rot_convert: negative count");
  if (n == 0) return MSMD_OK;
  MSMD_REQUIRE(in && out, "msmd_rot_convert: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool euler = kind == MSMD_ROT_EULER_TO_MATRIX || kind == MSMD_ROT_MATRIX_TO_EULER || kind == MSMD_ROT_EULER_TO_AA;
  if (euler) {
    const int a = convention / 9, b = (convention / 3) % 3, c = convention % 3;
    MSMD_REQUIRE(convention >= 0 && convention < 27 && b != a && b != c, "Invalid convention code %d.", convention);
  }
  switch (kind) {
#define CASE(K) case K: return launch_rot<K>(in, out, n, convention, st);
    CASE(MSMD_ROT_QUAT_TO_MATRIX) CASE(MSMD_ROT_MATRIX_TO_QUAT) CASE(MSMD_ROT_EULER_TO_MATRIX)
    CASE(MSMD_ROT_MATRIX_TO_EULER) CASE(MSMD_ROT_AA_TO_QUAT) CASE(MSMD_ROT_QUAT_TO_AA)
    CASE(MSMD_ROT_AA_TO_MATRIX) CASE(MSMD_ROT_MATRIX_TO_AA) CASE(MSMD_ROT_6D_TO_MATRIX)
    CASE(MSMD_ROT_MATRIX_TO_6D) CASE(MSMD_ROT_AA_TO_6D) CASE(MSMD_ROT_STANDARDIZE_QUAT)
    CASE(MSMD_ROT_QUAT_INVERT) CASE(MSMD_ROT_EULER_TO_AA) CASE(MSMD_ROT_RODRIGUES)
#undef CASE
    default:
      set_error("msmd_rot_convert: unknown kind %d", kind);
      return MSMD_ERR_INVALID;
  }
}

extern "C" int msmd_quat_binary(int op, const float* a, const float* b, float* out, int64_t n, void* stream) {
  MSMD_REQUIRE(n >= 0, "msmd_quat_binary: negative count");
  if (n == 0) return MSMD_OK;
  MSMD_REQUIRE(a && b && out, "msmd_quat_binary: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = (int)std::min<int64_t>((n + kRotBlock - 1) / kRotBlock, (int64_t)kNumSMs * 16);
  switch (op) {
    case MSMD_QUAT_RAW_MULTIPLY: quat_binary_kernel<MSMD_QUAT_RAW_MULTIPLY><<<blocks, kRotBlock, 0, st>>>(a, b, out, n); break;
    case MSMD_QUAT_MULTIPLY: quat_binary_kernel<MSMD_QUAT_MULTIPLY><<<blocks, kRotBlock, 0, st>>>(a, b, out, n); break;
    case MSMD_QUAT_APPLY: quat_binary_kernel<MSMD_QUAT_APPLY><<<blocks, kRotBlock, 0, st>>>(a, b, out, n); break;
    default: set_error("msmd_quat_binary: unknown op %d", op); return MSMD_ERR_INVALID;
  }
  MSMD_CHECK_LAUNCH();
  return MSMD_OK;
}
