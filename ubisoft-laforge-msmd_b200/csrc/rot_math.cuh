// Branch-light fp32 rotation math shared by rotconv.cu and flame.cu.
// Formulas follow /root/reference/utils/rotation_conversions.py and utils/lbs.py:270-301
// (cited per function); quirks are kept on purpose.
#pragma once
#include <cuda_runtime.h>

namespace msmd {

struct Mat3 { float m[9]; };   // row-major
struct Quat { float w, x, y, z; };
struct Vec3 { float x, y, z; };

__device__ __forceinline__ Mat3 mat3_mul(const Mat3& a, const Mat3& b) {
  Mat3 c;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      c.m[i * 3 + j] = a.m[i * 3 + 0] * b.m[0 * 3 + j] + a.m[i * 3 + 1] * b.m[1 * 3 + j] +
                       a.m[i * 3 + 2] * b.m[2 * 3 + j];
  return c;
}

// rotation_conversions.py:38-67
__device__ __forceinline__ Mat3 quat_to_matrix(const Quat& q) {
  const float r = q.w, i = q.x, j = q.y, k = q.z;
  const float two_s = __fdividef(2.0f, r * r + i * i + j * j + k * k);
  Mat3 o;
  o.m[0] = 1.0f - two_s * (j * j + k * k);
  o.m[1] = two_s * (i * j - k * r);
  o.m[2] = two_s * (i * k + j * r);
  o.m[3] = two_s * (i * j + k * r);
  o.m[4] = 1.0f - two_s * (i * i + k * k);
  o.m[5] = two_s * (j * k - i * r);
  o.m[6] = two_s * (i * k - j * r);
  o.m[7] = two_s * (j * k + i * r);
  o.m[8] = 1.0f - two_s * (i * i + j * j);
  return o;
}

// rotation_conversions.py:70-97
__device__ __forceinline__ float copysign_ref(float a, float b) { return ((a < 0.f) != (b < 0.f)) ? -a : a; }
// sqrt(max(0, x)) as x * rsqrt(x): 2 instructions, <= 2 ulp (sqrtf's IEEE fix-up and its denormal slow path are ~10)
__device__ __forceinline__ float sqrt_pos(float x) { return x > 0.f ? x * rsqrtf(x) : 0.f; }

// rotation_conversions.py:100-120 (old sqrt/copysign formula; lossy near 180 degrees by design)
__device__ __forceinline__ Quat matrix_to_quat(const Mat3& a) {
  const float m00 = a.m[0], m11 = a.m[4], m22 = a.m[8];
  Quat q;
  q.w = 0.5f * sqrt_pos(1.0f + m00 + m11 + m22);
  const float x = 0.5f * sqrt_pos(1.0f + m00 - m11 - m22);
  const float y = 0.5f * sqrt_pos(1.0f - m00 + m11 - m22);
  const float z = 0.5f * sqrt_pos(1.0f - m00 - m11 + m22);
  q.x = copysign_ref(x, a.m[7] - a.m[5]);
  q.y = copysign_ref(y, a.m[2] - a.m[6]);
  q.z = copysign_ref(z, a.m[3] - a.m[1]);
  return q;
}

// sin and cos to ~1 ulp in ~20 instructions: Cody-Waite reduction by pi/2 (two-term, FMA) + the cephes single-precision
// minimax polynomials on [-pi/4, pi/4].  sincosf costs ~45 and made the Euler kernels ALU-bound (profiles/r01_rot_ncu.txt:
// 19% of HBM bandwidth at 84% SM busy).  |x| > 1000 takes the library path (the reduction loses bits there).
static __device__ __noinline__ void sincos_slow(float x, float* sn, float* cs) { sincosf(x, sn, cs); }
__device__ __forceinline__ void sincos_fast(float x, float* sn, float* cs) {
  if (fabsf(x) > 1000.0f) { sincos_slow(x, sn, cs); return; }
  const float k = rintf(x * 0.6366197723675814f);
  float r = fmaf(k, -1.5707963705062866f, x);
  r = fmaf(k, 4.371139000186241e-8f, r);
  const float r2 = r * r;
  float ps = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = fmaf(ps, r2, -1.6666654611e-1f);
  const float s0 = fmaf(r * r2, ps, r);
  float pc = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = fmaf(pc, r2, 4.166664568298827e-2f);
  const float c0 = fmaf(r2 * r2, pc, fmaf(r2, -0.5f, 1.0f));
  const int q = (int)k;
  const float s1 = (q & 1) ? c0 : s0, c1 = (q & 1) ? s0 : c0;
  *sn = (q & 2) ? -s1 : s1;
  *cs = ((q + 1) & 2) ? -c1 : c1;
}

// atan2(y, x) for y >= 0 (result in [0, pi]): one division + a degree-8 polynomial in a^2 (|error| <= 9e-8 on [0, 1],
// fitted in tools/; the half angle of a quaternion is the only caller, tolerance 3e-5)
__device__ __forceinline__ float atan2_pos(float y, float x) {
  const float ax = fabsf(x);
  const float hi = fmaxf(y, ax), lo = fminf(y, ax);
  const float a = hi > 0.f ? __fdividef(lo, hi) : 0.f;
  const float t = a * a;
  float p = fmaf(t, 0.0029327620286494493f, -0.016413189470767975f);
  p = fmaf(p, t, 0.04327824339270592f);
  p = fmaf(p, t, -0.07556900382041931f);
  p = fmaf(p, t, 0.10667487233877182f);
  p = fmaf(p, t, -0.14211106300354004f);
  p = fmaf(p, t, 0.19993694126605988f);
  p = fmaf(p, t, -0.3333313763141632f);
  float r = fmaf(a * t, p, a);
  r = (y > ax) ? 1.5707963267948966f - r : r;
  return (x < 0.f) ? 3.14159265358979f - r : r;
}

// rotation_conversions.py:123-148; axis 0/1/2 = X/Y/Z
__device__ __forceinline__ Mat3 axis_rotation(int axis, float ang) {
  float s, c;
  sincos_fast(ang, &s, &c);
  Mat3 r;
  if (axis == 0) {
    r = {{1.f, 0.f, 0.f, 0.f, c, -s, 0.f, s, c}};
  } else if (axis == 1) {
    r = {{c, 0.f, s, 0.f, 1.f, 0.f, -s, 0.f, c}};
  } else {
    r = {{c, -s, 0.f, s, c, 0.f, 0.f, 0.f, 1.f}};
  }
  return r;
}

// rotation_conversions.py:151-173: (R_a(e0) @ R_b(e1)) @ R_c(e2); conv = a*9+b*3+c
__device__ __forceinline__ Mat3 euler_to_matrix(float e0, float e1, float e2, int conv) {
  const int a = conv / 9, b = (conv / 3) % 3, c = conv % 3;
  return mat3_mul(mat3_mul(axis_rotation(a, e0), axis_rotation(b, e1)), axis_rotation(c, e2));
}

// M <- M @ R_axis(angle) with the axis known at compile time: only two columns change (the zeros / ones of
// rotation_conversions.py:123-148 folded by hand - the generic 3x3 products were 54 multiplies + runtime branches)
template <int AXIS>
__device__ __forceinline__ void mul_axis_right(Mat3& M, float s, float c) {
  constexpr int A = AXIS == 0 ? 1 : 0, B = AXIS == 2 ? 1 : 2;      // the two columns that mix
  // X: colA' = colA c + colB s, colB' = -colA s + colB c;  Y (A=0,B=2): colA' = colA c - colB s, colB' = colA s + colB c;
  // Z (A=0,B=1): colA' = colA c + colB s, colB' = -colA s + colB c
  constexpr float sg = AXIS == 1 ? -1.0f : 1.0f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float a = M.m[r * 3 + A], b = M.m[r * 3 + B];
    M.m[r * 3 + A] = fmaf(a, c, sg * b * s);
    M.m[r * 3 + B] = fmaf(b, c, -sg * a * s);
  }
}
template <int CONV>
__device__ __forceinline__ Mat3 euler_to_matrix_c(float e0, float e1, float e2) {
  constexpr int a = CONV / 9, b = (CONV / 3) % 3, c = CONV % 3;
  float s0, c0, s1, c1, s2, c2;
  sincos_fast(e0, &s0, &c0);
  sincos_fast(e1, &s1, &c1);
  sincos_fast(e2, &s2, &c2);
  Mat3 M = {{1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f}};
  mul_axis_right<a>(M, s0, c0);
  mul_axis_right<b>(M, s1, c1);
  mul_axis_right<c>(M, s2, c2);
  return M;
}

// rotation_conversions.py:176-207. data = 3 values (a row or a column of R).
__device__ __forceinline__ float angle_from_tan(int axis, int other, const float* d, bool horizontal,
                                                bool tait_bryan) {
  int i1 = (axis == 0) ? 2 : (axis == 1 ? 0 : 1);
  int i2 = (axis == 0) ? 1 : (axis == 1 ? 2 : 0);
  if (horizontal) { int t = i1; i1 = i2; i2 = t; }
  const bool even = (axis == 0 && other == 1) || (axis == 1 && other == 2) || (axis == 2 && other == 0);
  if (horizontal == even) return atan2f(d[i1], d[i2]);
  if (tait_bryan) return atan2f(-d[i2], d[i1]);
  return atan2f(d[i2], -d[i1]);
}

// rotation_conversions.py:219-257
__device__ __forceinline__ Vec3 matrix_to_euler(const Mat3& r, int conv) {
  const int c0 = conv / 9, c1 = (conv / 3) % 3, c2 = conv % 3;
  const int i0 = c0, i2 = c2;
  const bool tb = i0 != i2;
  float central;
  if (tb) {
    const int d = i0 - i2;
    const float sgn = (d == -1 || d == 2) ? -1.0f : 1.0f;
    central = asinf(r.m[i0 * 3 + i2] * sgn);
  } else {
    central = acosf(r.m[i0 * 3 + i0]);
  }
  float col[3] = {r.m[0 * 3 + i2], r.m[1 * 3 + i2], r.m[2 * 3 + i2]};  // matrix[..., i2]
  float row[3] = {r.m[i0 * 3 + 0], r.m[i0 * 3 + 1], r.m[i0 * 3 + 2]};  // matrix[..., i0, :]
  Vec3 o;
  o.x = angle_from_tan(c0, c1, col, false, tb);
  o.y = central;
  o.z = angle_from_tan(c2, c1, row, true, tb);
  return o;
}

template <int CONV>
__device__ __forceinline__ Vec3 matrix_to_euler_c(const Mat3& r) { return matrix_to_euler(r, CONV); }

// sin(a/2)/a with the |a|<1e-6 series branch (rotation_conversions.py:462-475, :498-509)
__device__ __forceinline__ float half_sinc(float ang, float half) {
  const bool small = fabsf(ang) < 1e-6f;
  const float big = sinf(half) / (small ? 1.0f : ang);
  const float ser = 0.5f - (ang * ang) / 48.0f;
  return small ? ser : big;
}

// rotation_conversions.py:450-478
__device__ __forceinline__ Quat aa_to_quat(const Vec3& a) {
  const float ang = sqrt_pos(a.x * a.x + a.y * a.y + a.z * a.z);
  const float half = 0.5f * ang;
  float sh, ch;
  sincos_fast(half, &sh, &ch);
  const bool small = fabsf(ang) < 1e-6f;
  const float k = small ? 0.5f - (ang * ang) / 48.0f : __fdividef(sh, ang);       // sin(a/2)/a (:462-475)
  Quat q = {ch, a.x * k, a.y * k, a.z * k};
  return q;
}

// rotation_conversions.py:481-510
__device__ __forceinline__ Vec3 quat_to_aa(const Quat& q) {
  const float n = sqrt_pos(q.x * q.x + q.y * q.y + q.z * q.z);
  const float half = atan2_pos(n, q.w);
  const float ang = 2.0f * half;
  float sh, ch;
  sincos_fast(half, &sh, &ch);
  const bool small = fabsf(ang) < 1e-6f;
  // q / (sin(a/2)/a) as ONE division: inv = a / sin(a/2)  (series branch: 1 / (1/2 - a^2/48), :498-509)
  const float inv = small ? __fdividef(1.0f, 0.5f - (ang * ang) / 48.0f) : __fdividef(ang, sh);
  Vec3 v = {q.x * inv, q.y * inv, q.z * inv};
  return v;
}

// rotation_conversions.py:513-535 (F.normalize eps 1e-12)
__device__ __forceinline__ Mat3 rot6d_to_matrix(const float* d6) {
  const float a1x = d6[0], a1y = d6[1], a1z = d6[2], a2x = d6[3], a2y = d6[4], a2z = d6[5];
  float n1 = fmaxf(sqrtf(a1x * a1x + a1y * a1y + a1z * a1z), 1e-12f);
  const float b1x = a1x / n1, b1y = a1y / n1, b1z = a1z / n1;
  const float dp = b1x * a2x + b1y * a2y + b1z * a2z;
  float ux = a2x - dp * b1x, uy = a2y - dp * b1y, uz = a2z - dp * b1z;
  float n2 = fmaxf(sqrtf(ux * ux + uy * uy + uz * uz), 1e-12f);
  ux /= n2; uy /= n2; uz /= n2;
  Mat3 r = {{b1x, b1y, b1z, ux, uy, uz,
             b1y * uz - b1z * uy, b1z * ux - b1x * uz, b1x * uy - b1y * ux}};
  return r;
}

// utils/lbs.py:270-301 batch_rodrigues: angle = ||r + 1e-8|| (quirk, lbs.py:285)
__device__ __forceinline__ Mat3 rodrigues(float rx, float ry, float rz) {
  const float ex = rx + 1e-8f, ey = ry + 1e-8f, ez = rz + 1e-8f;
  const float angle = sqrtf(ex * ex + ey * ey + ez * ez);
  const float x = rx / angle, y = ry / angle, z = rz / angle;
  float s, c;
  sincos_fast(angle, &s, &c);
  const Mat3 K = {{0.f, -z, y, z, 0.f, -x, -y, x, 0.f}};
  const Mat3 KK = mat3_mul(K, K);
  const float omc = 1.0f - c;
  Mat3 r;
#pragma unroll
  for (int i = 0; i < 9; ++i) r.m[i] = ((i % 4 == 0) ? 1.0f : 0.0f) + s * K.m[i] + omc * KK.m[i];
  return r;
}

// rotation_conversions.py:341-359
__device__ __forceinline__ Quat quat_raw_mul(const Quat& a, const Quat& b) {
  Quat o;
  o.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  o.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  o.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
  o.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
  return o;
}

}  // namespace msmd
