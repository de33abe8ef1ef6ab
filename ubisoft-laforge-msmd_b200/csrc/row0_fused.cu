// Person-token (row 0) cross-attention block of a decoder layer as ONE kernel (nn.TransformerDecoderLayer._mha_block +
// norm2 for the query row that sees the whole memory, model.py:879-883 / :951-958):
//
//     x[s, 0, :] = LayerNorm2( x0c[s] + Wco . attn_h( Wq0 . x0c[s] + bq0 ; K_s, V_s ) + bco )
//
// It replaces four launches (q-projection GEMM, warp-per-head attention, out-projection GEMM, LayerNorm) that cost
// 23 us per layer on the critical path for S rows of work (S = sequences, 192 at config 3; 3 in the batch-1 latency
// regime, where the three saved launches per layer are 20% of the step).
//
// One thread-block CLUSTER of 8 CTAs per group of <= 16 sequences, CTA rank = attention head h:
//   phase 0  stage this head's weight slices in shared memory once per CTA: Wq0 rows [64h, 64h+64) (64 KB) and the
//            K-slice Wco[:, 64h:64h+64) (64 KB); L2-prefetch the head's K / V cache lines of the group's sequences;
//   phase 1  q_h = Wq0_h . x0c + b      [64 x 512] x [512 x n]   register tiles 4 outputs x 4 sequences, CUDA cores
//   phase 2  one warp per sequence: scores over the Tk memory keys, softmax, ctx_h = P V   (HBM-bound: the K/V cache)
//   phase 3  partial out-projection po = Wco[:, h-slice] . ctx_h   [512 x 64] x [64 x n]
//   phase 4  distributed-shared-memory reduction of the 8 partial projections: CTA h sums ITS 64 output columns from
//            all 8 CTAs, adds residual + bias, and the LayerNorm statistics are exchanged through DSMEM as well
//            (two-pass: mean, then centred second moment); each CTA writes its 128-byte column block of the row.
// Three cluster barriers per group.  16-bit storage format (bf16 / fp16) is a template parameter like everywhere else.
#include "denoiser_kernels.cuh"
#include "profile.cuh"
#include <cuda_fp16.h>

namespace msmd {
namespace {

constexpr int kThreads = 512, kNB = 16, kD = 512, kDh = 64, kHeads = 8;
constexpr int kWqPitch = kD + 8;        // halfs per staged Wq0 row (1040 B: 16-byte aligned, bank-staggered)
constexpr int kWoPitch = kDh + 8;       // halfs per staged Wco row slice (144 B)
constexpr int kOffWq = 0;
constexpr int kOffWo = kOffWq + kDh * kWqPitch * 2;
constexpr int kOffXs = kOffWo + kD * kWoPitch * 2;
constexpr int kOffQs = kOffXs + kNB * kD * 2;
constexpr int kOffCs = kOffQs + kNB * kDh * 4;
constexpr int kOffPo = kOffCs + kNB * kDh * 4;
constexpr int kOffSt = kOffPo + kNB * kD * 4;
constexpr int kSmem = kOffSt + 2 * kHeads * kNB * 4;
constexpr int kMaxGroup = 7;            // sequences per cluster: their 16 x 66-float attention partials share the 32 KB po buffer
static_assert(kMaxGroup * 16 * 66 * 4 <= kNB * kD * 4, "attention partials must fit the partial-projection buffer");
static_assert(kSmem <= 227 * 1024, "row0_fused shared memory");

struct Row0Params {
  const bf16* x0c;     // [S, 512]   LayerNorm1 output of the person rows
  const bf16* Wq;      // [512, 512] cross-attention query projection
  const float* bq;
  const bf16* kv;      // [S, Tk, 1024] memory K | V projections (per-window cache)
  const bf16* Wo;      // [512, 512] cross-attention out-projection
  const float* bo;
  const float* g;      // norm2
  const float* be;
  bf16* x;             // [S, T, 512] residual stream: row 0 of every sequence is written
  int S, T, Tk;
  unsigned long long* trace;   // -DMSMD_ROW0_TRACE builds: clock64 stamps of CTA 0 at the phase boundaries
};

template <bool F16>
__device__ __forceinline__ float2 up2(uint32_t w) {
  if constexpr (F16) return __half22float2(*reinterpret_cast<const __half2*>(&w));
  else return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <bool F16>
__device__ __forceinline__ uint32_t pk2(float a, float b) {
  if constexpr (F16) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}
template <bool F16>
__device__ __forceinline__ void up8(const uint4& u, float* f) {
  const float2 a = up2<F16>(u.x), b = up2<F16>(u.y), c = up2<F16>(u.z), d = up2<F16>(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ void r0_stamp(unsigned long long* tr, int ev) {
#ifdef MSMD_ROW0_TRACE
  if (tr != nullptr && blockIdx.x == 0 && threadIdx.x == 0) tr[ev] = clock64();
#endif
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ uint32_t ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `p` (a shared-memory location of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t remote_addr(const void* p, uint32_t rank) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ float2 ld_remote_f2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_remote_f(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

template <bool F16>
__global__ void __cluster_dims__(kHeads, 1, 1) __launch_bounds__(kThreads, 1) row0_fused_kernel(const Row0Params p) {
  extern __shared__ __align__(16) uint8_t sm[];
  griddep_launch();
  r0_stamp(p.trace, 0);
  bf16* wq = reinterpret_cast<bf16*>(sm + kOffWq);
  bf16* wo = reinterpret_cast<bf16*>(sm + kOffWo);
  bf16* xs = reinterpret_cast<bf16*>(sm + kOffXs);
  float* qs = reinterpret_cast<float*>(sm + kOffQs);
  float* cs = reinterpret_cast<float*>(sm + kOffCs);
  float* po = reinterpret_cast<float*>(sm + kOffPo);
  float* st = reinterpret_cast<float*>(sm + kOffSt);   // [2][kHeads][kNB]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = (int)ctarank();
  const int cid = blockIdx.x / kHeads, ncl = gridDim.x / kHeads;
  const int s_begin = (int)(((int64_t)p.S * cid) / ncl), s_end = (int)(((int64_t)p.S * (cid + 1)) / ncl);

  // ---- phase 0a: weight slices, asynchronously (cp.async: all 16 chunks of a thread in flight at once; they are
  // only needed by phase 1 / phase 3).  Independent of the previous kernel's output.
#pragma unroll
  for (int u = 0; u < kDh * (kD / 8) / kThreads; ++u) {        // Wq0 rows 64h..64h+63, 64 x 16-byte chunks each
    const int i = tid + u * kThreads, r = i >> 6, c = i & 63;
    cp_async16(wq + r * kWqPitch + c * 8, p.Wq + (int64_t)(h * kDh + r) * kD + c * 8);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
#pragma unroll
  for (int u = 0; u < kD * (kDh / 8) / kThreads; ++u) {        // Wco[:, 64h..64h+63]: 512 rows x 8 chunks
    const int i = tid + u * kThreads, r = i >> 3, c = i & 7;
    cp_async16(wo + r * kWoPitch + c * 8, p.Wo + (int64_t)r * kD + h * kDh + c * 8);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  griddep_wait();

  for (int s0 = s_begin; s0 < s_end; s0 += kMaxGroup) {
    const int n = min(kMaxGroup, s_end - s0);
    // ---- phase 0b: x0c rows of the group (zero-padded to kNB) + L2 prefetch of this head's K / V lines
    for (int i = tid; i < kNB * (kD / 8); i += kThreads) {
      const int s = i >> 6, c = i & 63;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (s < n) v = *reinterpret_cast<const uint4*>(p.x0c + (int64_t)(s0 + s) * kD + c * 8);
      *reinterpret_cast<uint4*>(xs + s * kD + c * 8) = v;
    }
    for (int i = tid; i < n * p.Tk * 2; i += kThreads) {
      const int s = i / (p.Tk * 2), r = i % (p.Tk * 2);
      const bf16* line = p.kv + ((int64_t)(s0 + s) * p.Tk + (r >> 1)) * (2 * kD) + (r & 1) * kD + h * kDh;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
    }
    asm volatile("cp.async.wait_group 1;" ::: "memory");     // Wq0 slice has landed (Wco may still be in flight)
    __syncthreads();
    r0_stamp(p.trace, 1);

    // ---- phase 1: q_h[s][o] = bq[64h+o] + sum_k Wq0[64h+o][k] x0c[s][k]
    {
      const int kpart = tid & 7, tile = tid >> 3, og = tile & 15, sg = tile >> 4;
      if (sg * 4 < n) {       // (warp-uniform: a warp holds 4 tiles of one sequence group) padded groups do no work
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 2
      for (int c = 0; c < 8; ++c) {
        const int ch = kpart + 8 * c;          // 16-byte chunk of the K axis: the 8 lanes of a tile read 128 contiguous bytes
        float xv[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j) up8<F16>(*reinterpret_cast<const uint4*>(xs + (sg * 4 + j) * kD + ch * 8), xv[j]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float wv[8];
          up8<F16>(*reinterpret_cast<const uint4*>(wq + (og * 4 + i) * kWqPitch + ch * 8), wv);
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[i][j] = fmaf(wv[k], xv[j][k], acc[i][j]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v = acc[i][j];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          if (kpart == 0) qs[(sg * 4 + j) * kDh + og * 4 + i] = v + __ldg(p.bq + h * kDh + og * 4 + i);
        }
      }
    }
    __syncthreads();

    // ---- phase 2: attention of head h.  The Tk <= 128 keys are ALWAYS cut into 16 parts of 8 keys, warp w = part w of
    // every sequence of the group: each part yields a partial (max, sum, P V) and the parts are merged in a fixed
    // order below - the arithmetic of a sequence does not depend on how many sequences share the cluster, so a clip's
    // codes are bit-identical whatever batch it is sampled in.  (All 16 warps stay busy even for a single sequence:
    // the batch-1 latency regime.)  Lane j < 8 scores key 8w + j; every lane owns two of the 64 head dims for P V.
    // The partials live in the (not yet used) partial-projection buffer: n <= kMaxGroup sequences per group.
    r0_stamp(p.trace, 2);
    float* parts = po;                              // [n][16 parts][66]: m, l, o(64)
    {
      const int part = warp, base = part * 8;
      const int nk = max(0, min(8, p.Tk - base));   // keys of this part
      for (int seq = 0; seq < n; ++seq) {
        const bf16* kbase = p.kv + ((int64_t)(s0 + seq) * p.Tk + base) * (2 * kD) + h * kDh;
        float dot = -INFINITY;
        uint4 kr[8];
        if (lane < nk) {
          const uint4* kp = reinterpret_cast<const uint4*>(kbase + (int64_t)lane * (2 * kD));
#pragma unroll
          for (int i = 0; i < 8; ++i) kr[i] = __ldg(kp + i);
        }
        // V rows of the part's keys: one coalesced 128-byte warp load per key, issued before the scores are needed
        const uint32_t* vbase = reinterpret_cast<const uint32_t*>(kbase + kD) + lane;
        uint32_t vr[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) vr[u] = u < nk ? __ldg(vbase + (int64_t)u * kD) : 0u;
        if (lane < nk) {
          float a = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float kf[8];
            up8<F16>(kr[i], kf);
            const float4 qa = *reinterpret_cast<const float4*>(qs + seq * kDh + 8 * i);
            const float4 qb = *reinterpret_cast<const float4*>(qs + seq * kDh + 8 * i + 4);
            a = fmaf(qa.x, kf[0], a); a = fmaf(qa.y, kf[1], a); a = fmaf(qa.z, kf[2], a); a = fmaf(qa.w, kf[3], a);
            a = fmaf(qb.x, kf[4], a); a = fmaf(qb.y, kf[5], a); a = fmaf(qb.z, kf[6], a); a = fmaf(qb.w, kf[7], a);
          }
          dot = a * 0.125f;
        }
        const float m = wmax(dot);
        const float pj = (dot == -INFINITY) ? 0.f : __expf(dot - m);
        const float l = wsum(pj);
        float oa = 0.f, ob = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float pu = __shfl_sync(0xffffffffu, pj, u);
          const float2 f = up2<F16>(vr[u]);
          oa = fmaf(pu, f.x, oa);
          ob = fmaf(pu, f.y, ob);
        }
        float* pt = parts + (seq * 16 + part) * 66;
        if (lane == 0) { pt[0] = m; pt[1] = l; }
        *reinterpret_cast<float2*>(pt + 2 + 2 * lane) = make_float2(oa, ob);
      }
    }
    __syncthreads();
    if (warp < kNB) {      // merge the 16 parts of sequence `warp` (flash-style rescale); padded sequences give zeros
      float2 o = make_float2(0.f, 0.f);
      if (warp < n) {
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < 16; ++w) M = fmaxf(M, parts[(warp * 16 + w) * 66]);
        float l = 0.f;
#pragma unroll
        for (int w = 0; w < 16; ++w) {
          const float* pt = parts + (warp * 16 + w) * 66;
          const float sc_w = pt[1] > 0.f ? __expf(pt[0] - M) : 0.f;
          const float2 t = *reinterpret_cast<const float2*>(pt + 2 + 2 * lane);
          l = fmaf(pt[1], sc_w, l);
          o.x = fmaf(t.x, sc_w, o.x);
          o.y = fmaf(t.y, sc_w, o.y);
        }
        const float inv = 1.0f / l;
        o.x *= inv; o.y *= inv;
      }
      *reinterpret_cast<float2*>(cs + warp * kDh + 2 * lane) = o;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");     // Wco slice
    __syncthreads();
    r0_stamp(p.trace, 3);

    // ---- phase 3: partial out-projection of this head's K-slice: po[s][c] = sum_{k<64} Wco[c][64h+k] ctx_h[s][k]
    {
      const int cg = tid & 127, sg = tid >> 7;      // rows cg, cg+128, cg+256, cg+384 (consecutive lanes -> consecutive rows)
      if (sg * 4 < n) {
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 2
      for (int ch = 0; ch < 8; ++ch) {
        float cv[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 a = *reinterpret_cast<const float4*>(cs + (sg * 4 + j) * kDh + ch * 8);
          const float4 b = *reinterpret_cast<const float4*>(cs + (sg * 4 + j) * kDh + ch * 8 + 4);
          cv[j][0] = a.x; cv[j][1] = a.y; cv[j][2] = a.z; cv[j][3] = a.w;
          cv[j][4] = b.x; cv[j][5] = b.y; cv[j][6] = b.z; cv[j][7] = b.w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float wv[8];
          up8<F16>(*reinterpret_cast<const uint4*>(wo + (cg + 128 * i) * kWoPitch + ch * 8), wv);
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[i][j] = fmaf(wv[k], cv[j][k], acc[i][j]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) po[(sg * 4 + j) * kD + cg + 128 * i] = acc[i][j];
      }
    }
    r0_stamp(p.trace, 4);
    cluster_barrier();     // every CTA's partial projection is complete and visible cluster-wide
    r0_stamp(p.trace, 5);

    // ---- phase 4: CTA h reduces ITS 64 output columns over the 8 partials (DSMEM), + residual + bias, LayerNorm
    {
      const int s = warp;                       // warp = sequence of the group (kNB = 16 warps)
      const int c = h * kDh + 2 * lane;         // this lane's two columns of the row
      float v0 = 0.f, v1 = 0.f;
#pragma unroll
      for (int r = 0; r < kHeads; ++r) {
        const float2 t = ld_remote_f2(remote_addr(po + s * kD + c, r));
        v0 += t.x; v1 += t.y;
      }
      const float2 xr = up2<F16>(*reinterpret_cast<const uint32_t*>(xs + s * kD + c));
      v0 += xr.x + __ldg(p.bo + c);
      v1 += xr.y + __ldg(p.bo + c + 1);
      const float part = wsum(v0 + v1);
      if (lane < kHeads) st_remote_f(remote_addr(st + (0 * kHeads + h) * kNB + s, lane), part);
      cluster_barrier();
      float mean = 0.f;
#pragma unroll
      for (int r = 0; r < kHeads; ++r) mean += st[(0 * kHeads + r) * kNB + s];
      mean *= (1.0f / kD);
      const float d0 = v0 - mean, d1 = v1 - mean;
      const float part2 = wsum(d0 * d0 + d1 * d1);
      if (lane < kHeads) st_remote_f(remote_addr(st + (1 * kHeads + h) * kNB + s, lane), part2);
      cluster_barrier();
      float var = 0.f;
#pragma unroll
      for (int r = 0; r < kHeads; ++r) var += st[(1 * kHeads + r) * kNB + s];
      const float rstd = rsqrtf(var * (1.0f / kD) + 1e-5f);
      if (s < n) {
        const float2 gg = __ldg(reinterpret_cast<const float2*>(p.g + c)), bb = __ldg(reinterpret_cast<const float2*>(p.be + c));
        *reinterpret_cast<uint32_t*>(p.x + (int64_t)(s0 + s) * p.T * kD + c) =
            pk2<F16>(d0 * rstd * gg.x + bb.x, d1 * rstd * gg.y + bb.y);
      }
    }
    r0_stamp(p.trace, 6);
    if (s0 + kMaxGroup < s_end) cluster_barrier();   // the next group overwrites po / st: every remote read of this one is done
  }
}

}  // namespace

int row0_fused_launch(const bf16* x0c, const bf16* Wq, const float* bq, const bf16* kv, const bf16* Wo, const float* bo,
                      const float* g, const float* be, bf16* x, int S, int T, int Tk, int H, int d, int fp16, cudaStream_t st) {
  MSMD_REQUIRE(d == kD && H == kHeads && Tk >= 1 && Tk <= 128,
               "row0_fused: built for d_model 512, 8 heads x 64 and <= 128 memory tokens (got %d, %d, %d)", d, H, Tk);
  Row0Params p{x0c, Wq, bq, kv, Wo, bo, g, be, x, S, T, Tk, nullptr};
#ifdef MSMD_ROW0_TRACE
  static unsigned long long* tbuf = nullptr;
  if (!tbuf) MSMD_CHECK_CUDA(cudaMalloc(&tbuf, 8 * 8));
  p.trace = tbuf;
#endif
  auto kern = fp16 ? row0_fused_kernel<true> : row0_fused_kernel<false>;
  static bool attr = false;
  if (!attr) {
    MSMD_CHECK_CUDA(cudaFuncSetAttribute(row0_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    MSMD_CHECK_CUDA(cudaFuncSetAttribute(row0_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr = true;
  }
  // clusters that can be co-resident (a cluster of 8 lives inside one GPC): more would only queue behind them
  static int max_clusters = 0;
  if (max_clusters == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kNumSMs / kHeads * kHeads);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kHeads; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, row0_fused_kernel<false>, &cfg) != cudaSuccess || nc < 1) {
      (void)cudaGetLastError();
      nc = kNumSMs / kHeads;
    }
    max_clusters = nc < kNumSMs / kHeads ? nc : kNumSMs / kHeads;
  }
  const int ncl = S < max_clusters ? S : max_clusters;
  ProfileScope prof("row0_fused", st);
  MSMD_CHECK_CUDA(launch_pdl(kern, dim3(ncl * kHeads), dim3(kThreads), kSmem, st, p));
  MSMD_CHECK_LAUNCH();
#ifdef MSMD_ROW0_TRACE
  {
    static int calls = 0;
    if (++calls == 40) {
      unsigned long long hst[8];
      MSMD_CHECK_CUDA(cudaStreamSynchronize(st));
      MSMD_CHECK_CUDA(cudaMemcpy(hst, tbuf, sizeof(hst), cudaMemcpyDeviceToHost));
      fprintf(stderr, "[row0 trace] S=%d clusters=%d cycles since start: staged+x %lld | q-proj %lld | attention %lld | out-proj %lld | "
              "cluster barrier %lld | reduce+LN %lld\n", S, ncl, (long long)(hst[1] - hst[0]), (long long)(hst[2] - hst[1]),
              (long long)(hst[3] - hst[2]), (long long)(hst[4] - hst[3]), (long long)(hst[5] - hst[4]), (long long)(hst[6] - hst[5]));
    }
  }
#endif
  return MSMD_OK;
}

}  // namespace msmd
