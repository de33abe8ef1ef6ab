// Multi-head self-attention of the HF Hubert / Wav2Vec2 encoder layers (utils/hubert.py:13-51, utils/wav2vec2.py:71-119 ->
// HubertAttention: 12 heads x 64, no mask, T = 50 tokens per audio-second: 600 for a 12 s window, 3000 for a 60 s clip) as a
// flash attention on tcgen05: the T x T score matrix never leaves the SM (the reference's eager path materialises
// 12 x [N,12,T,T] fp32 maps: 5.2 GB per clip at T = 3000).  Replaces the mma.sync kernel of round 1.
//
// One job = (clip, head, 256 queries) on a persistent CTA, 10 warps:
//   warp 0      TMA producer: the job's two 128-query Q tiles once, then K / V tiles of 128 keys through a 3-stage ring
//               (3-D tensor maps [clip][token][feature]: rows past the clip's T tokens are zero-filled on load and
//               clipped on store, so a tile never touches the next clip)
//   warp 1      MMA issuer (one thread): per key tile and query group g:  S_g = Q_g K^T  (M128 N128 K64, K-major A and B),
//               then, once the group's softmax has written P_g,  O_g = P_g V  (M128 N64 K128, V as an MN-major B
//               operand straight from its TMA tile).  PV_g(kt) is followed at once by S_g(kt+1), so the tensor pipe works
//               on group g's next scores while the group still folds O_g(kt) into its running output.
//   warps 2-5   softmax group 0, warps 6-9 group 1 (thread = query row): two passes over S_g in TMEM (row max, then
//               exp2 + row sum + bf16 P -> swizzled shared memory); the partial output O_g of the tile is read back and
//               accumulated in REGISTERS with the online-softmax rescale (acc = acc * 2^(m_old - m_new) + O_g), so TMEM
//               holds no state across key tiles and needs no correction pass.
// The two groups share every K / V tile (256 queries per K/V byte) and ping-pong on the SM's MUFU unit, which bounds the
// kernel: 128 x 128 exp2 per tile-step against 2 x 256 tensor cycles.
// q arrives pre-scaled by 1/sqrt(64) (folded into Wq at load, audio.cu).
#include "audio_kernels.cuh"
#include "profile.cuh"
#include "tc_common.cuh"

namespace msmd {
namespace {

using namespace tc;

constexpr int kBQ = 128, kBK = 128, kDh = 64;
constexpr int kKvStages = 3;
constexpr int kQBytes = kBQ * 128;                 // [128 rows][64 dims] bf16, SW128
constexpr int kKBytes = kBK * 128, kVBytes = kBK * 128;
constexpr int kStageBytes = kKBytes + kVBytes;
constexpr int kPBytes = 2 * kBQ * 128;             // two 64-key column blocks of [128 rows][128 B]
constexpr int kSmemBytes = 1024 + 2 * kQBytes + kKvStages * kStageBytes + 2 * kPBytes + 256;
constexpr int kThreads = 320;
constexpr int kGroupCols = 256;                    // TMEM columns per group: S [0,128), O [128,192)

struct FaParams {
  CUtensorMap qkv_map;     // [N][T][3*d] bf16, box {64, 128, 1}
  CUtensorMap out_map;     // [N][T][d]   bf16, box {64, 128, 1}
  int N, T, H, d, jobs, q_pairs, n_kt;
};

__device__ __forceinline__ float ex2(float x) {      // one MUFU.EX2 (exp2f adds a denormal-range fix-up around it)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(kThreads, 1) flash_attn_tc_kernel(const __grid_constant__ FaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_base = smem;
  uint8_t* kv_base = q_base + 2 * kQBytes;
  uint8_t* p_base = kv_base + kKvStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_base + 2 * kPBytes);
  uint64_t* qfull = bars;                       // [2] Q_g landed
  uint64_t* qempty = qfull + 2;                 // [2] every S MMA of the job has read Q_g
  uint64_t* kvfull = qempty + 2;                // [kKvStages]
  uint64_t* kvempty = kvfull + kKvStages;       // [kKvStages] both groups' P V of the tile are done
  uint64_t* sfull = kvempty + kKvStages;        // [2] S_g complete in TMEM
  uint64_t* pfull = sfull + 2;                  // [2] P_g written (4 warps)
  uint64_t* ofull = pfull + 2;                  // [2] O_g complete in TMEM (and P_g no longer read)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ofull + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.qkv_map);
    prefetch_tmap(&p.out_map);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&qfull[g], 1); mbar_init(&qempty[g], 1); mbar_init(&sfull[g], 1); mbar_init(&pfull[g], 4); mbar_init(&ofull[g], 1);
    }
    for (int s = 0; s < kKvStages; ++s) { mbar_init(&kvfull[s], 1); mbar_init(&kvempty[s], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // contiguous job range of this CTA (consecutive query pairs of one (clip, head) re-read its K / V from L2)
  const int job0 = (int)(((int64_t)p.jobs * blockIdx.x) / gridDim.x);
  const int job1 = (int)(((int64_t)p.jobs * (blockIdx.x + 1)) / gridDim.x);
  auto decode = [&](int job, int& n, int& h, int& q0) {
    const int per_clip = p.H * p.q_pairs;
    n = job / per_clip;
    const int r = job - n * per_clip;
    h = r / p.q_pairs;
    q0 = (r - h * p.q_pairs) * 2 * kBQ;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int kv_it = 0;
      uint32_t qn[2] = {0, 0};                     // Q tiles loaded so far per group
      for (int job = job0; job < job1; ++job) {
        int n, h, q0;
        decode(job, n, h, q0);
        for (int g = 0; g < 2; ++g) {
          if (q0 + g * kBQ >= p.T) continue;       // the clip has no second tile in this pair
          mbar_wait(&qempty[g], (qn[g]++ & 1) ^ 1);
          mbar_expect_tx(&qfull[g], kQBytes);
          tma_load_3d(q_base + g * kQBytes, &p.qkv_map, &qfull[g], h * kDh, q0 + g * kBQ, n);
        }
        for (int kt = 0; kt < p.n_kt; ++kt, ++kv_it) {
          const int st = kv_it % kKvStages;
          mbar_wait(&kvempty[st], ((kv_it / kKvStages) & 1) ^ 1);
          uint8_t* sk = kv_base + st * kStageBytes;
          mbar_expect_tx(&kvfull[st], kStageBytes);
          tma_load_3d(sk, &p.qkv_map, &kvfull[st], p.d + h * kDh, kt * kBK, n);
          tma_load_3d(sk + kKBytes, &p.qkv_map, &kvfull[st], 2 * p.d + h * kDh, kt * kBK, n);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_qk = make_idesc(1, kBQ, kBK);                    // bf16, K-major A and B
    constexpr uint32_t idesc_pv = make_idesc(1, kBQ, kDh) | (1u << 16);       // B (= V) is MN-major
    if (lane == 0) {
      int kv_it = 0;
      uint32_t use[2] = {0, 0};       // key tiles processed so far per group (phase of pfull / sfull / ofull)
      uint32_t qn[2] = {0, 0};        // Q tiles consumed so far per group
      for (int job = job0; job < job1; ++job) {
        int n, h, q0;
        decode(job, n, h, q0);
        const bool act[2] = {true, q0 + kBQ < p.T};
        auto issue_s = [&](int g, int st) {
          const uint32_t sq = smem_u32(q_base + g * kQBytes), sk = smem_u32(kv_base + st * kStageBytes);
          const uint64_t dq = make_smem_desc_sw128(sq), dk = make_smem_desc_sw128(sk);
#pragma unroll
          for (int k = 0; k < kDh / 16; ++k)
            umma<0>(tmem_base + g * kGroupCols, desc_advance(dq, k * 32), desc_advance(dk, k * 32), idesc_qk, k != 0);
          umma_commit(&sfull[g]);
        };
        // prologue: scores of the first key tile
        mbar_wait(&kvfull[kv_it % kKvStages], (kv_it / kKvStages) & 1);
        for (int g = 0; g < 2; ++g) {
          if (!act[g]) continue;
          mbar_wait(&qfull[g], qn[g]++ & 1);
          tc_fence_after();
          issue_s(g, kv_it % kKvStages);
          if (p.n_kt == 1) umma_commit(&qempty[g]);
        }
        for (int kt = 0; kt < p.n_kt; ++kt, ++kv_it) {
          const int st = kv_it % kKvStages;
          const bool more = kt + 1 < p.n_kt;
          if (more) mbar_wait(&kvfull[(kv_it + 1) % kKvStages], ((kv_it + 1) / kKvStages) & 1);
          for (int g = 0; g < 2; ++g) {
            if (!act[g]) continue;
            mbar_wait(&pfull[g], use[g] & 1);        // P_g(kt) is in shared memory (and S_g(kt) has been read out)
            tc_fence_after();
            const uint32_t sv = smem_u32(kv_base + st * kStageBytes + kKBytes);
            const uint32_t sp = smem_u32(p_base + g * kPBytes);
#pragma unroll
            for (int ks = 0; ks < kBK / 16; ++ks) {
              const uint64_t dp = make_smem_desc_sw128(sp + (ks >> 2) * (kBQ * 128) + (ks & 3) * 32);
              const uint64_t dv = make_smem_desc_sw128(sv + ks * 2048);      // 16 keys = two 8-key groups of 1024 B
              umma<0>(tmem_base + g * kGroupCols + 128, dp, dv, idesc_pv, ks != 0);
            }
            umma_commit(&ofull[g]);
            ++use[g];
            if (more) {
              issue_s(g, (kv_it + 1) % kKvStages);
              if (kt + 2 == p.n_kt) umma_commit(&qempty[g]);   // that was the job's last read of Q_g
            }
          }
          umma_commit(&kvempty[st]);                 // K (read by the S MMAs) and V of this stage are free
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax + output (thread = query row)
    const int g = (warp - 2) >> 2;
    const int qd = warp & 3;                       // TMEM lane quarter this warp may access
    const int r = qd * 32 + lane;                  // query row of the group's tile
    const uint32_t t_s = tmem_base + ((uint32_t)(qd * 32) << 16) + g * kGroupCols;
    const uint32_t t_o = t_s + 128;
    uint8_t* prow = p_base + g * kPBytes + r * 128;
    const int swz = r & 7;
    const float L2E = 1.4426950408889634f;
    const bool issuer = (warp - 2) % 4 == 0 && lane == 0;   // one thread per group owns the group's TMA stores
    uint32_t use = 0;
    for (int job = job0; job < job1; ++job) {
      int n, h, q0;
      decode(job, n, h, q0);
      if (q0 + g * kBQ >= p.T) continue;
      float m_run = -INFINITY, l_run = 0.f;
      float acc[kDh];
#pragma unroll
      for (int i = 0; i < kDh; ++i) acc[i] = 0.f;
      // the previous job's output store has finished reading this group's P buffer
      if (issuer) tma_store_wait_read<0>();
      asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
      for (int kt = 0; kt < p.n_kt; ++kt, ++use) {
        const uint32_t ph = use & 1;
        mbar_wait(&sfull[g], ph);
        tc_fence_after();
        const int kleft = p.T - kt * kBK;          // keys of this tile that exist (>= 128 except for the last tile)
        const bool full = kleft >= kBK;            // (warp-uniform) only the clip's last key tile needs masking
        // pass 1: row maximum.  Two 32-column TMEM reads are kept in flight (the second is issued before the first is used).
        float m_tile = -INFINITY;
        {
          uint32_t va[32], vb[32];
          tmem_ld32(t_s, va);
#pragma unroll
          for (int c = 0; c < kBK; c += 64) {
            tmem_ld_wait();
            tmem_ld32(t_s + c + 32, vb);
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; ++j) m_tile = fmaxf(m_tile, __uint_as_float(va[j]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) m_tile = fmaxf(m_tile, (c + j < kleft) ? __uint_as_float(va[j]) : -INFINITY);
            }
            tmem_ld_wait();
            if (c + 64 < kBK) tmem_ld32(t_s + c + 64, va);
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; ++j) m_tile = fmaxf(m_tile, __uint_as_float(vb[j]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) m_tile = fmaxf(m_tile, (c + 32 + j < kleft) ? __uint_as_float(vb[j]) : -INFINITY);
            }
          }
        }
        const float m_new = fmaxf(m_run, m_tile);
        const float alpha = ex2((m_run - m_new) * L2E);        // 0 for the first tile (m_run = -inf)
        const float mb = m_new * L2E;
        // pass 2: p = 2^(s log2e - m log2e), row sum, bf16 P -> swizzled shared memory (K-major A operand of P V)
        float l_tile = 0.f;
        auto emit = [&](const uint32_t (&v)[32], int c) {
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8) {
            uint32_t w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int j = k8 * 8 + 2 * u;
              float e0 = ex2(fmaf(__uint_as_float(v[j]), L2E, -mb));
              float e1 = ex2(fmaf(__uint_as_float(v[j + 1]), L2E, -mb));
              if (!full) {
                e0 = (c + j < kleft) ? e0 : 0.f;
                e1 = (c + j + 1 < kleft) ? e1 : 0.f;
              }
              l_tile += e0 + e1;
              w[u] = pack_bf2(e0, e1);
            }
            const int kc = (c >> 3) + k8;              // 16-byte chunk (8 keys) index within the 128-key row
            *reinterpret_cast<uint4*>(prow + (kc >> 3) * (kBQ * 128) + (((kc & 7) ^ swz) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        };
        {
          uint32_t va[32], vb[32];
          tmem_ld32(t_s, va);
#pragma unroll
          for (int c = 0; c < kBK; c += 64) {
            tmem_ld_wait();
            tmem_ld32(t_s + c + 32, vb);
            emit(va, c);
            tmem_ld_wait();
            if (c + 64 < kBK) tmem_ld32(t_s + c + 64, va);
            emit(vb, c + 32);
          }
        }
        fence_proxy_async_smem();     // P (generic-proxy writes) -> visible to the MMA's async-proxy reads
        tc_fence_before();            // ... and this warp's TMEM reads of S_g precede the next S MMA
        __syncwarp();
        if (lane == 0) mbar_arrive(&pfull[g]);
        // fold the tile's partial output into the running one while the tensor pipe computes it / the next scores
        m_run = m_new;
        l_run = fmaf(l_run, alpha, l_tile);
        mbar_wait(&ofull[g], ph);
        tc_fence_after();
        {
          uint32_t oa[32], ob[32];
          tmem_ld32(t_o, oa);
          tmem_ld32(t_o + 32, ob);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = fmaf(acc[j], alpha, __uint_as_float(oa[j]));
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[32 + j] = fmaf(acc[32 + j], alpha, __uint_as_float(ob[j]));
        }
        tc_fence_before();
      }
      // normalise, stage the [128 x 64] head slice in the (free again) P buffer, one TMA store (rows >= T are clipped)
      const float inv = 1.0f / l_run;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) w[u] = pack_bf2(acc[c * 8 + 2 * u] * inv, acc[c * 8 + 2 * u + 1] * inv);
        *reinterpret_cast<uint4*>(prow + ((c ^ swz) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
      if (issuer) {
        tma_store_3d(&p.out_map, p_base + g * kPBytes, h * kDh, q0 + g * kBQ, n);
        tma_store_commit();
      }
    }
    if (issuer) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int flash_attn_tc(const bf16* qkv, bf16* ctx, int N, int T, int H, cudaStream_t st) {
  MSMD_REQUIRE(N >= 1 && T >= 1 && H >= 1, "flash_attn: empty problem");
  FaParams p;
  memset(&p, 0, sizeof(p));
  const int d = H * kDh;
  int rc;
  {
    const uint64_t dims[3] = {(uint64_t)3 * d, (uint64_t)T, (uint64_t)N};
    const uint64_t str[2] = {(uint64_t)3 * d * 2, (uint64_t)T * 3 * d * 2};
    const uint32_t box[3] = {kDh, kBQ, 1};
    if ((rc = make_tmap(&p.qkv_map, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)d, (uint64_t)T, (uint64_t)N};
    const uint64_t str[2] = {(uint64_t)d * 2, (uint64_t)T * d * 2};
    const uint32_t box[3] = {kDh, kBQ, 1};
    if ((rc = make_tmap(&p.out_map, ctx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  p.N = N; p.T = T; p.H = H; p.d = d;
  p.q_pairs = cdiv(T, 2 * kBQ);
  p.n_kt = cdiv(T, kBK);
  p.jobs = N * H * p.q_pairs;
  static bool attr = false;
  if (!attr) {
    MSMD_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr = true;
  }
  ProfileScope prof("flash_attn", st);
  const int grid = p.jobs < kNumSMs ? p.jobs : kNumSMs;
  flash_attn_tc_kernel<<<grid, kThreads, kSmemBytes, st>>>(p);
  MSMD_CHECK_LAUNCH();
  return MSMD_OK;
}

}  // namespace msmd
