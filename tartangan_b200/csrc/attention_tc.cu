// Fused 2-D self-attention on tcgen05 / TMEM (reference models/blocks/attention.py:25-34):
//     beta = softmax(theta^T phi) over keys,  o = beta g        -- beta is never written to HBM.
// Per image: Q = theta [Nq][dk], K = pooled phi [Nk][dk], V = pooled g [Nk][dv], bf16, rows = positions (NHWC
// memory of the 1x1-conv outputs), O [Nq][dv] bf16, LSE [Nq] fp32 (row max + log of the row sum, for backward).
//
// Forward: one CTA = 128 query rows x all keys, K / V of the image resident in shared memory in the UMMA
// SWIZZLE_NONE layouts (16-byte units of 8 channels, planes per channel group), so that
//     S_j  = Q K_j^T      A = Q   (K-major, K = dk padded to 16),  B = K_j (K-major)         -> TMEM, 2 buffers
//     OP_j = P_j V_j      A = P_j (K-major, written by the softmax threads as bf16),         -> TMEM, 2 buffers
//                         B = V_j (MN-major: channels contiguous)
// warps 0-3: thread t owns query row t (TMEM lane t): online softmax in the exp2 domain, running output in
// registers (o = (o + OP_{j-1}) * alpha_j, so TMEM accumulators are never rescaled); warp 4: one elected thread
// issues every MMA.  The kernel is bound by MUFU.EX2 (Nq Nk exponentials per image, 16 / clk / SM), not by the
// tensor pipe: dk = C/8 is 8 or 16, i.e. the QK^T contraction is ONE k16 MMA step.
//
// Backward: one CTA = 128 keys x all query tiles of the image; everything is computed transposed so that thread t
// owns key row t:  S^T = K_j Q_i^T, dP^T = V_j dO_i^T  (TMEM), P^T = exp(S^T - lse_i), dS^T = P^T (dP^T - delta_i)
// (bf16, shared memory), then dV_j += P^T dO_i, dK_j += dS^T Q_i (TMEM accumulators over the whole loop) and
// dQ_i = dS K_j (the SAME dS^T image read MN-major), added to an fp32 dQ with red.global.
#include "tc_common.cuh"

#define ATT_LOG2E 1.4426950408889634f

__device__ __forceinline__ float att_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void att_cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void att_cp_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void att_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void att_cp_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ uint32_t att_pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------------ forward
__device__ __forceinline__ void att_group_sync(int g) { asm volatile("bar.sync %0, 160;" ::"r"(g + 1) : "memory"); }

// NG = query-tile groups per CTA (each: 4 softmax warps + 1 issuer warp, own Q / P buffers, own TMEM columns and
// barriers, sharing the resident K / V).  NG = 2 puts two warps on every SM sub-partition, so that one group's
// TMEM loads / barrier waits hide under the other group's MUFU work.
template <int DK8, int DV, int NKT, int NG>
__global__ void __launch_bounds__(NG * 160) attn_fwd_kernel(const bf16* __restrict__ Q, const bf16* __restrict__ K,
                                                            const bf16* __restrict__ V, bf16* __restrict__ O,
                                                            float* __restrict__ LSE, int Nq, int Nk, int tiles_per_cta) {
  constexpr int DK = DK8 * 8, DV8 = DV / 8, NT = NG * 160, TCOLS = 2 * NKT + 2 * DV;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sQ = smem;                                   // [NG][2 planes][128 rows] x 16 B
  uint8_t* sK = sQ + NG * 4096;                         // [2 planes][Nk] x 16 B
  uint8_t* sV = sK + (size_t)2 * Nk * 16;               // [DV8 planes][Nk] x 16 B
  uint8_t* sP = sV + (size_t)DV8 * Nk * 16;             // [NG][2 buffers][NKT/8 planes][128 rows] x 16 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + NG * 2 * NKT * 256);     // per group: s_full[2], p_full[2], o_full[2]
  uint64_t* v_ready = bars + 6 * NG;                     // V landed (one arrival per softmax thread)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6 * NG + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const bool issuer = warp >= 4 * NG;
  const int grp = issuer ? warp - 4 * NG : warp >> 2;
  const int row = tid & 127;                            // softmax threads: query row of the group's tile = TMEM lane
  const int qtiles = Nq >> 7, nkt = Nk / NKT;
  const long long t0 = (long long)blockIdx.x * tiles_per_cta;
  const int img = (int)(t0 / qtiles), qt0 = (int)(t0 % qtiles);
  Q += (long long)img * Nq * DK; K += (long long)img * Nk * DK; V += (long long)img * Nk * DV;
  O += (long long)img * Nq * DV; LSE += (long long)img * Nq;

  if (warp == 4 * NG) tmem_alloc(tmem_slot, 512u);
  if (tid == 0) {
    for (int g = 0; g < NG; ++g)
      for (int b = 0; b < 2; ++b) { mbar_init(&bars[6 * g + b], 1); mbar_init(&bars[6 * g + 2 + b], 128); mbar_init(&bars[6 * g + 4 + b], 1); }
    mbar_init(v_ready, 128 * NG);
    mbar_fence_init();
  }
  const uint32_t sK_a = smem_u32(sK), sV_a = smem_u32(sV);
  for (int u = tid; u < Nk * DK8; u += NT) {
    const int key = u / DK8, g = u - key * DK8;
    att_cp16(sK_a + (uint32_t)(g * Nk + key) * 16, K + (long long)key * DK + g * 8);
  }
  att_cp_commit();
  // V is not needed before the first P V product: the softmax threads fetch it as a second cp.async group and
  // publish it through v_ready after their first score tile
  if (!issuer)
    for (int u = tid; u < Nk * DV8; u += 128 * NG) {
      const int key = u / DV8, g = u - key * DV8;
      att_cp16(sV_a + (uint32_t)(g * Nk + key) * 16, V + (long long)key * DV + g * 8);
    }
  att_cp_commit();
  if (DK8 == 1) {          // the contraction is padded to one k16 step: the second channel group is zero
    for (int u = tid; u < Nk; u += NT) *reinterpret_cast<uint4*>(sK + (size_t)(Nk + u) * 16) = make_uint4(0u, 0u, 0u, 0u);
    if (!issuer) *reinterpret_cast<uint4*>(sQ + (size_t)grp * 4096 + (size_t)(128 + row) * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  att_cp_wait_but_one();
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base + (uint32_t)(grp * TCOLS), tO = tS + 2 * NKT;
  uint64_t* s_full = bars + 6 * grp;                    // [2] S_j landed in TMEM
  uint64_t* p_full = s_full + 2;                        // [2] P_j written (128 arrivals); S buffer free
  uint64_t* o_full = s_full + 4;                        // [2] OP_j landed; P buffer free
  uint8_t* sQg = sQ + (size_t)grp * 4096;
  uint8_t* sPg = sP + (size_t)grp * 2 * NKT * 256;
  const uint32_t sQ_a = smem_u32(sQg), sP_a = smem_u32(sPg);

  uint32_t ph_s = 0, ph_p = 0, ph_o = 0;                // bit b = parity the next wait on barrier b expects
  for (int t = grp; t < tiles_per_cta; t += NG) {
    const int q0 = (qt0 + t) * 128;
    if (!issuer) {
#pragma unroll
      for (int g = 0; g < DK8; ++g)
        *reinterpret_cast<uint4*>(sQg + (size_t)(g * 128 + row) * 16) =
            *reinterpret_cast<const uint4*>(Q + (long long)(q0 + row) * DK + g * 8);
    }
    fence_proxy_async_smem();
    att_group_sync(grp);
    if (issuer) {
      if (elect_one()) {
        tc_fence_after_sync();
        const uint32_t idS = umma_idesc_bf16(128, NKT, 0, 0), idO = umma_idesc_bf16(128, DV, 0, 1);
        const uint64_t qd = umma_desc(sQ_a, 128 * 16, 128);
        uint32_t php = ph_p;
        auto issue_S = [&](int j) {
          const uint64_t kd = umma_desc(sK_a + (uint32_t)(j * NKT) * 16, (uint32_t)Nk * 16, 128);
          umma_bf16(tS + (uint32_t)((j & 1) * NKT), qd, kd, idS, 0u);
          umma_commit(&s_full[j & 1]);
        };
        issue_S(0);
        if (nkt > 1) issue_S(1);
        for (int j = 0; j < nkt; ++j) {
          const int b = j & 1;
          mbar_wait(&p_full[b], (php >> b) & 1u); php ^= 1u << b;
          if (t == grp && j == 0) mbar_wait(v_ready, 0u);
          tc_fence_after_sync();
#pragma unroll
          for (int ks = 0; ks < NKT / 16; ++ks) {
            const uint64_t ad = umma_desc(sP_a + (uint32_t)(b * NKT * 256 + ks * 2 * 2048), 2048, 128);
            const uint64_t vd = umma_desc(sV_a + (uint32_t)(j * NKT + ks * 16) * 16, 128, (uint32_t)Nk * 16);
            umma_bf16(tO + (uint32_t)(b * DV), ad, vd, idO, ks > 0 ? 1u : 0u);
          }
          umma_commit(&o_full[b]);
          if (j + 2 < nkt) issue_S(j + 2);
        }
      }
      __syncwarp();
      // every lane tracks the parities so that any lane may be elected for the next tile
      ph_p ^= (uint32_t)(((nkt + 1) >> 1) & 1) | ((uint32_t)((nkt >> 1) & 1) << 1);
    } else {
      const uint32_t lane_t = (uint32_t)((warp & 3) * 32) << 16;
      float m = -INFINITY, l = 0.f;
      float o[DV];
#pragma unroll
      for (int c = 0; c < DV; ++c) o[c] = 0.f;
      for (int j = 0; j < nkt; ++j) {
        const int b = j & 1;
        mbar_wait(&s_full[b], (ph_s >> b) & 1u); ph_s ^= 1u << b;
        tc_fence_after_sync();
        const uint32_t ts = tS + lane_t + (uint32_t)(b * NKT);
        uint32_t s[NKT];                                 // the whole score row of this key tile, read from TMEM once
#pragma unroll
        for (int c = 0; c < NKT; c += 16) tmem_ld16(ts + c, s + c);
        tmem_ld_wait();
        float mx = m;
#pragma unroll
        for (int c = 0; c < NKT; ++c) mx = fmaxf(mx, __uint_as_float(s[c]));
        const float mxl = mx * ATT_LOG2E;
        const float alpha = att_ex2(m * ATT_LOG2E - mxl);
        float rs = 0.f;
        uint8_t* prow = sPg + (size_t)b * NKT * 256 + (size_t)row * 16;
#pragma unroll
        for (int c = 0; c < NKT; c += 8) {
          float p[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) { p[e] = att_ex2(fmaf(__uint_as_float(s[c + e]), ATT_LOG2E, -mxl)); rs += p[e]; }
          *reinterpret_cast<uint4*>(prow + (size_t)(c >> 3) * 2048) =
              make_uint4(att_pack2(p[0], p[1]), att_pack2(p[2], p[3]), att_pack2(p[4], p[5]), att_pack2(p[6], p[7]));
        }
        l = l * alpha + rs; m = mx;
        if (t == grp && j == 0) { att_cp_wait_all(); fence_proxy_async_smem(); mbar_arrive(v_ready); }
        tc_fence_before_sync();
        fence_proxy_async_smem();
        mbar_arrive(&p_full[b]);
        if (j > 0) {
          const int pb = b ^ 1;
          mbar_wait(&o_full[pb], (ph_o >> pb) & 1u); ph_o ^= 1u << pb;
          tc_fence_after_sync();
          // the running maximum settles after a few key tiles: skip the multiply when no row of the warp moved
          const bool rescale = __any_sync(0xffffffffu, alpha != 1.f);
#pragma unroll
          for (int c = 0; c < DV; c += 16) {
            uint32_t r[16];
            tmem_ld16(tO + lane_t + (uint32_t)(pb * DV + c), r);
            tmem_ld_wait();
            if (rescale) {
#pragma unroll
              for (int i = 0; i < 16; ++i) o[c + i] = (o[c + i] + __uint_as_float(r[i])) * alpha;
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) o[c + i] += __uint_as_float(r[i]);
            }
          }
        }
      }
      {
        const int pb = (nkt - 1) & 1;
        mbar_wait(&o_full[pb], (ph_o >> pb) & 1u); ph_o ^= 1u << pb;
        tc_fence_after_sync();
        const float inv = 1.f / l;
        bf16* orow = O + (long long)(q0 + row) * DV;
#pragma unroll
        for (int c = 0; c < DV; c += 16) {
          uint32_t r[16];
          tmem_ld16(tO + lane_t + (uint32_t)(pb * DV + c), r);
          tmem_ld_wait();
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = (o[c + i] + __uint_as_float(r[i])) * inv;
          *reinterpret_cast<uint4*>(orow + c) = make_uint4(att_pack2(f[0], f[1]), att_pack2(f[2], f[3]), att_pack2(f[4], f[5]), att_pack2(f[6], f[7]));
          *reinterpret_cast<uint4*>(orow + c + 8) = make_uint4(att_pack2(f[8], f[9]), att_pack2(f[10], f[11]), att_pack2(f[12], f[13]), att_pack2(f[14], f[15]));
        }
        LSE[q0 + row] = m + __logf(l);
      }
    }
    tc_fence_before_sync();
    att_group_sync(grp);                                 // the group's MMAs have all completed: sQ / TMEM may be reused
    tc_fence_after_sync();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4 * NG) tmem_dealloc(tmem_base, 512u);
}

// shared memory of the forward kernel: Q, resident K (padded to 16 channels) and V, P double buffers, barriers
static inline size_t attn_fwd_smem(int Nk, int dv, int nkt, int ng) {
  return (size_t)ng * 4096 + (size_t)32 * Nk + (size_t)2 * dv * Nk + (size_t)ng * 512 * nkt + 128;   // 13 barriers + TMEM slot
}
// (key tile, groups): two groups of 64-key tiles when they fit (TMEM: 2 x (128 + 2 dv) columns), else one group
static inline void attn_pick(int Nq, int Nk, int dv, int* nkt, int* ng) {
  *nkt = 0; *ng = 0;
  if ((Nq / 128) % 2 == 0 && Nk % 64 == 0 && attn_fwd_smem(Nk, dv, 64, 2) <= 232448) { *nkt = 64; *ng = 2; return; }
  if (Nk % 128 == 0 && attn_fwd_smem(Nk, dv, 128, 1) <= 232448) { *nkt = 128; *ng = 1; return; }
  if (Nk % 64 == 0 && attn_fwd_smem(Nk, dv, 64, 1) <= 232448) { *nkt = 64; *ng = 1; return; }
}
extern "C" int ttg_attn_supported(int Nq, int Nk, int dk, int dv) {
  if (Nq <= 0 || Nq % 128 != 0) return 0;
  if (!(dk == 8 || dk == 16) || !(dv == 32 || dv == 64)) return 0;
  if (Nk < 128 || Nk % 128 != 0) return 0;               // the backward kernel owns 128 keys per CTA
  int nkt, ng;
  attn_pick(Nq, Nk, dv, &nkt, &ng);
  return nkt != 0;
}

template <int DK8, int DV, int NKT, int NG>
static int attn_fwd_launch(const void* q, const void* k, const void* v, void* o, float* lse, int batch, int Nq, int Nk,
                           cudaStream_t st) {
  const size_t smem = attn_fwd_smem(Nk, DV, NKT, NG);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<DK8, DV, NKT, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return ttg_set_error(TTG_ERR_CUDA, "attn_fwd: %s", cudaGetErrorString(e));
    attr = true;
  }
  // consecutive query tiles of one image per CTA (K / V staged once): the smallest multiple of NG that fits one wave
  const int qtiles = Nq / 128;
  int tpc = qtiles;
  for (int d = NG; d <= qtiles; d += NG)
    if (qtiles % d == 0 && (long long)batch * (qtiles / d) <= ttg_num_sms()) { tpc = d; break; }
  const int grid = batch * (qtiles / tpc);
  attn_fwd_kernel<DK8, DV, NKT, NG><<<grid, NG * 160, smem, st>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, (bf16*)o, lse, Nq, Nk, tpc);
  TTG_CHECK_LAUNCH("attn_fwd");
  return TTG_OK;
}

extern "C" int ttg_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int batch, int Nq, int Nk,
                            int dk, int dv, void* stream) {
  if (!ttg_attn_supported(Nq, Nk, dk, dv))
    return ttg_set_error(TTG_ERR_UNSUPPORTED, "attn_fwd: unsupported shape Nq=%d Nk=%d dk=%d dv=%d", Nq, Nk, dk, dv);
  TTG_REQUIRE(batch > 0, "attn_fwd: empty batch");
  cudaStream_t st = (cudaStream_t)stream;
  int nkt, ng;
  attn_pick(Nq, Nk, dv, &nkt, &ng);
#define ATT_FWD(D8, DVV)                                                                                        \
  if (dk == D8 * 8 && dv == DVV) {                                                                              \
    if (ng == 2) return attn_fwd_launch<D8, DVV, 64, 2>(q, k, v, o, lse, batch, Nq, Nk, st);                    \
    return nkt == 128 ? attn_fwd_launch<D8, DVV, 128, 1>(q, k, v, o, lse, batch, Nq, Nk, st)                    \
                      : attn_fwd_launch<D8, DVV, 64, 1>(q, k, v, o, lse, batch, Nq, Nk, st);                    \
  }
  ATT_FWD(1, 32) ATT_FWD(2, 32) ATT_FWD(1, 64) ATT_FWD(2, 64)
#undef ATT_FWD
  return ttg_set_error(TTG_ERR_UNSUPPORTED, "attn_fwd: no instantiation");
}

// ------------------------------------------------------------------------------------------------ backward
// delta[row] = sum_c dO[row][c] O[row][c];  lse2[row] = lse[row] log2(e);  dQ accumulator zeroed.
template <int DV>
__global__ void attn_bwd_prep_kernel(const bf16* __restrict__ O, const bf16* __restrict__ dO, const float* __restrict__ lse,
                                     float* __restrict__ lse2, float* __restrict__ delta, float* __restrict__ dqacc,
                                     long long rows, int dk) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < DV; c += 8) {
      Vec<bf16> a, b; float fa[8], fb[8];
      a.load(O + r * DV + c); b.load(dO + r * DV + c); a.unpack(fa); b.unpack(fb);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc = fmaf(fa[e], fb[e], acc);
    }
    delta[r] = acc; lse2[r] = lse[r] * ATT_LOG2E;
    float4* z = reinterpret_cast<float4*>(dqacc + r * dk);
    for (int e = 0; e < dk / 4; ++e) z[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__global__ void attn_dq_cast_kernel(const float* __restrict__ acc, bf16* __restrict__ dq, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(acc)[2 * i], b = reinterpret_cast<const float4*>(acc)[2 * i + 1];
    reinterpret_cast<uint4*>(dq)[i] = make_uint4(att_pack2(a.x, a.y), att_pack2(a.z, a.w), att_pack2(b.x, b.y), att_pack2(b.z, b.w));
  }
}

template <int DK8, int DV>
__global__ void __launch_bounds__(192) attn_bwd_kernel(const bf16* __restrict__ Q, const bf16* __restrict__ K,
                                                       const bf16* __restrict__ V, const bf16* __restrict__ dO,
                                                       const float* __restrict__ lse2, const float* __restrict__ delta,
                                                       float* __restrict__ dQacc, bf16* __restrict__ dK, bf16* __restrict__ dV,
                                                       int Nq, int Nk) {
  constexpr int DK = DK8 * 8, DV8 = DV / 8;
  constexpr uint32_t PL = 2048;                        // one plane: 128 rows x 16 B
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sK = smem;                                   // [2 planes][128 keys]
  uint8_t* sV = sK + 2 * PL;                            // [DV8 planes][128 keys]
  uint8_t* sQ = sV + DV8 * PL;                          // [2 buffers][2 planes][128 queries]
  uint8_t* sdO = sQ + 2 * 2 * PL;                       // [2 buffers][DV8 planes][128 queries]
  float* sLse = reinterpret_cast<float*>(sdO + 2 * DV8 * PL);   // [2][128]
  float* sDel = sLse + 256;                             // [2][128]
  uint8_t* sPt = reinterpret_cast<uint8_t*>(sDel + 256);        // [16 planes (queries / 8)][128 keys]
  uint8_t* sdS = sPt + 16 * PL;                         // same shape
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + 16 * PL);
  uint64_t* ld_full = bars;          // [2] loader -> everyone (32 arrivals)
  uint64_t* ld_empty = bars + 2;     // [2] MMAs that read the buffer have completed
  uint64_t* s_full = bars + 4;       // S^T, dP^T in TMEM
  uint64_t* pd_full = bars + 5;      // P^T, dS^T in shared memory (128 arrivals)
  uint64_t* dq_full = bars + 6;      // dV, dK, dQ_i MMAs completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ktiles = Nk >> 7, nqt = Nq >> 7;
  const int img = blockIdx.x / ktiles, k0 = (blockIdx.x % ktiles) * 128;
  Q += (long long)img * Nq * DK; dO += (long long)img * Nq * DV; lse2 += (long long)img * Nq; delta += (long long)img * Nq;
  dQacc += (long long)img * Nq * DK;
  K += ((long long)img * Nk + k0) * DK; V += ((long long)img * Nk + k0) * DV;
  dK += ((long long)img * Nk + k0) * DK; dV += ((long long)img * Nk + k0) * DV;

  if (warp == 4) tmem_alloc(tmem_slot, 512u);
  if (tid == 0) {
    mbar_init(&ld_full[0], 32); mbar_init(&ld_full[1], 32); mbar_init(&ld_empty[0], 1); mbar_init(&ld_empty[1], 1);
    mbar_init(s_full, 1); mbar_init(pd_full, 128); mbar_init(dq_full, 1);
    mbar_fence_init();
  }
  const uint32_t sK_a = smem_u32(sK), sV_a = smem_u32(sV), sQ_a = smem_u32(sQ), sdO_a = smem_u32(sdO);
  const uint32_t sPt_a = smem_u32(sPt), sdS_a = smem_u32(sdS);
  for (int u = tid; u < 128 * DK8; u += 192) {
    const int key = u / DK8, g = u - key * DK8;
    att_cp16(sK_a + (uint32_t)(g * 128 + key) * 16, K + (long long)key * DK + g * 8);
  }
  for (int u = tid; u < 128 * DV8; u += 192) {
    const int key = u / DV8, g = u - key * DV8;
    att_cp16(sV_a + (uint32_t)(g * 128 + key) * 16, V + (long long)key * DV + g * 8);
  }
  if (DK8 == 1 && tid < 128) {
    *reinterpret_cast<uint4*>(sK + PL + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(sQ + PL + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(sQ + 3 * PL + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  att_cp_wait_all();
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tDP = tmem_base + 128, tDV = tmem_base + 256, tDK = tmem_base + 384, tDQ = tmem_base + 416;

  if (warp == 5) {
    // ---------------------------------------------------------------- loader: Q_i, dO_i, lse_i, delta_i
    for (int i = 0; i < nqt; ++i) {
      const int b = i & 1;
      if (i >= 2) mbar_wait(&ld_empty[b], (uint32_t)((i >> 1) - 1) & 1u);
      const long long q0 = (long long)i * 128;
      for (int u = lane; u < 128 * DK8; u += 32) {
        const int q = u / DK8, g = u - q * DK8;
        att_cp16(sQ_a + (uint32_t)(b * 2 * PL) + (uint32_t)(g * 128 + q) * 16, Q + (q0 + q) * DK + g * 8);
      }
      for (int u = lane; u < 128 * DV8; u += 32) {
        const int q = u / DV8, g = u - q * DV8;
        att_cp16(sdO_a + (uint32_t)(b * DV8 * PL) + (uint32_t)(g * 128 + q) * 16, dO + (q0 + q) * DV + g * 8);
      }
      att_cp16(smem_u32(sLse + b * 128 + lane * 4), lse2 + q0 + lane * 4);
      att_cp16(smem_u32(sDel + b * 128 + lane * 4), delta + q0 + lane * 4);
      att_cp_wait_all();
      fence_proxy_async_smem();
      mbar_arrive(&ld_full[b]);
    }
  } else if (warp == 4) {
    // ---------------------------------------------------------------- MMA issuer
    if (elect_one()) {
      const uint32_t idS = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t idV = umma_idesc_bf16(128, DV, 0, 1), idK = umma_idesc_bf16(128, 16, 0, 1), idQ = umma_idesc_bf16(128, 16, 1, 1);
      auto issue_scores = [&](int i) {
        const int b = i & 1;
        mbar_wait(&ld_full[b], (uint32_t)(i >> 1) & 1u);
        tc_fence_after_sync();
        umma_bf16(tS, umma_desc(sK_a, PL, 128), umma_desc(sQ_a + (uint32_t)(b * 2 * PL), PL, 128), idS, 0u);
#pragma unroll
        for (int ks = 0; ks < DV / 16; ++ks)
          umma_bf16(tDP, umma_desc(sV_a + (uint32_t)ks * 2 * PL, PL, 128),
                    umma_desc(sdO_a + (uint32_t)(b * DV8 * PL) + (uint32_t)ks * 2 * PL, PL, 128), idS, ks > 0 ? 1u : 0u);
        umma_commit(s_full);
      };
      issue_scores(0);
      for (int i = 0; i < nqt; ++i) {
        const int b = i & 1;
        mbar_wait(pd_full, (uint32_t)i & 1u);
        tc_fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)        // dV_j += P^T dO_i          (K = 16 queries per step)
          umma_bf16(tDV, umma_desc(sPt_a + (uint32_t)ks * 2 * PL, PL, 128),
                    umma_desc(sdO_a + (uint32_t)(b * DV8 * PL) + (uint32_t)ks * 256, 128, PL), idV, (i > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)        // dK_j += dS^T Q_i
          umma_bf16(tDK, umma_desc(sdS_a + (uint32_t)ks * 2 * PL, PL, 128),
                    umma_desc(sQ_a + (uint32_t)(b * 2 * PL) + (uint32_t)ks * 256, 128, PL), idK, (i > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)        // dQ_i = dS K_j             (K = 16 keys per step; dS^T image read MN-major)
          umma_bf16(tDQ, umma_desc(sdS_a + (uint32_t)ks * 256, 128, PL), umma_desc(sK_a + (uint32_t)ks * 256, 128, PL), idQ,
                    ks > 0 ? 1u : 0u);
        umma_commit(dq_full);
        umma_commit(&ld_empty[b]);
        if (i + 1 < nqt) issue_scores(i + 1);
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- thread t = key row t
    const uint32_t lane_t = (uint32_t)(warp * 32) << 16;
    for (int i = 0; i < nqt; ++i) {
      const int b = i & 1;
      mbar_wait(&ld_full[b], (uint32_t)(i >> 1) & 1u);         // lse / delta of this query tile are visible
      mbar_wait(s_full, (uint32_t)i & 1u);
      tc_fence_after_sync();
      const float4* l4 = reinterpret_cast<const float4*>(sLse + b * 128);
      const float4* d4 = reinterpret_cast<const float4*>(sDel + b * 128);
#pragma unroll 2
      for (int c = 0; c < 128; c += 16) {
        uint32_t rs[16], rp[16];
        tmem_ld16(tS + lane_t + (uint32_t)c, rs); tmem_ld16(tDP + lane_t + (uint32_t)c, rp);
        tmem_ld_wait();
        float p[16], ds[16];
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          const float4 L = l4[(c + e) >> 2], D = d4[(c + e) >> 2];
          p[e] = att_ex2(fmaf(__uint_as_float(rs[e]), ATT_LOG2E, -L.x));
          p[e + 1] = att_ex2(fmaf(__uint_as_float(rs[e + 1]), ATT_LOG2E, -L.y));
          p[e + 2] = att_ex2(fmaf(__uint_as_float(rs[e + 2]), ATT_LOG2E, -L.z));
          p[e + 3] = att_ex2(fmaf(__uint_as_float(rs[e + 3]), ATT_LOG2E, -L.w));
          ds[e] = p[e] * (__uint_as_float(rp[e]) - D.x);
          ds[e + 1] = p[e + 1] * (__uint_as_float(rp[e + 1]) - D.y);
          ds[e + 2] = p[e + 2] * (__uint_as_float(rp[e + 2]) - D.z);
          ds[e + 3] = p[e + 3] * (__uint_as_float(rp[e + 3]) - D.w);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const size_t off = (size_t)((c >> 3) + h) * PL + (size_t)tid * 16;
          *reinterpret_cast<uint4*>(sPt + off) = make_uint4(att_pack2(p[8 * h], p[8 * h + 1]), att_pack2(p[8 * h + 2], p[8 * h + 3]),
                                                            att_pack2(p[8 * h + 4], p[8 * h + 5]), att_pack2(p[8 * h + 6], p[8 * h + 7]));
          *reinterpret_cast<uint4*>(sdS + off) = make_uint4(att_pack2(ds[8 * h], ds[8 * h + 1]), att_pack2(ds[8 * h + 2], ds[8 * h + 3]),
                                                            att_pack2(ds[8 * h + 4], ds[8 * h + 5]), att_pack2(ds[8 * h + 6], ds[8 * h + 7]));
        }
      }
      tc_fence_before_sync();
      fence_proxy_async_smem();
      mbar_arrive(pd_full);
      mbar_wait(dq_full, (uint32_t)i & 1u);
      tc_fence_after_sync();
      {
        uint32_t r[16];
        tmem_ld16(tDQ + lane_t, r);                        // lane = query row of tile i
        tmem_ld_wait();
        float4* dst = reinterpret_cast<float4*>(dQacc + ((long long)i * 128 + tid) * DK);
#pragma unroll
        for (int e = 0; e < DK / 4; ++e)
          atomicAdd(dst + e, make_float4(__uint_as_float(r[4 * e]), __uint_as_float(r[4 * e + 1]), __uint_as_float(r[4 * e + 2]), __uint_as_float(r[4 * e + 3])));
      }
    }
    // dV_j, dK_j: lane = key row
    {
      bf16* vrow = dV + (long long)tid * DV;
#pragma unroll
      for (int c = 0; c < DV; c += 16) {
        uint32_t r[16];
        tmem_ld16(tDV + lane_t + (uint32_t)c, r);
        tmem_ld_wait();
        *reinterpret_cast<uint4*>(vrow + c) = make_uint4(att_pack2(__uint_as_float(r[0]), __uint_as_float(r[1])), att_pack2(__uint_as_float(r[2]), __uint_as_float(r[3])),
                                                         att_pack2(__uint_as_float(r[4]), __uint_as_float(r[5])), att_pack2(__uint_as_float(r[6]), __uint_as_float(r[7])));
        *reinterpret_cast<uint4*>(vrow + c + 8) = make_uint4(att_pack2(__uint_as_float(r[8]), __uint_as_float(r[9])), att_pack2(__uint_as_float(r[10]), __uint_as_float(r[11])),
                                                             att_pack2(__uint_as_float(r[12]), __uint_as_float(r[13])), att_pack2(__uint_as_float(r[14]), __uint_as_float(r[15])));
      }
      uint32_t r[16];
      tmem_ld16(tDK + lane_t, r);
      tmem_ld_wait();
      bf16* krow = dK + (long long)tid * DK;
#pragma unroll
      for (int e = 0; e < DK8; ++e)
        *reinterpret_cast<uint4*>(krow + 8 * e) = make_uint4(att_pack2(__uint_as_float(r[8 * e]), __uint_as_float(r[8 * e + 1])), att_pack2(__uint_as_float(r[8 * e + 2]), __uint_as_float(r[8 * e + 3])),
                                                             att_pack2(__uint_as_float(r[8 * e + 4]), __uint_as_float(r[8 * e + 5])), att_pack2(__uint_as_float(r[8 * e + 6]), __uint_as_float(r[8 * e + 7])));
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 512u);
}

extern "C" size_t ttg_attn_bwd_workspace_bytes(int batch, int Nq, int dk) {
  return ((size_t)batch * Nq * dk + (size_t)2 * batch * Nq) * sizeof(float);
}

template <int DK8, int DV>
static int attn_bwd_launch(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                           void* dq, void* dk_, void* dv_, int batch, int Nq, int Nk, float* ws, cudaStream_t st) {
  constexpr size_t smem = (size_t)(2 + DV / 8 + 4 + 2 * (DV / 8) + 32) * 2048 + 2048 + 128;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel<DK8, DV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return ttg_set_error(TTG_ERR_CUDA, "attn_bwd: %s", cudaGetErrorString(e));
    attr = true;
  }
  const long long rows = (long long)batch * Nq;
  float* dqacc = ws; float* lse2 = ws + rows * DK8 * 8; float* delta = lse2 + rows;
  attn_bwd_prep_kernel<DV><<<ttg_grid_for(rows, 256), 256, 0, st>>>((const bf16*)o, (const bf16*)dout, lse, lse2, delta, dqacc, rows, DK8 * 8);
  TTG_CHECK_LAUNCH("attn_bwd_prep");
  attn_bwd_kernel<DK8, DV><<<batch * (Nk / 128), 192, smem, st>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, (const bf16*)dout, lse2,
                                                                  delta, dqacc, (bf16*)dk_, (bf16*)dv_, Nq, Nk);
  TTG_CHECK_LAUNCH("attn_bwd");
  attn_dq_cast_kernel<<<ttg_grid_for(rows * DK8, 256), 256, 0, st>>>(dqacc, (bf16*)dq, rows * DK8);
  TTG_CHECK_LAUNCH("attn_dq_cast");
  return TTG_OK;
}

extern "C" int ttg_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                            void* dq, void* dk_out, void* dv_out, int batch, int Nq, int Nk, int dk, int dv, void* workspace,
                            void* stream) {
  if (!ttg_attn_supported(Nq, Nk, dk, dv))
    return ttg_set_error(TTG_ERR_UNSUPPORTED, "attn_bwd: unsupported shape Nq=%d Nk=%d dk=%d dv=%d", Nq, Nk, dk, dv);
  TTG_REQUIRE(batch > 0 && workspace != nullptr, "attn_bwd: empty batch / no workspace");
  cudaStream_t st = (cudaStream_t)stream;
#define ATT_BWD(D8, DVV) \
  if (dk == D8 * 8 && dv == DVV) return attn_bwd_launch<D8, DVV>(q, k, v, o, dout, lse, dq, dk_out, dv_out, batch, Nq, Nk, (float*)workspace, st);
  ATT_BWD(1, 32) ATT_BWD(2, 32) ATT_BWD(1, 64) ATT_BWD(2, 64)
#undef ATT_BWD
  return ttg_set_error(TTG_ERR_UNSUPPORTED, "attn_bwd: no instantiation");
}
