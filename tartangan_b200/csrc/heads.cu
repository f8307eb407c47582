// Heads and losses (all fp32, tiny tensors, latency-bound):
//   * IQN head: cosine tau embedding -> Linear(E->C) -> tanh -> multiply with pooled
//     features -> Linear(C->1), per-quantile prediction + mean over quantiles, in ONE
//     kernel (reference models/iqn.py:41-46,91-103 + blocks/discriminator.py:164-178);
//     its backward in one kernel.  The embedding (rows x C) is never written to HBM.
//   * quantile-Huber loss (models/iqn.py:111-130) forward / backward.
//   * BCE-with-logits mean (trainers/cnn.py:88,131,147) forward / backward.
//   * sum of squares (R1 penalty reduction, models/losses.py:27-29).
//   * small fp32 matmul with transposes (nn.Linear in blocks/generator.py:71 and
//     blocks/discriminator.py:137; closed under differentiation).
#include "common.cuh"

#define IQN_MAX_E 32

// tanh via one exp: |error| ~ 1e-7, far inside the fp32 parity tolerance, ~4x cheaper than tanhf
__device__ __forceinline__ float fast_tanh(float x) {
  const float e = __expf(2.f * fminf(fmaxf(x, -15.f), 15.f));
  return __fdividef(e - 1.f, e + 1.f);
}
// cos(tau*pi*(k+1)) for k < E: lane k computes one value, the warp shares them with shuffles
__device__ __forceinline__ void iqn_cosines(float tau, int E, int lane, float* cs) {
  const float mine = lane < E ? cosf(tau * 3.14159265358979323846f * (float)(lane + 1)) : 0.f;
#pragma unroll
  for (int k = 0; k < IQN_MAX_E; ++k) cs[k] = __shfl_sync(0xffffffffu, mine, k);
}

// ---------------------------------------------------------------- IQN head forward
// rows r = q*B + b (quantile-major, Appendix B.9).  One warp per row.
__global__ void __launch_bounds__(256) iqn_head_fwd_kernel(
    const float* __restrict__ feats, const float* __restrict__ taus, const float* __restrict__ We,
    const float* __restrict__ be, const float* __restrict__ wo, const float* __restrict__ bo,
    float* __restrict__ p_tau, int B, int R, int C, int E) {
  extern __shared__ float s_we[];   // [C][E+1] (odd stride: conflict-free) + be[C] + wo[C]
  const int ES = E + 1;
  float* s_be = s_we + C * ES; float* s_wo = s_be + C;
  for (int i = threadIdx.x; i < C * E; i += blockDim.x) s_we[(i / E) * ES + (i % E)] = We[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) { s_be[i] = be[i]; s_wo[i] = wo[i]; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int r = blockIdx.x * nwarp + warp; r < R; r += gridDim.x * nwarp) {
    const int b = r % B;
    const float tau = taus[r];
    float cs[IQN_MAX_E];
    iqn_cosines(tau, E, lane, cs);
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) {
      float pre = s_be[c];
#pragma unroll
      for (int k = 0; k < IQN_MAX_E; ++k) if (k < E) pre += cs[k] * s_we[c * ES + k];
      acc += feats[(long long)b * C + c] * fast_tanh(pre) * s_wo[c];
    }
    acc = warp_sum(acc);
    if (lane == 0) p_tau[r] = acc + (bo ? bo[0] : 0.f);
  }
}
__global__ void quantile_mean_kernel(const float* __restrict__ p_tau, float* __restrict__ p_mean, int B, int nq) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) { float s = 0.f; for (int q = 0; q < nq; ++q) s += p_tau[q * B + b]; p_mean[b] = s / (float)nq; }
}
// bo may be null (used by the double-backward, where the bias does not enter).
extern "C" int ttg_iqn_head_fwd(const float* feats, const float* taus, const float* We, const float* be, const float* wo,
                                const float* bo, float* p_tau, float* p_mean, int B, int nq, int C, int E, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  TTG_REQUIRE(E <= IQN_MAX_E, "iqn_head: embedding dims %d > %d", E, IQN_MAX_E);
  int R = B * nq;
  size_t smem = sizeof(float) * ((size_t)C * (E + 1) + 2 * C);
  TTG_REQUIRE(smem <= 200 * 1024, "iqn_head: C*E too large for shared memory");
  cudaFuncSetAttribute(iqn_head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  iqn_head_fwd_kernel<<<ttg_grid_for(R, 8, 1), 256, smem, st>>>(feats, taus, We, be, wo, bo, p_tau, B, R, C, E);
  TTG_CHECK_LAUNCH("iqn_head_fwd");
  if (p_mean) {
    quantile_mean_kernel<<<(B + 127) / 128, 128, 0, st>>>(p_tau, p_mean, B, nq);
    TTG_CHECK_LAUNCH("quantile_mean");
  }
  return TTG_OK;
}

// ---------------------------------------------------------------- IQN head + quantile mean + quantile-Huber loss: ONE kernel
// (blocks/discriminator.py:164-178 in one launch: models/iqn.py:91-103 embedding + mix, Linear(C->1), the mean over
// quantiles p_target (:174-175) and iqn_loss (:111-130)).  One warp owns a batch row b and walks its nq quantile rows
// r = q*B + b: the pooled features are read ONCE per batch row (registers), p_target needs no second pass over
// p_tau and the loss terms of the row are summed in the same registers.  loss must be zero on entry.
// Per batch row the cos(tau pi k) table of a chunk of <= 8 quantiles is written to a per-warp strip of shared memory
// once and read back as broadcast LDS.128, the embedding row We[c][:] of a lane's channel lives in registers while the
// lane walks the chunk: 5 vector loads per 20 FMAs (the first version shuffled every cosine to every lane for every
// row and re-read We per row: LDS / SHFL bound at 10 % of the fp32 rate, 940 us per 2^20 quantile rows).
#define IQN_QC 8
#define IQN_EP 20            // embedding dims of the table rows (E <= 20 takes the fast path; E = 20 in the reference)
template <int CPL>        // channels per lane = ceil(C / 32)
__global__ void __launch_bounds__(256) iqn_head_loss_fwd_kernel(
    const float* __restrict__ feats, const float* __restrict__ taus, const float* __restrict__ We,
    const float* __restrict__ be, const float* __restrict__ wo, const float* __restrict__ bo,
    const float* __restrict__ target, float* __restrict__ p_tau, float* __restrict__ p_mean,
    float* __restrict__ loss, int B, int nq, int C, int E, float k) {
  extern __shared__ __align__(16) float s_dyn[];   // cos tables [8 warps][IQN_QC][IQN_EP] + We [C][E+1] + be[C] + wo[C]
  float* s_cs_all = s_dyn;
  float* s_we = s_dyn + 8 * IQN_QC * IQN_EP;
  const int ES = E + 1;
  float* s_be = s_we + C * ES; float* s_wo = s_be + C;
  for (int i = threadIdx.x; i < C * E; i += blockDim.x) s_we[(i / E) * ES + (i % E)] = We[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) { s_be[i] = be[i]; s_wo[i] = wo[i]; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  float* s_cs = s_cs_all + warp * (IQN_QC * IQN_EP);
  const float bias_o = bo ? bo[0] : 0.f;
  float lacc = 0.f;
  for (int b = blockIdx.x * nwarp + warp; b < B; b += gridDim.x * nwarp) {
    float f[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) { const int c = lane + 32 * j; f[j] = c < C ? feats[(long long)b * C + c] * s_wo[c] : 0.f; }
    const float tgt = target ? target[b] : 0.f;
    float msum = 0.f;
    for (int q0 = 0; q0 < nq; q0 += IQN_QC) {
      const int nqc = min(IQN_QC, nq - q0);
      __syncwarp();
      for (int i = lane; i < nqc * IQN_EP; i += 32) {
        const int q = i / IQN_EP, kk = i - q * IQN_EP;
        s_cs[i] = kk < E ? cosf(taus[(q0 + q) * B + b] * 3.14159265358979323846f * (float)(kk + 1)) : 0.f;
      }
      __syncwarp();
      float acc[IQN_QC];
#pragma unroll
      for (int q = 0; q < IQN_QC; ++q) acc[q] = 0.f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int c = lane + 32 * j;
        if (c < C) {
          float we[IQN_EP];
#pragma unroll
          for (int kk = 0; kk < IQN_EP; ++kk) we[kk] = kk < E ? s_we[c * ES + kk] : 0.f;
          const float bec = s_be[c];
#pragma unroll
          for (int q = 0; q < IQN_QC; ++q) {
            if (q < nqc) {
              float pre = bec;
              const float4* cs4 = reinterpret_cast<const float4*>(s_cs + q * IQN_EP);
#pragma unroll
              for (int v = 0; v < IQN_EP / 4; ++v) {
                const float4 cv = cs4[v];
                pre += cv.x * we[4 * v] + cv.y * we[4 * v + 1] + cv.z * we[4 * v + 2] + cv.w * we[4 * v + 3];
              }
              acc[q] += f[j] * fast_tanh(pre);
            }
          }
        }
      }
#pragma unroll
      for (int q = 0; q < IQN_QC; ++q) {
        if (q < nqc) {
          const float p = warp_sum(acc[q]) + bias_o;
          msum += p;
          if (lane == 0) {
            const int r = (q0 + q) * B + b;
            p_tau[r] = p;
            if (target) {
              const float err = tgt - p, mag = fabsf(err);
              const float h = mag <= k ? 0.5f * err * err : k * (mag - 0.5f * k);
              lacc += fabsf(taus[r] - (err < 0.f ? 1.f : 0.f)) * h;
            }
          }
        }
      }
    }
    if (lane == 0 && p_mean) p_mean[b] = msum / (float)nq;
  }
  if (target && loss) {
    __shared__ float s_l[8];
    if (lane == 0) s_l[warp] = lacc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < nwarp; ++w) t += s_l[w];
      atomicAdd(loss, t / (float)B);
    }
  }
}
extern "C" int ttg_iqn_head_loss_fwd(const float* feats, const float* taus, const float* We, const float* be, const float* wo,
                                     const float* bo, const float* target, float* p_tau, float* p_mean, float* loss, int B,
                                     int nq, int C, int E, float k, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  TTG_REQUIRE(E <= IQN_EP, "iqn_head_loss: embedding dims %d > %d", E, IQN_EP);
  TTG_REQUIRE(C >= 1 && C <= 256, "iqn_head_loss: feature dims %d outside [1, 256]", C);
  TTG_REQUIRE(p_tau != nullptr, "iqn_head_loss: p_tau is required (saved for backward)");
  const size_t smem = sizeof(float) * (8 * IQN_QC * IQN_EP + (size_t)C * (E + 1) + 2 * C);
  if (target && loss) cudaMemsetAsync(loss, 0, sizeof(float), st);
  int grid = (B + 7) / 8;
  if (grid > ttg_num_sms() * 8) grid = ttg_num_sms() * 8;
#define TTG_IQN(CPL)                                                                                                    \
  do {                                                                                                                  \
    cudaFuncSetAttribute(iqn_head_loss_fwd_kernel<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
    iqn_head_loss_fwd_kernel<CPL><<<grid, 256, smem, st>>>(feats, taus, We, be, wo, bo, target, p_tau, p_mean, loss, B, nq, C, E, k); \
  } while (0)
  if (C <= 32) TTG_IQN(1); else if (C <= 64) TTG_IQN(2); else if (C <= 128) TTG_IQN(4); else TTG_IQN(8);
#undef TTG_IQN
  TTG_CHECK_LAUNCH("iqn_head_loss_fwd");
  return TTG_OK;
}

// ---------------------------------------------------------------- IQN head backward
// g[r] = cotangent of p_tau[r] (the mean-over-quantiles cotangent is folded in by the caller).
// Lane = channel; each warp walks a slice of the batch; per-channel partials reduced through
// shared memory then fp32 atomics (outputs zeroed here).
// g == nullptr: the cotangent of p_tau[r] is rebuilt on the fly from the cotangents of the head's two outputs,
//   g[r] = g_pmean[b] / nq  +  gloss * d iqn_loss / d p_tau[r]          (either may be null)
// which folds quantile_huber_bwd, the broadcast of the quantile mean's cotangent and their sum into this kernel.
__global__ void __launch_bounds__(256) iqn_head_bwd_kernel(
    const float* __restrict__ g, const float* __restrict__ feats, const float* __restrict__ taus,
    const float* __restrict__ We, const float* __restrict__ be, const float* __restrict__ wo,
    float* __restrict__ gf, float* __restrict__ gWe, float* __restrict__ gbe, float* __restrict__ gwo,
    float* __restrict__ gbo, int B, int nq, int C, int E, const float* __restrict__ g_pmean = nullptr,
    const float* __restrict__ gloss = nullptr, const float* __restrict__ p_tau = nullptr,
    const float* __restrict__ target = nullptr, float hk = 1.f) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const bool live = c < C;
  float we[IQN_MAX_E], gwe[IQN_MAX_E];
#pragma unroll
  for (int k = 0; k < IQN_MAX_E; ++k) { we[k] = (live && k < E) ? We[c * E + k] : 0.f; gwe[k] = 0.f; }
  const float bec = live ? be[c] : 0.f, woc = live ? wo[c] : 0.f;
  float a_gwo = 0.f, a_gbe = 0.f, a_gbo = 0.f;
  for (int b = blockIdx.y * nwarp + warp; b < B; b += gridDim.y * nwarp) {
    const float f = live ? feats[(long long)b * C + c] : 0.f;
    float a_gf = 0.f;
    for (int q = 0; q < nq; ++q) {
      const int r = q * B + b;
      const float tau = taus[r];
      float gr;
      if (g) gr = g[r];
      else {
        gr = g_pmean ? g_pmean[b] / (float)nq : 0.f;
        if (gloss) {
          const float err = target[b] - p_tau[r];
          const float w = fabsf(tau - (err < 0.f ? 1.f : 0.f));
          const float dh = fabsf(err) <= hk ? err : (err > 0.f ? hk : -hk);
          gr -= gloss[0] / (float)B * w * dh;
        }
      }
      float pre = bec, cs[IQN_MAX_E];
      iqn_cosines(tau, E, lane, cs);
#pragma unroll
      for (int k = 0; k < IQN_MAX_E; ++k) pre += cs[k] * we[k];
      const float e = fast_tanh(pre);
      a_gf += gr * e * woc;
      a_gwo += gr * f * e;
      const float gpre = gr * f * woc * (1.f - e * e);
      a_gbe += gpre;
#pragma unroll
      for (int k = 0; k < IQN_MAX_E; ++k) gwe[k] += gpre * cs[k];
      if (blockIdx.x == 0 && lane == 0) a_gbo += gr;
    }
    if (live && gf) gf[(long long)b * C + c] = a_gf;
  }
  if (live) {
    if (gwo) atomicAdd(&gwo[c], a_gwo);
    if (gbe) atomicAdd(&gbe[c], a_gbe);
    if (gWe) {
#pragma unroll
      for (int k = 0; k < IQN_MAX_E; ++k) if (k < E) atomicAdd(&gWe[c * E + k], gwe[k]);
    }
  }
  if (gbo && blockIdx.x == 0 && lane == 0) atomicAdd(gbo, a_gbo);
}
extern "C" int ttg_iqn_head_bwd(const float* g, const float* feats, const float* taus, const float* We, const float* be,
                                const float* wo, float* gf, float* gWe, float* gbe, float* gwo, float* gbo, int B, int nq,
                                int C, int E, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  TTG_REQUIRE(E <= IQN_MAX_E, "iqn_head: embedding dims %d > %d", E, IQN_MAX_E);
  if (gWe) cudaMemsetAsync(gWe, 0, sizeof(float) * (size_t)C * E, st);
  if (gbe) cudaMemsetAsync(gbe, 0, sizeof(float) * C, st);
  if (gwo) cudaMemsetAsync(gwo, 0, sizeof(float) * C, st);
  if (gbo) cudaMemsetAsync(gbo, 0, sizeof(float), st);
  int gy = (B + 7) / 8; if (gy > 64) gy = 64;
  dim3 grid((C + 31) / 32, gy);
  iqn_head_bwd_kernel<<<grid, 256, 0, st>>>(g, feats, taus, We, be, wo, gf, gWe, gbe, gwo, gbo, B, nq, C, E);
  TTG_CHECK_LAUNCH("iqn_head_bwd");
  return TTG_OK;
}

// backward of ttg_iqn_head_loss_fwd in ONE kernel: cotangents g_pmean [B] and gloss [1] (either may be NULL)
extern "C" int ttg_iqn_head_loss_bwd(const float* g_pmean, const float* gloss, const float* p_tau, const float* target,
                                     const float* feats, const float* taus, const float* We, const float* be, const float* wo,
                                     float* gf, float* gWe, float* gbe, float* gwo, float* gbo, int B, int nq, int C, int E,
                                     float k, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  TTG_REQUIRE(E <= IQN_MAX_E, "iqn_head_loss: embedding dims %d > %d", E, IQN_MAX_E);
  TTG_REQUIRE(gloss == nullptr || (p_tau != nullptr && target != nullptr), "iqn_head_loss_bwd: the loss cotangent needs p_tau and target");
  if (gWe) cudaMemsetAsync(gWe, 0, sizeof(float) * (size_t)C * E, st);
  if (gbe) cudaMemsetAsync(gbe, 0, sizeof(float) * C, st);
  if (gwo) cudaMemsetAsync(gwo, 0, sizeof(float) * C, st);
  if (gbo) cudaMemsetAsync(gbo, 0, sizeof(float), st);
  int gy = (B + 7) / 8; if (gy > 64) gy = 64;
  dim3 grid((C + 31) / 32, gy);
  iqn_head_bwd_kernel<<<grid, 256, 0, st>>>(nullptr, feats, taus, We, be, wo, gf, gWe, gbe, gwo, gbo, B, nq, C, E, g_pmean, gloss, p_tau,
                                            target, k);
  TTG_CHECK_LAUNCH("iqn_head_loss_bwd");
  return TTG_OK;
}

// ---------------------------------------------------------------- quantile Huber loss
__device__ __forceinline__ float block_sum_256(float v) {
  __shared__ float s[8];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < 8 ? s[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) t = warp_sum(t);
  __syncthreads();
  return t;   // valid in thread 0 (and warp 0)
}
// loss = (1/B) * sum_{q,b} |tau - 1[err<0]| * huber_k(err),  err = target[b] - p[q,b]
__global__ void __launch_bounds__(256) quantile_huber_fwd_kernel(const float* __restrict__ p, const float* __restrict__ target,
                                                                   const float* __restrict__ taus, float* __restrict__ loss,
                                                                   int B, int R, float k) {
  float acc = 0.f;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x) {
    float err = target[r % B] - p[r];
    float mag = fabsf(err);
    float h = mag <= k ? 0.5f * err * err : k * (mag - 0.5f * k);
    acc += fabsf(taus[r] - (err < 0.f ? 1.f : 0.f)) * h;
  }
  acc = block_sum_256(acc);
  if (threadIdx.x == 0) atomicAdd(loss, acc / (float)B);
}
__global__ void quantile_huber_bwd_kernel(const float* __restrict__ p, const float* __restrict__ target,
                                          const float* __restrict__ taus, const float* __restrict__ gloss,
                                          float* __restrict__ gp, int B, int R, float k) {
  const float gl = gloss[0] / (float)B;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < R; r += gridDim.x * blockDim.x) {
    float err = target[r % B] - p[r];
    float w = fabsf(taus[r] - (err < 0.f ? 1.f : 0.f));
    float dh = fabsf(err) <= k ? err : (err > 0.f ? k : -k);
    gp[r] = -gl * w * dh;
  }
}
extern "C" int ttg_quantile_huber_fwd(const float* p_tau, const float* target, const float* taus, float* loss, int B, int nq,
                                      float k, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(loss, 0, sizeof(float), st);
  quantile_huber_fwd_kernel<<<ttg_grid_for(B * nq, 256, 1), 256, 0, st>>>(p_tau, target, taus, loss, B, B * nq, k);
  TTG_CHECK_LAUNCH("quantile_huber_fwd");
  return TTG_OK;
}
extern "C" int ttg_quantile_huber_bwd(const float* p_tau, const float* target, const float* taus, const float* gloss,
                                      float* gp, int B, int nq, float k, void* stream) {
  quantile_huber_bwd_kernel<<<ttg_grid_for(B * nq, 256, 1), 256, 0, (cudaStream_t)stream>>>(p_tau, target, taus, gloss, gp, B, B * nq, k);
  TTG_CHECK_LAUNCH("quantile_huber_bwd");
  return TTG_OK;
}

// ---------------------------------------------------------------- BCE with logits (mean)
__global__ void __launch_bounds__(256) bce_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                      float* __restrict__ loss, int n) {
  float acc = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float v = x[i];
    acc += fmaxf(v, 0.f) - v * y[i] + log1pf(expf(-fabsf(v)));
  }
  acc = block_sum_256(acc);
  if (threadIdx.x == 0) atomicAdd(loss, acc / (float)n);
}
__global__ void bce_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gloss,
                               float* __restrict__ gx, int n) {
  const float gl = gloss[0] / (float)n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    gx[i] = gl * (1.f / (1.f + expf(-x[i])) - y[i]);
}
extern "C" int ttg_bce_logits_fwd(const float* x, const float* y, float* loss, int n, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(loss, 0, sizeof(float), st);
  bce_fwd_kernel<<<ttg_grid_for(n, 256, 1), 256, 0, st>>>(x, y, loss, n);
  TTG_CHECK_LAUNCH("bce_logits_fwd");
  return TTG_OK;
}
extern "C" int ttg_bce_logits_bwd(const float* x, const float* y, const float* gloss, float* gx, int n, void* stream) {
  bce_bwd_kernel<<<ttg_grid_for(n, 256, 1), 256, 0, (cudaStream_t)stream>>>(x, y, gloss, gx, n);
  TTG_CHECK_LAUNCH("bce_logits_bwd");
  return TTG_OK;
}

// ---------------------------------------------------------------- sum of squares: out = scale * sum x^2
__global__ void __launch_bounds__(256) sqsum_kernel(const float* __restrict__ x, double* __restrict__ ws, long long n) {
  float acc = 0.f;
  const long long n4 = n / 4;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = x4[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) acc += x[i] * x[i];
  acc = block_sum_256(acc);
  if (threadIdx.x == 0) atomicAdd(ws, (double)acc);
}
__global__ void sqsum_finalize_kernel(const double* ws, float scale, float* out) { out[0] = (float)(ws[0] * (double)scale); }
extern "C" int ttg_sqsum_f32(const float* x, long long n, float scale, float* out, void* workspace, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  TTG_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "sqsum: input must be 16-byte aligned");
  double* ws = (double*)workspace;
  cudaMemsetAsync(ws, 0, sizeof(double), st);
  sqsum_kernel<<<ttg_grid_for(n / 4 + 1, 1024, 4), 256, 0, st>>>(x, ws, n);
  TTG_CHECK_LAUNCH("sqsum");
  sqsum_finalize_kernel<<<1, 1, 0, st>>>(ws, scale, out);
  TTG_CHECK_LAUNCH("sqsum_finalize");
  return TTG_OK;
}

// ---------------------------------------------------------------- small fp32 matmul
// C[M,N] = op(A)[M,K] * op(B)[K,N] (+ bias[N]);  row-major.  64x64 output tile per 256-thread block, 4x4 outputs per
// thread, K in steps of 16 through shared memory (the generator's input Linear 256 x 256 -> 2048 and its two gradient
// products: 61 us with the one-output-per-thread 16x16 kernel of round 1).  Operand tiles are read along their
// contiguous dimension whatever the transpose flags are.
#define MM_BM 64
#define MM_BN 64
#define MM_BK 16
__global__ void __launch_bounds__(256) matmul_f32_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                         const float* __restrict__ bias, float* __restrict__ Cm, int M,
                                                         int N, int K, int ta, int tb) {
  __shared__ float sa[MM_BK][MM_BM + 4], sb[MM_BK][MM_BN + 4];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * MM_BM, n0 = blockIdx.x * MM_BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += MM_BK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = tid + r * 256;
      int m, k;
      if (ta) { m = idx % MM_BM; k = idx / MM_BM; } else { k = idx % MM_BK; m = idx / MM_BK; }
      const int gm = m0 + m, gk = k0 + k;
      sa[k][m] = (gm < M && gk < K) ? (ta ? A[(long long)gk * M + gm] : A[(long long)gm * K + gk]) : 0.f;
      int n, kb;
      if (tb) { kb = idx % MM_BK; n = idx / MM_BK; } else { n = idx % MM_BN; kb = idx / MM_BN; }
      const int gn = n0 + n, gkb = k0 + kb;
      sb[kb][n] = (gn < N && gkb < K) ? (tb ? Bm[(long long)gn * K + gkb] : Bm[(long long)gkb * N + gn]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < MM_BK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sa[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sb[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col < N) Cm[(long long)row * N + col] = acc[i][j] + (bias ? bias[col] : 0.f);
    }
  }
}
extern "C" int ttg_matmul_f32(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int transA,
                              int transB, void* stream) {
  TTG_REQUIRE(M > 0 && N > 0 && K > 0, "matmul: empty operand");
  dim3 grid((N + MM_BN - 1) / MM_BN, (M + MM_BM - 1) / MM_BM);
  matmul_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, B, bias, C, M, N, K, transA, transB);
  TTG_CHECK_LAUNCH("matmul_f32");
  return TTG_OK;
}
// out[n] = scale * sum_m x[m,n]
__global__ void colsum_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int M, int N, float scale) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < N) { float s = 0.f; for (int m = 0; m < M; ++m) s += x[(long long)m * N + n]; out[n] = s * scale; }
}
extern "C" int ttg_colsum_f32(const float* x, float* out, int M, int N, float scale, void* stream) {
  colsum_f32_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(x, out, M, N, scale);
  TTG_CHECK_LAUNCH("colsum_f32");
  return TTG_OK;
}
// out[m,n] = scale * g[n]
__global__ void rowbcast_f32_kernel(const float* __restrict__ g, float* __restrict__ out, int M, int N, float scale) {
  long long total = (long long)M * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) out[i] = scale * g[i % N];
}
extern "C" int ttg_rowbcast_f32(const float* g, float* out, int M, int N, float scale, void* stream) {
  rowbcast_f32_kernel<<<ttg_grid_for((long long)M * N, 1024), 256, 0, (cudaStream_t)stream>>>(g, out, M, N, scale);
  TTG_CHECK_LAUNCH("rowbcast_f32");
  return TTG_OK;
}
