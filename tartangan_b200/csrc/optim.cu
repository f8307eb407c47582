// Fused flat Adam step (+ EMA of the target generator) over one contiguous fp32 parameter
// buffer.  Replaces the per-tensor lerp_/addcmul_/sqrt/addcdiv_ chains of torch.optim.Adam
// (reference trainers/cnn.py:84-85, trainers/iqn.py:84-85: betas=(0, 0.999), eps 1e-8) and the
// per-parameter EMA loop of update_target_generator (trainers/cnn.py:158-165).  One HBM stream:
// reads p,g,m,v(,t) once, writes p,m,v(,t) once.
#include "common.cuh"

__global__ void adam_tick_kernel(float* step) { step[0] += 1.f; }

__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v,
                                                        float* __restrict__ ema, long long n, float lr, float b1,
                                                        float b2, float eps, float ema_lr,
                                                        const float* __restrict__ step_ptr) {
  const float step = step_ptr[0];
  const float bc1 = 1.f - powf(b1, step);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, step));
  const float step_size = lr / bc1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i];
    float mi = m[i] + (gi - m[i]) * (1.f - b1);          // exp_avg.lerp_(grad, 1-beta1)
    float vi = v[i] * b2 + (1.f - b2) * gi * gi;         // exp_avg_sq.mul_(b2).addcmul_(g,g,1-b2)
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    float pi = p[i] - step_size * (mi / denom);
    m[i] = mi; v[i] = vi; p[i] = pi;
    if (ema) { float t = ema[i]; ema[i] = t + (pi - t) * ema_lr; }
  }
}

// step: device float, incremented here before use (graph-capturable, no host state).
extern "C" int ttg_adam_flat(float* p, const float* g, float* m, float* v, float* ema, long long n, float lr, float beta1,
                             float beta2, float eps, float ema_lr, float* step, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  adam_tick_kernel<<<1, 1, 0, st>>>(step);
  TTG_CHECK_LAUNCH("adam_tick");
  adam_flat_kernel<<<ttg_grid_for(n, 1024, 4), 256, 0, st>>>(p, g, m, v, ema, n, lr, beta1, beta2, eps, ema_lr, step);
  TTG_CHECK_LAUNCH("adam_flat");
  return TTG_OK;
}

// target += (src - target) * lr   (EMA without an optimiser step)
__global__ void ema_flat_kernel(float* __restrict__ t, const float* __restrict__ s, long long n, float lr) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float a = t[i]; t[i] = a + (s[i] - a) * lr;
  }
}
extern "C" int ttg_ema_flat(float* target, const float* src, long long n, float lr, void* stream) {
  if (n == 0) return TTG_OK;
  ema_flat_kernel<<<ttg_grid_for(n, 1024, 4), 256, 0, (cudaStream_t)stream>>>(target, src, n, lr);
  TTG_CHECK_LAUNCH("ema_flat");
  return TTG_OK;
}
