// Generic NHWC channel-wise drivers: a per-channel reduction over rows and an
// element-wise map with per-channel parameters.  Both are pure HBM streams:
// 16-byte vector accesses, coalesced along the channel (fastest) dimension,
// grid sized in multiples of the SM count (grid-stride loops).
#pragma once
#include "common.cuh"

// Op concept for reductions:
//   static constexpr int NIN, NACC;
//   const T* in[NIN];
//   __device__ void acc(const float* v /*NIN*/, int c, float* a /*NACC*/) const;
template <typename T, int V, class Op>
__global__ void __launch_bounds__(256) chan_reduce_kernel(Op op, long long M, int C, double* __restrict__ out) {
  extern __shared__ float s_acc[];   // [NACC][C]
  for (int i = threadIdx.x; i < Op::NACC * C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int groups = C / V;
  const int rpi = blockDim.x / groups;           // rows per block iteration
  const int g = threadIdx.x % groups, rsub = threadIdx.x / groups;
  float a[Op::NACC][V];
#pragma unroll
  for (int k = 0; k < Op::NACC; ++k)
#pragma unroll
    for (int j = 0; j < V; ++j) a[k][j] = 0.f;
  if (rsub < rpi) {
    for (long long r = (long long)blockIdx.x * rpi + rsub; r < M; r += (long long)gridDim.x * rpi) {
      const long long base = r * C + (long long)g * V;
      float v[Op::NIN][V];
#pragma unroll
      for (int t = 0; t < Op::NIN; ++t) {
        if constexpr (V == 1) v[t][0] = to_f(op.in[t][base]);
        else { Vec<T> q; q.load(op.in[t] + base); q.unpack(v[t]); }
      }
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float vin[Op::NIN], acc[Op::NACC];
#pragma unroll
        for (int t = 0; t < Op::NIN; ++t) vin[t] = v[t][j];
#pragma unroll
        for (int k = 0; k < Op::NACC; ++k) acc[k] = a[k][j];
        op.acc(vin, g * V + j, acc);
#pragma unroll
        for (int k = 0; k < Op::NACC; ++k) a[k][j] = acc[k];
      }
    }
#pragma unroll
    for (int k = 0; k < Op::NACC; ++k)
#pragma unroll
      for (int j = 0; j < V; ++j) atomicAdd(&s_acc[k * C + g * V + j], a[k][j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Op::NACC * C; i += blockDim.x) atomicAdd(&out[i], (double)s_acc[i]);
}

template <typename T> static inline bool ttg_vec_ok(int C, const void* const* ptrs, int n) {
  if (C % Vec<T>::N) return false;
  for (int i = 0; i < n; ++i)
    if (reinterpret_cast<uintptr_t>(ptrs[i]) & 15) return false;
  return true;
}

// out must hold NACC*C doubles; zeroed here.
template <typename T, class Op>
static int launch_chan_reduce(const char* name, Op op, long long M, int C, double* out, cudaStream_t st) {
  cudaMemsetAsync(out, 0, sizeof(double) * Op::NACC * C, st);
  const void* ptrs[Op::NIN];
  for (int i = 0; i < Op::NIN; ++i) ptrs[i] = op.in[i];
  size_t smem = sizeof(float) * Op::NACC * C;
  if (ttg_vec_ok<T>(C, ptrs, Op::NIN) && C / Vec<T>::N <= 256) {
    constexpr int V = Vec<T>::N;
    int rpi = 256 / (C / V);
    int grid = ttg_grid_for(M, rpi * 8, 4);
    chan_reduce_kernel<T, V, Op><<<grid, 256, smem, st>>>(op, M, C, out);
  } else {
    if (C > 256) return ttg_set_error(TTG_ERR_UNSUPPORTED, "%s: C=%d needs C%%%d==0 and 16B alignment", name, C, Vec<T>::N);
    int rpi = 256 / C;
    int grid = ttg_grid_for(M, rpi * 8, 4);
    chan_reduce_kernel<T, 1, Op><<<grid, 256, smem, st>>>(op, M, C, out);
  }
  TTG_CHECK_LAUNCH(name);
  return TTG_OK;
}

// Op concept for maps:
//   static constexpr int NIN, NOUT;  const T* in[NIN];  T* out[NOUT];
//   __device__ void apply(const float* v, int c, float* o) const;
template <typename T, int V, class Op>
__global__ void __launch_bounds__(256) chan_map_kernel(Op op, long long nvec, int C) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const long long base = i * V;
    const int c0 = (int)(base % C);
    float v[Op::NIN][V], o[Op::NOUT][V];
#pragma unroll
    for (int t = 0; t < Op::NIN; ++t) {
      if constexpr (V == 1) v[t][0] = to_f(op.in[t][base]);
      else { Vec<T> q; q.load(op.in[t] + base); q.unpack(v[t]); }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float vin[Op::NIN], vout[Op::NOUT];
#pragma unroll
      for (int t = 0; t < Op::NIN; ++t) vin[t] = v[t][j];
      op.apply(vin, c0 + j, vout);
#pragma unroll
      for (int t = 0; t < Op::NOUT; ++t) o[t][j] = vout[t];
    }
#pragma unroll
    for (int t = 0; t < Op::NOUT; ++t) {
      if constexpr (V == 1) op.out[t][base] = from_f<T>(o[t][0]);
      else { Vec<T> q; q.pack(o[t]); q.store(op.out[t] + base); }
    }
  }
}

template <typename T, class Op>
static int launch_chan_map(const char* name, Op op, long long n, int C, cudaStream_t st) {
  const void* ptrs[Op::NIN + Op::NOUT];
  for (int i = 0; i < Op::NIN; ++i) ptrs[i] = op.in[i];
  for (int i = 0; i < Op::NOUT; ++i) ptrs[Op::NIN + i] = op.out[i];
  if (n == 0) return TTG_OK;
  if (ttg_vec_ok<T>(C, ptrs, Op::NIN + Op::NOUT)) {
    constexpr int V = Vec<T>::N;
    long long nvec = n / V;
    chan_map_kernel<T, V, Op><<<ttg_grid_for(nvec, 256 * 2), 256, 0, st>>>(op, nvec, C);
  } else {
    chan_map_kernel<T, 1, Op><<<ttg_grid_for(n, 256 * 4), 256, 0, st>>>(op, n, C);
  }
  TTG_CHECK_LAUNCH(name);
  return TTG_OK;
}
