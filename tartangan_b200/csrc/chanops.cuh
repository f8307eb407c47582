// Generic NHWC channel-wise drivers: a per-channel reduction over rows and an
// element-wise map with per-channel parameters.  Both are pure HBM streams:
// 16-byte vector accesses, coalesced along the channel (fastest) dimension,
// grid sized in multiples of the SM count (grid-stride loops).
#pragma once
#include "common.cuh"
#include <type_traits>

// Vector width per (dtype, Op): 16-byte vectors unless the Op asks for 4-element bf16 vectors (`BF16_V = 4`).
// Measured on B200 (bn_act_bwd, M = 4 Mi, C = 16): 8 channels per thread (80 / 105 registers, 3 / 2 blocks per SM)
// 182 us; 4 channels per thread (60 / 63 registers, 4 blocks per SM) 203 us -> every op keeps 16-byte vectors.
template <class Op, class = void> struct OpBf16V { static constexpr int value = 8; };
template <class Op> struct OpBf16V<Op, std::void_t<decltype(Op::BF16_V)>> { static constexpr int value = Op::BF16_V; };
template <typename T, class Op> struct ChanV { static constexpr int value = Vec<T>::N; };
template <class Op> struct ChanV<bf16, Op> { static constexpr int value = OpBf16V<Op>::value; };

// Op concept for reductions:
//   static constexpr int NIN, NACC;   const T* in[NIN];
//   template <int V> struct P { ... };                       per-thread cache of per-channel parameters
//   template <int V> __device__ void load(int c0, P<V>&) const;   (channels c0..c0+V-1, loaded ONCE per thread)
//   template <int V> __device__ void acc(const float* v /*NIN*/, int j, const P<V>&, float* a /*NACC*/) const;
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
    typename Op::template P<V> prm;
    op.template load<V>(g * V, prm);
    constexpr int U = 2;                           // rows in flight per thread (4 measured slower: registers)
    const long long rstep = (long long)gridDim.x * rpi;
    long long r0 = (long long)blockIdx.x * rpi + rsub;
    for (; r0 + (U - 1) * rstep < M; r0 += U * rstep) {
      float v[U][Op::NIN][V];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long base = (r0 + u * rstep) * C + (long long)g * V;
#pragma unroll
        for (int t = 0; t < Op::NIN; ++t) {
          if constexpr (V == 1) v[u][t][0] = to_f(op.in[t][base]);
          else { VecN<T, V> q; q.load(op.in[t] + base); q.unpack(v[u][t]); }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < V; ++j) {
          float vin[Op::NIN], acc[Op::NACC];
#pragma unroll
          for (int t = 0; t < Op::NIN; ++t) vin[t] = v[u][t][j];
#pragma unroll
          for (int k = 0; k < Op::NACC; ++k) acc[k] = a[k][j];
          op.template acc<V>(vin, j, prm, acc);
#pragma unroll
          for (int k = 0; k < Op::NACC; ++k) a[k][j] = acc[k];
        }
    }
    for (; r0 < M; r0 += rstep) {
      const long long base = r0 * C + (long long)g * V;
      float v[Op::NIN][V];
#pragma unroll
      for (int t = 0; t < Op::NIN; ++t) {
        if constexpr (V == 1) v[t][0] = to_f(op.in[t][base]);
        else { VecN<T, V> q; q.load(op.in[t] + base); q.unpack(v[t]); }
      }
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float vin[Op::NIN], acc[Op::NACC];
#pragma unroll
        for (int t = 0; t < Op::NIN; ++t) vin[t] = v[t][j];
#pragma unroll
        for (int k = 0; k < Op::NACC; ++k) acc[k] = a[k][j];
        op.template acc<V>(vin, j, prm, acc);
#pragma unroll
        for (int k = 0; k < Op::NACC; ++k) a[k][j] = acc[k];
      }
    }
  }
  // block reduction: lanes that own the same channel group are first combined with warp shuffles
  // (power-of-two group counts), then one shared-memory atomic per (warp, channel)
  const bool pow2 = (groups & (groups - 1)) == 0 && groups <= 32;
  if (pow2) {
#pragma unroll
    for (int k = 0; k < Op::NACC; ++k)
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float t = a[k][j];
        for (int o = 16; o >= groups; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        a[k][j] = t;
      }
  }
  if (rsub < rpi && (!pow2 || (threadIdx.x & 31) < groups)) {
#pragma unroll
    for (int k = 0; k < Op::NACC; ++k)
#pragma unroll
      for (int j = 0; j < V; ++j) atomicAdd(&s_acc[k * C + g * V + j], a[k][j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Op::NACC * C; i += blockDim.x) atomicAdd(&out[i], (double)s_acc[i]);
}

template <typename T> static inline bool ttg_vec_ok(int C, const void* const* ptrs, int n, int vec = Vec<T>::N) {
  if (C % vec) return false;
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
  constexpr int VV = ChanV<T, Op>::value;
  if (ttg_vec_ok<T>(C, ptrs, Op::NIN, VV) && C / VV <= 256) {
    constexpr int V = VV;
    int rpi = 256 / (C / V);
    int grid = ttg_grid_occ(chan_reduce_kernel<T, V, Op>, M, rpi * 4, 256, smem);
    chan_reduce_kernel<T, V, Op><<<grid, 256, smem, st>>>(op, M, C, out);
  } else {
    if (C > 256) return ttg_set_error(TTG_ERR_UNSUPPORTED, "%s: C=%d needs C%%%d==0 and 16B alignment", name, C, Vec<T>::N);
    int rpi = 256 / C;
    int grid = ttg_grid_occ(chan_reduce_kernel<T, 1, Op>, M, rpi * 8, 256, smem);
    chan_reduce_kernel<T, 1, Op><<<grid, 256, smem, st>>>(op, M, C, out);
  }
  TTG_CHECK_LAUNCH(name);
  return TTG_OK;
}

// Op concept for maps:
//   static constexpr int NIN, NOUT;  const T* in[NIN];  T* out[NOUT];
//   template <int V> struct P;  load<V>(c0, P&);  apply<V>(const float* v, int j, const P<V>&, float* o)
// When the grid stride keeps every thread on the same channel group (`invariant`), the per-channel
// parameters are loaded once per thread instead of once per element.
// `reverse`: walk the tensor from its end.  A map pass that follows a reduction over the same tensors then starts
// with the part of them the reduction touched last, i.e. the part that is still resident in the 126 MB L2.
template <typename T, int V, class Op>
__global__ void __launch_bounds__(256) chan_map_kernel(Op op, long long nvec, int C, int invariant, int reverse) {
  constexpr int U = 2;                             // vectors in flight per thread
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  typename Op::template P<V> prm;
  const long long last = nvec - 1;
  if (invariant && i0 < nvec) op.template load<V>((int)(((reverse ? last - i0 : i0) * V) % C), prm);
  for (; i0 < nvec; i0 += U * stride) {
    float v[U][Op::NIN][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long ii = i0 + u * stride;
      const long long i = reverse ? last - ii : ii;
      if (ii < nvec) {
#pragma unroll
        for (int t = 0; t < Op::NIN; ++t) {
          if constexpr (V == 1) v[u][t][0] = to_f(op.in[t][i]);
          else { VecN<T, V> q; q.load(op.in[t] + i * V); q.unpack(v[u][t]); }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long ii = i0 + u * stride;
      const long long i = reverse ? last - ii : ii;
      if (ii < nvec) {
        const long long base = i * V;
        if (!invariant) op.template load<V>((int)(base % C), prm);
        float o[Op::NOUT][V];
#pragma unroll
        for (int j = 0; j < V; ++j) {
          float vin[Op::NIN], vout[Op::NOUT];
#pragma unroll
          for (int t = 0; t < Op::NIN; ++t) vin[t] = v[u][t][j];
          op.template apply<V>(vin, j, prm, vout);
#pragma unroll
          for (int t = 0; t < Op::NOUT; ++t) o[t][j] = vout[t];
        }
#pragma unroll
        for (int t = 0; t < Op::NOUT; ++t) {
          if constexpr (V == 1) op.out[t][base] = from_f<T>(o[t][0]);
          else { VecN<T, V> q; q.pack(o[t]); q.store(op.out[t] + base); }
        }
      }
    }
  }
}

template <typename T, class Op>
static int launch_chan_map(const char* name, Op op, long long n, int C, cudaStream_t st, int reverse = 0) {
  const void* ptrs[Op::NIN + Op::NOUT];
  for (int i = 0; i < Op::NIN; ++i) ptrs[i] = op.in[i];
  for (int i = 0; i < Op::NOUT; ++i) ptrs[Op::NIN + i] = op.out[i];
  if (n == 0) return TTG_OK;
  if (ttg_vec_ok<T>(C, ptrs, Op::NIN + Op::NOUT, ChanV<T, Op>::value)) {
    constexpr int V = ChanV<T, Op>::value;
    long long nvec = n / V;
    int grid = ttg_grid_occ(chan_map_kernel<T, V, Op>, nvec, 256 * 2);
    int inv = ((long long)grid * 256 * V) % C == 0;
    chan_map_kernel<T, V, Op><<<grid, 256, 0, st>>>(op, nvec, C, inv, reverse);
  } else {
    int grid = ttg_grid_occ(chan_map_kernel<T, 1, Op>, n, 256 * 4);
    int inv = ((long long)grid * 256) % C == 0;
    chan_map_kernel<T, 1, Op><<<grid, 256, 0, st>>>(op, n, C, inv, reverse);
  }
  TTG_CHECK_LAUNCH(name);
  return TTG_OK;
}
