// Shared helpers for the tartangan_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>

#define TTG_OK 0
#define TTG_ERR_ARG 1
#define TTG_ERR_CUDA 2
#define TTG_ERR_UNSUPPORTED 3

#define TTG_F32 0
#define TTG_BF16 1

typedef __nv_bfloat16 bf16;

extern "C" const char* ttg_last_error(void);
int ttg_set_error(int code, const char* fmt, ...);

#define TTG_CHECK_LAUNCH(name)                                                      \
  do {                                                                              \
    cudaError_t e__ = cudaPeekAtLastError();                                        \
    if (e__ != cudaSuccess) {                                                       \
      (void)cudaGetLastError();                                                     \
      return ttg_set_error(TTG_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e__));  \
    }                                                                               \
  } while (0)

#define TTG_REQUIRE(cond, ...)                                   \
  do {                                                           \
    if (!(cond)) return ttg_set_error(TTG_ERR_ARG, __VA_ARGS__); \
  } while (0)

// Dispatch on the activation dtype code.
#define TTG_DISPATCH(dtype, ...)                                                       \
  do {                                                                                 \
    if ((dtype) == TTG_F32) { typedef float T; __VA_ARGS__; }                          \
    else if ((dtype) == TTG_BF16) { typedef bf16 T; __VA_ARGS__; }                     \
    else return ttg_set_error(TTG_ERR_ARG, "bad dtype code %d", (int)(dtype));         \
  } while (0)

static inline int ttg_num_sms() { return 148; }   // B200
static inline int ttg_grid_for(long long work_items, int per_block, int max_waves = 8) {
  long long b = (work_items + per_block - 1) / per_block;
  long long cap = (long long)ttg_num_sms() * max_waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// Grid for a grid-stride kernel: enough blocks for the work, at most ONE full wave at the kernel's real occupancy
// (a fixed "8 blocks per SM" cap leaves a ragged second wave when registers limit a kernel to 5-6 blocks per SM).
#include <mutex>
#include <unordered_map>
static inline int ttg_blocks_per_sm(const void* fn, int block, size_t smem) {
  static std::mutex mu;
  static std::unordered_map<unsigned long long, int> cache;
  const unsigned long long key = (unsigned long long)(uintptr_t)fn ^ ((unsigned long long)block << 48) ^ ((unsigned long long)smem << 20);
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, block, smem) != cudaSuccess || n < 1) { (void)cudaGetLastError(); n = 4; }
  cache[key] = n;
  return n;
}
template <class Kern>
static inline int ttg_grid_occ(Kern kernel, long long work_items, int per_block, int block = 256, size_t smem = 0) {
  long long b = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)ttg_num_sms() * ttg_blocks_per_sm((const void*)kernel, block, smem);
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ---------------------------------------------------------------- device side
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 16-byte vectors: 4 floats or 8 bf16.
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  float4 raw;
  __device__ __forceinline__ void load(const float* p) { raw = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = raw; }
  __device__ __forceinline__ void unpack(float* f) const { f[0] = raw.x; f[1] = raw.y; f[2] = raw.z; f[3] = raw.w; }
  __device__ __forceinline__ void pack(const float* f) { raw = make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct Vec<bf16> {
  static constexpr int N = 8;
  uint4 raw;
  __device__ __forceinline__ void load(const bf16* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(bf16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ void unpack(float* f) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  }
  __device__ __forceinline__ void pack(const float* f) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  }
};

// N-element vectors (N chosen per kernel: 16-byte vectors by default, 8-byte bf16 vectors for register-heavy ops)
template <typename T, int N> struct VecN;
template <> struct VecN<float, 4> : Vec<float> {};
template <> struct VecN<bf16, 8> : Vec<bf16> {};
template <> struct VecN<bf16, 4> {
  static constexpr int N = 4;
  uint2 raw;
  __device__ __forceinline__ void load(const bf16* p) { raw = *reinterpret_cast<const uint2*>(p); }
  __device__ __forceinline__ void store(bf16* p) const { *reinterpret_cast<uint2*>(p) = raw; }
  __device__ __forceinline__ void unpack(float* f) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 2; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  }
  __device__ __forceinline__ void pack(const float* f) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 2; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }
__device__ __forceinline__ float lrelu_mask(float v, float slope) { return v > 0.f ? 1.f : slope; }
