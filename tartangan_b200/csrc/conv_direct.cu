// Exact (fp32-accumulate, CUDA-core) direct convolution over NHWC activations,
// k in {1,3}, stride 1, pad k/2.  This is the parity path (fp32 mode) and the
// path for channel counts the tensor-core kernels do not take (RGB layers:
// Cin==3 first D conv / Cout==3 last G conv; channels not a multiple of 16).
// Replaces aten::convolution / convolution_backward for every nn.Conv2d in
// reference models/blocks/{generator,discriminator,attention}.py.
//   fprop : y = conv(x, Wp)          Wp packed [tap][Cin][Cout] fp32
//   dgrad : the same kernel with Wp packed from the flipped/transposed filter
//   wgrad : gw[Cout][Cin][k][k] = sum_pixels gy[p,co] * x[p+tap,ci]
// Optional nearest x2 upsample folded into the input addressing (up=1): the
// kernel reads x at (y>>1, x>>1), so the upsampled tensor is never written.
#include "common.cuh"

#define CD_TH 8
#define CD_TW 16
#define CD_CK 8      // input-channel chunk staged in shared memory
#define CD_CO 16     // output channels per block

// mode 0: wp[tap][ci][co] = w[co][ci][ky][kx]            (fprop; Cin_p=Cin, Cout_p=Cout)
// mode 1: wp[tap][co][ci] = w[co][ci][k-1-ky][k-1-kx]    (dgrad; Cin_p=Cout, Cout_p=Cin)
__global__ void pack_weight_direct_kernel(const float* __restrict__ w, float* __restrict__ wp, int Cout, int Cin, int k, int mode) {
  int total = Cout * Cin * k * k;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int kx = i % k, ky = (i / k) % k, ci = (i / (k * k)) % Cin, co = i / (k * k * Cin);
    float v = w[i];
    if (mode == 0) wp[((ky * k + kx) * Cin + ci) * Cout + co] = v;
    else wp[(((k - 1 - ky) * k + (k - 1 - kx)) * Cout + co) * Cin + ci] = v;
  }
}
extern "C" int ttg_pack_weight_direct(const float* w, float* wp, int Cout, int Cin, int ksize, int mode, void* stream) {
  int total = Cout * Cin * ksize * ksize;
  pack_weight_direct_kernel<<<ttg_grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(w, wp, Cout, Cin, ksize, mode);
  TTG_CHECK_LAUNCH("pack_weight_direct");
  return TTG_OK;
}

template <typename TI, typename TO, int K>
__global__ void __launch_bounds__(CD_TH * CD_TW) conv_direct_kernel(
    const TI* __restrict__ x, const float* __restrict__ wp, const float* __restrict__ bias, TO* __restrict__ y,
    int H, int W, int Cin, int Cout, int up) {
  constexpr int HALO = K / 2;
  constexpr int SH = CD_TH + 2 * HALO, SW = CD_TW + 2 * HALO;
  __shared__ float s_x[CD_CK][SH][SW + 1];
  __shared__ __align__(16) float s_w[K * K][CD_CK][CD_CO];
  const int tiles_x = (W + CD_TW - 1) / CD_TW;
  const int tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
  const int n = blockIdx.z, co0 = blockIdx.y * CD_CO;
  const int tx = threadIdx.x % CD_TW, ty = threadIdx.x / CD_TW;
  const int oy = tile_y * CD_TH + ty, ox = tile_x * CD_TW + tx;
  const int Hi = H >> up, Wi = W >> up;
  float acc[CD_CO];
#pragma unroll
  for (int j = 0; j < CD_CO; ++j) acc[j] = 0.f;

  for (int c0 = 0; c0 < Cin; c0 += CD_CK) {
    // stage the input halo tile: idx -> (pixel, ci) with ci fastest (NHWC contiguous)
    for (int i = threadIdx.x; i < SH * SW * CD_CK; i += blockDim.x) {
      int ci = i % CD_CK, p = i / CD_CK;
      int sx = p % SW, sy = p / SW;
      int gy = tile_y * CD_TH + sy - HALO, gx = tile_x * CD_TW + sx - HALO;
      float v = 0.f;
      if (gy >= 0 && gy < H && gx >= 0 && gx < W && c0 + ci < Cin)
        v = to_f(x[(((long long)n * Hi + (gy >> up)) * Wi + (gx >> up)) * Cin + c0 + ci]);
      s_x[ci][sy][sx] = v;
    }
    for (int i = threadIdx.x; i < K * K * CD_CK * CD_CO; i += blockDim.x) {
      int co = i % CD_CO, ci = (i / CD_CO) % CD_CK, tap = i / (CD_CO * CD_CK);
      float v = 0.f;
      if (c0 + ci < Cin && co0 + co < Cout) v = wp[((long long)tap * Cin + c0 + ci) * Cout + co0 + co];
      s_w[tap][ci][co] = v;
    }
    __syncthreads();
#pragma unroll
    for (int ky = 0; ky < K; ++ky)
#pragma unroll
      for (int kx = 0; kx < K; ++kx)
#pragma unroll
        for (int ci = 0; ci < CD_CK; ++ci) {
          float xv = s_x[ci][ty + ky][tx + kx];
          const float4* wv = reinterpret_cast<const float4*>(&s_w[ky * K + kx][ci][0]);
#pragma unroll
          for (int q = 0; q < CD_CO / 4; ++q) {
            float4 w4 = wv[q];
            acc[4 * q + 0] += xv * w4.x; acc[4 * q + 1] += xv * w4.y;
            acc[4 * q + 2] += xv * w4.z; acc[4 * q + 3] += xv * w4.w;
          }
        }
    __syncthreads();
  }
  if (oy < H && ox < W) {
    TO* dst = y + (((long long)n * H + oy) * W + ox) * Cout + co0;
#pragma unroll
    for (int j = 0; j < CD_CO; ++j)
      if (co0 + j < Cout) dst[j] = from_f<TO>(acc[j] + (bias ? bias[co0 + j] : 0.f));
  }
}

template <typename TI, typename TO>
static int launch_conv_direct(const void* x, const float* wp, const float* bias, void* y, int N, int H, int W, int Cin,
                              int Cout, int k, int up, cudaStream_t st) {
  dim3 grid(((W + CD_TW - 1) / CD_TW) * ((H + CD_TH - 1) / CD_TH), (Cout + CD_CO - 1) / CD_CO, N);
  if (k == 3) conv_direct_kernel<TI, TO, 3><<<grid, CD_TH * CD_TW, 0, st>>>((const TI*)x, wp, bias, (TO*)y, H, W, Cin, Cout, up);
  else conv_direct_kernel<TI, TO, 1><<<grid, CD_TH * CD_TW, 0, st>>>((const TI*)x, wp, bias, (TO*)y, H, W, Cin, Cout, up);
  TTG_CHECK_LAUNCH("conv2d_direct");
  return TTG_OK;
}

// x: [N, H>>up, W>>up, Cin] (dtype_in), y: [N,H,W,Cout] (dtype_out), wp fp32 packed, bias fp32 or null.
extern "C" int ttg_conv2d_direct(const void* x, const float* wp, const float* bias, void* y, int N, int H, int W, int Cin,
                                 int Cout, int ksize, int up, int dtype_in, int dtype_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  TTG_REQUIRE(ksize == 1 || ksize == 3, "conv2d_direct: ksize %d unsupported (1 or 3)", ksize);
  TTG_REQUIRE(N > 0 && N <= 65535, "conv2d_direct: batch %d out of range", N);
  TTG_REQUIRE(up == 0 || (H % 2 == 0 && W % 2 == 0), "conv2d_direct: upsample needs even output size");
  if (dtype_in == TTG_F32 && dtype_out == TTG_F32) return launch_conv_direct<float, float>(x, wp, bias, y, N, H, W, Cin, Cout, ksize, up, st);
  if (dtype_in == TTG_BF16 && dtype_out == TTG_BF16) return launch_conv_direct<bf16, bf16>(x, wp, bias, y, N, H, W, Cin, Cout, ksize, up, st);
  if (dtype_in == TTG_BF16 && dtype_out == TTG_F32) return launch_conv_direct<bf16, float>(x, wp, bias, y, N, H, W, Cin, Cout, ksize, up, st);
  if (dtype_in == TTG_F32 && dtype_out == TTG_BF16) return launch_conv_direct<float, bf16>(x, wp, bias, y, N, H, W, Cin, Cout, ksize, up, st);
  return ttg_set_error(TTG_ERR_ARG, "conv2d_direct: bad dtypes");
}

// ------------------------------------------------------------------ wgrad
// Block = 16 output channels x 16 input channels (x K*K taps in registers), looping over the
// pixel tiles of its slice.  Split-K over blockIdx.z: with a workspace every split writes its partial sums to its own
// slab and a second kernel adds the slabs in a FIXED order (bitwise repeatable, like the reference's CPU path);
// without one (legacy entry point) the partial sums are added with fp32 atomics into gw (OIHW, zeroed here).
#define WG_C 16
template <typename TX, typename TG, int K>
__global__ void __launch_bounds__(WG_C * WG_C) conv_wgrad_direct_kernel(
    const TX* __restrict__ x, const TG* __restrict__ gy, float* __restrict__ gw, float* __restrict__ partial,
    int N, int H, int W, int Cin, int Cout, int up, int tiles_per_img, int nsplit) {
  constexpr int HALO = K / 2;
  constexpr int SH = CD_TH + 2 * HALO, SW = CD_TW + 2 * HALO;
  __shared__ float s_x[SH * SW][WG_C + 1];
  __shared__ float s_g[CD_TH * CD_TW][WG_C + 1];
  const int ci = threadIdx.x % WG_C, co = threadIdx.x / WG_C;
  const int ci0 = blockIdx.x * WG_C, co0 = blockIdx.y * WG_C;
  const int tiles_x = (W + CD_TW - 1) / CD_TW;
  const int Hi = H >> up, Wi = W >> up;
  float acc[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t) acc[t] = 0.f;
  const long long total_tiles = (long long)N * tiles_per_img;
  for (long long tile = blockIdx.z; tile < total_tiles; tile += nsplit) {
    int n = (int)(tile / tiles_per_img), tt = (int)(tile % tiles_per_img);
    int tile_x = tt % tiles_x, tile_y = tt / tiles_x;
    for (int i = threadIdx.x; i < SH * SW * WG_C; i += blockDim.x) {
      int c = i % WG_C, p = i / WG_C;
      int sx = p % SW, sy = p / SW;
      int py = tile_y * CD_TH + sy - HALO, px = tile_x * CD_TW + sx - HALO;
      float v = 0.f;
      if (py >= 0 && py < H && px >= 0 && px < W && ci0 + c < Cin)
        v = to_f(x[(((long long)n * Hi + (py >> up)) * Wi + (px >> up)) * Cin + ci0 + c]);
      s_x[p][c] = v;
    }
    for (int i = threadIdx.x; i < CD_TH * CD_TW * WG_C; i += blockDim.x) {
      int c = i % WG_C, p = i / WG_C;
      int py = tile_y * CD_TH + p / CD_TW, px = tile_x * CD_TW + p % CD_TW;
      float v = 0.f;
      if (py < H && px < W && co0 + c < Cout) v = to_f(gy[(((long long)n * H + py) * W + px) * Cout + co0 + c]);
      s_g[p][c] = v;
    }
    __syncthreads();
    for (int py = 0; py < CD_TH; ++py)
#pragma unroll 4
      for (int px = 0; px < CD_TW; ++px) {
        float g = s_g[py * CD_TW + px][co];
#pragma unroll
        for (int ky = 0; ky < K; ++ky)
#pragma unroll
          for (int kx = 0; kx < K; ++kx) acc[ky * K + kx] += g * s_x[(py + ky) * SW + px + kx][ci];
      }
    __syncthreads();
  }
  if (ci0 + ci < Cin && co0 + co < Cout) {
    const long long off = ((long long)(co0 + co) * Cin + ci0 + ci) * K * K;
    if (partial) {
      float* dst = partial + (long long)blockIdx.z * Cout * Cin * K * K + off;
#pragma unroll
      for (int t = 0; t < K * K; ++t) dst[t] = acc[t];
    } else {
      float* dst = gw + off;
#pragma unroll
      for (int t = 0; t < K * K; ++t) atomicAdd(dst + t, acc[t]);
    }
  }
}
// gw[i] = sum over the splits, in split order (fixed summation order)
__global__ void wgrad_direct_reduce_kernel(const float* __restrict__ partial, float* __restrict__ gw, long long n, int nsplit) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = 0.f;
  for (int s = 0; s < nsplit; ++s) a += partial[(long long)s * n + i];
  gw[i] = a;
}
static int wgrad_direct_nsplit(int N, int H, int W, int Cin, int Cout) {
  int tiles = ((W + CD_TW - 1) / CD_TW) * ((H + CD_TH - 1) / CD_TH);
  int gx = (Cin + WG_C - 1) / WG_C, gyb = (Cout + WG_C - 1) / WG_C;
  long long total_tiles = (long long)N * tiles;
  int nsplit = (ttg_num_sms() * 6 + gx * gyb - 1) / (gx * gyb);
  if (nsplit > total_tiles) nsplit = (int)total_tiles;
  if (nsplit > 65535) nsplit = 65535;
  if (nsplit < 1) nsplit = 1;
  return nsplit;
}

template <typename TX, typename TG>
static int launch_wgrad_direct(const void* x, const void* gy, float* gw, int N, int H, int W, int Cin, int Cout, int k,
                               int up, float* partial, cudaStream_t st) {
  int tiles = ((W + CD_TW - 1) / CD_TW) * ((H + CD_TH - 1) / CD_TH);
  int gx = (Cin + WG_C - 1) / WG_C, gyb = (Cout + WG_C - 1) / WG_C;
  const int nsplit = wgrad_direct_nsplit(N, H, W, Cin, Cout);
  dim3 grid(gx, gyb, nsplit);
  const long long n = (long long)Cout * Cin * k * k;
  if (!partial) cudaMemsetAsync(gw, 0, sizeof(float) * (size_t)n, st);
  if (k == 3) conv_wgrad_direct_kernel<TX, TG, 3><<<grid, WG_C * WG_C, 0, st>>>((const TX*)x, (const TG*)gy, gw, partial, N, H, W, Cin, Cout, up, tiles, nsplit);
  else conv_wgrad_direct_kernel<TX, TG, 1><<<grid, WG_C * WG_C, 0, st>>>((const TX*)x, (const TG*)gy, gw, partial, N, H, W, Cin, Cout, up, tiles, nsplit);
  TTG_CHECK_LAUNCH("conv2d_wgrad_direct");
  if (partial) {
    wgrad_direct_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(partial, gw, n, nsplit);
    TTG_CHECK_LAUNCH("conv2d_wgrad_direct_reduce");
  }
  return TTG_OK;
}

// x: [N, H>>up, W>>up, Cin], gy: [N,H,W,Cout]; gw: fp32 [Cout][Cin][k][k] (overwritten).
extern "C" size_t ttg_conv2d_wgrad_direct_workspace_bytes(int N, int H, int W, int Cin, int Cout, int ksize) {
  return sizeof(float) * (size_t)wgrad_direct_nsplit(N, H, W, Cin, Cout) * Cout * Cin * ksize * ksize;
}
// workspace != NULL (ttg_conv2d_wgrad_direct_workspace_bytes): deterministic fixed-order split-K reduction
extern "C" int ttg_conv2d_wgrad_direct_det(const void* x, const void* gy, float* gw, int N, int H, int W, int Cin, int Cout,
                                           int ksize, int up, int dtype_x, int dtype_gy, void* workspace, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = (float*)workspace;
  TTG_REQUIRE(ksize == 1 || ksize == 3, "conv2d_wgrad_direct: ksize %d unsupported", ksize);
  if (dtype_x == TTG_F32 && dtype_gy == TTG_F32) return launch_wgrad_direct<float, float>(x, gy, gw, N, H, W, Cin, Cout, ksize, up, ws, st);
  if (dtype_x == TTG_BF16 && dtype_gy == TTG_BF16) return launch_wgrad_direct<bf16, bf16>(x, gy, gw, N, H, W, Cin, Cout, ksize, up, ws, st);
  if (dtype_x == TTG_F32 && dtype_gy == TTG_BF16) return launch_wgrad_direct<float, bf16>(x, gy, gw, N, H, W, Cin, Cout, ksize, up, ws, st);
  if (dtype_x == TTG_BF16 && dtype_gy == TTG_F32) return launch_wgrad_direct<bf16, float>(x, gy, gw, N, H, W, Cin, Cout, ksize, up, ws, st);
  return ttg_set_error(TTG_ERR_ARG, "conv2d_wgrad_direct: bad dtypes");
}
extern "C" int ttg_conv2d_wgrad_direct(const void* x, const void* gy, float* gw, int N, int H, int W, int Cin, int Cout,
                                       int ksize, int up, int dtype_x, int dtype_gy, void* stream) {
  return ttg_conv2d_wgrad_direct_det(x, gy, gw, N, H, W, Cin, Cout, ksize, up, dtype_x, dtype_gy, nullptr, stream);
}
