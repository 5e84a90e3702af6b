// The two 3-channel convolutions of UNet2DModel (conv_in 3->C, conv_out C->3; diffusers unet_2d.py, built at
// /root/reference/generator_model/PolypGeneratorModel.py:25-48).  K = 27 resp. N = 3 is far below a UMMA tile,
// and both layers are HBM-bound (SURVEY.md §7 "hard parts"), so they are plain SIMT kernels that also do the
// NCHW-fp32 <-> NHWC-bf16 layout change at the model boundary for free.
#include "common.cuh"

#include "../../include/ddpm_b200.h"

namespace ddpm {

constexpr int kScThreads = 256;

// out[n,h,w,co] = bias[co] + sum_{k<cin, tap} x[n,k,h+dh,w+dw] * w[co*s_co + tap'*s_tap + k*s_ci]
// thread = (pixel, 8-channel group).  smem weights: [tap*cin + k][cout] fp32.
__global__ void __launch_bounds__(kScThreads)
conv3_to_c_kernel(const float* __restrict__ x, const float* __restrict__ w, long long s_co, long long s_tap,
                  long long s_ci, int flip, const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                  long long ldo, int N, int H, int W, int cin, int cout, long long total) {
  extern __shared__ float ws[];  // [9*cin][cout]
  const int KK = 9 * cin;
  for (int i = threadIdx.x; i < KK * cout; i += blockDim.x) {
    const int co = i % cout, kk = i / cout;
    const int tap = kk / cin, k = kk - tap * cin;
    const int tp = flip ? 8 - tap : tap;
    ws[i] = w[co * s_co + tp * s_tap + k * s_ci];
  }
  __syncthreads();
  const int G = cout / 8;
  const long long hw = static_cast<long long>(H) * W;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int g = static_cast<int>(i % G);
    const long long pix = i / G;
    const int n = static_cast<int>(pix / hw);
    const int rem = static_cast<int>(pix - n * hw);
    const int h = rem / W, wq = rem - h * W;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = bias ? bias[g * 8 + e] : 0.f;
    for (int tap = 0; tap < 9; ++tap) {
      const int hh = h + tap / 3 - 1, ww = wq + tap % 3 - 1;
      if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
      for (int k = 0; k < cin; ++k) {
        const float xv = x[(static_cast<long long>(n) * cin + k) * hw + static_cast<long long>(hh) * W + ww];
        const float4* wp = reinterpret_cast<const float4*>(ws + (tap * cin + k) * cout + g * 8);
        const float4 w0 = wp[0], w1 = wp[1];
        acc[0] += xv * w0.x; acc[1] += xv * w0.y; acc[2] += xv * w0.z; acc[3] += xv * w0.w;
        acc[4] += xv * w1.x; acc[5] += xv * w1.y; acc[6] += xv * w1.z; acc[7] += xv * w1.w;
      }
    }
    *reinterpret_cast<bf16x8*>(out + pix * ldo + g * 8) = pack8(acc);
  }
}

// out[n,co,h,w] = bias[co] + sum_{tap, ci} a[n,h+dh,w+dw,ci] * w[co][tap][ci];  one warp per pixel.
// COUT <= 4.  smem weights [cout][9][cin] fp32.
template <int COUT>
__global__ void __launch_bounds__(kScThreads)
conv_c_to_3_kernel(const __nv_bfloat16* __restrict__ a, long long lda, const float* __restrict__ w,
                   const float* __restrict__ bias, float* __restrict__ out, int N, int H, int W, int cin,
                   long long npix) {
  extern __shared__ float ws[];
  for (int i = threadIdx.x; i < COUT * 9 * cin; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long hw = static_cast<long long>(H) * W;
  const long long warp0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long pix = warp0; pix < npix; pix += nwarps) {
    const int n = static_cast<int>(pix / hw);
    const int rem = static_cast<int>(pix - n * hw);
    const int h = rem / W, wq = rem - h * W;
    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
      const int hh = h + tap / 3 - 1, ww = wq + tap % 3 - 1;
      if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
      const __nv_bfloat16* ap = a + ((static_cast<long long>(n) * H + hh) * W + ww) * lda;
      for (int c0 = lane * 8; c0 < cin; c0 += 256) {
        float f[8];
        unpack8(*reinterpret_cast<const bf16x8*>(ap + c0), f);
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
          const float* wp = ws + (c * 9 + tap) * cin + c0;
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[c] += f[e] * wp[e];
        }
      }
    }
#pragma unroll
    for (int c = 0; c < COUT; ++c) {
      const float v = warp_sum(acc[c]);
      if (lane == 0) out[(static_cast<long long>(n) * COUT + c) * hw + rem] = v + (bias ? bias[c] : 0.f);
    }
  }
}

// dw[c*s_c + tap'*s_tap + k*s_k] += sum_pix big[pix, c] * small[n, k, pix + tap]
// Persistent CTAs; each CTA walks pixel tiles of kTile pixels, stages big[tile][cbig] (fp32) and the small
// patch [tile][9*ks] in smem; thread t owns channel c = t % cbig and a slice of the 9*ks (tap,k) columns.
constexpr int kWgTile = 64;
__global__ void __launch_bounds__(kScThreads)
conv3_wgrad_kernel(const __nv_bfloat16* __restrict__ big, long long ldbig, int cbig, const float* __restrict__ small_,
                   int ks, float* __restrict__ dw, long long s_c, long long s_tap, long long s_k, int flip,
                   float* __restrict__ dbias_small, int N, int H, int W, long long npix) {
  extern __shared__ float sm[];
  const int KK = 9 * ks;
  float* sbig = sm;                       // [kWgTile][cbig]
  float* ssm = sm + kWgTile * cbig;       // [kWgTile][KK]
  const int parts = blockDim.x / cbig;    // column slices
  const int c = threadIdx.x % cbig, part = threadIdx.x / cbig;
  const int cols_per = (KK + parts - 1) / parts;
  const int col0 = part * cols_per;
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  float bsum = 0.f;   // thread j < ks accumulates sum of small[:, j, :] over its tiles (centre tap)
  const long long hw = static_cast<long long>(H) * W;
  const long long ntiles = (npix + kWgTile - 1) / kWgTile;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long p0 = tile * kWgTile;
    __syncthreads();
    for (int i = threadIdx.x; i < kWgTile * cbig / 8; i += blockDim.x) {
      const int pl = i / (cbig / 8), v = i - pl * (cbig / 8);
      float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (p0 + pl < npix) unpack8(*reinterpret_cast<const bf16x8*>(big + (p0 + pl) * ldbig + v * 8), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) sbig[pl * cbig + v * 8 + e] = f[e];
    }
    for (int i = threadIdx.x; i < kWgTile * KK; i += blockDim.x) {
      const int pl = i / KK, kk = i - pl * KK;
      const int tap = kk / ks, k = kk - tap * ks;
      float v = 0.f;
      const long long pix = p0 + pl;
      if (pix < npix) {
        const int n = static_cast<int>(pix / hw);
        const int rem = static_cast<int>(pix - n * hw);
        const int h = rem / W + tap / 3 - 1, wq = rem % W + tap % 3 - 1;
        if (h >= 0 && h < H && wq >= 0 && wq < W)
          v = small_[(static_cast<long long>(n) * ks + k) * hw + static_cast<long long>(h) * W + wq];
      }
      ssm[i] = v;
    }
    __syncthreads();
    if (part < parts) {
      for (int pl = 0; pl < kWgTile; ++pl) {
        const float b = sbig[pl * cbig + c];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < cols_per && col0 + j < KK) acc[j] += b * ssm[pl * KK + col0 + j];
      }
    }
    if (dbias_small && threadIdx.x < ks) {
      for (int pl = 0; pl < kWgTile; ++pl) bsum += ssm[pl * KK + 4 * ks + threadIdx.x];  // centre tap = the pixel itself
    }
  }
  if (part < parts) {
    for (int j = 0; j < 16; ++j) {
      if (j < cols_per && col0 + j < KK) {
        const int kk = col0 + j;
        const int tap = kk / ks, k = kk - tap * ks;
        const int tp = flip ? 8 - tap : tap;
        atomicAdd(&dw[c * s_c + tp * s_tap + k * s_k], acc[j]);
      }
    }
  }
  if (dbias_small && threadIdx.x < ks) atomicAdd(&dbias_small[threadIdx.x], bsum);
}

}  // namespace ddpm

using namespace ddpm;

extern "C" int ddpm_conv3_to_c(const float* x, const float* w, long long w_sco, long long w_stap, long long w_sci,
                               int flip, const float* bias, void* out, long long ldo, int n, int h, int wd, int cin,
                               int cout, void* stream) {
  DDPM_REQUIRE(x && w && out && n > 0 && h > 0 && wd > 0, "ddpm_conv3_to_c: bad argument");
  DDPM_REQUIRE(cin >= 1 && cin <= 4 && cout % 8 == 0 && cout <= 512 && ldo % 8 == 0,
               "ddpm_conv3_to_c: unsupported channels cin=%d cout=%d", cin, cout);
  const size_t smem = sizeof(float) * 9 * cin * cout;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    DDPM_CUDA(cudaFuncSetAttribute(conv3_to_c_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const long long total = static_cast<long long>(n) * h * wd * (cout / 8);
  long long blocks = (total + kScThreads - 1) / kScThreads;
  if (blocks > kNumSMs * 8LL) blocks = kNumSMs * 8LL;
  conv3_to_c_kernel<<<static_cast<int>(blocks), kScThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      x, w, w_sco, w_stap, w_sci, flip, bias, static_cast<__nv_bfloat16*>(out), ldo, n, h, wd, cin, cout, total);
  return check_launch("conv3_to_c_kernel");
}

extern "C" int ddpm_conv_c_to_3(const void* a, long long lda, const float* w, const float* bias, float* out, int n,
                                int h, int wd, int cin, int cout, void* stream) {
  DDPM_REQUIRE(a && w && out && n > 0 && h > 0 && wd > 0, "ddpm_conv_c_to_3: bad argument");
  DDPM_REQUIRE(cout >= 1 && cout <= 4 && cin % 8 == 0 && cin <= 1024 && lda % 8 == 0,
               "ddpm_conv_c_to_3: unsupported channels cin=%d cout=%d", cin, cout);
  const size_t smem = sizeof(float) * cout * 9 * cin;
  const long long npix = static_cast<long long>(n) * h * wd;
  long long blocks = (npix * 32 + kScThreads - 1) / kScThreads;
  if (blocks > kNumSMs * 8LL) blocks = kNumSMs * 8LL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* ap = static_cast<const __nv_bfloat16*>(a);
#define LAUNCH_C3(CO)                                                                                           \
  {                                                                                                             \
    static size_t configured = 0;                                                                               \
    if (smem > 48 * 1024 && smem > configured) {                                                                \
      DDPM_CUDA(cudaFuncSetAttribute(conv_c_to_3_kernel<CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      configured = smem;                                                                                        \
    }                                                                                                           \
    conv_c_to_3_kernel<CO><<<static_cast<int>(blocks), kScThreads, smem, st>>>(ap, lda, w, bias, out, n, h, wd, cin, npix); \
  }
  switch (cout) {
    case 1: LAUNCH_C3(1); break;
    case 2: LAUNCH_C3(2); break;
    case 3: LAUNCH_C3(3); break;
    default: LAUNCH_C3(4); break;
  }
#undef LAUNCH_C3
  return check_launch("conv_c_to_3_kernel");
}

extern "C" int ddpm_conv3_wgrad(const void* big, long long ldbig, int cbig, const float* small_, int ksmall, float* dw,
                                long long s_c, long long s_tap, long long s_k, int flip, float* dbias_small, int n,
                                int h, int wd, void* stream) {
  DDPM_REQUIRE(big && small_ && dw && n > 0 && h > 0 && wd > 0, "ddpm_conv3_wgrad: bad argument");
  DDPM_REQUIRE(ksmall >= 1 && ksmall <= 4 && cbig % 8 == 0 && cbig <= kScThreads && kScThreads % cbig == 0 &&
                   ldbig % 8 == 0,
               "ddpm_conv3_wgrad: unsupported channels cbig=%d ksmall=%d", cbig, ksmall);
  const int parts = kScThreads / cbig;
  DDPM_REQUIRE((9 * ksmall + parts - 1) / parts <= 16, "ddpm_conv3_wgrad: too many columns per thread");
  const size_t smem = sizeof(float) * kWgTile * (cbig + 9 * ksmall);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    DDPM_CUDA(cudaFuncSetAttribute(conv3_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const long long npix = static_cast<long long>(n) * h * wd;
  long long blocks = (npix + kWgTile - 1) / kWgTile;
  if (blocks > kNumSMs * 4LL) blocks = kNumSMs * 4LL;
  conv3_wgrad_kernel<<<static_cast<int>(blocks), kScThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(big), ldbig, cbig, small_, ksmall, dw, s_c, s_tap, s_k, flip, dbias_small, n, h,
      wd, npix);
  return check_launch("conv3_wgrad_kernel");
}
