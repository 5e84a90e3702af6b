// The two 3-channel convolutions of UNet2DModel (conv_in 3->C, conv_out C->3; diffusers unet_2d.py, built at
// /root/reference/generator_model/PolypGeneratorModel.py:25-48).  K = 27 resp. N = 3 is far below a UMMA tile,
// and both layers are HBM-bound (SURVEY.md §7 "hard parts"), so they are plain SIMT kernels that also do the
// NCHW-fp32 <-> NHWC-bf16 layout change at the model boundary for free.
//
// Mapping: the 128-channel side is spread over the 32 lanes of a warp (4 channels = one 8-byte access per lane), so
// NHWC accesses are fully coalesced and the per-lane weights (27 x 4 resp. 9 x 3 x 4 floats) stay in registers; the
// 3-channel side is warp-uniform (broadcast loads / shuffle reductions).
#include "common.cuh"

#include "../../include/ddpm_b200.h"

namespace ddpm {

constexpr int kScThreads = 256;

// out[n,h,w,co] = bias[co] + sum_{k<CIN, tap} x[n,k,h+dh,w+dw] * w[co*s_co + tap'*s_tap + k*s_ci]
// warp = one pixel per iteration, lane = 4 output channels (chunk blockIdx.y of 128 channels)
template <int CIN>
__global__ void __launch_bounds__(kScThreads)
conv3_to_c_kernel(const float* __restrict__ x, const float* __restrict__ w, long long s_co, long long s_tap,
                  long long s_ci, int flip, const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                  long long ldo, int N, int H, int W, int cout, long long npix) {
  const int lane = threadIdx.x & 31;
  const int c0 = blockIdx.y * 128 + lane * 4;
  const bool active = c0 < cout;
  float wr[9][CIN][4];
  float b4[4] = {0.f, 0.f, 0.f, 0.f};
  if (active) {
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int tp = flip ? 8 - tap : tap;
#pragma unroll
      for (int k = 0; k < CIN; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) wr[tap][k][j] = w[(c0 + j) * s_co + tp * s_tap + k * s_ci];
    }
    if (bias) {
#pragma unroll
      for (int j = 0; j < 4; ++j) b4[j] = bias[c0 + j];
    }
  }
  const long long hw = static_cast<long long>(H) * W;
  const long long warp0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long pix = warp0; pix < npix; pix += nwarps) {
    const int n = static_cast<int>(pix / hw);
    const int rem = static_cast<int>(pix - n * hw);
    const int h = rem / W, wq = rem - h * W;
    const float* xn = x + static_cast<long long>(n) * CIN * hw;
    float xv[9][CIN];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int hh = h + tap / 3 - 1, ww = wq + tap % 3 - 1;
      const bool in = hh >= 0 && hh < H && ww >= 0 && ww < W;
#pragma unroll
      for (int k = 0; k < CIN; ++k) xv[tap][k] = in ? __ldg(xn + k * hw + static_cast<long long>(hh) * W + ww) : 0.f;
    }
    if (!active) continue;
    float acc[4] = {b4[0], b4[1], b4[2], b4[3]};
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
      for (int k = 0; k < CIN; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = fmaf(xv[tap][k], wr[tap][k][j], acc[j]);
    *reinterpret_cast<bf16x4*>(out + pix * ldo + c0) = pack4(acc);
  }
}

// out[n,co,h,w] = bias[co] + sum_{tap, ci} a[n,h+dh,w+dw,ci] * w[co][tap][ci]
// warp = 4 consecutive pixels of a row per iteration (shared 3x6 input window), lane = 4 input channels of a
// 128-channel chunk; 4*COUT partial sums are reduced over the warp with shuffles.
template <int COUT>
__global__ void __launch_bounds__(kScThreads)
conv_c_to_3_kernel(const __nv_bfloat16* __restrict__ a, long long lda, const float* __restrict__ w,
                   const float* __restrict__ bias, float* __restrict__ out, int N, int H, int W, int cin,
                   long long ngroups, int groups_per_row) {
  const int lane = threadIdx.x & 31;
  const int nchunks = (cin + 127) / 128;
  float wr[9][COUT][4];
  auto load_w = [&](int chunk) {
    const int c = chunk * 128 + lane * 4;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
      for (int co = 0; co < COUT; ++co)
#pragma unroll
        for (int j = 0; j < 4; ++j) wr[tap][co][j] = (c + j < cin) ? w[(co * 9 + tap) * cin + c + j] : 0.f;
  };
  if (nchunks == 1) load_w(0);
  const long long hw = static_cast<long long>(H) * W;
  const long long warp0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long g = warp0; g < ngroups; g += nwarps) {
    const long long rowid = g / groups_per_row;             // n * H + h
    const int w0 = static_cast<int>(g - rowid * groups_per_row) * 4;
    const int n = static_cast<int>(rowid / H);
    const int h = static_cast<int>(rowid - static_cast<long long>(n) * H);
    float acc[4][COUT];
#pragma unroll
    for (int px = 0; px < 4; ++px)
#pragma unroll
      for (int co = 0; co < COUT; ++co) acc[px][co] = 0.f;
    for (int chunk = 0; chunk < nchunks; ++chunk) {
      if (nchunks > 1) load_w(chunk);
      const int c = chunk * 128 + lane * 4;
      const bool cact = c < cin;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int hh = h + r - 1;
        if (hh < 0 || hh >= H) continue;
        const __nv_bfloat16* rowp = a + (static_cast<long long>(n) * H + hh) * W * lda + c;
        float f[6][4];
#pragma unroll
        for (int col = 0; col < 6; ++col) {
          const int ww = w0 + col - 1;
          if (cact && ww >= 0 && ww < W) {
            const bf16x4 raw = *reinterpret_cast<const bf16x4*>(rowp + static_cast<long long>(ww) * lda);
            unpack4(raw, f[col]);
          } else {
            f[col][0] = f[col][1] = f[col][2] = f[col][3] = 0.f;
          }
        }
#pragma unroll
        for (int col = 0; col < 6; ++col)
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const int px = col - s;
            if (px >= 0 && px < 4) {
#pragma unroll
              for (int co = 0; co < COUT; ++co) {
                const float* wp = wr[r * 3 + s][co];
                acc[px][co] = fmaf(f[col][0], wp[0], fmaf(f[col][1], wp[1], fmaf(f[col][2], wp[2],
                              fmaf(f[col][3], wp[3], acc[px][co]))));
              }
            }
          }
      }
    }
#pragma unroll
    for (int px = 0; px < 4; ++px)
#pragma unroll
      for (int co = 0; co < COUT; ++co) {
        const float v = warp_sum(acc[px][co]);
        if (lane == 0 && w0 + px < W)
          out[(static_cast<long long>(n) * COUT + co) * hw + static_cast<long long>(h) * W + w0 + px] =
              v + (bias ? bias[co] : 0.f);
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
  DDPM_REQUIRE(cin >= 1 && cin <= 4 && cout % 4 == 0 && cout <= 2048 && ldo % 4 == 0,
               "ddpm_conv3_to_c: unsupported channels cin=%d cout=%d", cin, cout);
  const long long npix = static_cast<long long>(n) * h * wd;
  long long blocks = (npix * 32 + kScThreads - 1) / kScThreads;
  if (blocks > kNumSMs * 8LL) blocks = kNumSMs * 8LL;
  dim3 grid(static_cast<unsigned>(blocks), (cout + 127) / 128);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* op = static_cast<__nv_bfloat16*>(out);
  switch (cin) {
    case 1: conv3_to_c_kernel<1><<<grid, kScThreads, 0, st>>>(x, w, w_sco, w_stap, w_sci, flip, bias, op, ldo, n, h, wd, cout, npix); break;
    case 2: conv3_to_c_kernel<2><<<grid, kScThreads, 0, st>>>(x, w, w_sco, w_stap, w_sci, flip, bias, op, ldo, n, h, wd, cout, npix); break;
    case 3: conv3_to_c_kernel<3><<<grid, kScThreads, 0, st>>>(x, w, w_sco, w_stap, w_sci, flip, bias, op, ldo, n, h, wd, cout, npix); break;
    default: conv3_to_c_kernel<4><<<grid, kScThreads, 0, st>>>(x, w, w_sco, w_stap, w_sci, flip, bias, op, ldo, n, h, wd, cout, npix); break;
  }
  return check_launch("conv3_to_c_kernel");
}

extern "C" int ddpm_conv_c_to_3(const void* a, long long lda, const float* w, const float* bias, float* out, int n,
                                int h, int wd, int cin, int cout, void* stream) {
  DDPM_REQUIRE(a && w && out && n > 0 && h > 0 && wd > 0, "ddpm_conv_c_to_3: bad argument");
  DDPM_REQUIRE(cout >= 1 && cout <= 4 && cin % 4 == 0 && cin <= 2048 && lda % 4 == 0,
               "ddpm_conv_c_to_3: unsupported channels cin=%d cout=%d", cin, cout);
  const int gpr = (wd + 3) / 4;
  const long long ngroups = static_cast<long long>(n) * h * gpr;
  long long blocks = (ngroups * 32 + kScThreads - 1) / kScThreads;
  if (blocks > kNumSMs * 8LL) blocks = kNumSMs * 8LL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* ap = static_cast<const __nv_bfloat16*>(a);
  const int g = static_cast<int>(blocks);
  switch (cout) {
    case 1: conv_c_to_3_kernel<1><<<g, kScThreads, 0, st>>>(ap, lda, w, bias, out, n, h, wd, cin, ngroups, gpr); break;
    case 2: conv_c_to_3_kernel<2><<<g, kScThreads, 0, st>>>(ap, lda, w, bias, out, n, h, wd, cin, ngroups, gpr); break;
    case 3: conv_c_to_3_kernel<3><<<g, kScThreads, 0, st>>>(ap, lda, w, bias, out, n, h, wd, cin, ngroups, gpr); break;
    default: conv_c_to_3_kernel<4><<<g, kScThreads, 0, st>>>(ap, lda, w, bias, out, n, h, wd, cin, ngroups, gpr); break;
  }
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
