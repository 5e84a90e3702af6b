// Device-side input transform: uint8 HWC frames -> fp32 NCHW in [-1, 1].
//
// Replaces, per batch instead of per image on one CPU core, the transform of
// /root/reference/generator_model/PolypDiffusionDataset.py:52-59
//     Resize((S, S)) -> RandomHorizontalFlip() -> ToTensor() -> Normalize([0.5], [0.5])
// whose Resize on a PIL image is Pillow's two-pass antialiased triangle resampler in 8-bit fixed point
// (Resample.c: 22 fractional bits, 8-bit intermediate between the horizontal and the vertical pass).  The kernels keep
// that arithmetic exactly -- same integer coefficient tables (built on the host as Pillow builds them), same rounding,
// same 8-bit intermediate, same fp32 division order -- so the result is bit-identical to the reference transform
// (tests/test_preprocess.py checks against Pillow / torchvision themselves).
//
// HBM-bound byte work: pass 1 reads the frame once and writes an [H][S] intermediate, pass 2 reads it (L2-resident at
// these sizes) and writes the fp32 tensor; neighbouring threads walk neighbouring pixels, so global accesses coalesce.
#include "common.cuh"

#include "../../include/ddpm_b200.h"

namespace ddpm {

constexpr int kPrecisionBits = 32 - 8 - 2;

__device__ __forceinline__ int clip8(int acc) { return min(max(acc >> kPrecisionBits, 0), 255); }

// dst[b][r][xx][ch] = clip8(round + sum_i src[b][r][xmin(xx) + i][ch] * k[xx][i])
template <int C>
__global__ void __launch_bounds__(256)
resize_h_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, long long rows, int w, int out_w,
                const int* __restrict__ bounds, const int* __restrict__ coeffs, int ksize) {
  const long long total = rows * out_w;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / out_w;
    const int xx = static_cast<int>(i - r * out_w);
    const int xmin = bounds[2 * xx], n = bounds[2 * xx + 1];
    const int* k = coeffs + static_cast<long long>(xx) * ksize;
    const unsigned char* p = src + (r * w + xmin) * C;
    int acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 1 << (kPrecisionBits - 1);
    for (int j = 0; j < n; ++j) {
      const int kj = k[j];
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] += static_cast<int>(p[j * C + c]) * kj;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) dst[i * C + c] = static_cast<unsigned char>(clip8(acc[c]));
  }
}

// out[b][ch][yy][flip ? w-1-x : x] = ((clip8(round + sum_i src[b][ymin(yy) + i][x][ch] * k[yy][i]) / 255) - 0.5) / 0.5
template <int C>
__global__ void __launch_bounds__(256)
resize_v_normalize_kernel(const unsigned char* __restrict__ src, float* __restrict__ out, int b, int h, int w, int out_h,
                          const int* __restrict__ bounds, const int* __restrict__ coeffs, int ksize,
                          const unsigned char* __restrict__ flip) {
  const long long per_img = static_cast<long long>(out_h) * w;
  const long long total = per_img * b;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int n_img = static_cast<int>(i / per_img);
    const long long rem = i - n_img * per_img;
    const int yy = static_cast<int>(rem / w);
    const int x = static_cast<int>(rem - static_cast<long long>(yy) * w);
    const int ymin = bounds[2 * yy], n = bounds[2 * yy + 1];
    const int* k = coeffs + static_cast<long long>(yy) * ksize;
    const unsigned char* p = src + ((static_cast<long long>(n_img) * h + ymin) * w + x) * C;
    int acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 1 << (kPrecisionBits - 1);
    for (int j = 0; j < n; ++j) {
      const int kj = k[j];
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] += static_cast<int>(p[static_cast<long long>(j) * w * C + c]) * kj;
    }
    const int xo = (flip != nullptr && flip[n_img]) ? w - 1 - x : x;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float v = __fdiv_rn(static_cast<float>(clip8(acc[c])), 255.0f);             // ToTensor
      out[((static_cast<long long>(n_img) * C + c) * out_h + yy) * w + xo] = __fdiv_rn(__fsub_rn(v, 0.5f), 0.5f);
    }
  }
}

static unsigned grid_for(long long total) {
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(kNumSMs) * 16;
  return static_cast<unsigned>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace ddpm

using namespace ddpm;

extern "C" int ddpm_resize_h_u8(const unsigned char* src, unsigned char* dst, long long rows, int w, int c, int out_w,
                                const int* bounds, const int* coeffs, int ksize, void* stream) {
  DDPM_REQUIRE(src && dst && bounds && coeffs, "ddpm_resize_h_u8: null pointer");
  DDPM_REQUIRE(rows >= 0 && w > 0 && out_w > 0 && ksize > 0 && (c == 1 || c == 3),
               "ddpm_resize_h_u8: bad shape (rows=%lld w=%d c=%d out_w=%d ksize=%d)", rows, w, c, out_w, ksize);
  if (rows == 0) return DDPM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned g = grid_for(rows * out_w);
  if (c == 3) resize_h_kernel<3><<<g, 256, 0, st>>>(src, dst, rows, w, out_w, bounds, coeffs, ksize);
  else resize_h_kernel<1><<<g, 256, 0, st>>>(src, dst, rows, w, out_w, bounds, coeffs, ksize);
  return check_launch("resize_h_kernel");
}

extern "C" int ddpm_resize_v_normalize(const unsigned char* src, float* out, int b, int h, int w, int c, int out_h,
                                       const int* bounds, const int* coeffs, int ksize, const unsigned char* flip,
                                       void* stream) {
  DDPM_REQUIRE(src && out && bounds && coeffs, "ddpm_resize_v_normalize: null pointer");
  DDPM_REQUIRE(b >= 0 && h > 0 && w > 0 && out_h > 0 && ksize > 0 && (c == 1 || c == 3),
               "ddpm_resize_v_normalize: bad shape (b=%d h=%d w=%d c=%d out_h=%d ksize=%d)", b, h, w, c, out_h, ksize);
  if (b == 0) return DDPM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned g = grid_for(static_cast<long long>(b) * out_h * w);
  if (c == 3) resize_v_normalize_kernel<3><<<g, 256, 0, st>>>(src, out, b, h, w, out_h, bounds, coeffs, ksize, flip);
  else resize_v_normalize_kernel<1><<<g, 256, 0, st>>>(src, out, b, h, w, out_h, bounds, coeffs, ksize, flip);
  return check_launch("resize_v_normalize_kernel");
}
