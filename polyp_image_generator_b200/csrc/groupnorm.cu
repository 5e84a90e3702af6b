// GroupNorm(32)+SiLU forward / backward on NHWC bf16, fp32 statistics.
// Replaces at::native::group_norm + a separate silu kernel per site (71 sites per UNet forward), and folds
// torch.cat([h, skip], 1) into the load path: the input may be split over two tensors along channels.
//
// All kernels are HBM-bound streamers.  Thread mapping (all four kernels): a CTA owns a chunk of pixels of ONE
// sample; thread (v, pl) owns the 8-channel vector v (one 16-byte access) and walks pixels pl, pl+ppb, ... of the
// chunk.  Per-channel parameters (gamma, beta, mean, rstd, group coefficients) therefore live in registers for the
// whole loop, the inner loop is 16-byte loads -> 8 FMAs (+ SiLU) -> 16-byte store with kUnroll pixels in flight,
// and all index arithmetic is 32-bit.  Cross-CTA combines use fp32 atomics on tiny [N][groups] / [N][C] buffers.
#include "common.cuh"

#include "../../include/ddpm_b200.h"

namespace ddpm {

constexpr int kGnThreads = 256;
constexpr int kMaxC = 2048;
constexpr int kUnroll = 4;

struct GnSrc {
  const __nv_bfloat16* x0;
  const __nv_bfloat16* x1;
  long long ld0, ld1;
  int c0, c1;
};

// pointer to this thread's 8-channel vector at pixel 0 of sample n, and its pixel stride (elements)
__device__ __forceinline__ const __nv_bfloat16* gn_base(const GnSrc& s, int n, int hw, int c, long long* ld) {
  if (c < s.c0) {
    *ld = s.ld0;
    return s.x0 + static_cast<long long>(n) * hw * s.ld0 + c;
  }
  *ld = s.ld1;
  return s.x1 + static_cast<long long>(n) * hw * s.ld1 + (c - s.c0);
}

__device__ __forceinline__ void gn_mean_rstd(const float* stats, int n, int g, int groups, float inv_m, float eps,
                                             float* mean, float* rstd) {
  const float2 st = *reinterpret_cast<const float2*>(stats + (static_cast<long long>(n) * groups + g) * 2);
  const float mu = st.x * inv_m;
  const float var = fmaxf(st.y * inv_m - mu * mu, 0.f);
  *mean = mu;
  *rstd = rsqrtf(var + eps);
}

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// ---- statistics ------------------------------------------------------------------------------------
// grid (chunks, N); block = V * ppb threads (V = C/8 vectors per pixel, ppb pixels in flight)
__global__ void __launch_bounds__(kGnThreads)
gn_stats_kernel(GnSrc s, int hw, int cpg, int groups, float* __restrict__ stats, int pix_per_block, int V) {
  __shared__ float sm[64 * 2];
  for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int n = blockIdx.y;
  const int v = threadIdx.x % V, pl = threadIdx.x / V, ppb = blockDim.x / V;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(hw, p_begin + pix_per_block);
  long long ld;
  const __nv_bfloat16* xp = gn_base(s, n, hw, v * 8, &ld);
  float sum[8], sq[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) sum[e] = sq[e] = 0.f;
  int p = p_begin + pl;
  for (; p + (kUnroll - 1) * ppb < p_end; p += kUnroll * ppb) {
    bf16x8 raw[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) raw[u] = *reinterpret_cast<const bf16x8*>(xp + (p + u * ppb) * ld);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      float f[8];
      unpack8(raw[u], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        sum[e] += f[e];
        sq[e] = fmaf(f[e], f[e], sq[e]);
      }
    }
  }
  for (; p < p_end; p += ppb) {
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(xp + p * ld), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sum[e] += f[e];
      sq[e] = fmaf(f[e], f[e], sq[e]);
    }
  }
  // fold the 8 channels into their groups
  int g_prev = (v * 8) / cpg;
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int g = (v * 8 + e) / cpg;
    if (g != g_prev) {
      atomicAdd(&sm[g_prev * 2], a);
      atomicAdd(&sm[g_prev * 2 + 1], b);
      a = b = 0.f;
      g_prev = g;
    }
    a += sum[e];
    b += sq[e];
  }
  atomicAdd(&sm[g_prev * 2], a);
  atomicAdd(&sm[g_prev * 2 + 1], b);
  __syncthreads();
  for (int i = threadIdx.x; i < groups * 2; i += blockDim.x)
    atomicAdd(&stats[static_cast<long long>(n) * groups * 2 + i], sm[i]);
}

// ---- forward apply -----------------------------------------------------------------------------------
// y = act(x * a + b) with a = rstd*gamma, b = beta - mean*rstd*gamma (per channel, in registers)
template <bool SILU>
__global__ void __launch_bounds__(kGnThreads)
gn_apply_kernel(GnSrc s, int hw, int cpg, int groups, const float* __restrict__ stats, float eps,
                const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                long long ldy, int pix_per_block, int V) {
  const float inv_m = 1.0f / (static_cast<float>(cpg) * static_cast<float>(hw));
  const int n = blockIdx.y;
  const int v = threadIdx.x % V, pl = threadIdx.x / V, ppb = blockDim.x / V;
  const int c = v * 8;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(hw, p_begin + pix_per_block);
  long long ld;
  const __nv_bfloat16* xp = gn_base(s, n, hw, c, &ld);
  __nv_bfloat16* yp = y + static_cast<long long>(n) * hw * ldy + c;
  float ka[8], kb[8];
  {
    int g_prev = -1;
    float mean = 0.f, rstd = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int g = (c + e) / cpg;
      if (g != g_prev) {
        gn_mean_rstd(stats, n, g, groups, inv_m, eps, &mean, &rstd);
        g_prev = g;
      }
      ka[e] = rstd * gamma[c + e];
      kb[e] = beta[c + e] - mean * ka[e];
    }
  }
  int p = p_begin + pl;
  for (; p + (kUnroll - 1) * ppb < p_end; p += kUnroll * ppb) {
    bf16x8 raw[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) raw[u] = *reinterpret_cast<const bf16x8*>(xp + (p + u * ppb) * ld);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      float f[8];
      unpack8(raw[u], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float z = fmaf(f[e], ka[e], kb[e]);
        f[e] = SILU ? z * fast_sigmoid(z) : z;
      }
      *reinterpret_cast<bf16x8*>(yp + (p + u * ppb) * ldy) = pack8(f);
    }
  }
  for (; p < p_end; p += ppb) {
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(xp + p * ld), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float z = fmaf(f[e], ka[e], kb[e]);
      f[e] = SILU ? z * fast_sigmoid(z) : z;
    }
    *reinterpret_cast<bf16x8*>(yp + p * ldy) = pack8(f);
  }
}

// dz = dy * act'(z) with z = xhat*gamma + beta
template <bool SILU>
__device__ __forceinline__ float gn_dz(float dy, float z) {
  if (!SILU) return dy;
  const float sg = fast_sigmoid(z);
  return dy * sg * fmaf(z, 1.0f - sg, 1.0f);
}

// ---- backward pass 1: per-(n, c) sums of dz and dz*xhat ---------------------------------------------------
// grid (chunks, N); dynamic smem: C*2 floats.  The loop accumulates the raw moments S1 = sum dz, S2 = sum dz*x with
// only the affine z = x*kz1 + kz0 in registers; sum dz*xhat = rstd*S2 - mean*rstd*S1 is formed once at the end.
template <bool SILU>
__global__ void __launch_bounds__(kGnThreads, 3)
gn_bwd_reduce_kernel(GnSrc s, int hw, int cpg, int groups, const float* __restrict__ stats, float eps,
                     const float* __restrict__ gamma, const float* __restrict__ beta,
                     const __nv_bfloat16* __restrict__ dy, long long lddy, float* __restrict__ sums /*[N][C][2]*/,
                     int pix_per_block, int V) {
  extern __shared__ float smc[];
  const int C = V * 8;
  for (int i = threadIdx.x; i < C * 2; i += blockDim.x) smc[i] = 0.f;
  __syncthreads();
  const float inv_m = 1.0f / (static_cast<float>(cpg) * static_cast<float>(hw));
  const int n = blockIdx.y;
  const int v = threadIdx.x % V, pl = threadIdx.x / V, ppb = blockDim.x / V;
  const int c = v * 8;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(hw, p_begin + pix_per_block);
  long long ld;
  const __nv_bfloat16* xp = gn_base(s, n, hw, c, &ld);
  const __nv_bfloat16* dp = dy + static_cast<long long>(n) * hw * lddy + c;
  float kz1[8], kz0[8], A[8], B[8];   // z = x*kz1 + kz0
  {
    int g_prev = -1;
    float mean = 0.f, rstd = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int g = (c + e) / cpg;
      if (g != g_prev) {
        gn_mean_rstd(stats, n, g, groups, inv_m, eps, &mean, &rstd);
        g_prev = g;
      }
      const float ga = gamma[c + e];
      kz1[e] = rstd * ga;
      kz0[e] = fmaf(-mean * rstd, ga, beta[c + e]);
      A[e] = B[e] = 0.f;
    }
  }
  int p = p_begin + pl;
  for (; p + (kUnroll - 1) * ppb < p_end; p += kUnroll * ppb) {
    bf16x8 rx[kUnroll], rd[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      rx[u] = *reinterpret_cast<const bf16x8*>(xp + (p + u * ppb) * ld);
      rd[u] = *reinterpret_cast<const bf16x8*>(dp + (p + u * ppb) * lddy);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      float f[8], d[8];
      unpack8(rx[u], f);
      unpack8(rd[u], d);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float dz = gn_dz<SILU>(d[e], fmaf(f[e], kz1[e], kz0[e]));
        A[e] += dz;
        B[e] = fmaf(dz, f[e], B[e]);
      }
    }
  }
  for (; p < p_end; p += ppb) {
    float f[8], d[8];
    unpack8(*reinterpret_cast<const bf16x8*>(xp + p * ld), f);
    unpack8(*reinterpret_cast<const bf16x8*>(dp + p * lddy), d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float dz = gn_dz<SILU>(d[e], fmaf(f[e], kz1[e], kz0[e]));
      A[e] += dz;
      B[e] = fmaf(dz, f[e], B[e]);
    }
  }
  {
    int g_prev = -1;
    float mean = 0.f, rstd = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int g = (c + e) / cpg;
      if (g != g_prev) {
        gn_mean_rstd(stats, n, g, groups, inv_m, eps, &mean, &rstd);
        g_prev = g;
      }
      atomicAdd(&smc[(c + e) * 2], A[e]);
      atomicAdd(&smc[(c + e) * 2 + 1], rstd * (B[e] - mean * A[e]));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 2; i += blockDim.x)
    atomicAdd(&sums[static_cast<long long>(n) * C * 2 + i], smc[i]);
}

// ---- backward finalize: group coefficients + dgamma/dbeta ------------------------------------------------
// blocks [0, N): coef[n][g] = (sum_c gamma*A, sum_c gamma*B);  blocks [N, N + ceil(C/256)): dgamma/dbeta
__global__ void __launch_bounds__(kGnThreads)
gn_bwd_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ gamma, int N, int C, int cpg,
                       int groups, float* __restrict__ coef, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  if (static_cast<int>(blockIdx.x) < N) {
    const int n = blockIdx.x;
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
      float s1 = 0.f, s2 = 0.f;
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        const float2 ab = *reinterpret_cast<const float2*>(sums + (static_cast<long long>(n) * C + c) * 2);
        s1 += gamma[c] * ab.x;
        s2 += gamma[c] * ab.y;
      }
      coef[(static_cast<long long>(n) * groups + g) * 2] = s1;
      coef[(static_cast<long long>(n) * groups + g) * 2 + 1] = s2;
    }
  } else {
    const int c = (blockIdx.x - N) * blockDim.x + threadIdx.x;
    if (c < C && (dgamma || dbeta)) {
      float a = 0.f, b = 0.f;
      for (int n = 0; n < N; ++n) {
        const float2 ab = *reinterpret_cast<const float2*>(sums + (static_cast<long long>(n) * C + c) * 2);
        a += ab.x;
        b += ab.y;
      }
      if (dbeta) dbeta[c] += a;
      if (dgamma) dgamma[c] += b;
    }
  }
}

// ---- backward pass 2: dx = dz*k1 - k2 - xhat*k3 (+ addends), k1 = rstd*gamma, k2 = rstd*s1/m, k3 = rstd*s2/m
//      (xhat = x*rstd - mean*rstd is folded into the coefficients so only four 8-vectors stay in registers)
struct GnDst {
  __nv_bfloat16* d0;
  __nv_bfloat16* d1;
  long long ld0, ld1;
  const __nv_bfloat16* add0;
  const __nv_bfloat16* add1;
  long long lda0, lda1;
};

// dx = dz*k1 - x*k4 - k5 with z = x*kz1 + kz0, k1 = rstd*gamma, k4 = rstd^2*s2/m, k5 = rstd*s1/m - mean*rstd^2*s2/m
template <bool SILU>
__global__ void __launch_bounds__(kGnThreads, 3)
gn_bwd_apply_kernel(GnSrc s, int hw, int cpg, int groups, const float* __restrict__ stats, float eps,
                    const float* __restrict__ gamma, const float* __restrict__ beta,
                    const __nv_bfloat16* __restrict__ dy, long long lddy, const float* __restrict__ coef, GnDst o,
                    int pix_per_block, int V) {
  const float inv_m = 1.0f / (static_cast<float>(cpg) * static_cast<float>(hw));
  const int n = blockIdx.y;
  const int v = threadIdx.x % V, pl = threadIdx.x / V, ppb = blockDim.x / V;
  const int c = v * 8;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(hw, p_begin + pix_per_block);
  long long ld;
  const __nv_bfloat16* xp = gn_base(s, n, hw, c, &ld);
  const long long pix0 = static_cast<long long>(n) * hw;
  const __nv_bfloat16* dp = dy + pix0 * lddy + c;
  const __nv_bfloat16* a0p = o.add0 ? o.add0 + pix0 * o.lda0 + c : nullptr;
  const __nv_bfloat16* a1p = o.add1 ? o.add1 + pix0 * o.lda1 + c : nullptr;
  __nv_bfloat16* op;
  long long ldo;
  if (c < s.c0) {
    op = o.d0 + pix0 * o.ld0 + c;
    ldo = o.ld0;
  } else {
    op = o.d1 ? o.d1 + pix0 * o.ld1 + (c - s.c0) : nullptr;
    ldo = o.ld1;
  }
  if (op == nullptr) return;   // gradient of this source not requested
  float kz1[8], kz0[8], k4[8], k5[8];
  {
    int g_prev = -1;
    float mean = 0.f, rstd = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int g = (c + e) / cpg;
      if (g != g_prev) {
        gn_mean_rstd(stats, n, g, groups, inv_m, eps, &mean, &rstd);
        const float2 cf = *reinterpret_cast<const float2*>(coef + (static_cast<long long>(n) * groups + g) * 2);
        s1 = cf.x * inv_m;
        s2 = cf.y * inv_m;
        g_prev = g;
      }
      const float ga = gamma[c + e];
      kz1[e] = rstd * ga;                                  // also k1
      kz0[e] = fmaf(-mean * rstd, ga, beta[c + e]);
      k4[e] = rstd * rstd * s2;
      k5[e] = rstd * s1 - mean * k4[e];
    }
  }
  for (int p = p_begin + pl; p < p_end; p += 2 * ppb) {
    const bool two = p + ppb < p_end;
    bf16x8 rx[2], rd[2], ra0[2], ra1[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 0 || two) {
        const int q = p + u * ppb;
        rx[u] = *reinterpret_cast<const bf16x8*>(xp + q * ld);
        rd[u] = *reinterpret_cast<const bf16x8*>(dp + q * lddy);
        if (a0p) ra0[u] = *reinterpret_cast<const bf16x8*>(a0p + q * o.lda0);
        if (a1p) ra1[u] = *reinterpret_cast<const bf16x8*>(a1p + q * o.lda1);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 0 || two) {
        float f[8], d[8], r[8];
        unpack8(rx[u], f);
        unpack8(rd[u], d);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float dz = gn_dz<SILU>(d[e], fmaf(f[e], kz1[e], kz0[e]));
          r[e] = fmaf(dz, kz1[e], -fmaf(f[e], k4[e], k5[e]));
        }
        if (a0p) {
          float a[8];
          unpack8(ra0[u], a);
#pragma unroll
          for (int e = 0; e < 8; ++e) r[e] += a[e];
        }
        if (a1p) {
          float a[8];
          unpack8(ra1[u], a);
#pragma unroll
          for (int e = 0; e < 8; ++e) r[e] += a[e];
        }
        *reinterpret_cast<bf16x8*>(op + (p + u * ppb) * ldo) = pack8(r);
      }
    }
  }
}

static int gn_check(const void* x0, int c0, long long ld0, const void* x1, int c1, long long ld1, int n, int hw,
                    int groups, const char* who) {
  if (!x0 || n <= 0 || hw <= 0 || groups <= 0 || groups > 64) {
    set_last_error("%s: bad argument (n=%d hw=%d groups=%d)", who, n, hw, groups);
    return DDPM_ERR_INVALID;
  }
  const int C = c0 + c1;
  if (c0 <= 0 || c0 % 8 || c1 < 0 || c1 % 8 || (c1 > 0 && !x1) || C % groups || C > kMaxC || ld0 % 8 ||
      (c1 > 0 && ld1 % 8) || C / 8 > kGnThreads) {
    set_last_error("%s: unsupported channel configuration c0=%d c1=%d groups=%d ld0=%lld ld1=%lld", who, c0, c1,
                   groups, ld0, ld1);
    return DDPM_ERR_INVALID;
  }
  return DDPM_OK;
}

// Pixel chunks per sample.  CTAs are kept SMALL (min_iters loop trips per thread) so that the grid is many waves
// deep (<= kMaxCtas): with a handful of large CTAs per SM the last, partially filled wave cost up to 30 % of the
// kernel (1216 CTAs over 888 resident slots = 1.37 waves).  `waves` is unused now and kept for call-site clarity.
constexpr long long kMaxCtas = 148LL * 64;
static void gn_geometry(int C, int hw, int n, int waves, int min_iters, int* V, int* threads, int* pix_per_block,
                        int* chunks) {
  (void)waves;
  *V = C / 8;
  int ppb = kGnThreads / *V;
  if (ppb < 1) ppb = 1;
  *threads = *V * ppb;
  long long ppblk = static_cast<long long>(ppb) * min_iters;
  while (static_cast<long long>(n) * ((hw + ppblk - 1) / ppblk) > kMaxCtas) ppblk += static_cast<long long>(ppb) * min_iters;
  if (ppblk > hw) ppblk = hw;
  *pix_per_block = static_cast<int>(ppblk);
  *chunks = static_cast<int>((hw + ppblk - 1) / ppblk);
}

}  // namespace ddpm

using namespace ddpm;

extern "C" int ddpm_gn_stats(const void* x0, int c0, long long ld0, const void* x1, int c1, long long ld1, int n,
                             int hw, int groups, float* stats, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = gn_check(x0, c0, ld0, x1, c1, ld1, n, hw, groups, "ddpm_gn_stats")) return e;
  DDPM_REQUIRE(stats, "ddpm_gn_stats: stats is null");
  const int C = c0 + c1;
  GnSrc s{static_cast<const __nv_bfloat16*>(x0), static_cast<const __nv_bfloat16*>(x1), ld0, ld1, c0, c1};
  int V, threads, ppblk, chunks;
  gn_geometry(C, hw, n, 4, 4 * kUnroll, &V, &threads, &ppblk, &chunks);
  DDPM_CUDA(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * groups * n, stream));
  gn_stats_kernel<<<dim3(chunks, n), threads, 0, stream>>>(s, hw, C / groups, groups, stats, ppblk, V);
  return check_launch("gn_stats_kernel");
}

extern "C" int ddpm_gn_apply(const void* x0, int c0, long long ld0, const void* x1, int c1, long long ld1, int n,
                             int hw, int groups, const float* stats, float eps, const float* gamma,
                             const float* beta, int silu, void* y, long long ldy, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = gn_check(x0, c0, ld0, x1, c1, ld1, n, hw, groups, "ddpm_gn_apply")) return e;
  DDPM_REQUIRE(stats && gamma && beta && y && ldy % 8 == 0, "ddpm_gn_apply: bad argument");
  const int C = c0 + c1;
  GnSrc s{static_cast<const __nv_bfloat16*>(x0), static_cast<const __nv_bfloat16*>(x1), ld0, ld1, c0, c1};
  int V, threads, ppblk, chunks;
  gn_geometry(C, hw, n, 4, 2 * kUnroll, &V, &threads, &ppblk, &chunks);
  __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(y);
  if (silu)
    gn_apply_kernel<true><<<dim3(chunks, n), threads, 0, stream>>>(s, hw, C / groups, groups, stats, eps, gamma, beta,
                                                                   yp, ldy, ppblk, V);
  else
    gn_apply_kernel<false><<<dim3(chunks, n), threads, 0, stream>>>(s, hw, C / groups, groups, stats, eps, gamma, beta,
                                                                    yp, ldy, ppblk, V);
  return check_launch("gn_apply_kernel");
}

extern "C" int ddpm_gn_bwd(const void* x0, int c0, long long ld0, const void* x1, int c1, long long ld1, int n,
                           int hw, int groups, const float* stats, float eps, const float* gamma, const float* beta,
                           int silu, const void* dy, long long lddy, const void* add0, long long ldadd0,
                           const void* add1, long long ldadd1, void* dx0, long long lddx0, void* dx1, long long lddx1,
                           float* dgamma, float* dbeta, float* ws, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = gn_check(x0, c0, ld0, x1, c1, ld1, n, hw, groups, "ddpm_gn_bwd")) return e;
  DDPM_REQUIRE(stats && gamma && beta && dy && dx0 && ws, "ddpm_gn_bwd: null pointer argument");
  DDPM_REQUIRE(lddy % 8 == 0 && lddx0 % 8 == 0 && (c1 == 0 || !dx1 || lddx1 % 8 == 0) &&
                   (!add0 || ldadd0 % 8 == 0) && (!add1 || ldadd1 % 8 == 0),
               "ddpm_gn_bwd: strides must be multiples of 8");
  const int C = c0 + c1;
  GnSrc s{static_cast<const __nv_bfloat16*>(x0), static_cast<const __nv_bfloat16*>(x1), ld0, ld1, c0, c1};
  const __nv_bfloat16* dyp = static_cast<const __nv_bfloat16*>(dy);
  int V, threads, ppblk, chunks;
  gn_geometry(C, hw, n, 4, 8, &V, &threads, &ppblk, &chunks);
  float* sums = ws;
  float* coef = ws + static_cast<long long>(n) * C * 2;
  DDPM_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * C * n, stream));
  if (silu)
    gn_bwd_reduce_kernel<true><<<dim3(chunks, n), threads, C * 2 * sizeof(float), stream>>>(
        s, hw, C / groups, groups, stats, eps, gamma, beta, dyp, lddy, sums, ppblk, V);
  else
    gn_bwd_reduce_kernel<false><<<dim3(chunks, n), threads, C * 2 * sizeof(float), stream>>>(
        s, hw, C / groups, groups, stats, eps, gamma, beta, dyp, lddy, sums, ppblk, V);
  if (int e = check_launch("gn_bwd_reduce_kernel")) return e;
  gn_bwd_finalize_kernel<<<n + (C + kGnThreads - 1) / kGnThreads, kGnThreads, 0, stream>>>(sums, gamma, n, C,
                                                                                          C / groups, groups, coef,
                                                                                          dgamma, dbeta);
  if (int e = check_launch("gn_bwd_finalize_kernel")) return e;
  GnDst o{static_cast<__nv_bfloat16*>(dx0), static_cast<__nv_bfloat16*>(dx1), lddx0, lddx1,
          static_cast<const __nv_bfloat16*>(add0), static_cast<const __nv_bfloat16*>(add1), ldadd0, ldadd1};
  gn_geometry(C, hw, n, 4, 4, &V, &threads, &ppblk, &chunks);
  if (silu)
    gn_bwd_apply_kernel<true><<<dim3(chunks, n), threads, 0, stream>>>(s, hw, C / groups, groups, stats, eps, gamma,
                                                                       beta, dyp, lddy, coef, o, ppblk, V);
  else
    gn_bwd_apply_kernel<false><<<dim3(chunks, n), threads, 0, stream>>>(s, hw, C / groups, groups, stats, eps, gamma,
                                                                        beta, dyp, lddy, coef, o, ppblk, V);
  return check_launch("gn_bwd_apply_kernel");
}
