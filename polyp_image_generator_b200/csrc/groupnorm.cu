// GroupNorm(32)+SiLU forward / backward on NHWC bf16, fp32 statistics.
// Replaces at::native::group_norm + a separate silu kernel per site (71 sites per UNet forward), and folds
// torch.cat([h, skip], 1) into the load path: the input may be split over two tensors along channels.
//
// All kernels are HBM-bound streamers: one 16-byte vector (8 channels) per thread per pixel, coalesced along
// the channel dimension, warp-shuffle / shared-memory reductions, fp32 atomics for the cross-CTA combine.
#include "common.cuh"

#include "../../include/ddpm_b200.h"

namespace ddpm {

constexpr int kGnThreads = 256;
constexpr int kMaxC = 2048;

struct GnSrc {
  const __nv_bfloat16* x0;
  const __nv_bfloat16* x1;
  long long ld0, ld1;
  int c0, c1;
};

__device__ __forceinline__ bf16x8 gn_load(const GnSrc& s, long long pix, int c) {
  if (c < s.c0) return *reinterpret_cast<const bf16x8*>(s.x0 + pix * s.ld0 + c);
  return *reinterpret_cast<const bf16x8*>(s.x1 + pix * s.ld1 + (c - s.c0));
}

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
  float sum[8], sq[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) sum[e] = sq[e] = 0.f;
  if (pl < ppb) {
    for (int p = p_begin + pl; p < p_end; p += ppb) {
      float f[8];
      unpack8(gn_load(s, static_cast<long long>(n) * hw + p, v * 8), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        sum[e] += f[e];
        sq[e] += f[e] * f[e];
      }
    }
    // fold the 8 channels into their groups (a vector spans at most two groups for cpg >= 4)
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
  }
  __syncthreads();
  for (int i = threadIdx.x; i < groups * 2; i += blockDim.x)
    atomicAdd(&stats[static_cast<long long>(n) * groups * 2 + i], sm[i]);
}

__device__ __forceinline__ void gn_mean_rstd(const float* stats, int n, int g, int groups, float inv_m, float eps,
                                             float* mean, float* rstd) {
  const float2 st = *reinterpret_cast<const float2*>(stats + (static_cast<long long>(n) * groups + g) * 2);
  const float mu = st.x * inv_m;
  const float var = fmaxf(st.y * inv_m - mu * mu, 0.f);
  *mean = mu;
  *rstd = rsqrtf(var + eps);
}

// ---- forward apply -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGnThreads)
gn_apply_kernel(GnSrc s, int hw, long long total_vec, int cpg, int groups, const float* __restrict__ stats, float eps,
                const float* __restrict__ gamma, const float* __restrict__ beta, int silu,
                __nv_bfloat16* __restrict__ y, long long ldy, int V) {
  const float inv_m = 1.0f / (static_cast<float>(cpg) * static_cast<float>(hw));
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total_vec; i += stride) {
    const int v = static_cast<int>(i % V);
    const long long pix = i / V;
    const int n = static_cast<int>(pix / hw);
    const int c = v * 8;
    float f[8];
    unpack8(gn_load(s, pix, c), f);
    int g_prev = -1;
    float mean = 0.f, rstd = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int g = (c + e) / cpg;
      if (g != g_prev) {
        gn_mean_rstd(stats, n, g, groups, inv_m, eps, &mean, &rstd);
        g_prev = g;
      }
      float z = (f[e] - mean) * rstd * gamma[c + e] + beta[c + e];
      f[e] = silu ? silu_f(z) : z;
    }
    *reinterpret_cast<bf16x8*>(y + pix * ldy + c) = pack8(f);
  }
}

// ---- backward pass 1: per-(n, c) sums of dz and dz*xhat ---------------------------------------------------
// grid (chunks, N); dynamic smem: C*2 floats
__global__ void __launch_bounds__(kGnThreads)
gn_bwd_reduce_kernel(GnSrc s, int hw, int cpg, int groups, const float* __restrict__ stats, float eps,
                     const float* __restrict__ gamma, const float* __restrict__ beta, int silu,
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
  if (pl < ppb) {
    float mean[8], rstd[8], ga[8], be[8], A[8], B[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      gn_mean_rstd(stats, n, (c + e) / cpg, groups, inv_m, eps, &mean[e], &rstd[e]);
      ga[e] = gamma[c + e];
      be[e] = beta[c + e];
      A[e] = B[e] = 0.f;
    }
    for (int p = p_begin + pl; p < p_end; p += ppb) {
      const long long pix = static_cast<long long>(n) * hw + p;
      float f[8], d[8];
      unpack8(gn_load(s, pix, c), f);
      unpack8(*reinterpret_cast<const bf16x8*>(dy + pix * lddy + c), d);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xh = (f[e] - mean[e]) * rstd[e];
        float dz = d[e];
        if (silu) dz *= silu_grad_f(xh * ga[e] + be[e]);
        A[e] += dz;
        B[e] += dz * xh;
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      atomicAdd(&smc[(c + e) * 2], A[e]);
      atomicAdd(&smc[(c + e) * 2 + 1], B[e]);
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

// ---- backward pass 2: dx ------------------------------------------------------------------------------
struct GnDst {
  __nv_bfloat16* d0;
  __nv_bfloat16* d1;
  long long ld0, ld1;
  const __nv_bfloat16* add0;
  const __nv_bfloat16* add1;
  long long lda0, lda1;
};

__global__ void __launch_bounds__(kGnThreads)
gn_bwd_apply_kernel(GnSrc s, int hw, long long total_vec, int cpg, int groups, const float* __restrict__ stats,
                    float eps, const float* __restrict__ gamma, const float* __restrict__ beta, int silu,
                    const __nv_bfloat16* __restrict__ dy, long long lddy, const float* __restrict__ coef, GnDst o,
                    int V) {
  const float inv_m = 1.0f / (static_cast<float>(cpg) * static_cast<float>(hw));
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total_vec; i += stride) {
    const int v = static_cast<int>(i % V);
    const long long pix = i / V;
    const int n = static_cast<int>(pix / hw);
    const int c = v * 8;
    float f[8], d[8], r[8];
    unpack8(gn_load(s, pix, c), f);
    unpack8(*reinterpret_cast<const bf16x8*>(dy + pix * lddy + c), d);
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
      const float xh = (f[e] - mean) * rstd;
      float dz = d[e];
      if (silu) dz *= silu_grad_f(xh * ga + beta[c + e]);
      r[e] = rstd * (dz * ga - s1 - xh * s2);
    }
    if (o.add0) {
      float a[8];
      unpack8(*reinterpret_cast<const bf16x8*>(o.add0 + pix * o.lda0 + c), a);
#pragma unroll
      for (int e = 0; e < 8; ++e) r[e] += a[e];
    }
    if (o.add1) {
      float a[8];
      unpack8(*reinterpret_cast<const bf16x8*>(o.add1 + pix * o.lda1 + c), a);
#pragma unroll
      for (int e = 0; e < 8; ++e) r[e] += a[e];
    }
    if (c < s.c0)
      *reinterpret_cast<bf16x8*>(o.d0 + pix * o.ld0 + c) = pack8(r);
    else if (o.d1)
      *reinterpret_cast<bf16x8*>(o.d1 + pix * o.ld1 + (c - s.c0)) = pack8(r);
  }
}

static int gn_check(const void* x0, int c0, long long ld0, const void* x1, int c1, long long ld1, int n, int hw,
                    int groups, const char* who) {
  if (!x0 || n <= 0 || hw <= 0 || groups <= 0 || groups > 64) {
    set_last_error("%s: bad argument (n=%d hw=%d groups=%d)", who, n, hw, groups);
    return DDPM_ERR_INVALID;
  }
  const int C = c0 + c1;
  if (c0 <= 0 || c0 % 8 || c1 < 0 || c1 % 8 || (c1 > 0 && !x1) || C % groups || C > kMaxC || ld0 % 8 || (c1 > 0 && ld1 % 8)) {
    set_last_error("%s: unsupported channel configuration c0=%d c1=%d groups=%d ld0=%lld ld1=%lld", who, c0, c1,
                   groups, ld0, ld1);
    return DDPM_ERR_INVALID;
  }
  if ((C / groups) < 4 && (C / groups) != 1 && (C / groups) != 2) {
    set_last_error("%s: channels per group %d unsupported", who, C / groups);
    return DDPM_ERR_INVALID;
  }
  return DDPM_OK;
}

static void gn_geometry(int C, int hw, int n, int* V, int* threads, int* pix_per_block, int* chunks) {
  *V = C / 8;
  int ppb = kGnThreads / *V;
  if (ppb < 1) ppb = 1;
  *threads = *V * ppb;
  // enough CTAs for ~4 waves, at least 8*ppb pixels per CTA
  long long want = (4LL * kNumSMs + n - 1) / n;
  long long ppblk = (hw + want - 1) / want;
  const long long min_ppblk = static_cast<long long>(ppb) * 8;
  if (ppblk < min_ppblk) ppblk = min_ppblk;
  if (ppblk > hw) ppblk = hw;
  *pix_per_block = static_cast<int>(ppblk);
  *chunks = static_cast<int>((hw + ppblk - 1) / ppblk);
}

static int stream_blocks(long long items) {
  long long b = (items + kGnThreads - 1) / kGnThreads;
  const long long cap = static_cast<long long>(kNumSMs) * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace ddpm

using namespace ddpm;

extern "C" int ddpm_gn_stats(const void* x0, int c0, long long ld0, const void* x1, int c1, long long ld1, int n,
                             int hw, int groups, float* stats, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = gn_check(x0, c0, ld0, x1, c1, ld1, n, hw, groups, "ddpm_gn_stats")) return e;
  DDPM_REQUIRE(stats, "ddpm_gn_stats: stats is null");
  const int C = c0 + c1;
  DDPM_REQUIRE(C / 8 <= kGnThreads, "ddpm_gn_stats: C=%d too large", C);
  GnSrc s{static_cast<const __nv_bfloat16*>(x0), static_cast<const __nv_bfloat16*>(x1), ld0, ld1, c0, c1};
  int V, threads, ppblk, chunks;
  gn_geometry(C, hw, n, &V, &threads, &ppblk, &chunks);
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
  const int V = C / 8;
  const long long total_vec = static_cast<long long>(n) * hw * V;
  gn_apply_kernel<<<stream_blocks(total_vec), kGnThreads, 0, stream>>>(s, hw, total_vec, C / groups, groups, stats, eps,
                                                                       gamma, beta, silu,
                                                                       static_cast<__nv_bfloat16*>(y), ldy, V);
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
  DDPM_REQUIRE(C / 8 <= kGnThreads, "ddpm_gn_bwd: C=%d too large", C);
  GnSrc s{static_cast<const __nv_bfloat16*>(x0), static_cast<const __nv_bfloat16*>(x1), ld0, ld1, c0, c1};
  int V, threads, ppblk, chunks;
  gn_geometry(C, hw, n, &V, &threads, &ppblk, &chunks);
  float* sums = ws;
  float* coef = ws + static_cast<long long>(n) * C * 2;
  DDPM_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * C * n, stream));
  gn_bwd_reduce_kernel<<<dim3(chunks, n), threads, C * 2 * sizeof(float), stream>>>(
      s, hw, C / groups, groups, stats, eps, gamma, beta, silu, static_cast<const __nv_bfloat16*>(dy), lddy, sums,
      ppblk, V);
  if (int e = check_launch("gn_bwd_reduce_kernel")) return e;
  gn_bwd_finalize_kernel<<<n + (C + kGnThreads - 1) / kGnThreads, kGnThreads, 0, stream>>>(sums, gamma, n, C,
                                                                                          C / groups, groups, coef,
                                                                                          dgamma, dbeta);
  if (int e = check_launch("gn_bwd_finalize_kernel")) return e;
  GnDst o{static_cast<__nv_bfloat16*>(dx0), static_cast<__nv_bfloat16*>(dx1), lddx0, lddx1,
          static_cast<const __nv_bfloat16*>(add0), static_cast<const __nv_bfloat16*>(add1), ldadd0, ldadd1};
  const long long total_vec = static_cast<long long>(n) * hw * V;
  gn_bwd_apply_kernel<<<stream_blocks(total_vec), kGnThreads, 0, stream>>>(
      s, hw, total_vec, C / groups, groups, stats, eps, gamma, beta, silu, static_cast<const __nv_bfloat16*>(dy), lddy,
      coef, o, V);
  return check_launch("gn_bwd_apply_kernel");
}
