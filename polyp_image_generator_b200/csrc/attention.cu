// Self-attention core for the UNet attention blocks (diffusers Attention + AttnProcessor2_0, i.e.
// F.scaled_dot_product_attention).  The reference configuration has
// head_dim = 8 (64 heads x 8) and at most 256 tokens, which is below the bf16 UMMA K = 16 and latency-bound,
// so the core is a SIMT kernel: one CTA per (sample, head), one thread per token, K/V staged in shared memory
// in fp32, exp2-domain softmax with fp32 statistics.  The q/k/v/out projections run on the tcgen05 GEMM.
#include "common.cuh"

#include "../../include/ddpm_b200.h"

namespace ddpm {

constexpr float kLog2e = 1.4426950408889634f;

// qkv row layout: [q (heads*D) | k (heads*D) | v (heads*D)]
template <int D>
__global__ void attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, long long ldqkv, __nv_bfloat16* __restrict__ o,
                                long long ldo, float* __restrict__ lse, int T, int heads, float scale_log2) {
  extern __shared__ float sm[];
  float* ks = sm;            // [T][D]
  float* vs = sm + T * D;    // [T][D]
  const int b = blockIdx.x / heads, hd = blockIdx.x - b * heads;
  const int C = heads * D;
  const __nv_bfloat16* base = qkv + static_cast<long long>(b) * T * ldqkv + hd * D;
  for (int i = threadIdx.x; i < T * D; i += blockDim.x) {
    const int j = i / D, e = i - j * D;
    ks[i] = __bfloat162float(base[j * ldqkv + C + e]);
    vs[i] = __bfloat162float(base[j * ldqkv + 2 * C + e]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    float q[D], acc[D];
#pragma unroll
    for (int e = 0; e < D; ++e) {
      q[e] = __bfloat162float(base[i * ldqkv + e]) * scale_log2;
      acc[e] = 0.f;
    }
    float m = -INFINITY;
    for (int j = 0; j < T; ++j) {
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < D; ++e) s += q[e] * ks[j * D + e];
      m = fmaxf(m, s);
    }
    float l = 0.f;
    for (int j = 0; j < T; ++j) {
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < D; ++e) s += q[e] * ks[j * D + e];
      const float p = exp2f(s - m);
      l += p;
#pragma unroll
      for (int e = 0; e < D; ++e) acc[e] += p * vs[j * D + e];
    }
    const float inv = 1.0f / l;
    __nv_bfloat16* op = o + (static_cast<long long>(b) * T + i) * ldo + hd * D;
#pragma unroll
    for (int e = 0; e < D; ++e) op[e] = __float2bfloat16(acc[e] * inv);
    lse[(static_cast<long long>(b) * heads + hd) * T + i] = m + log2f(l);  // log2-domain logsumexp
  }
}

template <int D>
__global__ void attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, long long ldqkv,
                                const __nv_bfloat16* __restrict__ o, long long ldo,
                                const __nv_bfloat16* __restrict__ d_o, long long lddo, const float* __restrict__ lse,
                                __nv_bfloat16* __restrict__ dqkv, long long lddqkv, int T, int heads, float scale,
                                float scale_log2) {
  extern __shared__ float sm[];
  float* qs = sm;                 // [T][D]  (pre-scaled by scale*log2e)
  float* ks = qs + T * D;
  float* vs = ks + T * D;
  float* dos = vs + T * D;
  float* Ls = dos + T * D;        // [T]
  float* Ds = Ls + T;             // [T]  D_i = dO_i . O_i
  const int b = blockIdx.x / heads, hd = blockIdx.x - b * heads;
  const int C = heads * D;
  const __nv_bfloat16* base = qkv + static_cast<long long>(b) * T * ldqkv + hd * D;
  for (int i = threadIdx.x; i < T * D; i += blockDim.x) {
    const int j = i / D, e = i - j * D;
    qs[i] = __bfloat162float(base[j * ldqkv + e]) * scale_log2;
    ks[i] = __bfloat162float(base[j * ldqkv + C + e]);
    vs[i] = __bfloat162float(base[j * ldqkv + 2 * C + e]);
    dos[i] = __bfloat162float(d_o[(static_cast<long long>(b) * T + j) * lddo + hd * D + e]);
  }
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    Ls[i] = lse[(static_cast<long long>(b) * heads + hd) * T + i];
    float dsum = 0.f;
    for (int e = 0; e < D; ++e)
      dsum += __bfloat162float(d_o[(static_cast<long long>(b) * T + i) * lddo + hd * D + e]) *
              __bfloat162float(o[(static_cast<long long>(b) * T + i) * ldo + hd * D + e]);
    Ds[i] = dsum;
  }
  __syncthreads();
  __nv_bfloat16* dbase = dqkv + static_cast<long long>(b) * T * lddqkv + hd * D;
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    // as query i: dq_i = scale * sum_j dS_ij k_j
    {
      float q[D], dov[D], dq[D];
#pragma unroll
      for (int e = 0; e < D; ++e) { q[e] = qs[i * D + e]; dov[e] = dos[i * D + e]; dq[e] = 0.f; }
      const float L = Ls[i], Di = Ds[i];
      for (int j = 0; j < T; ++j) {
        float s = 0.f, dp = 0.f;
#pragma unroll
        for (int e = 0; e < D; ++e) { s += q[e] * ks[j * D + e]; dp += dov[e] * vs[j * D + e]; }
        const float ds = exp2f(s - L) * (dp - Di);
#pragma unroll
        for (int e = 0; e < D; ++e) dq[e] += ds * ks[j * D + e];
      }
#pragma unroll
      for (int e = 0; e < D; ++e) dbase[i * lddqkv + e] = __float2bfloat16(dq[e] * scale);
    }
    // as key j = i: dv_j = sum_i p_ij dO_i ; dk_j = scale * sum_i dS_ij q_i
    {
      const int j = i;
      float k[D], v[D], dk[D], dv[D];
#pragma unroll
      for (int e = 0; e < D; ++e) { k[e] = ks[j * D + e]; v[e] = vs[j * D + e]; dk[e] = 0.f; dv[e] = 0.f; }
      for (int ii = 0; ii < T; ++ii) {
        float s = 0.f, dp = 0.f;
#pragma unroll
        for (int e = 0; e < D; ++e) { s += qs[ii * D + e] * k[e]; dp += dos[ii * D + e] * v[e]; }
        const float p = exp2f(s - Ls[ii]);
        const float ds = p * (dp - Ds[ii]);
#pragma unroll
        for (int e = 0; e < D; ++e) { dv[e] += p * dos[ii * D + e]; dk[e] += ds * qs[ii * D + e]; }
      }
      // qs is pre-scaled by scale*log2e: dk = scale * sum dS * q_raw = sum dS * qs / log2e
#pragma unroll
      for (int e = 0; e < D; ++e) {
        dbase[j * lddqkv + C + e] = __float2bfloat16(dk[e] * (1.0f / kLog2e));
        dbase[j * lddqkv + 2 * C + e] = __float2bfloat16(dv[e]);
      }
    }
  }
}

static int attn_threads(int t) {
  int th = ((t + 31) / 32) * 32;
  if (th > 256) th = 256;
  if (th < 32) th = 32;
  return th;
}

}  // namespace ddpm

using namespace ddpm;

#define ATTN_DISPATCH(D_, ...)          \
  switch (D_) {                         \
    case 8: { constexpr int D = 8; __VA_ARGS__; } break;   \
    case 16: { constexpr int D = 16; __VA_ARGS__; } break; \
    case 32: { constexpr int D = 32; __VA_ARGS__; } break; \
    case 64: { constexpr int D = 64; __VA_ARGS__; } break; \
    default:                            \
      set_last_error("attention head_dim=%d unsupported (8/16/32/64)", D_); \
      return DDPM_ERR_UNSUPPORTED;      \
  }

extern "C" int ddpm_attn_fwd(const void* qkv, long long ldqkv, void* o, long long ldo, float* lse, int b, int t,
                             int heads, int d, float scale, void* stream) {
  DDPM_REQUIRE(qkv && o && lse && b > 0 && t > 0 && heads > 0, "ddpm_attn_fwd: bad argument");
  const size_t smem = sizeof(float) * 2 * t * d;
  DDPM_REQUIRE(smem <= 200 * 1024, "ddpm_attn_fwd: t=%d d=%d does not fit shared memory", t, d);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ATTN_DISPATCH(d, {
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
      DDPM_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
    attn_fwd_kernel<D><<<b * heads, attn_threads(t), smem, st>>>(static_cast<const __nv_bfloat16*>(qkv), ldqkv,
                                                                static_cast<__nv_bfloat16*>(o), ldo, lse, t, heads,
                                                                scale * kLog2e);
  });
  return check_launch("attn_fwd_kernel");
}

extern "C" int ddpm_attn_bwd(const void* qkv, long long ldqkv, const void* o, long long ldo, const void* d_o,
                             long long lddo, const float* lse, void* dqkv, long long lddqkv, int b, int t, int heads,
                             int d, float scale, void* stream) {
  DDPM_REQUIRE(qkv && o && d_o && lse && dqkv && b > 0 && t > 0 && heads > 0, "ddpm_attn_bwd: bad argument");
  const size_t smem = sizeof(float) * (4 * t * d + 2 * t);
  DDPM_REQUIRE(smem <= 200 * 1024, "ddpm_attn_bwd: t=%d d=%d does not fit shared memory", t, d);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ATTN_DISPATCH(d, {
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
      DDPM_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
    attn_bwd_kernel<D><<<b * heads, attn_threads(t), smem, st>>>(
        static_cast<const __nv_bfloat16*>(qkv), ldqkv, static_cast<const __nv_bfloat16*>(o), ldo,
        static_cast<const __nv_bfloat16*>(d_o), lddo, lse, static_cast<__nv_bfloat16*>(dqkv), lddqkv, t, heads, scale,
        scale * kLog2e);
  });
  return check_launch("attn_bwd_kernel");
}
