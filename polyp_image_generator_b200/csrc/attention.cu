// Self-attention core for the UNet attention blocks (diffusers Attention + AttnProcessor2_0, i.e.
// F.scaled_dot_product_attention).  The reference configuration has
// head_dim = 8 (64 heads x 8) and at most 256 tokens, which is below the bf16 UMMA K = 16 and latency-bound,
// so the core is a SIMT kernel: one CTA per (sample, head), one thread per token, K/V staged in shared memory
// in fp32, exp2-domain softmax with fp32 statistics.  The q/k/v/out projections run on the tcgen05 GEMM.
#include "common.cuh"

#include "../../include/ddpm_b200.h"

namespace ddpm {

constexpr float kLog2e = 1.4426950408889634f;

// 8 bf16 (one 128-bit access) -> 8 floats at dst (16-byte aligned shared or local memory)
__device__ __forceinline__ void ld8_bf16(const __nv_bfloat16* src, float* dst, float scale = 1.0f) {
  float f[8];
  unpack8(*reinterpret_cast<const bf16x8*>(src), f);
  *reinterpret_cast<float4*>(dst) = make_float4(f[0] * scale, f[1] * scale, f[2] * scale, f[3] * scale);
  *reinterpret_cast<float4*>(dst + 4) = make_float4(f[4] * scale, f[5] * scale, f[6] * scale, f[7] * scale);
}

// qkv row layout: [q (heads*D) | k (heads*D) | v (heads*D)];  D % 8 == 0, every row pointer 16-byte aligned
template <int D>
__global__ void attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, long long ldqkv, __nv_bfloat16* __restrict__ o,
                                long long ldo, float* __restrict__ lse, int T, int heads, float scale_log2) {
  extern __shared__ __align__(16) float sm[];
  float* ks = sm;            // [T][D]
  float* vs = sm + T * D;    // [T][D]
  const int b = blockIdx.x / heads, hd = blockIdx.x - b * heads;
  const int C = heads * D;
  const __nv_bfloat16* base = qkv + static_cast<long long>(b) * T * ldqkv + hd * D;
  for (int i = threadIdx.x; i < T * (D / 8); i += blockDim.x) {
    const int j = i / (D / 8), e = (i - j * (D / 8)) * 8;
    ld8_bf16(base + j * ldqkv + C + e, ks + j * D + e);
    ld8_bf16(base + j * ldqkv + 2 * C + e, vs + j * D + e);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    float q[D], acc[D];
#pragma unroll
    for (int e = 0; e < D; e += 8) ld8_bf16(base + i * ldqkv + e, q + e, scale_log2);
#pragma unroll
    for (int e = 0; e < D; ++e) acc[e] = 0.f;
    float m = -INFINITY;
    for (int j = 0; j < T; ++j) {
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < D; e += 4) {
        const float4 k4 = *reinterpret_cast<const float4*>(ks + j * D + e);
        s += q[e] * k4.x + q[e + 1] * k4.y + q[e + 2] * k4.z + q[e + 3] * k4.w;
      }
      m = fmaxf(m, s);
    }
    float l = 0.f;
    for (int j = 0; j < T; ++j) {
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < D; e += 4) {
        const float4 k4 = *reinterpret_cast<const float4*>(ks + j * D + e);
        s += q[e] * k4.x + q[e + 1] * k4.y + q[e + 2] * k4.z + q[e + 3] * k4.w;
      }
      const float p = exp2f(s - m);
      l += p;
#pragma unroll
      for (int e = 0; e < D; e += 4) {
        const float4 v4 = *reinterpret_cast<const float4*>(vs + j * D + e);
        acc[e] += p * v4.x; acc[e + 1] += p * v4.y; acc[e + 2] += p * v4.z; acc[e + 3] += p * v4.w;
      }
    }
    const float inv = 1.0f / l;
    __nv_bfloat16* op = o + (static_cast<long long>(b) * T + i) * ldo + hd * D;
#pragma unroll
    for (int e = 0; e < D; e += 8) {
      float f[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) f[u] = acc[e + u] * inv;
      *reinterpret_cast<bf16x8*>(op + e) = pack8(f);
    }
    lse[(static_cast<long long>(b) * heads + hd) * T + i] = m + log2f(l);  // log2-domain logsumexp
  }
}

// fp32-faithful mode (include/ddpm_b200.h): the same core on SPLIT rows -- qkv rows are [q k v hi (3C) | q k v lo (3C)],
// o rows [hi (C) | lo (C)], value = hi + lo; exact exp2f / division.  Inference only (no log-sum-exp output).
__device__ __forceinline__ void ld8_split(const __nv_bfloat16* hp, int lo_off, float* dst, float scale = 1.0f) {
  float f[8], g[8];
  unpack8(*reinterpret_cast<const bf16x8*>(hp), f);
  unpack8(*reinterpret_cast<const bf16x8*>(hp + lo_off), g);
#pragma unroll
  for (int u = 0; u < 8; ++u) dst[u] = (f[u] + g[u]) * scale;
}

template <int D>
__global__ void attn_fwd_split_kernel(const __nv_bfloat16* __restrict__ qkv, long long ldqkv,
                                      __nv_bfloat16* __restrict__ o, long long ldo, int T, int heads, float scale_log2) {
  extern __shared__ __align__(16) float sm[];
  float* ks = sm;            // [T][D]
  float* vs = sm + T * D;    // [T][D]
  const int b = blockIdx.x / heads, hd = blockIdx.x - b * heads;
  const int C = heads * D;
  const __nv_bfloat16* base = qkv + static_cast<long long>(b) * T * ldqkv + hd * D;
  for (int i = threadIdx.x; i < T * (D / 8); i += blockDim.x) {
    const int j = i / (D / 8), e = (i - j * (D / 8)) * 8;
    ld8_split(base + j * ldqkv + C + e, 3 * C, ks + j * D + e);
    ld8_split(base + j * ldqkv + 2 * C + e, 3 * C, vs + j * D + e);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    float q[D], acc[D];
#pragma unroll
    for (int e = 0; e < D; e += 8) ld8_split(base + i * ldqkv + e, 3 * C, q + e, scale_log2);
#pragma unroll
    for (int e = 0; e < D; ++e) acc[e] = 0.f;
    float m = -INFINITY;
    for (int j = 0; j < T; ++j) {
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < D; ++e) s = fmaf(q[e], ks[j * D + e], s);
      m = fmaxf(m, s);
    }
    float l = 0.f;
    for (int j = 0; j < T; ++j) {
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < D; ++e) s = fmaf(q[e], ks[j * D + e], s);
      const float p = exp2f(s - m);
      l += p;
#pragma unroll
      for (int e = 0; e < D; ++e) acc[e] = fmaf(p, vs[j * D + e], acc[e]);
    }
    __nv_bfloat16* op = o + (static_cast<long long>(b) * T + i) * ldo + hd * D;
#pragma unroll
    for (int e = 0; e < D; e += 8) {
      float f[8], h[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) f[u] = acc[e + u] / l;
      const bf16x8 hi = pack8(f);
      unpack8(hi, h);
#pragma unroll
      for (int u = 0; u < 8; ++u) f[u] -= h[u];
      *reinterpret_cast<bf16x8*>(op + e) = hi;
      *reinterpret_cast<bf16x8*>(op + C + e) = pack8(f);
    }
  }
}

template <int D>
__global__ void attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, long long ldqkv,
                                const __nv_bfloat16* __restrict__ o, long long ldo,
                                const __nv_bfloat16* __restrict__ d_o, long long lddo, const float* __restrict__ lse,
                                __nv_bfloat16* __restrict__ dqkv, long long lddqkv, int T, int heads, float scale,
                                float scale_log2) {
  extern __shared__ __align__(16) float sm[];
  float* qs = sm;                 // [T][D]  (pre-scaled by scale*log2e)
  float* ks = qs + T * D;
  float* vs = ks + T * D;
  float* dos = vs + T * D;
  float* Ls = dos + T * D;        // [T]
  float* Ds = Ls + T;             // [T]  D_i = dO_i . O_i
  const int b = blockIdx.x / heads, hd = blockIdx.x - b * heads;
  const int C = heads * D;
  const __nv_bfloat16* base = qkv + static_cast<long long>(b) * T * ldqkv + hd * D;
  for (int i = threadIdx.x; i < T * (D / 8); i += blockDim.x) {
    const int j = i / (D / 8), e = (i - j * (D / 8)) * 8;
    ld8_bf16(base + j * ldqkv + e, qs + j * D + e, scale_log2);
    ld8_bf16(base + j * ldqkv + C + e, ks + j * D + e);
    ld8_bf16(base + j * ldqkv + 2 * C + e, vs + j * D + e);
    ld8_bf16(d_o + (static_cast<long long>(b) * T + j) * lddo + hd * D + e, dos + j * D + e);
  }
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    Ls[i] = lse[(static_cast<long long>(b) * heads + hd) * T + i];
    float dsum = 0.f;
#pragma unroll
    for (int e = 0; e < D; e += 8) {
      float a8[8], b8[8];
      unpack8(*reinterpret_cast<const bf16x8*>(d_o + (static_cast<long long>(b) * T + i) * lddo + hd * D + e), a8);
      unpack8(*reinterpret_cast<const bf16x8*>(o + (static_cast<long long>(b) * T + i) * ldo + hd * D + e), b8);
#pragma unroll
      for (int u = 0; u < 8; ++u) dsum += a8[u] * b8[u];
    }
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
        float kk[D];
#pragma unroll
        for (int e = 0; e < D; e += 4) {
          const float4 k4 = *reinterpret_cast<const float4*>(ks + j * D + e);
          const float4 v4 = *reinterpret_cast<const float4*>(vs + j * D + e);
          kk[e] = k4.x; kk[e + 1] = k4.y; kk[e + 2] = k4.z; kk[e + 3] = k4.w;
          s += q[e] * k4.x + q[e + 1] * k4.y + q[e + 2] * k4.z + q[e + 3] * k4.w;
          dp += dov[e] * v4.x + dov[e + 1] * v4.y + dov[e + 2] * v4.z + dov[e + 3] * v4.w;
        }
        const float ds = exp2f(s - L) * (dp - Di);
#pragma unroll
        for (int e = 0; e < D; ++e) dq[e] += ds * kk[e];
      }
#pragma unroll
      for (int e = 0; e < D; e += 8) {
        float f[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) f[u] = dq[e + u] * scale;
        *reinterpret_cast<bf16x8*>(dbase + i * lddqkv + e) = pack8(f);
      }
    }
    // as key j = i: dv_j = sum_i p_ij dO_i ; dk_j = scale * sum_i dS_ij q_i
    {
      const int j = i;
      float k[D], v[D], dk[D], dv[D];
#pragma unroll
      for (int e = 0; e < D; ++e) { k[e] = ks[j * D + e]; v[e] = vs[j * D + e]; dk[e] = 0.f; dv[e] = 0.f; }
      for (int ii = 0; ii < T; ++ii) {
        float s = 0.f, dp = 0.f;
        float qq[D], dd[D];
#pragma unroll
        for (int e = 0; e < D; e += 4) {
          const float4 q4 = *reinterpret_cast<const float4*>(qs + ii * D + e);
          const float4 d4 = *reinterpret_cast<const float4*>(dos + ii * D + e);
          qq[e] = q4.x; qq[e + 1] = q4.y; qq[e + 2] = q4.z; qq[e + 3] = q4.w;
          dd[e] = d4.x; dd[e + 1] = d4.y; dd[e + 2] = d4.z; dd[e + 3] = d4.w;
          s += q4.x * k[e] + q4.y * k[e + 1] + q4.z * k[e + 2] + q4.w * k[e + 3];
          dp += d4.x * v[e] + d4.y * v[e + 1] + d4.z * v[e + 2] + d4.w * v[e + 3];
        }
        const float p = exp2f(s - Ls[ii]);
        const float ds = p * (dp - Ds[ii]);
#pragma unroll
        for (int e = 0; e < D; ++e) { dv[e] += p * dd[e]; dk[e] += ds * qq[e]; }
      }
      // qs is pre-scaled by scale*log2e: dk = scale * sum dS * q_raw = sum dS * qs / log2e
#pragma unroll
      for (int e = 0; e < D; e += 8) {
        float fk[8], fv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { fk[u] = dk[e + u] * (1.0f / kLog2e); fv[u] = dv[e + u]; }
        *reinterpret_cast<bf16x8*>(dbase + j * lddqkv + C + e) = pack8(fk);
        *reinterpret_cast<bf16x8*>(dbase + j * lddqkv + 2 * C + e) = pack8(fv);
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
  DDPM_REQUIRE(ldqkv % 8 == 0 && ldo % 8 == 0 && d % 8 == 0, "ddpm_attn_fwd: rows must be 16-byte aligned");
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

extern "C" int ddpm_attn_fwd_split(const void* qkv, long long ldqkv, void* o, long long ldo, int b, int t, int heads,
                                   int d, float scale, void* stream) {
  DDPM_REQUIRE(qkv && o && b > 0 && t > 0 && heads > 0, "ddpm_attn_fwd_split: bad argument");
  DDPM_REQUIRE(ldqkv % 8 == 0 && ldo % 8 == 0 && d % 8 == 0 && ldqkv >= 6LL * heads * d && ldo >= 2LL * heads * d,
               "ddpm_attn_fwd_split: rows must be 16-byte aligned split rows [hi | lo]");
  const size_t smem = sizeof(float) * 2 * t * d;
  DDPM_REQUIRE(smem <= 200 * 1024, "ddpm_attn_fwd_split: t=%d d=%d does not fit shared memory", t, d);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ATTN_DISPATCH(d, {
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
      DDPM_CUDA(cudaFuncSetAttribute(attn_fwd_split_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
    attn_fwd_split_kernel<D><<<b * heads, attn_threads(t), smem, st>>>(static_cast<const __nv_bfloat16*>(qkv), ldqkv,
                                                                      static_cast<__nv_bfloat16*>(o), ldo, t, heads,
                                                                      scale * kLog2e);
  });
  return check_launch("attn_fwd_split_kernel");
}

extern "C" int ddpm_attn_bwd(const void* qkv, long long ldqkv, const void* o, long long ldo, const void* d_o,
                             long long lddo, const float* lse, void* dqkv, long long lddqkv, int b, int t, int heads,
                             int d, float scale, void* stream) {
  DDPM_REQUIRE(qkv && o && d_o && lse && dqkv && b > 0 && t > 0 && heads > 0, "ddpm_attn_bwd: bad argument");
  DDPM_REQUIRE(ldqkv % 8 == 0 && ldo % 8 == 0 && lddo % 8 == 0 && lddqkv % 8 == 0 && d % 8 == 0,
               "ddpm_attn_bwd: rows must be 16-byte aligned");
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
