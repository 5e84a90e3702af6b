// GroupNorm(32)+SiLU forward / backward on NHWC bf16, fp32 statistics.
// Replaces at::native::group_norm + a separate silu kernel per site (71 sites per UNet forward), and folds
// torch.cat([h, skip], 1) into the load path: the input may be split over two tensors along channels.
//
// Thread mapping (all kernels): a CTA owns a slice of pixels of ONE sample; thread (v, pl) owns the 8-channel vector
// v (one 128-bit access) and walks pixels pl, pl+ppb, ... of the slice, so per-channel coefficients live in
// registers for the whole loop.
//
// ncu (profiles/r1_groupnorm.md) showed these streamers are NOT bandwidth-bound but issue-bound: at bf16 the SiLU
// derivative costs ~26 instructions and 4 MUFU ops per element against 6-10 bytes of traffic.  Hence
//   * all arithmetic is packed fp32x2 (FFMA2/FADD2/FMUL2 on sm_100a: two lanes per issue slot),
//   * sigmoid(z) = 0.5 + 0.5*tanh(z/2): ONE MUFU op (tanh.approx.f32) instead of ex2 + rcp,
//   * the two-phase ops run as "team" kernels (below) whose second phase re-reads from L2 instead of HBM.
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

struct GnDst {
  __nv_bfloat16* d0;
  __nv_bfloat16* d1;
  long long ld0, ld1;
  const __nv_bfloat16* add0;
  const __nv_bfloat16* add1;
  long long lda0, lda1;
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

__device__ __forceinline__ void gn_mean_rstd_of(float2 st, float inv_m, float eps, float* mean, float* rstd) {
  const float mu = st.x * inv_m;
  const float var = fmaxf(st.y * inv_m - mu * mu, 0.f);
  *mean = mu;
  *rstd = rsqrtf(var + eps);
}

// ---- packed fp32x2 helpers --------------------------------------------------------------------------------------
struct f2x4 {
  float2 p[4];   // 8 channels as 4 pairs
};
__device__ __forceinline__ f2x4 unpack8p(const bf16x8& r) {
  f2x4 o;
  o.p[0] = make_float2(bf16lo_f(r.x), bf16hi_f(r.x));
  o.p[1] = make_float2(bf16lo_f(r.y), bf16hi_f(r.y));
  o.p[2] = make_float2(bf16lo_f(r.z), bf16hi_f(r.z));
  o.p[3] = make_float2(bf16lo_f(r.w), bf16hi_f(r.w));
  return o;
}
__device__ __forceinline__ bf16x8 pack8p(const f2x4& v) {
  return make_uint4(pack2_bf16_(v.p[0].x, v.p[0].y), pack2_bf16_(v.p[1].x, v.p[1].y),
                    pack2_bf16_(v.p[2].x, v.p[2].y), pack2_bf16_(v.p[3].x, v.p[3].y));
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sigmoid of a pair: 0.5 + 0.5 * tanh(z / 2)
__device__ __forceinline__ float2 sigmoid2(float2 z) {
  const float2 h = __fmul2_rn(z, make_float2(0.5f, 0.5f));
  const float2 t = make_float2(tanh_approx(h.x), tanh_approx(h.y));
  return __ffma2_rn(t, make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));
}
template <bool SILU>
__device__ __forceinline__ float2 act2(float2 z) {
  if (!SILU) return z;
  return __fmul2_rn(z, sigmoid2(z));
}
// dz = dy * act'(z);  silu'(z) = s * (1 + z * (1 - s))
template <bool SILU>
__device__ __forceinline__ float2 dz2(float2 dy, float2 z) {
  if (!SILU) return dy;
  const float2 s = sigmoid2(z);
  const float2 om = __ffma2_rn(s, make_float2(-1.f, -1.f), make_float2(1.f, 1.f));
  const float2 u = __ffma2_rn(z, om, make_float2(1.f, 1.f));
  return __fmul2_rn(__fmul2_rn(dy, s), u);
}

// ---- statistics (standalone) ----------------------------------------------------------------------------------
// accumulate (sum, sumsq) of this thread's 8 channels over pixels p_first, p_first + ppb, ... < p_end
__device__ __forceinline__ void gn_accumulate_moments(const __nv_bfloat16* xp, long long ld, int p_first, int p_end,
                                                      int ppb, f2x4* sum, f2x4* sq) {
#pragma unroll
  for (int j = 0; j < 4; ++j) sum->p[j] = sq->p[j] = make_float2(0.f, 0.f);
  int p = p_first;
  for (; p + (kUnroll - 1) * ppb < p_end; p += kUnroll * ppb) {
    bf16x8 raw[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) raw[u] = *reinterpret_cast<const bf16x8*>(xp + (p + u * ppb) * ld);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const f2x4 f = unpack8p(raw[u]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sum->p[j] = __fadd2_rn(sum->p[j], f.p[j]);
        sq->p[j] = __ffma2_rn(f.p[j], f.p[j], sq->p[j]);
      }
    }
  }
  for (; p < p_end; p += ppb) {
    const f2x4 f = unpack8p(*reinterpret_cast<const bf16x8*>(xp + p * ld));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      sum->p[j] = __fadd2_rn(sum->p[j], f.p[j]);
      sq->p[j] = __ffma2_rn(f.p[j], f.p[j], sq->p[j]);
    }
  }
}

// fold this thread's 8 channel moments into the per-group shared accumulators sm[g*2 + {0,1}]
__device__ __forceinline__ void gn_fold_groups(const f2x4& sum, const f2x4& sq, int c, int cpg, float* sm) {
  const float s8[8] = {sum.p[0].x, sum.p[0].y, sum.p[1].x, sum.p[1].y, sum.p[2].x, sum.p[2].y, sum.p[3].x, sum.p[3].y};
  const float q8[8] = {sq.p[0].x, sq.p[0].y, sq.p[1].x, sq.p[1].y, sq.p[2].x, sq.p[2].y, sq.p[3].x, sq.p[3].y};
  int g_prev = c / cpg;
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int g = (c + e) / cpg;
    if (g != g_prev) {
      atomicAdd(&sm[g_prev * 2], a);
      atomicAdd(&sm[g_prev * 2 + 1], b);
      a = b = 0.f;
      g_prev = g;
    }
    a += s8[e];
    b += q8[e];
  }
  atomicAdd(&sm[g_prev * 2], a);
  atomicAdd(&sm[g_prev * 2 + 1], b);
}

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
  f2x4 sum, sq;
  gn_accumulate_moments(xp, ld, p_begin + pl, p_end, ppb, &sum, &sq);
  gn_fold_groups(sum, sq, v * 8, cpg, sm);
  __syncthreads();
  for (int i = threadIdx.x; i < groups * 2; i += blockDim.x)
    atomicAdd(&stats[static_cast<long long>(n) * groups * 2 + i], sm[i]);
}

// ---- forward apply (standalone) ---------------------------------------------------------------------------------
// y = act(x * ka + kb) with ka = rstd*gamma, kb = beta - mean*rstd*gamma (per channel, in registers)
__device__ __forceinline__ void gn_apply_coefs(const float* stats_n, int c, int cpg, float inv_m, float eps,
                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                               bool bypass_l1, f2x4* ka, f2x4* kb) {
  float a8[8], b8[8];
  int g_prev = -1;
  float mean = 0.f, rstd = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int g = (c + e) / cpg;
    if (g != g_prev) {
      const float2* sp = reinterpret_cast<const float2*>(stats_n + g * 2);
      gn_mean_rstd_of(bypass_l1 ? __ldcg(sp) : *sp, inv_m, eps, &mean, &rstd);
      g_prev = g;
    }
    a8[e] = rstd * gamma[c + e];
    b8[e] = beta[c + e] - mean * a8[e];
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    ka->p[j] = make_float2(a8[2 * j], a8[2 * j + 1]);
    kb->p[j] = make_float2(b8[2 * j], b8[2 * j + 1]);
  }
}

template <bool SILU>
__device__ __forceinline__ void gn_apply_stream(const __nv_bfloat16* xp, long long ld, __nv_bfloat16* yp, long long ldy,
                                                int p_first, int p_end, int ppb, const f2x4& ka, const f2x4& kb) {
  int p = p_first;
  for (; p + (kUnroll - 1) * ppb < p_end; p += kUnroll * ppb) {
    bf16x8 raw[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) raw[u] = *reinterpret_cast<const bf16x8*>(xp + (p + u * ppb) * ld);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      f2x4 f = unpack8p(raw[u]);
#pragma unroll
      for (int j = 0; j < 4; ++j) f.p[j] = act2<SILU>(__ffma2_rn(f.p[j], ka.p[j], kb.p[j]));
      *reinterpret_cast<bf16x8*>(yp + (p + u * ppb) * ldy) = pack8p(f);
    }
  }
  for (; p < p_end; p += ppb) {
    f2x4 f = unpack8p(*reinterpret_cast<const bf16x8*>(xp + p * ld));
#pragma unroll
    for (int j = 0; j < 4; ++j) f.p[j] = act2<SILU>(__ffma2_rn(f.p[j], ka.p[j], kb.p[j]));
    *reinterpret_cast<bf16x8*>(yp + p * ldy) = pack8p(f);
  }
}

template <bool SILU>
__global__ void __launch_bounds__(kGnThreads, 4)
gn_apply_kernel(GnSrc s, int hw, int cpg, int groups, const float* __restrict__ stats, float eps,
                const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                long long ldy, int pix_per_block, int V, float* __restrict__ coef /*[N][C/2][4] or NULL*/) {
  const float inv_m = 1.0f / (static_cast<float>(cpg) * static_cast<float>(hw));
  const int n = blockIdx.y;
  const int v = threadIdx.x % V, pl = threadIdx.x / V, ppb = blockDim.x / V;
  const int c = v * 8;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(hw, p_begin + pix_per_block);
  long long ld;
  const __nv_bfloat16* xp = gn_base(s, n, hw, c, &ld);
  __nv_bfloat16* yp = y + static_cast<long long>(n) * hw * ldy + c;
  f2x4 ka, kb;
  gn_apply_coefs(stats + static_cast<long long>(n) * groups * 2, c, cpg, inv_m, eps, gamma, beta, false, &ka, &kb);
  if (coef != nullptr && blockIdx.x == 0 && pl == 0) {     // affine table for the GroupNorm-backward conv fusion
    float4* cp = reinterpret_cast<float4*>(coef) + (static_cast<long long>(n) * V * 8 + c) / 2;
#pragma unroll
    for (int j = 0; j < 4; ++j) cp[j] = make_float4(ka.p[j].x, ka.p[j].y, kb.p[j].x, kb.p[j].y);
  }
  gn_apply_stream<SILU>(xp, ld, yp, ldy, p_begin + pl, p_end, ppb, ka, kb);
}

// stats[n][g] = (sum, sum of squares) over the group's channels of the per-(sample, 4-channel granule) moments that the
// producing conv epilogues accumulated (ddpm_conv_args.out_csum: [n][c / 4][2]); the input may be a channel concat of two
// tensors.  Group sizes are multiples of 4 channels (host-checked), so granules never straddle a group.
__global__ void __launch_bounds__(256)
gn_stats_from_csum_kernel(const float* __restrict__ cs0, int c0, const float* __restrict__ cs1, int c1, int n, int groups,
                          float* __restrict__ stats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * groups) return;
  const int b = i / groups, g = i - b * groups;
  const int gpg = (c0 + c1) / groups / 4;       // granules per group
  const int q0 = c0 / 4, q1 = c1 / 4;
  float a = 0.f, q = 0.f;
  for (int k = g * gpg; k < (g + 1) * gpg; ++k) {
    const float2 m = k < q0 ? *reinterpret_cast<const float2*>(cs0 + (static_cast<long long>(b) * q0 + k) * 2)
                            : *reinterpret_cast<const float2*>(cs1 + (static_cast<long long>(b) * q1 + (k - q0)) * 2);
    a += m.x;
    q += m.y;
  }
  stats[i * 2] = a;
  stats[i * 2 + 1] = q;
}

// =====================================================================================================
// Fused two-phase kernels ("teams"): one persistent, cooperatively launched grid.  A team of `team_size` CTAs owns
// one sample at a time: phase 1 reduces (statistics resp. backward sums) over the team's pixel slices, the team
// meets at a global-memory barrier, phase 2 re-reads the SAME slices -- now L2 hits, because only
// teams * sample_bytes (<= the L2 budget) are live between the phases -- and writes the result.  HBM traffic drops
// from 6 to 4 B/elem (forward) and from 10 to 6 B/elem (backward).  CTAs of different teams share SMs, so one
// team's barrier wait is covered by another team's streaming.
// =====================================================================================================
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// all threads of the CTA call this; `counter` is zero-initialised per (launch, sample)
__device__ __forceinline__ void gn_team_barrier(int* counter, int expected) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(counter, 1);
    unsigned spins = 0;
    while (ld_acquire_gpu(counter) < expected) {
      __nanosleep(40);
      if (++spins > (1u << 26)) {
        printf("ddpm_b200: GroupNorm team barrier timed out (block %d)\n", blockIdx.x);
        __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
}

struct GnTeam {
  int N, hw, cpg, groups, V, team_size, pix_per_cta;
  float eps;
  int slab_ch;   // solo launches (team_size == 1) only: channels per CTA, blockIdx.y = slab (whole groups); 0 = all
};

// ---- forward: statistics + normalise/affine/SiLU --------------------------------------------------------------
template <bool SILU>
__global__ void __launch_bounds__(kGnThreads, 4)
gn_fwd_fused_kernel(GnSrc s, GnTeam t, float* __restrict__ stats /*[N][groups][2], zeroed*/,
                    int* __restrict__ counters /*[N], zeroed*/, const float* __restrict__ gamma,
                    const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, long long ldy,
                    float* __restrict__ coef /*[N][C/2][4] or NULL*/) {
  __shared__ float sm[64 * 2];
  const int team = blockIdx.x / t.team_size, rank = blockIdx.x - team * t.team_size;
  const int nteams = gridDim.x / t.team_size;
  const int V = t.V, hw = t.hw, cpg = t.cpg, groups = t.groups;
  // solo launches may cut a sample into channel slabs of whole groups (blockIdx.y): groups are independent, so the
  // slabs need no exchange at all and a 64-sample batch of small maps fills the GPU instead of 64 SMs
  const int Vt = t.slab_ch ? t.slab_ch / 8 : V;                   // threads per pixel
  const int c_base = t.slab_ch ? static_cast<int>(blockIdx.y) * t.slab_ch : 0;
  const int v = threadIdx.x % Vt, pl = threadIdx.x / Vt, ppb = blockDim.x / Vt;
  const int c = c_base + v * 8;
  const int p_begin = rank * t.pix_per_cta;
  const int p_end = min(hw, p_begin + t.pix_per_cta);
  const float inv_m = 1.0f / (static_cast<float>(cpg) * static_cast<float>(hw));
  for (int n = team; n < t.N; n += nteams) {
    for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    long long ld;
    const __nv_bfloat16* xp = gn_base(s, n, hw, c, &ld);
    {
      f2x4 sum, sq;
      gn_accumulate_moments(xp, ld, p_begin + pl, p_end, ppb, &sum, &sq);
      gn_fold_groups(sum, sq, c, cpg, sm);
    }
    __syncthreads();
    float* st = stats + static_cast<long long>(n) * groups * 2;
    f2x4 ka, kb;
    if (t.team_size == 1) {
      // small maps: the CTA owns the whole sample (or slab) -- statistics stay in shared memory, written out once
      const int i_lo = (c_base / cpg) * 2, i_hi = t.slab_ch ? ((c_base + t.slab_ch) / cpg) * 2 : groups * 2;
      for (int i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) st[i] = sm[i];
      gn_apply_coefs(sm, c, cpg, inv_m, t.eps, gamma, beta, false, &ka, &kb);
    } else {
      for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) atomicAdd(&st[i], sm[i]);
      gn_team_barrier(&counters[n], t.team_size);
      gn_apply_coefs(st, c, cpg, inv_m, t.eps, gamma, beta, true, &ka, &kb);
    }
    if (coef != nullptr && rank == 0 && pl == 0) {
      float4* cp = reinterpret_cast<float4*>(coef) + (static_cast<long long>(n) * V * 8 + c) / 2;
#pragma unroll
      for (int j = 0; j < 4; ++j) cp[j] = make_float4(ka.p[j].x, ka.p[j].y, kb.p[j].x, kb.p[j].y);
    }
    __nv_bfloat16* yp = y + static_cast<long long>(n) * hw * ldy + c;
    gn_apply_stream<SILU>(xp, ld, yp, ldy, p_begin + pl, p_end, ppb, ka, kb);
    __syncthreads();   // sm is reused by the next sample
  }
}

// ---- backward: per-channel sums -> group coefficients -> dx -------------------------------------------------------
// dynamic smem: C*2 floats (per-channel sums of this CTA) + groups*2 floats (group coefficients)
// z = x*kz1 + kz0;  dx = dz*kz1 + x*nk4 + nk5 (+ addends) with nk4 = -rstd^2*s2/m, nk5 = mean*rstd^2*s2/m - rstd*s1/m
template <bool SILU>
__global__ void __launch_bounds__(kGnThreads, 2)
gn_bwd_fused_kernel(GnSrc s, GnTeam t, const float* __restrict__ stats, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const __nv_bfloat16* __restrict__ dy, long long lddy,
                    float* __restrict__ sums /*[N][C][2], zeroed*/, int* __restrict__ counters /*[N], zeroed*/,
                    GnDst o) {
  extern __shared__ float smc[];
  const int team = blockIdx.x / t.team_size, rank = blockIdx.x - team * t.team_size;
  const int nteams = gridDim.x / t.team_size;
  const int V = t.V, hw = t.hw, cpg = t.cpg, groups = t.groups;
  const int C = V * 8;
  float* scoef = smc + C * 2;
  const int Vt = t.slab_ch ? t.slab_ch / 8 : V;                   // channel slabs of a solo launch: see the forward
  const int c_base = t.slab_ch ? static_cast<int>(blockIdx.y) * t.slab_ch : 0;
  const int v = threadIdx.x % Vt, pl = threadIdx.x / Vt, ppb = blockDim.x / Vt;
  const int c = c_base + v * 8;
  const int p_begin = rank * t.pix_per_cta;
  const int p_end = min(hw, p_begin + t.pix_per_cta);
  const float inv_m = 1.0f / (static_cast<float>(cpg) * static_cast<float>(hw));
  const bool writes = (c < s.c0) || (o.d1 != nullptr);
  for (int n = team; n < t.N; n += nteams) {
    for (int i = threadIdx.x; i < C * 2; i += blockDim.x) smc[i] = 0.f;
    __syncthreads();
    long long ld;
    const __nv_bfloat16* xp = gn_base(s, n, hw, c, &ld);
    const long long pix0 = static_cast<long long>(n) * hw;
    const __nv_bfloat16* dp = dy + pix0 * lddy + c;
    const float* stats_n = stats + static_cast<long long>(n) * groups * 2;
    f2x4 kz1, kz0;
    gn_apply_coefs(stats_n, c, cpg, inv_m, t.eps, gamma, beta, false, &kz1, &kz0);
    // ---- phase 1: S1 = sum dz, S2 = sum dz*x ----
    {
      f2x4 A, B;
#pragma unroll
      for (int j = 0; j < 4; ++j) A.p[j] = B.p[j] = make_float2(0.f, 0.f);
      int p = p_begin + pl;
      for (; p + ppb < p_end; p += 2 * ppb) {
        bf16x8 rx[2], rd[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          rx[u] = *reinterpret_cast<const bf16x8*>(xp + (p + u * ppb) * ld);
          rd[u] = *reinterpret_cast<const bf16x8*>(dp + (p + u * ppb) * lddy);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const f2x4 f = unpack8p(rx[u]), d = unpack8p(rd[u]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 dz = dz2<SILU>(d.p[j], __ffma2_rn(f.p[j], kz1.p[j], kz0.p[j]));
            A.p[j] = __fadd2_rn(A.p[j], dz);
            B.p[j] = __ffma2_rn(dz, f.p[j], B.p[j]);
          }
        }
      }
      for (; p < p_end; p += ppb) {
        const f2x4 f = unpack8p(*reinterpret_cast<const bf16x8*>(xp + p * ld));
        const f2x4 d = unpack8p(*reinterpret_cast<const bf16x8*>(dp + p * lddy));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 dz = dz2<SILU>(d.p[j], __ffma2_rn(f.p[j], kz1.p[j], kz0.p[j]));
          A.p[j] = __fadd2_rn(A.p[j], dz);
          B.p[j] = __ffma2_rn(dz, f.p[j], B.p[j]);
        }
      }
      const float a8[8] = {A.p[0].x, A.p[0].y, A.p[1].x, A.p[1].y, A.p[2].x, A.p[2].y, A.p[3].x, A.p[3].y};
      const float b8[8] = {B.p[0].x, B.p[0].y, B.p[1].x, B.p[1].y, B.p[2].x, B.p[2].y, B.p[3].x, B.p[3].y};
      int g_prev = -1;
      float mean = 0.f, rstd = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int g = (c + e) / cpg;
        if (g != g_prev) {
          gn_mean_rstd_of(*reinterpret_cast<const float2*>(stats_n + g * 2), inv_m, t.eps, &mean, &rstd);
          g_prev = g;
        }
        atomicAdd(&smc[(c + e) * 2], a8[e]);
        atomicAdd(&smc[(c + e) * 2 + 1], rstd * (b8[e] - mean * a8[e]));   // sum dz*xhat
      }
    }
    __syncthreads();
    float* sn = sums + static_cast<long long>(n) * C * 2;
    const bool solo = t.team_size == 1;
    const int c_end = t.slab_ch ? c_base + t.slab_ch : C;        // this CTA's channels: [c_base, c_end)
    if (solo) {
      for (int i = c_base * 2 + threadIdx.x; i < c_end * 2; i += blockDim.x) sn[i] = smc[i];   // kept for dgamma / dbeta
    } else {
      for (int i = threadIdx.x; i < C * 2; i += blockDim.x) atomicAdd(&sn[i], smc[i]);
      gn_team_barrier(&counters[n], t.team_size);
    }
    // ---- group coefficients: (sum_c gamma*A, sum_c gamma*B) / m ----
    for (int g = c_base / cpg + threadIdx.x; g < c_end / cpg; g += blockDim.x) {
      float s1 = 0.f, s2 = 0.f;
      for (int cc = g * cpg; cc < (g + 1) * cpg; ++cc) {
        const float2 ab = solo ? *reinterpret_cast<const float2*>(smc + cc * 2)
                               : __ldcg(reinterpret_cast<const float2*>(sn + cc * 2));
        const float ga = gamma[cc];
        s1 = fmaf(ga, ab.x, s1);
        s2 = fmaf(ga, ab.y, s2);
      }
      scoef[g * 2] = s1 * inv_m;
      scoef[g * 2 + 1] = s2 * inv_m;
    }
    __syncthreads();
    // ---- phase 2 ----
    if (writes) {
      f2x4 nk4, nk5;
      {
        float k4[8], k5[8];
        int g_prev = -1;
        float mean = 0.f, rstd = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int g = (c + e) / cpg;
          if (g != g_prev) {
            gn_mean_rstd_of(*reinterpret_cast<const float2*>(stats_n + g * 2), inv_m, t.eps, &mean, &rstd);
            s1 = scoef[g * 2];
            s2 = scoef[g * 2 + 1];
            g_prev = g;
          }
          k4[e] = -rstd * rstd * s2;
          k5[e] = -rstd * s1 - mean * k4[e];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          nk4.p[j] = make_float2(k4[2 * j], k4[2 * j + 1]);
          nk5.p[j] = make_float2(k5[2 * j], k5[2 * j + 1]);
        }
      }
      const __nv_bfloat16* a0p = o.add0 ? o.add0 + pix0 * o.lda0 + c : nullptr;
      const __nv_bfloat16* a1p = o.add1 ? o.add1 + pix0 * o.lda1 + c : nullptr;
      __nv_bfloat16* op;
      long long ldo;
      if (c < s.c0) {
        op = o.d0 + pix0 * o.ld0 + c;
        ldo = o.ld0;
      } else {
        op = o.d1 + pix0 * o.ld1 + (c - s.c0);
        ldo = o.ld1;
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
            const f2x4 f = unpack8p(rx[u]), d = unpack8p(rd[u]);
            f2x4 r;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 dz = dz2<SILU>(d.p[j], __ffma2_rn(f.p[j], kz1.p[j], kz0.p[j]));
              r.p[j] = __ffma2_rn(dz, kz1.p[j], __ffma2_rn(f.p[j], nk4.p[j], nk5.p[j]));
            }
            if (a0p) {
              const f2x4 a = unpack8p(ra0[u]);
#pragma unroll
              for (int j = 0; j < 4; ++j) r.p[j] = __fadd2_rn(r.p[j], a.p[j]);
            }
            if (a1p) {
              const f2x4 a = unpack8p(ra1[u]);
#pragma unroll
              for (int j = 0; j < 4; ++j) r.p[j] = __fadd2_rn(r.p[j], a.p[j]);
            }
            *reinterpret_cast<bf16x8*>(op + (p + u * ppb) * ldo) = pack8p(r);
          }
        }
      }
    }
    __syncthreads();   // smc / scoef are reused by the next sample
  }
}

// dgamma[c] += sum_n sums[n][c][1], dbeta[c] += sum_n sums[n][c][0].  One WARP per channel, lanes over samples
// (a thread-per-channel loop over N serialised 64 dependent-latency loads: 12-27 us per launch, 112 launches a step).
__global__ void __launch_bounds__(kGnThreads)
gn_bwd_dparam_kernel(const float* __restrict__ sums, int N, int C, float* __restrict__ dgamma,
                     float* __restrict__ dbeta) {
  const int c = blockIdx.x * (kGnThreads / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int n = lane; n < N; n += 32) {
    const float2 ab = *reinterpret_cast<const float2*>(sums + (static_cast<long long>(n) * C + c) * 2);
    a += ab.x;
    b += ab.y;
  }
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) {
    if (dbeta) dbeta[c] += a;
    if (dgamma) dgamma[c] += b;
  }
}

// ---- backward, second half only (first half fused into the dgrad conv epilogue, conv_epilogue.cuh) ---------------
// grid (chunks, N).  sums[n][c] = (S1 = sum dz, S2 = sum dz*x) raw moments.  dx = dz*k1 + x*nk4 + nk5 (+ addends).
__global__ void __launch_bounds__(kGnThreads, 3)
gn_bwd_apply_kernel(GnSrc s, int hw, int cpg, int groups, const float* __restrict__ stats, float eps,
                    const float* __restrict__ gamma, const __nv_bfloat16* __restrict__ dz, long long lddz,
                    const float* __restrict__ sums, GnDst o, int pix_per_block, int V,
                    float* __restrict__ out_nc /*[N][ld_nc] += sum_pix dx, or NULL*/, long long ld_nc,
                    float* __restrict__ out_c /*[C] += sum_{n,pix} dx, or NULL*/) {
  __shared__ float scoef[64 * 2];
  extern __shared__ float scol[];   // [C] per-channel sums of this CTA's dx (only when out_nc / out_c)
  const bool want_sums = (out_nc != nullptr) || (out_c != nullptr);
  if (want_sums)
    for (int i = threadIdx.x; i < V * 8; i += blockDim.x) scol[i] = 0.f;
  const float inv_m = 1.0f / (static_cast<float>(cpg) * static_cast<float>(hw));
  const int n = blockIdx.y;
  const int C = V * 8;
  const float* stats_n = stats + static_cast<long long>(n) * groups * 2;
  const float* sn = sums + static_cast<long long>(n) * C * 2;
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    float mean, rstd;
    gn_mean_rstd_of(*reinterpret_cast<const float2*>(stats_n + g * 2), inv_m, eps, &mean, &rstd);
    float s1 = 0.f, s2 = 0.f;
    for (int cc = g * cpg; cc < (g + 1) * cpg; ++cc) {
      const float2 ab = *reinterpret_cast<const float2*>(sn + cc * 2);
      const float ga = gamma[cc];
      s1 = fmaf(ga, ab.x, s1);
      s2 = fmaf(ga, rstd * (ab.y - mean * ab.x), s2);
    }
    scoef[g * 2] = s1 * inv_m;
    scoef[g * 2 + 1] = s2 * inv_m;
  }
  __syncthreads();
  const int v = threadIdx.x % V, pl = threadIdx.x / V, ppb = blockDim.x / V;
  const int c = v * 8;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(hw, p_begin + pix_per_block);
  const bool writes = (c < s.c0) || (o.d1 != nullptr);   // else: gradient of this source not requested
  if (!writes && !want_sums) return;
  long long ld;
  const __nv_bfloat16* xp = gn_base(s, n, hw, c, &ld);
  const long long pix0 = static_cast<long long>(n) * hw;
  const __nv_bfloat16* dp = dz + pix0 * lddz + c;
  f2x4 k1, nk4, nk5, csum;
#pragma unroll
  for (int j = 0; j < 4; ++j) csum.p[j] = make_float2(0.f, 0.f);
  {
    float a1[8], a4[8], a5[8];
    int g_prev = -1;
    float mean = 0.f, rstd = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int g = (c + e) / cpg;
      if (g != g_prev) {
        gn_mean_rstd_of(*reinterpret_cast<const float2*>(stats_n + g * 2), inv_m, eps, &mean, &rstd);
        s1 = scoef[g * 2];
        s2 = scoef[g * 2 + 1];
        g_prev = g;
      }
      a1[e] = rstd * gamma[c + e];
      a4[e] = -rstd * rstd * s2;
      a5[e] = -rstd * s1 - mean * a4[e];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      k1.p[j] = make_float2(a1[2 * j], a1[2 * j + 1]);
      nk4.p[j] = make_float2(a4[2 * j], a4[2 * j + 1]);
      nk5.p[j] = make_float2(a5[2 * j], a5[2 * j + 1]);
    }
  }
  const __nv_bfloat16* a0p = o.add0 ? o.add0 + pix0 * o.lda0 + c : nullptr;
  const __nv_bfloat16* a1p = o.add1 ? o.add1 + pix0 * o.lda1 + c : nullptr;
  __nv_bfloat16* op = nullptr;
  long long ldo = 0;
  if (c < s.c0) {
    op = o.d0 + pix0 * o.ld0 + c;
    ldo = o.ld0;
  } else if (o.d1 != nullptr) {
    op = o.d1 + pix0 * o.ld1 + (c - s.c0);
    ldo = o.ld1;
  }
  for (int p = p_begin + pl; p < p_end; p += 2 * ppb) {
    const bool two = p + ppb < p_end;
    bf16x8 rx[2], rd[2], ra0[2], ra1[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 0 || two) {
        const int q = p + u * ppb;
        rx[u] = *reinterpret_cast<const bf16x8*>(xp + q * ld);
        rd[u] = *reinterpret_cast<const bf16x8*>(dp + q * lddz);
        if (a0p) ra0[u] = *reinterpret_cast<const bf16x8*>(a0p + q * o.lda0);
        if (a1p) ra1[u] = *reinterpret_cast<const bf16x8*>(a1p + q * o.lda1);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 0 || two) {
        const f2x4 f = unpack8p(rx[u]), d = unpack8p(rd[u]);
        f2x4 r;
#pragma unroll
        for (int j = 0; j < 4; ++j) r.p[j] = __ffma2_rn(d.p[j], k1.p[j], __ffma2_rn(f.p[j], nk4.p[j], nk5.p[j]));
        if (a0p) {
          const f2x4 a = unpack8p(ra0[u]);
#pragma unroll
          for (int j = 0; j < 4; ++j) r.p[j] = __fadd2_rn(r.p[j], a.p[j]);
        }
        if (a1p) {
          const f2x4 a = unpack8p(ra1[u]);
#pragma unroll
          for (int j = 0; j < 4; ++j) r.p[j] = __fadd2_rn(r.p[j], a.p[j]);
        }
        if (op != nullptr) *reinterpret_cast<bf16x8*>(op + (p + u * ppb) * ldo) = pack8p(r);
        if (want_sums) {
#pragma unroll
          for (int j = 0; j < 4; ++j) csum.p[j] = __fadd2_rn(csum.p[j], r.p[j]);
        }
      }
    }
  }
  if (want_sums) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(&scol[c + 2 * j], csum.p[j].x);
      atomicAdd(&scol[c + 2 * j + 1], csum.p[j].y);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < V * 8; i += blockDim.x) {
      const float v = scol[i];
      if (out_nc) atomicAdd(&out_nc[static_cast<long long>(n) * ld_nc + i], v);
      if (out_c) atomicAdd(&out_c[i], v);
    }
  }
}

// dgamma[c] += sum_n rstd*(S2 - mean*S1), dbeta[c] += sum_n S1   (raw moments from the conv epilogue); warp per channel
__global__ void __launch_bounds__(kGnThreads)
gn_bwd_dparam_raw_kernel(const float* __restrict__ sums, const float* __restrict__ stats, int N, int C, int cpg,
                         int groups, int hw, float eps, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * (kGnThreads / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  const float inv_m = 1.0f / (static_cast<float>(cpg) * static_cast<float>(hw));
  const int g = c / cpg;
  float a = 0.f, b = 0.f;
  for (int n = lane; n < N; n += 32) {
    float mean, rstd;
    gn_mean_rstd_of(*reinterpret_cast<const float2*>(stats + (static_cast<long long>(n) * groups + g) * 2), inv_m, eps,
                    &mean, &rstd);
    const float2 ab = *reinterpret_cast<const float2*>(sums + (static_cast<long long>(n) * C + c) * 2);
    a += ab.x;
    b += rstd * (ab.y - mean * ab.x);
  }
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) {
    if (dbeta) dbeta[c] += a;
    if (dgamma) dgamma[c] += b;
  }
}

// =====================================================================================================
// fp32-faithful mode (include/ddpm_b200.h): the same statistics / apply passes on SPLIT tensors -- a row holds C "hi"
// channels then C "lo" channels, value = hi + lo -- with exact sigmoid arithmetic.  Thread mapping as above; each
// thread moves its 8-channel vector twice (hi and lo).  Inference only: no backward, no conv fusions.
// =====================================================================================================
__device__ __forceinline__ void gn_load_split(const __nv_bfloat16* hp, int c_src, float* f) {
  float lo[8];
  unpack8(*reinterpret_cast<const bf16x8*>(hp), f);
  unpack8(*reinterpret_cast<const bf16x8*>(hp + c_src), lo);
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] += lo[j];
}

__global__ void __launch_bounds__(kGnThreads)
gn_stats_split_kernel(GnSrc s, int hw, int cpg, int groups, float* __restrict__ stats, int pix_per_block, int V) {
  __shared__ float sm[64 * 2];
  for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int n = blockIdx.y;
  const int v = threadIdx.x % V, pl = threadIdx.x / V, ppb = blockDim.x / V;
  const int c = v * 8;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(hw, p_begin + pix_per_block);
  long long ld;
  const __nv_bfloat16* xp = gn_base(s, n, hw, c, &ld);
  const int c_src = c < s.c0 ? s.c0 : s.c1;
  float sum[8], sq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sum[j] = sq[j] = 0.f;
  for (int p = p_begin + pl; p < p_end; p += ppb) {
    float f[8];
    gn_load_split(xp + p * ld, c_src, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sum[j] += f[j];
      sq[j] = fmaf(f[j], f[j], sq[j]);
    }
  }
  f2x4 s4, q4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s4.p[j] = make_float2(sum[2 * j], sum[2 * j + 1]);
    q4.p[j] = make_float2(sq[2 * j], sq[2 * j + 1]);
  }
  gn_fold_groups(s4, q4, c, cpg, sm);
  __syncthreads();
  for (int i = threadIdx.x; i < groups * 2; i += blockDim.x)
    atomicAdd(&stats[static_cast<long long>(n) * groups * 2 + i], sm[i]);
}

template <bool SILU>
__global__ void __launch_bounds__(kGnThreads)
gn_apply_split_kernel(GnSrc s, int hw, int cpg, int groups, const float* __restrict__ stats, float eps,
                      const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                      long long ldy, int pix_per_block, int V) {
  const float inv_m = 1.0f / (static_cast<float>(cpg) * static_cast<float>(hw));
  const int n = blockIdx.y;
  const int v = threadIdx.x % V, pl = threadIdx.x / V, ppb = blockDim.x / V;
  const int c = v * 8, C = V * 8;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(hw, p_begin + pix_per_block);
  long long ld;
  const __nv_bfloat16* xp = gn_base(s, n, hw, c, &ld);
  const int c_src = c < s.c0 ? s.c0 : s.c1;
  __nv_bfloat16* yp = y + static_cast<long long>(n) * hw * ldy + c;
  f2x4 ka, kb;
  gn_apply_coefs(stats + static_cast<long long>(n) * groups * 2, c, cpg, inv_m, eps, gamma, beta, false, &ka, &kb);
  const float a8[8] = {ka.p[0].x, ka.p[0].y, ka.p[1].x, ka.p[1].y, ka.p[2].x, ka.p[2].y, ka.p[3].x, ka.p[3].y};
  const float b8[8] = {kb.p[0].x, kb.p[0].y, kb.p[1].x, kb.p[1].y, kb.p[2].x, kb.p[2].y, kb.p[3].x, kb.p[3].y};
  for (int p = p_begin + pl; p < p_end; p += ppb) {
    float f[8], hi[8];
    gn_load_split(xp + p * ld, c_src, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float z = fmaf(f[j], a8[j], b8[j]);
      f[j] = SILU ? z / (1.0f + expf(-z)) : z;
    }
    const bf16x8 h = pack8(f);
    unpack8(h, hi);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] -= hi[j];
    *reinterpret_cast<bf16x8*>(yp + p * ldy) = h;
    *reinterpret_cast<bf16x8*>(yp + p * ldy + C) = pack8(f);
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

// Pixel chunks per sample for the standalone kernels.  CTAs are kept SMALL (min_iters loop trips per thread) so the
// grid is many waves deep (<= kMaxCtas): with a handful of large CTAs per SM the last, partially filled wave cost
// up to 30 % of the kernel (1216 CTAs over 888 resident slots = 1.37 waves).
constexpr long long kMaxCtas = 148LL * 64;
static void gn_geometry(int C, int hw, int n, int min_iters, int* V, int* threads, int* pix_per_block, int* chunks) {
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

// Team geometry for the fused kernels.  `resident` = co-resident CTAs of the kernel (occupancy * SMs);
// `sample_bytes` = bytes of one sample that phase 2 re-reads.  Teams are sized so that teams * sample_bytes stays
// within the L2 budget and every CTA still has >= 2 loop trips.
// Channel slab of a solo launch: whole GroupNorm groups and whole 8-channel vectors (a multiple of lcm(cpg, 8)), rows of
// at least 128 bytes, and no more slabs than it takes to put ~3 CTAs on every SM.  Returns C when one slab is enough.
static int gn_slab_channels(int C, int cpg, int n) {
  if (env_int("DDPM_GN_SLABS", 1) == 0) return C;
  int u = cpg;
  while (u % 8) u += cpg;
  if (C % u) return C;
  const int units = C / u;
  int want = (3 * kNumSMs + n - 1) / n;
  if (want < 1) want = 1;
  int best = C;
  for (int k = 1; k <= units; ++k) {
    if (units % k) continue;
    const int ch = k * u;
    if (ch < 64 && k < units) continue;
    best = ch;
    if (units / k <= want) break;
  }
  return best;
}

static void gn_team_geometry(int C, int hw, int n, int resident, double sample_bytes, GnTeam* t, int* threads,
                             int* grid, int* grid_y) {
  const int V = C / 8;
  *grid_y = 1;
  t->slab_ch = 0;
  int ppb = kGnThreads / V;
  if (ppb < 1) ppb = 1;
  *threads = V * ppb;
  const double budget = static_cast<double>(env_int("DDPM_GN_L2_MB", 64)) * 1048576.0;
  int max_teams = static_cast<int>(budget / sample_bytes);
  if (max_teams < 1) max_teams = 1;
  // a tensor that fits L2 as a whole gets one team per sample: a single barrier per CTA instead of one per round
  if (sample_bytes * n <= static_cast<double>(env_int("DDPM_GN_L2_WHOLE_MB", 100)) * 1048576.0) max_teams = n;
  int teams = n < max_teams ? n : max_teams;
  if (teams > resident) teams = resident;
  int team_size = resident / teams;
  const int max_ts = (hw + 2 * ppb - 1) / (2 * ppb);   // >= 2 pixels per thread row
  if (team_size > max_ts) team_size = max_ts;
  if (team_size < 1) team_size = 1;
  teams = resident / team_size;
  if (teams > n) teams = n;
  const int slab = gn_slab_channels(C, t->cpg, n);
  const bool solo_whole = static_cast<long long>(hw) * C <= env_int("DDPM_GN_SOLO_ELEMS", 65536);
  const bool solo_slabs = slab < C && static_cast<long long>(hw) * slab <= env_int("DDPM_GN_SOLO_SLAB_ELEMS", 32768);
  if (solo_whole || solo_slabs) {
    team_size = 1;                       // latency-sized sample: one CTA, no atomics / barrier / memsets
    teams = resident < n ? resident : n;
    if (slab < C) {                      // ... per channel slab
      const int slabs = C / slab;
      t->slab_ch = slab;
      *grid_y = slabs;
      teams = resident / slabs < n ? resident / slabs : n;
      if (teams < 1) teams = 1;
      const int vt = slab / 8;
      int ppb_s = kGnThreads / vt;
      if (ppb_s < 1) ppb_s = 1;
      *threads = vt * ppb_s;
    }
  }
  int ppc = (hw + team_size - 1) / team_size;
  ppc = (ppc + ppb - 1) / ppb * ppb;
  t->V = V;
  t->team_size = team_size;
  t->pix_per_cta = ppc;
  *grid = teams * team_size;
}

template <typename K>
static int gn_resident_ctas(K kern, int threads, size_t smem) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem) != cudaSuccess || occ < 1) occ = 1;
  int dev = 0, sms = kNumSMs;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return occ * sms;
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
  gn_geometry(C, hw, n, 4 * kUnroll, &V, &threads, &ppblk, &chunks);
  DDPM_CUDA(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * groups * n, stream));
  gn_stats_kernel<<<dim3(chunks, n), threads, 0, stream>>>(s, hw, C / groups, groups, stats, ppblk, V);
  return check_launch("gn_stats_kernel");
}

extern "C" int ddpm_gn_apply(const void* x0, int c0, long long ld0, const void* x1, int c1, long long ld1, int n,
                             int hw, int groups, const float* stats, float eps, const float* gamma,
                             const float* beta, int silu, void* y, long long ldy, float* coef, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = gn_check(x0, c0, ld0, x1, c1, ld1, n, hw, groups, "ddpm_gn_apply")) return e;
  DDPM_REQUIRE(stats && gamma && beta && y && ldy % 8 == 0, "ddpm_gn_apply: bad argument");
  const int C = c0 + c1;
  GnSrc s{static_cast<const __nv_bfloat16*>(x0), static_cast<const __nv_bfloat16*>(x1), ld0, ld1, c0, c1};
  int V, threads, ppblk, chunks;
  gn_geometry(C, hw, n, 8 * kUnroll, &V, &threads, &ppblk, &chunks);
  __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(y);
  if (silu)
    gn_apply_kernel<true><<<dim3(chunks, n), threads, 0, stream>>>(s, hw, C / groups, groups, stats, eps, gamma, beta,
                                                                   yp, ldy, ppblk, V, coef);
  else
    gn_apply_kernel<false><<<dim3(chunks, n), threads, 0, stream>>>(s, hw, C / groups, groups, stats, eps, gamma, beta,
                                                                    yp, ldy, ppblk, V, coef);
  return check_launch("gn_apply_kernel");
}

extern "C" int ddpm_gn_stats_from_csum(const float* csum0, int c0, const float* csum1, int c1, int n, int groups,
                                       float* stats, void* stream) {
  DDPM_REQUIRE(csum0 && stats && n > 0 && c0 > 0 && c1 >= 0 && (c1 == 0 || csum1) && groups > 0 &&
                   (c0 + c1) % groups == 0 && c0 % 4 == 0 && c1 % 4 == 0 && ((c0 + c1) / groups) % 4 == 0,
               "ddpm_gn_stats_from_csum: bad argument (c0=%d c1=%d groups=%d; channels per group must be a multiple "
               "of 4)", c0, c1, groups);
  gn_stats_from_csum_kernel<<<(n * groups + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(csum0, c0, csum1, c1,
                                                                                                     n, groups, stats);
  return check_launch("gn_stats_from_csum_kernel");
}

extern "C" int ddpm_gn_fwd(const void* x0, int c0, long long ld0, const void* x1, int c1, long long ld1, int n,
                           int hw, int groups, float eps, const float* gamma, const float* beta, int silu,
                           float* stats, void* y, long long ldy, float* coef, int* ws, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = gn_check(x0, c0, ld0, x1, c1, ld1, n, hw, groups, "ddpm_gn_fwd")) return e;
  DDPM_REQUIRE(stats && gamma && beta && y && ws && ldy % 8 == 0, "ddpm_gn_fwd: bad argument");
  const int C = c0 + c1;
  GnSrc s{static_cast<const __nv_bfloat16*>(x0), static_cast<const __nv_bfloat16*>(x1), ld0, ld1, c0, c1};
  GnTeam t;
  t.N = n; t.hw = hw; t.cpg = C / groups; t.groups = groups; t.eps = eps;
  static int resident[2] = {0, 0};
  const int ki = silu ? 1 : 0;
  if (!resident[ki])
    resident[ki] = silu ? gn_resident_ctas(gn_fwd_fused_kernel<true>, kGnThreads, 0)
                        : gn_resident_ctas(gn_fwd_fused_kernel<false>, kGnThreads, 0);
  int threads, grid, grid_y;
  gn_team_geometry(C, hw, n, resident[ki], 2.0 * hw * C, &t, &threads, &grid, &grid_y);
  if (t.team_size > 1) {
    DDPM_CUDA(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * groups * n, stream));
    DDPM_CUDA(cudaMemsetAsync(ws, 0, sizeof(int) * n, stream));
  }
  __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(y);
  void* args[] = {&s, &t, &stats, &ws, &gamma, &beta, &yp, &ldy, &coef};
  const void* fn = silu ? reinterpret_cast<const void*>(gn_fwd_fused_kernel<true>)
                        : reinterpret_cast<const void*>(gn_fwd_fused_kernel<false>);
  // co-residency (cooperative launch) is only needed when CTAs of a team wait for each other; a solo-CTA launch is an
  // ordinary kernel node, which a captured step can overlap with its neighbours
  if (t.team_size > 1)
    DDPM_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(threads), args, 0, stream));
  else
    DDPM_CUDA(cudaLaunchKernel(fn, dim3(grid, grid_y), dim3(threads), args, 0, stream));
  return check_launch("gn_fwd_fused_kernel");
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
  GnDst o{static_cast<__nv_bfloat16*>(dx0), static_cast<__nv_bfloat16*>(dx1), lddx0, lddx1,
          static_cast<const __nv_bfloat16*>(add0), static_cast<const __nv_bfloat16*>(add1), ldadd0, ldadd1};
  float* sums = ws;
  int* counters = reinterpret_cast<int*>(ws + static_cast<long long>(n) * C * 2);
  GnTeam t;
  t.N = n; t.hw = hw; t.cpg = C / groups; t.groups = groups; t.eps = eps;
  const size_t smem = sizeof(float) * (C * 2 + groups * 2);
  static int resident[2] = {0, 0};
  const int ki = silu ? 1 : 0;
  if (!resident[ki])
    resident[ki] = silu ? gn_resident_ctas(gn_bwd_fused_kernel<true>, kGnThreads, sizeof(float) * (kMaxC * 2 + 128))
                        : gn_resident_ctas(gn_bwd_fused_kernel<false>, kGnThreads, sizeof(float) * (kMaxC * 2 + 128));
  int threads, grid, grid_y;
  const double sample_bytes = 2.0 * hw * C * (2 + (add0 ? 1 : 0) + (add1 ? 1 : 0));
  gn_team_geometry(C, hw, n, resident[ki], sample_bytes, &t, &threads, &grid, &grid_y);
  if (t.team_size > 1)
    DDPM_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * static_cast<size_t>(n) * C * 2 + sizeof(int) * n, stream));
  void* args[] = {&s, &t, &stats, &gamma, &beta, &dyp, &lddy, &sums, &counters, &o};
  const void* fn = silu ? reinterpret_cast<const void*>(gn_bwd_fused_kernel<true>)
                        : reinterpret_cast<const void*>(gn_bwd_fused_kernel<false>);
  if (t.team_size > 1)
    DDPM_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(threads), args, smem, stream));
  else
    DDPM_CUDA(cudaLaunchKernel(fn, dim3(grid, grid_y), dim3(threads), args, smem, stream));
  if (int e = check_launch("gn_bwd_fused_kernel")) return e;
  if (dgamma || dbeta) {
    gn_bwd_dparam_kernel<<<(C + 7) / 8, kGnThreads, 0, stream>>>(sums, n, C, dgamma, dbeta);
    return check_launch("gn_bwd_dparam_kernel");
  }
  return DDPM_OK;
}

// The dgamma / dbeta reductions on their own, so a caller can issue them off the critical path (another stream): call
// ddpm_gn_bwd / ddpm_gn_bwd_apply with dgamma = dbeta = NULL, then this.  stats == NULL: `sums` are the centred sums
// ddpm_gn_bwd leaves in its workspace; otherwise the raw moments of the conv-epilogue fusion.
extern "C" int ddpm_gn_bwd_dparams(const float* sums, const float* stats, int n, int c, int groups, int hw, float eps,
                                   float* dgamma, float* dbeta, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DDPM_REQUIRE(sums && n > 0 && c > 0 && groups > 0 && c % groups == 0 && hw > 0, "ddpm_gn_bwd_dparams: bad argument");
  if (!dgamma && !dbeta) return DDPM_OK;
  if (stats == nullptr) {
    gn_bwd_dparam_kernel<<<(c + 7) / 8, kGnThreads, 0, stream>>>(sums, n, c, dgamma, dbeta);
    return check_launch("gn_bwd_dparam_kernel");
  }
  gn_bwd_dparam_raw_kernel<<<(c + 7) / 8, kGnThreads, 0, stream>>>(sums, stats, n, c, c / groups, groups, hw, eps,
                                                                   dgamma, dbeta);
  return check_launch("gn_bwd_dparam_raw_kernel");
}

extern "C" int ddpm_gn_bwd_apply(const void* x0, int c0, long long ld0, const void* x1, int c1, long long ld1, int n,
                                 int hw, int groups, const float* stats, float eps, const float* gamma, const void* dz,
                                 long long lddz, const float* sums, const void* add0, long long ldadd0,
                                 const void* add1, long long ldadd1, void* dx0, long long lddx0, void* dx1,
                                 long long lddx1, float* dgamma, float* dbeta, float* out_nc, long long ld_nc,
                                 float* out_c, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = gn_check(x0, c0, ld0, x1, c1, ld1, n, hw, groups, "ddpm_gn_bwd_apply")) return e;
  DDPM_REQUIRE(stats && gamma && dz && sums && dx0, "ddpm_gn_bwd_apply: null pointer argument");
  DDPM_REQUIRE(lddz % 8 == 0 && lddx0 % 8 == 0 && (c1 == 0 || !dx1 || lddx1 % 8 == 0) &&
                   (!add0 || ldadd0 % 8 == 0) && (!add1 || ldadd1 % 8 == 0),
               "ddpm_gn_bwd_apply: strides must be multiples of 8");
  const int C = c0 + c1;
  GnSrc s{static_cast<const __nv_bfloat16*>(x0), static_cast<const __nv_bfloat16*>(x1), ld0, ld1, c0, c1};
  GnDst o{static_cast<__nv_bfloat16*>(dx0), static_cast<__nv_bfloat16*>(dx1), lddx0, lddx1,
          static_cast<const __nv_bfloat16*>(add0), static_cast<const __nv_bfloat16*>(add1), ldadd0, ldadd1};
  int V, threads, ppblk, chunks;
  const bool want_sums = out_nc || out_c;
  // 32 loop trips per thread: measured sweep (4..64) -- below that the per-CTA prologue (group coefficients, 24
  // coefficient registers) dominates, 0.198 ms -> 0.165 ms at 128^2 x 128 channels
  gn_geometry(C, hw, n, env_int("DDPM_GN_BWD_APPLY_ITERS", 32), &V, &threads, &ppblk, &chunks);
  gn_bwd_apply_kernel<<<dim3(chunks, n), threads, want_sums ? C * sizeof(float) : 0, stream>>>(
      s, hw, C / groups, groups, stats, eps, gamma, static_cast<const __nv_bfloat16*>(dz), lddz, sums, o, ppblk, V,
      out_nc, ld_nc, out_c);
  if (int e = check_launch("gn_bwd_apply_kernel")) return e;
  if (dgamma || dbeta) {
    gn_bwd_dparam_raw_kernel<<<(C + 7) / 8, kGnThreads, 0, stream>>>(sums, stats, n, C, C / groups, groups, hw, eps,
                                                                     dgamma, dbeta);
    return check_launch("gn_bwd_dparam_raw_kernel");
  }
  return DDPM_OK;
}

extern "C" int ddpm_gn_stats_split(const void* x0, int c0, long long ld0, const void* x1, int c1, long long ld1, int n,
                                   int hw, int groups, float* stats, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = gn_check(x0, c0, ld0, x1, c1, ld1, n, hw, groups, "ddpm_gn_stats_split")) return e;
  DDPM_REQUIRE(stats && ld0 >= 2LL * c0 && (c1 == 0 || ld1 >= 2LL * c1), "ddpm_gn_stats_split: rows must hold [hi | lo]");
  const int C = c0 + c1;
  GnSrc s{static_cast<const __nv_bfloat16*>(x0), static_cast<const __nv_bfloat16*>(x1), ld0, ld1, c0, c1};
  int V, threads, ppblk, chunks;
  gn_geometry(C, hw, n, 16, &V, &threads, &ppblk, &chunks);
  DDPM_CUDA(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * groups * n, stream));
  gn_stats_split_kernel<<<dim3(chunks, n), threads, 0, stream>>>(s, hw, C / groups, groups, stats, ppblk, V);
  return check_launch("gn_stats_split_kernel");
}

extern "C" int ddpm_gn_apply_split(const void* x0, int c0, long long ld0, const void* x1, int c1, long long ld1, int n,
                                   int hw, int groups, const float* stats, float eps, const float* gamma,
                                   const float* beta, int silu, void* y, long long ldy, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int e = gn_check(x0, c0, ld0, x1, c1, ld1, n, hw, groups, "ddpm_gn_apply_split")) return e;
  const int C = c0 + c1;
  DDPM_REQUIRE(stats && gamma && beta && y && ldy % 8 == 0 && ldy >= 2LL * C && ld0 >= 2LL * c0 &&
                   (c1 == 0 || ld1 >= 2LL * c1),
               "ddpm_gn_apply_split: bad argument (rows must hold [hi | lo])");
  GnSrc s{static_cast<const __nv_bfloat16*>(x0), static_cast<const __nv_bfloat16*>(x1), ld0, ld1, c0, c1};
  int V, threads, ppblk, chunks;
  gn_geometry(C, hw, n, 16, &V, &threads, &ppblk, &chunks);
  __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(y);
  if (silu)
    gn_apply_split_kernel<true><<<dim3(chunks, n), threads, 0, stream>>>(s, hw, C / groups, groups, stats, eps, gamma,
                                                                         beta, yp, ldy, ppblk, V);
  else
    gn_apply_split_kernel<false><<<dim3(chunks, n), threads, 0, stream>>>(s, hw, C / groups, groups, stats, eps, gamma,
                                                                          beta, yp, ldy, ppblk, V);
  return check_launch("gn_apply_split_kernel");
}
