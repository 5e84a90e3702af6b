// Single-pass fp32 elementwise kernels of the DDPM path: add_noise, MSE fwd+bwd, scheduler step, uint8
// post-processing.  HBM-bound; 16-byte vector accesses, grid sized to the problem (streaming, no reuse).
// Replaces ~6 / ~3 / ~30 TensorIterator launches of the reference.
#include "common.cuh"

#include <cstdarg>
#include <cstdio>

#include "../../include/ddpm_b200.h"

namespace ddpm {

static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s launch failed: %s", what, cudaGetErrorString(e));
    return DDPM_ERR_CUDA;
  }
  return DDPM_OK;
}

constexpr int kEwThreads = 256;

static inline int ew_blocks(long long work_items) {
  long long b = (work_items + kEwThreads - 1) / kEwThreads;
  const long long cap = static_cast<long long>(kNumSMs) * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

// ---- add_noise ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEwThreads)
add_noise_kernel(const float* __restrict__ x0, const float* __restrict__ noise, const long long* __restrict__ t,
                 const float* __restrict__ sa_tab, const float* __restrict__ sb_tab, float* __restrict__ out,
                 long long per_sample, long long total_vec, int vec_per_sample_ok, int T) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  if (vec_per_sample_ok) {
    const long long vps = per_sample >> 2;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total_vec; i += stride) {
      const long long b = i / vps;
      long long tt = t[b];
      tt = tt < 0 ? 0 : (tt >= T ? T - 1 : tt);
      const float a = sa_tab[tt], s = sb_tab[tt];
      const float4 x = reinterpret_cast<const float4*>(x0)[i];
      const float4 z = reinterpret_cast<const float4*>(noise)[i];
      float4 o;
      o.x = __fadd_rn(__fmul_rn(a, x.x), __fmul_rn(s, z.x));
      o.y = __fadd_rn(__fmul_rn(a, x.y), __fmul_rn(s, z.y));
      o.z = __fadd_rn(__fmul_rn(a, x.z), __fmul_rn(s, z.z));
      o.w = __fadd_rn(__fmul_rn(a, x.w), __fmul_rn(s, z.w));
      reinterpret_cast<float4*>(out)[i] = o;
    }
  } else {
    const long long total = total_vec;  // scalar count in this mode
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
      const long long b = i / per_sample;
      long long tt = t[b];
      tt = tt < 0 ? 0 : (tt >= T ? T - 1 : tt);
      out[i] = __fadd_rn(__fmul_rn(sa_tab[tt], x0[i]), __fmul_rn(sb_tab[tt], noise[i]));
    }
  }
}

// ---- MSE forward + backward ----------------------------------------------------------------------
__global__ void __launch_bounds__(kEwThreads)
mse_fwd_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target, float* __restrict__ loss_sum,
                   float* __restrict__ dpred, long long n, float two_over_n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long nvec = n >> 2;
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const float4 p = reinterpret_cast<const float4*>(pred)[i];
    const float4 q = reinterpret_cast<const float4*>(target)[i];
    float4 d = make_float4(p.x - q.x, p.y - q.y, p.z - q.z, p.w - q.w);
    acc += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
    if (dpred) reinterpret_cast<float4*>(dpred)[i] = make_float4(d.x * two_over_n, d.y * two_over_n, d.z * two_over_n, d.w * two_over_n);
  }
  for (long long i = (nvec << 2) + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = pred[i] - target[i];
    acc += d * d;
    if (dpred) dpred[i] = d * two_over_n;
  }
  __shared__ float red[kEwThreads / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < kEwThreads / 32 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(loss_sum, v);
  }
}

__global__ void __launch_bounds__(kEwThreads)
scale_by_scalar_kernel(float* __restrict__ x, const float* __restrict__ scale, long long n) {
  const float s = *scale;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long nvec = n >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float4 v = reinterpret_cast<float4*>(x)[i];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    reinterpret_cast<float4*>(x)[i] = v;
  }
  for (long long i = (nvec << 2) + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    x[i] *= s;
}

// ---- scheduler step --------------------------------------------------------------------------------
struct StepCoef {
  float sa, sb, c0, ct, sigma, clip;
};

__device__ __forceinline__ float step_one(float e, float x, float z, bool has_z, const StepCoef& k, float* x0_out) {
  float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(k.sb, e)), k.sa);
  if (k.clip > 0.f) x0 = fminf(fmaxf(x0, -k.clip), k.clip);
  *x0_out = x0;
  float mu = __fadd_rn(__fmul_rn(k.c0, x0), __fmul_rn(k.ct, x));
  if (has_z) mu = __fadd_rn(mu, __fmul_rn(k.sigma, z));
  return mu;
}

__global__ void __launch_bounds__(kEwThreads)
scheduler_step_kernel(const float* __restrict__ eps, const float* __restrict__ x, const float* __restrict__ z,
                      float* __restrict__ prev, float* __restrict__ pred_x0, long long n, StepCoef k) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long nvec = n >> 2;
  const bool has_z = z != nullptr;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const float4 e = reinterpret_cast<const float4*>(eps)[i];
    const float4 xv = reinterpret_cast<const float4*>(x)[i];
    float4 zv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_z) zv = reinterpret_cast<const float4*>(z)[i];
    float4 o, p0;
    o.x = step_one(e.x, xv.x, zv.x, has_z, k, &p0.x);
    o.y = step_one(e.y, xv.y, zv.y, has_z, k, &p0.y);
    o.z = step_one(e.z, xv.z, zv.z, has_z, k, &p0.z);
    o.w = step_one(e.w, xv.w, zv.w, has_z, k, &p0.w);
    reinterpret_cast<float4*>(prev)[i] = o;
    if (pred_x0) reinterpret_cast<float4*>(pred_x0)[i] = p0;
  }
  for (long long i = (nvec << 2) + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float p0;
    prev[i] = step_one(eps[i], x[i], has_z ? z[i] : 0.f, has_z, k, &p0);
    if (pred_x0) pred_x0[i] = p0;
  }
}

// ---- DDIM step (strided sampler on the same UNet) --------------------------------------
//   x0 = clamp((x - sb*eps)/sa);  eps' = use_clipped ? (x - sa*x0)/sb : eps;  prev = sap*x0 + dir*eps' (+ sigma*z)
// same op order (and no FMA contraction) as the torch expressions of diffusers' DDIMScheduler.step, so bit-identical
struct DdimCoef {
  float sa, sb, sap, dir, sigma, clip;
  int use_clipped;
};
__device__ __forceinline__ float ddim_one(float e, float x, float z, bool has_z, const DdimCoef& k, float* x0_out) {
  float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(k.sb, e)), k.sa);
  if (k.clip > 0.f) x0 = fminf(fmaxf(x0, -k.clip), k.clip);
  *x0_out = x0;
  float pe = e;
  if (k.use_clipped) pe = __fdiv_rn(__fsub_rn(x, __fmul_rn(k.sa, x0)), k.sb);
  float prev = __fadd_rn(__fmul_rn(k.sap, x0), __fmul_rn(k.dir, pe));
  if (has_z) prev = __fadd_rn(prev, __fmul_rn(k.sigma, z));
  return prev;
}
__global__ void __launch_bounds__(kEwThreads)
ddim_step_kernel(const float* __restrict__ eps, const float* __restrict__ x, const float* __restrict__ z,
                 float* __restrict__ prev, float* __restrict__ pred_x0, long long n, DdimCoef k) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long nvec = n >> 2;
  const bool has_z = z != nullptr;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const float4 e = reinterpret_cast<const float4*>(eps)[i];
    const float4 xv = reinterpret_cast<const float4*>(x)[i];
    float4 zv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_z) zv = reinterpret_cast<const float4*>(z)[i];
    float4 o, p0;
    o.x = ddim_one(e.x, xv.x, zv.x, has_z, k, &p0.x);
    o.y = ddim_one(e.y, xv.y, zv.y, has_z, k, &p0.y);
    o.z = ddim_one(e.z, xv.z, zv.z, has_z, k, &p0.z);
    o.w = ddim_one(e.w, xv.w, zv.w, has_z, k, &p0.w);
    reinterpret_cast<float4*>(prev)[i] = o;
    if (pred_x0) reinterpret_cast<float4*>(pred_x0)[i] = p0;
  }
  for (long long i = (nvec << 2) + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float p0;
    prev[i] = ddim_one(eps[i], x[i], has_z ? z[i] : 0.f, has_z, k, &p0);
    if (pred_x0) pred_x0[i] = p0;
  }
}

// Philox4x32-10, counter = (element-quad index, offset), key = seed
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = (static_cast<float>(a) + 1.0f) * 2.3283064365386963e-10f;  // (0,1]
  const float u2 = static_cast<float>(b) * 2.3283064365386963e-10f;
  const float r = sqrtf(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  return make_float2(r * c, r * s);
}

__global__ void __launch_bounds__(kEwThreads)
scheduler_step_philox_kernel(const float* __restrict__ eps, const float* __restrict__ x, float* __restrict__ prev,
                             long long n, StepCoef k, unsigned long long seed, unsigned long long offset) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long nvec = (n + 3) >> 2;
  const bool has_z = k.sigma != 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float zz[4] = {0.f, 0.f, 0.f, 0.f};
    if (has_z) {
      const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(i), static_cast<uint32_t>(i >> 32),
                                               static_cast<uint32_t>(offset), static_cast<uint32_t>(offset >> 32)),
                                    make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
      const float2 g0 = box_muller(r.x, r.y), g1 = box_muller(r.z, r.w);
      zz[0] = g0.x; zz[1] = g0.y; zz[2] = g1.x; zz[3] = g1.y;
    }
    const long long base = i << 2;
    if (base + 3 < n) {
      const float4 e = reinterpret_cast<const float4*>(eps)[i];
      const float4 xv = reinterpret_cast<const float4*>(x)[i];
      float4 o;
      float p0;
      o.x = step_one(e.x, xv.x, zz[0], has_z, k, &p0);
      o.y = step_one(e.y, xv.y, zz[1], has_z, k, &p0);
      o.z = step_one(e.z, xv.z, zz[2], has_z, k, &p0);
      o.w = step_one(e.w, xv.w, zz[3], has_z, k, &p0);
      reinterpret_cast<float4*>(prev)[i] = o;
    } else {
      for (int j = 0; j < 4 && base + j < n; ++j) {
        float p0;
        prev[base + j] = step_one(eps[base + j], x[base + j], zz[j], has_z, k, &p0);
      }
    }
  }
}

// ---- uint8 post-processing -----------------------------------------------------------------------
__global__ void __launch_bounds__(kEwThreads)
to_uint8_nhwc_kernel(const float* __restrict__ x, unsigned char* __restrict__ out, int n, int c, int h, int w) {
  const long long total = static_cast<long long>(n) * h * w * c;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long hw = static_cast<long long>(h) * w;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int ch = static_cast<int>(i % c);
    const long long pix = i / c;
    const long long b = pix / hw, p = pix - b * hw;
    float v = x[(b * c + ch) * hw + p];
    v = __fadd_rn(__fdiv_rn(v, 2.0f), 0.5f);
    v = fminf(fmaxf(v, 0.f), 1.f);
    out[i] = static_cast<unsigned char>(rintf(__fmul_rn(v, 255.0f)));
  }
}

// ---- UniPC multistep sampler (train_with_lora_all_classes.py:314) -------------------------------
// Two single-pass kernels with the scalars (sigma / alpha / lambda arithmetic, B(h), rho) computed on the host as 0-dim
// fp32 tensors exactly as diffusers does; every elementwise expression below keeps torch's op order and rounding
// (explicit _rn intrinsics: no FMA contraction), so the sampler is bit-identical to the oracle's torch expressions.
//   x0  = (x - sigma_t * eps) / alpha_t                                  (convert_model_output, epsilon / predict_x0)
//   out = (cx * x - cm * m0) - cb * res,   res = rho0 * ((m1 - m0) / rk)  [has_d1]  (+)  rho_t * (mt - m0)  [has_t]
__global__ void __launch_bounds__(kEwThreads)
unipc_x0_kernel(const float* __restrict__ eps, const float* __restrict__ x, float* __restrict__ out, long long n,
                float sigma_t, float alpha_t) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = __fdiv_rn(__fsub_rn(x[i], __fmul_rn(sigma_t, eps[i])), alpha_t);
}

__global__ void __launch_bounds__(kEwThreads)
unipc_update_kernel(const float* __restrict__ x, const float* __restrict__ m0, const float* __restrict__ m1,
                    const float* __restrict__ mt, float* __restrict__ out, long long n, float cx, float cm, float cb,
                    float rk, float rho0, float rho_t, int has_d1, int has_t) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float a = m0[i];
    const float base = __fsub_rn(__fmul_rn(cx, x[i]), __fmul_rn(cm, a));
    float res = 0.f;
    if (has_d1) res = __fmul_rn(rho0, __fdiv_rn(__fsub_rn(m1[i], a), rk));
    if (has_t) {
      const float dt = __fmul_rn(rho_t, __fsub_rn(mt[i], a));
      res = has_d1 ? __fadd_rn(res, dt) : dt;
    }
    out[i] = (has_d1 || has_t) ? __fsub_rn(base, __fmul_rn(cb, res)) : base;
  }
}

// ---- fused global-norm clip + AdamW over the flat parameter arena -----------------------
// train_from_scratch.py:106-108: clip_grad_norm_(params, 1.0); optimizer.step().  With every parameter, gradient and
// moment in ONE flat fp32 buffer the update is two streaming kernels (sum of squares; update) instead of ~30
// multi-tensor launches; the clip coefficient is applied on the fly, so the gradients are read once.
// Deterministic: every block writes its partial sum to a fixed slot, the block that finishes last adds the slots in
// index order.  (An atomicAdd per block made the global norm -- and with it the clip coefficient -- differ in the last
// bit between the ranks of a data-parallel job that hold bit-identical gradients, so the replicas drifted apart:
// tests/nccl_worker.py.)  The scratch below is per device and not re-entrant across streams; the optimizer step is the
// only caller and runs on one stream.
constexpr int kSumsqMaxBlocks = kNumSMs * 16;
__device__ float g_sumsq_partials[kSumsqMaxBlocks];
__device__ unsigned int g_sumsq_ticket = 0;

__global__ void __launch_bounds__(kEwThreads)
sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long nvec = n >> 2;
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (long long i = (nvec << 2) + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    acc += x[i] * x[i];
  __shared__ float red[kEwThreads / 32];
  __shared__ bool last;
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < kEwThreads / 32 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) {
      g_sumsq_partials[blockIdx.x] = v;
      __threadfence();
      last = atomicAdd(&g_sumsq_ticket, 1u) == gridDim.x - 1;
    }
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  // fixed-order final reduction: thread t adds slots t, t + 256, ... then the usual shuffle tree
  float tot = 0.f;
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += kEwThreads) tot += __ldcg(&g_sumsq_partials[i]);
  tot = warp_sum(tot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tot;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < kEwThreads / 32 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) {
      *out += v;
      g_sumsq_ticket = 0;
    }
  }
}

// scal[0] = step (incremented here), scal[1] = clip coefficient, scal[2] = 1 - beta1^step, scal[3] = 1 - beta2^step
__global__ void adamw_tick_kernel(float* __restrict__ scal, const float* __restrict__ gnorm_sq, float max_norm,
                                  float beta1, float beta2) {
  const float step = scal[0] + 1.0f;
  scal[0] = step;
  float coef = 1.0f;
  if (gnorm_sq != nullptr && max_norm > 0.f) coef = fminf(1.0f, max_norm / (sqrtf(*gnorm_sq) + 1e-6f));
  scal[1] = coef;
  scal[2] = 1.0f - powf(beta1, step);
  scal[3] = 1.0f - powf(beta2, step);
}

__global__ void __launch_bounds__(kEwThreads)
adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  long long n, const float* __restrict__ scal, const float* __restrict__ lr_dev, float lr_host,
                  float beta1, float beta2, float eps, float wd) {
  const float lr = lr_dev ? *lr_dev : lr_host;
  const float coef = scal[1];
  const float step_size = lr / scal[2];
  const float inv_sqrt_bc2 = rsqrtf(scal[3]);
  const float decay = 1.0f - lr * wd;
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long nvec = n >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pa = reinterpret_cast<float*>(&pp);
    const float* ga = reinterpret_cast<const float*>(&gg);
    float* ma = reinterpret_cast<float*>(&mm);
    float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = ga[j] * coef;
      const float pj = pa[j] * decay;
      const float mj = ma[j] + (gr - ma[j]) * omb1;
      const float vj = va[j] * beta2 + gr * gr * omb2;
      const float denom = sqrtf(vj) * inv_sqrt_bc2 + eps;
      pa[j] = pj - step_size * (mj / denom);
      ma[j] = mj;
      va[j] = vj;
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (long long i = (nvec << 2) + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gr = g[i] * coef;
    const float pj = p[i] * decay;
    const float mj = m[i] + (gr - m[i]) * omb1;
    const float vj = v[i] * beta2 + gr * gr * omb2;
    p[i] = pj - step_size * (mj / (sqrtf(vj) * inv_sqrt_bc2 + eps));
    m[i] = mj;
    v[i] = vj;
  }
}

}  // namespace ddpm

using namespace ddpm;

extern "C" const char* ddpm_last_error(void) { return g_err; }
extern "C" int ddpm_abi_version(void) { return DDPM_ABI_VERSION; }

extern "C" int ddpm_add_noise(const float* x0, const float* noise, const long long* t, const float* sqrt_ac,
                              const float* sqrt_1mac, float* out, int batch, long long per_sample,
                              int num_train_timesteps, void* stream) {
  DDPM_REQUIRE(x0 && noise && t && sqrt_ac && sqrt_1mac && out, "ddpm_add_noise: null pointer argument");
  DDPM_REQUIRE(batch >= 0 && per_sample >= 0 && num_train_timesteps > 0, "ddpm_add_noise: bad sizes");
  const long long total = static_cast<long long>(batch) * per_sample;
  if (total == 0) return DDPM_OK;
  const bool vec_ok = (per_sample % 4 == 0) && ((reinterpret_cast<uintptr_t>(x0) | reinterpret_cast<uintptr_t>(noise) |
                                                 reinterpret_cast<uintptr_t>(out)) % 16 == 0);
  const long long items = vec_ok ? total / 4 : total;
  add_noise_kernel<<<ew_blocks(items), kEwThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      x0, noise, t, sqrt_ac, sqrt_1mac, out, per_sample, items, vec_ok ? 1 : 0, num_train_timesteps);
  return check_launch("add_noise_kernel");
}

extern "C" int ddpm_mse_fwd_bwd(const float* pred, const float* target, float* loss_sum, float* dpred, long long n,
                                void* stream) {
  DDPM_REQUIRE(pred && target && loss_sum, "ddpm_mse_fwd_bwd: null pointer argument");
  DDPM_REQUIRE(n > 0, "ddpm_mse_fwd_bwd: n must be positive");
  DDPM_REQUIRE((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(target) |
                reinterpret_cast<uintptr_t>(dpred)) % 16 == 0, "ddpm_mse_fwd_bwd: pointers must be 16-byte aligned");
  mse_fwd_bwd_kernel<<<ew_blocks(n / 4 + 1), kEwThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      pred, target, loss_sum, dpred, n, 2.0f / static_cast<float>(n));
  return check_launch("mse_fwd_bwd_kernel");
}

extern "C" int ddpm_scale_by_device_scalar(float* x, const float* scale, long long n, void* stream) {
  DDPM_REQUIRE(x && scale && n >= 0, "ddpm_scale_by_device_scalar: bad argument");
  DDPM_REQUIRE(reinterpret_cast<uintptr_t>(x) % 16 == 0, "ddpm_scale_by_device_scalar: x must be 16-byte aligned");
  if (n == 0) return DDPM_OK;
  scale_by_scalar_kernel<<<ew_blocks(n / 4 + 1), kEwThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, scale, n);
  return check_launch("scale_by_scalar_kernel");
}

extern "C" int ddpm_scheduler_step(const float* eps, const float* x, const float* z, float* prev, float* pred_x0,
                                   long long n, float sqrt_alpha_prod, float sqrt_beta_prod, float c0, float ct,
                                   float sigma, float clip, void* stream) {
  DDPM_REQUIRE(eps && x && prev && n >= 0, "ddpm_scheduler_step: bad argument");
  DDPM_REQUIRE((reinterpret_cast<uintptr_t>(eps) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(z) |
                reinterpret_cast<uintptr_t>(prev) | reinterpret_cast<uintptr_t>(pred_x0)) % 16 == 0,
               "ddpm_scheduler_step: pointers must be 16-byte aligned");
  if (n == 0) return DDPM_OK;
  StepCoef k{sqrt_alpha_prod, sqrt_beta_prod, c0, ct, sigma, clip};
  scheduler_step_kernel<<<ew_blocks(n / 4 + 1), kEwThreads, 0, static_cast<cudaStream_t>(stream)>>>(eps, x, z, prev,
                                                                                                  pred_x0, n, k);
  return check_launch("scheduler_step_kernel");
}

extern "C" int ddpm_ddim_step(const float* eps, const float* x, const float* z, float* prev, float* pred_x0, long long n,
                              float sqrt_alpha_prod, float sqrt_beta_prod, float sqrt_alpha_prod_prev, float dir_coef,
                              float sigma, float clip, int use_clipped_model_output, void* stream) {
  DDPM_REQUIRE(eps && x && prev && n >= 0, "ddpm_ddim_step: bad argument");
  DDPM_REQUIRE((reinterpret_cast<uintptr_t>(eps) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(z) |
                reinterpret_cast<uintptr_t>(prev) | reinterpret_cast<uintptr_t>(pred_x0)) % 16 == 0,
               "ddpm_ddim_step: pointers must be 16-byte aligned");
  if (n == 0) return DDPM_OK;
  DdimCoef k{sqrt_alpha_prod, sqrt_beta_prod, sqrt_alpha_prod_prev, dir_coef, sigma, clip, use_clipped_model_output};
  ddim_step_kernel<<<ew_blocks(n / 4 + 1), kEwThreads, 0, static_cast<cudaStream_t>(stream)>>>(eps, x, z, prev, pred_x0,
                                                                                               n, k);
  return check_launch("ddim_step_kernel");
}

extern "C" int ddpm_scheduler_step_philox(const float* eps, const float* x, float* prev, long long n,
                                          float sqrt_alpha_prod, float sqrt_beta_prod, float c0, float ct,
                                          float sigma, float clip, unsigned long long seed,
                                          unsigned long long offset, void* stream) {
  DDPM_REQUIRE(eps && x && prev && n >= 0, "ddpm_scheduler_step_philox: bad argument");
  DDPM_REQUIRE((reinterpret_cast<uintptr_t>(eps) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(prev)) % 16 == 0,
               "ddpm_scheduler_step_philox: pointers must be 16-byte aligned");
  if (n == 0) return DDPM_OK;
  StepCoef k{sqrt_alpha_prod, sqrt_beta_prod, c0, ct, sigma, clip};
  scheduler_step_philox_kernel<<<ew_blocks(n / 4 + 1), kEwThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      eps, x, prev, n, k, seed, offset);
  return check_launch("scheduler_step_philox_kernel");
}

extern "C" int ddpm_to_uint8_nhwc(const float* x, unsigned char* out, int n, int c, int h, int w, void* stream) {
  DDPM_REQUIRE(x && out && n >= 0 && c > 0 && h > 0 && w > 0, "ddpm_to_uint8_nhwc: bad argument");
  if (n == 0) return DDPM_OK;
  to_uint8_nhwc_kernel<<<ew_blocks(static_cast<long long>(n) * c * h * w), kEwThreads, 0,
                         static_cast<cudaStream_t>(stream)>>>(x, out, n, c, h, w);
  return check_launch("to_uint8_nhwc_kernel");
}

// ---- LoRA dropout (peft lora_dropout on the adapter branch; train_with_lora_all_classes.py:316-322) ----------
// The keep mask is a pure function of (seed, offset, element index): Philox4x32-10 yields eight 16-bit lanes per
// 8-element vector, element kept iff lane >= p*65536.  Nothing is stored; backward regenerates the same mask.
namespace ddpm {
__device__ __forceinline__ void dropout_mask8(long long vec_idx, unsigned long long seed, unsigned long long offset,
                                              uint32_t thresh, float keep_scale, float* m) {
  const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(vec_idx), static_cast<uint32_t>(vec_idx >> 32),
                                           static_cast<uint32_t>(offset), static_cast<uint32_t>(offset >> 32)),
                                make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[2 * i] = (w[i] & 0xFFFFu) >= thresh ? keep_scale : 0.f;
    m[2 * i + 1] = (w[i] >> 16) >= thresh ? keep_scale : 0.f;
  }
}

// out = (add ? add : 0) + x * mask / (1 - p)     (forward: add == NULL; backward: x = dy, add = upstream grad)
__global__ void __launch_bounds__(kEwThreads)
dropout_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ add,
               __nv_bfloat16* __restrict__ out, long long nvec, uint32_t thresh, float keep_scale,
               unsigned long long seed, unsigned long long offset, const unsigned long long* __restrict__ tick) {
  // tick: device-resident step counter (upper Philox counter word), so a CUDA-graph replay of the step draws a
  // fresh mask although seed / offset are baked into the captured launch
  if (tick != nullptr) offset += *tick << 32;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float f[8], m[8];
    unpack8(reinterpret_cast<const bf16x8*>(x)[i], f);
    dropout_mask8(i, seed, offset, thresh, keep_scale, m);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] *= m[e];
    if (add) {
      float a[8];
      unpack8(reinterpret_cast<const bf16x8*>(add)[i], a);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] += a[e];
    }
    reinterpret_cast<bf16x8*>(out)[i] = pack8(f);
  }
}
}  // namespace ddpm

extern "C" int ddpm_dropout(const void* x, const void* add, void* out, long long n, float p, unsigned long long seed,
                            unsigned long long offset, const unsigned long long* tick, void* stream) {
  DDPM_REQUIRE(x && out && n >= 0 && n % 8 == 0 && p >= 0.f && p < 1.f, "ddpm_dropout: bad argument (n %% 8, 0<=p<1)");
  DDPM_REQUIRE((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(add) | reinterpret_cast<uintptr_t>(out)) % 16 == 0,
               "ddpm_dropout: pointers must be 16-byte aligned");
  if (n == 0) return DDPM_OK;
  const uint32_t thresh = static_cast<uint32_t>(p * 65536.0f + 0.5f);
  dropout_kernel<<<ew_blocks(n / 8), kEwThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(add), static_cast<__nv_bfloat16*>(out),
      n / 8, thresh, 1.0f / (1.0f - p), seed, offset, tick);
  return check_launch("dropout_kernel");
}

extern "C" int ddpm_sumsq_f32(const float* x, long long n, float* out, void* stream) {
  DDPM_REQUIRE(x && out && n >= 0, "ddpm_sumsq_f32: bad argument");
  DDPM_REQUIRE(reinterpret_cast<uintptr_t>(x) % 16 == 0, "ddpm_sumsq_f32: x must be 16-byte aligned");
  if (n == 0) return DDPM_OK;
  sumsq_kernel<<<ew_blocks(n / 4 + 1), kEwThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, n, out);
  return check_launch("sumsq_kernel");
}

extern "C" int ddpm_adamw_flat(float* p, const float* g, float* m, float* v, long long n, float* scal,
                               const float* gnorm_sq, float max_norm, const float* lr_dev, float lr, float beta1,
                               float beta2, float eps, float weight_decay, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DDPM_REQUIRE(p && g && m && v && scal && n > 0, "ddpm_adamw_flat: bad argument");
  DDPM_REQUIRE((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                reinterpret_cast<uintptr_t>(v)) % 16 == 0, "ddpm_adamw_flat: buffers must be 16-byte aligned");
  adamw_tick_kernel<<<1, 1, 0, stream>>>(scal, gnorm_sq, max_norm, beta1, beta2);
  if (int e = check_launch("adamw_tick_kernel")) return e;
  adamw_flat_kernel<<<ew_blocks(n / 4 + 1), kEwThreads, 0, stream>>>(p, g, m, v, n, scal, lr_dev, lr, beta1, beta2, eps,
                                                                     weight_decay);
  return check_launch("adamw_flat_kernel");
}

extern "C" int ddpm_unipc_x0(const float* eps, const float* x, float* out, long long n, float sigma_t, float alpha_t,
                             void* stream) {
  DDPM_REQUIRE(eps && x && out && n >= 0, "ddpm_unipc_x0: bad argument");
  if (n == 0) return DDPM_OK;
  unipc_x0_kernel<<<ew_blocks(n), kEwThreads, 0, static_cast<cudaStream_t>(stream)>>>(eps, x, out, n, sigma_t, alpha_t);
  return check_launch("unipc_x0_kernel");
}

extern "C" int ddpm_unipc_update(const float* x, const float* m0, const float* m1, const float* mt, float* out,
                                 long long n, float cx, float cm, float cb, float rk, float rho0, float rho_t,
                                 void* stream) {
  DDPM_REQUIRE(x && m0 && out && n >= 0, "ddpm_unipc_update: bad argument");
  if (n == 0) return DDPM_OK;
  unipc_update_kernel<<<ew_blocks(n), kEwThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      x, m0, m1, mt, out, n, cx, cm, cb, rk, rho0, rho_t, m1 != nullptr ? 1 : 0, mt != nullptr ? 1 : 0);
  return check_launch("unipc_update_kernel");
}
