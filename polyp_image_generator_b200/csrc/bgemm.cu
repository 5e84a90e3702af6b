// Batched bf16 GEMM on tcgen05 for the wide-head self-attention blocks (head_dim > 64, e.g. the single 512-wide head
// of the google/ddpm-celebahq-256 architecture that train_with_lora_*.py fine-tunes -- SURVEY.md §8 row a3.4, App. A.5),
// plus the row-softmax kernels that sit between the GEMMs.
//
//   C[z] = alpha * A[z] (m x k) * B[z] (k x n),   z = (batch, head),  bf16 operands, fp32 accumulation in TMEM
//
// Replaces what AttnProcessor2_0's F.scaled_dot_product_attention dispatches to on the reference path
// (diffusers Attention blocks of the UNet2DModel built at /root/reference/generator_model/PolypGeneratorModel.py:25-58).
//
// Each operand is described by a row stride and a layout flag, so every product of the attention forward and backward
// (Q K^T, P V, dO V^T, P^T dO, dS K, dS^T Q) reads q / k / v / dO / P / dS straight out of the buffers they already
// live in -- nothing is transposed or copied:
//   *_mn = 0 ("K-major"):  the reduction index is contiguous      A(i, kk) = a[i*lda + kk]   B(kk, j) = b[j*ldb + kk]
//   *_mn = 1 ("MN-major"): the output index is contiguous          A(i, kk) = a[kk*lda + i]   B(kk, j) = b[kk*ldb + j]
// TMA (4-D maps: inner, rows, head, batch; SWIZZLE_128B) stages 128 x 64 operand blocks in a 4-deep ring; rows and
// reduction indices past the matrix edge are zero-filled by TMA, so token counts need not be multiples of the tile.
// One CTA per 128 x 128 output tile: warp 0 TMA, warp 1 MMA issue, warps 2-5 epilogue (TMEM -> alpha -> bf16 / fp32).
#include "common.cuh"

#include "../../include/ddpm_b200.h"

namespace ddpm {

constexpr int kBgM = 128, kBgN = 128, kBgK = 64, kBgStages = 4;
constexpr int kBgThreads = 192;
constexpr int kBgTileBytes = 128 * 64 * 2;                       // one operand block: 16 KB
constexpr int kBgStageBytes = 2 * kBgTileBytes;
constexpr int kBgSmem = kBgStages * kBgStageBytes + 256 + 1024;

struct BgemmParams {
  int m, n, k, heads;
  int a_mn, b_mn;
  float alpha;
  __nv_bfloat16* c16;
  float* c32;
  long long ldc, c_head, c_batch;
};

__global__ void __launch_bounds__(kBgThreads, 1)
bgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ BgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kBgStages * kBgStageBytes);
  uint64_t* empty_bar = full_bar + kBgStages;
  uint64_t* tmem_full_bar = empty_bar + kBgStages;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * kBgN, m0 = blockIdx.y * kBgM;
  const int zb = blockIdx.z / p.heads, zh = blockIdx.z - zb * p.heads;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kBgStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_base_slot, kBgN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const int iters = (p.k + kBgK - 1) / kBgK;

  if (warp == 0) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % kBgStages;
      const uint32_t ph = (it / kBgStages) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      if (elect_one()) {
        uint8_t* sa = smem + s * kBgStageBytes;
        uint8_t* sb = sa + kBgTileBytes;
        mbar_expect_tx(&full_bar[s], kBgStageBytes);
        const int k0 = it * kBgK;
        if (p.a_mn) {   // two [64 k-rows][64 m] boxes, 8 KB apart
          tma_load_4d(sa, &tmA, &full_bar[s], m0, k0, zh, zb);
          tma_load_4d(sa + kBgTileBytes / 2, &tmA, &full_bar[s], m0 + 64, k0, zh, zb);
        } else {        // one [128 m-rows][64 k] box
          tma_load_4d(sa, &tmA, &full_bar[s], k0, m0, zh, zb);
        }
        if (p.b_mn) {
          tma_load_4d(sb, &tmB, &full_bar[s], n0, k0, zh, zb);
          tma_load_4d(sb + kBgTileBytes / 2, &tmB, &full_bar[s], n0 + 64, k0, zh, zb);
        } else {
          tma_load_4d(sb, &tmB, &full_bar[s], k0, n0, zh, zb);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(kBgM, kBgN, p.a_mn != 0, p.b_mn != 0);
    // K-major: 8-row groups 1024 B apart, a K = 16 step is 32 B inside the swizzle row.
    // MN-major: 64-wide MN blocks 8 KB apart, 8-row K groups 1024 B apart, a K = 16 step is 16 rows = 2 KB.
    const uint32_t a_lbo = p.a_mn ? kBgTileBytes / 2 : 16, b_lbo = p.b_mn ? kBgTileBytes / 2 : 16;
    const uint32_t a_step = (p.a_mn ? 16 * 128 : 32) >> 4, b_step = (p.b_mn ? 16 * 128 : 32) >> 4;
    for (int it = 0; it < iters; ++it) {
      const int s = it % kBgStages;
      const uint32_t ph = (it / kBgStages) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem + s * kBgStageBytes);
      const uint64_t da0 = make_smem_desc_sw128(a_addr, a_lbo, 1024);
      const uint64_t db0 = make_smem_desc_sw128(a_addr + kBgTileBytes, b_lbo, 1024);
      const uint32_t first = it != 0 ? 1u : 0u;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < kBgK / 16; ++kk)
          umma_bf16(tmem_base, da0 + static_cast<uint64_t>(kk * a_step), db0 + static_cast<uint64_t>(kk * b_step), idesc,
                    kk == 0 ? first : 1u);
        umma_commit(&empty_bar[s]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(tmem_full_bar);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    const bool rvalid = row < p.m;
    const long long off = static_cast<long long>(zb) * p.c_batch + static_cast<long long>(zh) * p.c_head +
                          static_cast<long long>(row) * p.ldc;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < kBgN / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, r);
      tmem_ld_wait();
      const int col = n0 + c * 32;
      if (!rvalid || col >= p.n) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
      if (p.c32 != nullptr) {
        float* dst = p.c32 + off + col;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (col + j + 4 <= p.n) {
            *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
            for (int e = 0; e < 4; ++e)
              if (col + j + e < p.n) dst[j + e] = v[j + e];
          }
        }
      } else {
        __nv_bfloat16* dst = p.c16 + off + col;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (col + j + 8 <= p.n) {
            *reinterpret_cast<bf16x8*>(dst + j) = pack8(v + j);
          } else {
            for (int e = 0; e < 8; ++e)
              if (col + j + e < p.n) dst[j + e] = __float2bfloat16(v[j + e]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kBgN);
}

// ---- row softmax between the GEMMs: one warp per (batch, head, query) row --------------------------------------
// P = softmax(S) (S already carries the d^-1/2 scale); S fp32 [rows][lds], P bf16 [rows][ldp]
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ s, long long lds, __nv_bfloat16* __restrict__ pout, long long ldp,
                    long long rows, int t) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* sr = s + row * lds;
  float mx = -INFINITY;
  for (int j = lane; j < t; j += 32) mx = fmaxf(mx, sr[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int j = lane; j < t; j += 32) sum += __expf(sr[j] - mx);
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  __nv_bfloat16* pr = pout + row * ldp;
  for (int j = lane; j < t; j += 32) pr[j] = __float2bfloat16(__expf(sr[j] - mx) * inv);
}

// dS = scale * P o (dP - sum_j P_j dP_j); P bf16, dP fp32, dS bf16 (same leading dimension as P)
__global__ void __launch_bounds__(256)
softmax_rows_bwd_kernel(const __nv_bfloat16* __restrict__ pin, long long ldp, const float* __restrict__ dp,
                        long long lddp, __nv_bfloat16* __restrict__ ds, long long rows, int t, float scale) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const __nv_bfloat16* pr = pin + row * ldp;
  const float* dr = dp + row * lddp;
  float dot = 0.f;
  for (int j = lane; j < t; j += 32) dot += __bfloat162float(pr[j]) * dr[j];
  dot = warp_sum(dot);
  __nv_bfloat16* o = ds + row * ldp;
  for (int j = lane; j < t; j += 32) o[j] = __float2bfloat16(scale * __bfloat162float(pr[j]) * (dr[j] - dot));
}

static int make_operand_map(CUtensorMap* out, const void* ptr, int mn_major, long long rows, long long k, long long ld,
                            long long head_stride, long long batch_stride, int heads, int batch) {
  // K-major: (k, rows, head, batch), box (64, 128);  MN-major: (rows, k, head, batch), box (64, 64)
  const long long dims[4] = {mn_major ? rows : k, mn_major ? k : rows, heads, batch};
  // a size-1 dimension's stride is never used for addressing but must still satisfy the encoder (multiple of 16 B)
  const long long st[3] = {ld, heads > 1 ? head_stride : ld, batch > 1 ? batch_stride : ld};
  const int box[4] = {64, mn_major ? 64 : 128, 1, 1};
  return make_map4(out, ptr, dims, st, box);
}

}  // namespace ddpm

using namespace ddpm;

extern "C" int ddpm_bgemm(const void* a, long long lda, long long a_head, long long a_batch, int a_mn, const void* b,
                          long long ldb, long long b_head, long long b_batch, int b_mn, void* c, long long ldc,
                          long long c_head, long long c_batch, int c_f32, int m, int n, int k, int heads, int batch,
                          float alpha, void* stream) {
  DDPM_REQUIRE(a && b && c, "ddpm_bgemm: null pointer");
  DDPM_REQUIRE(m > 0 && n > 0 && k > 0 && heads > 0 && batch > 0, "ddpm_bgemm: bad shape m=%d n=%d k=%d heads=%d batch=%d",
               m, n, k, heads, batch);
  DDPM_REQUIRE(static_cast<long long>(heads) * batch <= 65535, "ddpm_bgemm: heads*batch=%lld exceeds gridDim.z",
               static_cast<long long>(heads) * batch);
  const int vec = c_f32 ? 4 : 8;
  DDPM_REQUIRE(ldc % vec == 0 && c_head % vec == 0 && c_batch % vec == 0 &&
                   (reinterpret_cast<uintptr_t>(c) & 15) == 0,
               "ddpm_bgemm: output rows must be 16-byte aligned (ldc=%lld)", ldc);
  CUtensorMap tmA, tmB;
  if (int e = make_operand_map(&tmA, a, a_mn, m, k, lda, a_head, a_batch, heads, batch)) return e;
  if (int e = make_operand_map(&tmB, b, b_mn, n, k, ldb, b_head, b_batch, heads, batch)) return e;
  BgemmParams p;
  p.m = m; p.n = n; p.k = k; p.heads = heads;
  p.a_mn = a_mn ? 1 : 0; p.b_mn = b_mn ? 1 : 0;
  p.alpha = alpha;
  p.c16 = c_f32 ? nullptr : static_cast<__nv_bfloat16*>(c);
  p.c32 = c_f32 ? static_cast<float*>(c) : nullptr;
  p.ldc = ldc; p.c_head = c_head; p.c_batch = c_batch;
  static bool attr_set = false;
  if (!attr_set) {
    DDPM_CUDA(cudaFuncSetAttribute(bgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBgSmem));
    attr_set = true;
  }
  dim3 grid((n + kBgN - 1) / kBgN, (m + kBgM - 1) / kBgM, heads * batch);
  bgemm_kernel<<<grid, kBgThreads, kBgSmem, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, p);
  return check_launch("bgemm_kernel");
}

extern "C" int ddpm_softmax_rows(const float* s, long long lds, void* p, long long ldp, long long rows, int t,
                                 void* stream) {
  DDPM_REQUIRE(s && p && rows > 0 && t > 0 && lds >= t && ldp >= t, "ddpm_softmax_rows: bad argument");
  const int wpb = 8;
  softmax_rows_kernel<<<static_cast<unsigned>((rows + wpb - 1) / wpb), wpb * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      s, lds, static_cast<__nv_bfloat16*>(p), ldp, rows, t);
  return check_launch("softmax_rows_kernel");
}

extern "C" int ddpm_softmax_rows_bwd(const void* p, long long ldp, const float* dp, long long lddp, void* ds,
                                     long long rows, int t, float scale, void* stream) {
  DDPM_REQUIRE(p && dp && ds && rows > 0 && t > 0 && ldp >= t && lddp >= t, "ddpm_softmax_rows_bwd: bad argument");
  const int wpb = 8;
  softmax_rows_bwd_kernel<<<static_cast<unsigned>((rows + wpb - 1) / wpb), wpb * 32, 0,
                            static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(p), ldp, dp, lddp,
                                                                 static_cast<__nv_bfloat16*>(ds), rows, t, scale);
  return check_launch("softmax_rows_bwd_kernel");
}
