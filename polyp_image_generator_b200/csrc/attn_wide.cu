// Fused wide-head self-attention forward on tcgen05 (head_dim a multiple of 128, up to 256 tokens): the 16x16 / 8x8
// attention blocks of the google/ddpm-celebahq-256 architecture -- ONE head of 512 channels, T = 256 / 64 tokens at
// 256^2 -- that train_with_lora_*.py fine-tunes (north_star "the mid/low-resolution
// self-attention blocks as a fused tcgen05 attention kernel").  Replaces what AttnProcessor2_0's
// F.scaled_dot_product_attention dispatches to on the reference path (UNet2DModel of
// /root/reference/generator_model/PolypGeneratorModel.py:25-58; LoRA loop train_with_lora_all_classes.py:120-180).
//
// One CTA per (sample, head, block of 128 queries); S and P never leave the SM:
//   phase 1  S[128 x T] = Q K^T           tcgen05.mma over d in 64-channel steps, Q / K staged by TMA (K-major boxes),
//                                         fp32 accumulator in TMEM columns 0 .. 255
//   phase 2  P = softmax(scale * S)       the four epilogue warps read their rows from TMEM (one thread = one query),
//                                         write P as bf16 straight into shared memory in the SWIZZLE_128B K-major
//                                         layout a TMA load would have produced (the A operand of phase 3), and -- for
//                                         training -- also to global memory: the backward still runs on the batched
//                                         GEMMs of bgemm.cu and reads these probabilities
//   phase 3  O[128 x d] = P V             128 output channels at a time, V staged by TMA as an MN-major operand
//                                         (tokens = reduction index), accumulators double-buffered in TMEM columns
//                                         256 .. 511 so the store of one chunk overlaps the MMAs of the next
// The unfused path (three launches, S fp32 and P bf16 round-tripping through L2 / HBM) stays for T > 256 and for head
// sizes that are not multiples of 128.
#include "common.cuh"

#include "../../include/ddpm_b200.h"

namespace ddpm {

constexpr int kAwThreads = 192;                 // warp 0: TMA, warp 1: MMA + TMEM owner, warps 2-5: softmax / epilogue
constexpr int kAwStages = 3;
constexpr int kAwTile = 128 * 64 * 2;           // one 128-row x 64-element operand block: 16 KB
constexpr int kAwStageBytes = 3 * kAwTile;      // phase 1: Q block + two K blocks; phase 3: one V tile in the first 16 KB
constexpr int kAwPBytes = 4 * kAwTile;          // P: 128 queries x 256 keys, four K-major 64-key blocks
constexpr int kAwSmem = kAwStages * kAwStageBytes + kAwPBytes + 256 + 1024;
constexpr uint32_t kAwOCol = 256;               // TMEM column of the first O accumulator

struct AttnWideParams {
  int t, d, heads;
  float scale;
  __nv_bfloat16* o;
  long long ldo;
  __nv_bfloat16* p;                             // NULL: inference, probabilities are not kept
  long long ldp, p_head, p_batch;
};

__global__ void __launch_bounds__(kAwThreads, 1)
attn_wide_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnWideParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* psm = smem + kAwStages * kAwStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(psm + kAwPBytes);
  uint64_t* empty_bar = full_bar + kAwStages;
  uint64_t* s_full = empty_bar + kAwStages;
  uint64_t* p_ready = s_full + 1;
  uint64_t* o_full = p_ready + 1;      // [2]
  uint64_t* o_empty = o_full + 2;      // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int zb = blockIdx.y / p.heads, zh = blockIdx.y - zb * p.heads;
  const int nh = p.t > 128 ? 2 : 1;               // 128-key halves of S
  const int dsteps = p.d / 64;                    // phase-1 reduction steps
  const int kblocks = (p.t + 63) / 64;            // phase-3 reduction steps (64 keys each)
  const int nchunks = p.d / 128;                  // phase-3 output chunks

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int s = 0; s < kAwStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_base_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    int s = 0;
    uint32_t ph = 1;
    for (int kc = 0; kc < dsteps; ++kc) {
      mbar_wait(&empty_bar[s], ph);
      if (elect_one()) {
        uint8_t* sa = smem + s * kAwStageBytes;
        mbar_expect_tx(&full_bar[s], static_cast<uint32_t>((1 + nh) * kAwTile));
        tma_load_4d(sa, &tmQ, &full_bar[s], kc * 64, m0, zh, zb);
        tma_load_4d(sa + kAwTile, &tmK, &full_bar[s], kc * 64, 0, zh, zb);
        if (nh == 2) tma_load_4d(sa + 2 * kAwTile, &tmK, &full_bar[s], kc * 64, 128, zh, zb);
      }
      __syncwarp();
      if (++s == kAwStages) { s = 0; ph ^= 1; }
    }
    for (int c = 0; c < nchunks; ++c) {
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&empty_bar[s], ph);
        if (elect_one()) {
          uint8_t* sa = smem + s * kAwStageBytes;
          mbar_expect_tx(&full_bar[s], static_cast<uint32_t>(kAwTile));
          // V tile [64 keys][128 channels] as two MN-major boxes of 64 channels, 8 KB apart
          tma_load_4d(sa, &tmV, &full_bar[s], c * 128, kb * 64, zh, zb);
          tma_load_4d(sa + kAwTile / 2, &tmV, &full_bar[s], c * 128 + 64, kb * 64, zh, zb);
        }
        __syncwarp();
        if (++s == kAwStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, false, false);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, false, true);
    int s = 0;
    uint32_t ph = 0;
    for (int kc = 0; kc < dsteps; ++kc) {
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem + s * kAwStageBytes);
      const uint64_t da0 = make_smem_desc_sw128(a_addr, 16, 1024);
      const uint64_t db0 = make_smem_desc_sw128(a_addr + kAwTile, 16, 1024);
      const uint64_t db1 = make_smem_desc_sw128(a_addr + 2 * kAwTile, 16, 1024);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {            // K = 16 step: 32 bytes inside the swizzle row
          const uint32_t accum = (kc | kk) != 0 ? 1u : 0u;
          umma_bf16(tmem_base, da0 + static_cast<uint64_t>(kk * 2), db0 + static_cast<uint64_t>(kk * 2), idesc_qk, accum);
          if (nh == 2)
            umma_bf16(tmem_base + 128, da0 + static_cast<uint64_t>(kk * 2), db1 + static_cast<uint64_t>(kk * 2), idesc_qk,
                      accum);
        }
        umma_commit(&empty_bar[s]);
      }
      __syncwarp();
      if (++s == kAwStages) { s = 0; ph ^= 1; }
    }
    if (elect_one()) umma_commit(s_full);
    __syncwarp();
    mbar_wait(p_ready, 0);                          // P is in shared memory (writers fenced it to the async proxy)
    tc_fence_after();
    const uint32_t p_addr = smem_u32(psm);
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      mbar_wait(&o_empty[buf], (((c >> 1) & 1) ^ 1));   // the epilogue has drained this accumulator (free at first use)
      tc_fence_after();
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint64_t da0 = make_smem_desc_sw128(p_addr + kb * kAwTile, 16, 1024);
        // MN-major B: 64-channel blocks 8 KB apart, 8-key groups 1024 B apart, a K = 16 step is 16 rows = 2 KB
        const uint64_t db0 = make_smem_desc_sw128(smem_u32(smem + s * kAwStageBytes), kAwTile / 2, 1024);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tmem_base + kAwOCol + buf * 128, da0 + static_cast<uint64_t>(kk * 2),
                      db0 + static_cast<uint64_t>(kk * 128), idesc_pv, (kb | kk) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[s]);
        }
        __syncwarp();
        if (++s == kAwStages) { s = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(&o_full[buf]);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;                         // TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;                  // query row of the tile = TMEM lane
    const int grow = m0 + row;
    const bool rvalid = grow < p.t;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int nch = (p.t + 31) / 32;                // 32-column chunks of S that hold keys
    mbar_wait(s_full, 0);
    tc_fence_after();
    float mx = -INFINITY;
    for (int c = 0; c < nch; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(trow + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c * 32 + j < p.t) mx = fmaxf(mx, __uint_as_float(r[j]) * p.scale);
    }
    float sum = 0.f;
    for (int c = 0; c < nch; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(trow + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c * 32 + j < p.t) sum += __expf(__uint_as_float(r[j]) * p.scale - mx);
    }
    const float inv = 1.0f / sum;
    __nv_bfloat16* prow = (p.p != nullptr && rvalid)
                              ? p.p + static_cast<long long>(zb) * p.p_batch + static_cast<long long>(zh) * p.p_head +
                                    static_cast<long long>(grow) * p.ldp
                              : nullptr;
    uint8_t* srow = psm + (row >> 3) * 1024 + (row & 7) * 128;
    for (int c = 0; c < 2 * kblocks; ++c) {         // every 32-key chunk of the 64-key blocks phase 3 will read
      float v[32];
      if (c < nch) {
        uint32_t r[32];
        tmem_ld_32x32(trow + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          v[j] = (c * 32 + j < p.t) ? __expf(__uint_as_float(r[j]) * p.scale - mx) * inv : 0.f;
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      uint8_t* blk = srow + (c >> 1) * kAwTile;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bf16x8 w = pack8(v + 8 * i);
        const int chunk = (c & 1) * 4 + i;          // 16-byte chunk of the 128-byte row; swizzled with the row index
        *reinterpret_cast<bf16x8*>(blk + ((chunk ^ (row & 7)) << 4)) = w;
        if (prow != nullptr && c * 32 + 8 * i < p.ldp) *reinterpret_cast<bf16x8*>(prow + c * 32 + 8 * i) = w;
      }
    }
    fence_proxy_async();                            // generic-proxy stores -> visible to the tensor core's async proxy
    tc_fence_before();
    mbar_arrive(p_ready);

    __nv_bfloat16* orow = p.o + (static_cast<long long>(zb) * p.t + grow) * p.ldo + static_cast<long long>(zh) * p.d;
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      mbar_wait(&o_full[buf], (c >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t r[32];
        tmem_ld_32x32(trow + kAwOCol + buf * 128 + cc * 32, r);
        tmem_ld_wait();
        if (rvalid) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
#pragma unroll
          for (int j = 0; j < 32; j += 8) *reinterpret_cast<bf16x8*>(orow + c * 128 + cc * 32 + j) = pack8(v + j);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (elect_one()) mbar_arrive(&o_empty[buf]);
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static int make_attn_map(CUtensorMap* out, const void* ptr, int mn_major, int t, int d, long long ld, int heads, int batch) {
  // K-major (Q, K): (channel, token, head, sample), box 64 x 128;  MN-major (V): same dims, box 64 channels x 64 tokens
  const long long dims[4] = {d, t, heads, batch};
  const long long st[3] = {ld, heads > 1 ? static_cast<long long>(d) : ld, batch > 1 ? static_cast<long long>(t) * ld : ld};
  const int box[4] = {64, mn_major ? 64 : 128, 1, 1};
  return make_map4(out, ptr, dims, st, box);
}

}  // namespace ddpm

using namespace ddpm;

extern "C" int ddpm_attn_wide_supported(int t, int heads, int d) {
  return (t > 0 && t <= 256 && d >= 128 && d % 128 == 0 && heads > 0) ? 1 : 0;
}

extern "C" int ddpm_attn_wide_fwd(const void* qkv, long long ldqkv, void* o, long long ldo, void* probs, long long ldp,
                                  int b, int t, int heads, int d, float scale, void* stream) {
  DDPM_REQUIRE(qkv && o && b > 0, "ddpm_attn_wide_fwd: bad argument");
  DDPM_REQUIRE(ddpm_attn_wide_supported(t, heads, d), "ddpm_attn_wide_fwd: t=%d heads=%d d=%d is outside the fused kernel "
               "(t <= 256, d a multiple of 128)", t, heads, d);
  DDPM_REQUIRE(static_cast<long long>(heads) * b <= 65535, "ddpm_attn_wide_fwd: heads*batch exceeds gridDim.y");
  DDPM_REQUIRE(ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0, "ddpm_attn_wide_fwd: o must be 16-byte aligned");
  DDPM_REQUIRE(!probs || (ldp % 8 == 0 && ldp >= t && (reinterpret_cast<uintptr_t>(probs) & 15) == 0),
               "ddpm_attn_wide_fwd: probabilities need a 16-byte aligned row stride >= t (ldp=%lld)", ldp);
  const long long C = static_cast<long long>(heads) * d;
  const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(qkv);
  CUtensorMap tmQ, tmK, tmV;
  if (int e = make_attn_map(&tmQ, base, 0, t, d, ldqkv, heads, b)) return e;
  if (int e = make_attn_map(&tmK, base + C, 0, t, d, ldqkv, heads, b)) return e;
  if (int e = make_attn_map(&tmV, base + 2 * C, 1, t, d, ldqkv, heads, b)) return e;
  AttnWideParams p;
  p.t = t; p.d = d; p.heads = heads; p.scale = scale;
  p.o = static_cast<__nv_bfloat16*>(o); p.ldo = ldo;
  p.p = static_cast<__nv_bfloat16*>(probs);
  p.ldp = ldp; p.p_head = static_cast<long long>(t) * ldp; p.p_batch = static_cast<long long>(heads) * t * ldp;
  static bool attr_set = false;
  if (!attr_set) {
    DDPM_CUDA(cudaFuncSetAttribute(attn_wide_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAwSmem));
    attr_set = true;
  }
  dim3 grid((t + 127) / 128, heads * b);
  attn_wide_fwd_kernel<<<grid, kAwThreads, kAwSmem, static_cast<cudaStream_t>(stream)>>>(tmQ, tmK, tmV, p);
  return check_launch("attn_wide_fwd_kernel");
}
