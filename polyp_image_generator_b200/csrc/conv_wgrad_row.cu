// Row-resident weight gradient of the 3x3 stride-1 convolutions (W >= 16, W % 16 == 0) on tcgen05.
//
// Why: the generic wgrad kernel (conv_igemm.cu) gives one CTA one (tap, 128 cout, 128 cin) tile, so every dY tile is
// re-read once per tap and 64 KB of operands are fetched per 512 tensor-clocks (128 B/clk/SM) -- it is pinned on L2
// bandwidth at ~690 TFLOP/s in the training step.
//
// How: a CTA owns one KERNEL ROW r (taps (r,0), (r,1), (r,2)), one 128-cout x 128-cin weight tile and a split-K range
// of pixel segments.  A segment is `kw` consecutive pixels of `hb` consecutive image rows (hb = 1 for W >= 128, else
// hb * W ~ 128 so that narrow maps still feed K = 128 per stage).  Per segment it loads
//     A = dY[n, h, w0 : w0+kw, cout tile]                      (two 64-channel TMA boxes, MN-major: pixels = K)
//     B = X [n, h+r-1, w0-1 : w0+kw+1, cin tile]               (two boxes WITH a one-pixel halo on each side; TMA zero
//                                                               fills w = -1, w = W and rows outside the image)
// ONCE and issues the three taps from it by starting the SWIZZLE_128B B-descriptor 0, 1 or 2 rows (128 B) into the
// box -- the same shifted-descriptor trick as conv_halo.cu, here on an MN-major operand.  Three fp32 accumulators
// (3 x 128 TMEM columns) stay resident over the whole split-K range and are flushed with red.global.add.v4.f32.
// Operand traffic drops to 64 KB per 1536 tensor-clocks (42 B/clk/SM), a 3x cut.
#include "common.cuh"

#include <cstring>

#include "../../include/ddpm_b200.h"

namespace ddpm {

constexpr int kWrThreads = 192;   // warp0: TMA producer, warp1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int kWrMaxStages = 8;

struct WgradRowParams {
  int N, H, W;
  int kw;                  // pixels per segment row, multiple of 16
  int hb;                  // image rows per segment (K of one stage = hb * kw)
  int segs_per_row;        // W / kw
  int hsegs;               // H / hb
  int total_segs;          // N * hsegs * segs_per_row
  int segs_per_split;
  int cin_tiles, cout_tiles;
  int kb0;                 // 64-channel blocks in X source 0
  int Cin_total;
  int stages;
  uint32_t a_box_bytes;    // hb * kw * 128
  uint32_t b_box_bytes;    // hb * (kw + 2) * 128
  uint32_t b_box_stride;   // b_box_bytes rounded up to 1024
  uint32_t stage_bytes;
  float* dw;
  long long ldw;
  int noflush;             // timing experiment only (DDPM_WGRAD_NOFLUSH=1): skip the red.add epilogue
};

__global__ void __launch_bounds__(kWrThreads, 1)
conv_wgrad_row_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX0,
                      const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ WgradRowParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.stages) * p.stage_bytes);
  uint64_t* empty_bar = full_bar + kWrMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kWrMaxStages;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int r = blockIdx.x / p.cin_tiles;            // kernel row 0..2
  const int cit = blockIdx.x - r * p.cin_tiles;      // 128-channel cin tile
  const int cot = blockIdx.y;                        // 128-channel cout tile
  const int seg_begin = blockIdx.z * p.segs_per_split;
  const int seg_end = min(p.total_segs, seg_begin + p.segs_per_split);
  const int n_iters = seg_end - seg_begin;
  const int S = p.stages;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmX0);
    tma_prefetch_desc(&tmX1);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_base_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const uint32_t a_bytes = 2u * p.a_box_bytes;

  if (warp == 0) {
    const int cb = cit * 2;                                  // first 64-channel block of the cin tile
    const bool src0 = cb < p.kb0;
    const CUtensorMap* mx = src0 ? &tmX0 : &tmX1;
    const int cx = (src0 ? cb : cb - p.kb0) * 64;
    // segment coordinates, ring slot and phase advance incrementally: no per-iteration division in the issuing warps
    int sw = seg_begin % p.segs_per_row;
    int hs = (seg_begin / p.segs_per_row) % p.hsegs;
    int n = seg_begin / (p.segs_per_row * p.hsegs);
    int s = 0;
    uint32_t ph = 1;
    for (int it = 0; it < n_iters; ++it) {
      const int h = hs * p.hb;
      const int w0 = sw * p.kw;
      mbar_wait(&empty_bar[s], ph);
      if (elect_one()) {
        uint8_t* sa = smem + static_cast<size_t>(s) * p.stage_bytes;
        uint8_t* sb = sa + a_bytes;
        mbar_expect_tx(&full_bar[s], a_bytes + 2u * p.b_box_bytes);
        tma_load_4d(sa, &tmY, &full_bar[s], cot * 128, w0, h, n);
        tma_load_4d(sa + p.a_box_bytes, &tmY, &full_bar[s], cot * 128 + 64, w0, h, n);
        tma_load_4d(sb, mx, &full_bar[s], cx, w0 - 1, h + r - 1, n);
        tma_load_4d(sb + p.b_box_stride, mx, &full_bar[s], cx + 64, w0 - 1, h + r - 1, n);
      }
      __syncwarp();
      if (++sw == p.segs_per_row) {
        sw = 0;
        if (++hs == p.hsegs) { hs = 0; ++n; }
      }
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // whole warp runs the loop (warp-uniform descriptors stay in uniform registers); one elected lane issues
    constexpr uint32_t idesc = make_idesc_bf16(128, 128, true, true);
    const int ksteps = p.kw / 16;
    const uint64_t da_base = make_smem_desc_sw128(smem_u32(smem), p.a_box_bytes, 1024);
    const uint64_t db_base = make_smem_desc_sw128(smem_u32(smem) + a_bytes, p.b_box_stride, 1024);
    const uint64_t stage_step = p.stage_bytes >> 4;
    uint64_t da0 = da_base, db0 = db_base;
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < n_iters; ++it) {
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      const uint32_t first = it != 0 ? 1u : 0u;
      if (elect_one()) {
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          uint32_t acc = first;
          for (int hi = 0; hi < p.hb; ++hi) {
            // descriptor start offsets in 16-byte units: (rows * 128 B) >> 4 = rows * 8
            const uint64_t ao = static_cast<uint64_t>(hi * p.kw) * 8u;
            const uint64_t bo = static_cast<uint64_t>(hi * (p.kw + 2) + t) * 8u;
#pragma unroll 4
            for (int k = 0; k < ksteps; ++k) {
              umma_bf16(tmem_base + t * 128, da0 + ao + static_cast<uint64_t>(k * 128),
                        db0 + bo + static_cast<uint64_t>(k * 128), idesc, acc);
              acc = 1u;
            }
          }
        }
        umma_commit(&empty_bar[s]);
      }
      __syncwarp();
      da0 += stage_step;
      db0 += stage_step;
      if (++s == S) { s = 0; ph ^= 1; da0 = da_base; db0 = db_base; }
    }
    if (elect_one()) umma_commit(tmem_full_bar);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int co = cot * 128 + q * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (n_iters > 0 && !p.noflush) {
#pragma unroll 1
      for (int t = 0; t < 3; ++t) {
        float* row = p.dw + static_cast<long long>(co) * p.ldw + (r * 3 + t) * p.Cin_total + cit * 128;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * 128 + c * 32, v);
          tmem_ld_wait();
          float* dst = row + c * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j])),
                         "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                         : "memory");
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// returns 1 when the problem is not eligible (caller falls back to the generic kernel)
int launch_wgrad_row(const ::ddpm_wgrad_args* a, cudaStream_t stream) {
  if (env_int("DDPM_WGRAD_ROW", 1) == 0) return 1;
  if (a->ntaps != 9 || (a->src_n != 0 && a->src_n != a->n)) return 1;
  const int cin_total = a->c0 + a->c1;
  for (int t = 0; t < 9; ++t) {
    if (a->tap_dn[t] != 0 || a->tap_dh[t] != t / 3 - 1 || a->tap_dw[t] != t % 3 - 1 || a->tap_wk[t] != t * cin_total)
      return 1;
  }
  const int W = a->w;
  if (W < env_int("DDPM_WGRAD_ROW_MIN_W", 16) || W % 16) return 1;
  // pixels per segment row: the largest multiple of 16 up to 128 that divides W (128 at 128^2 / 256^2, 112 at 224^2)
  int kw = W < 128 ? W : 128;
  while (kw >= 16 && W % kw) kw -= 16;
  if (kw < 16 || (W >= 128 && kw < 64)) return 1;
  int hb = 1;
  if (W < 128) {
    hb = 128 / W;
    if (hb > a->h) hb = a->h;
    while (hb > 1 && a->h % hb) --hb;
  }
  if (a->cout % 128 || a->c0 % 128 || a->c1 % 128) return 1;

  WgradRowParams p;
  std::memset(&p, 0, sizeof(p));
  p.N = a->n; p.H = a->h; p.W = W;
  p.kw = kw;
  p.hb = hb;
  p.segs_per_row = W / kw;
  p.hsegs = a->h / hb;
  p.total_segs = a->n * p.hsegs * p.segs_per_row;
  p.cin_tiles = cin_total / 128;
  p.cout_tiles = a->cout / 128;
  p.kb0 = a->c0 / 64;
  p.Cin_total = cin_total;
  p.a_box_bytes = static_cast<uint32_t>(hb * kw) * 128u;
  p.b_box_bytes = static_cast<uint32_t>(hb * (kw + 2)) * 128u;
  p.b_box_stride = (p.b_box_bytes + 1023u) & ~1023u;
  if (p.a_box_bytes % 1024u) return 1;     // TMA destinations and descriptor bases stay 1024-byte aligned
  p.stage_bytes = 2u * p.a_box_bytes + 2u * p.b_box_stride;
  const size_t budget = 227 * 1024 - 1024 - 256;
  int stages = static_cast<int>(budget / p.stage_bytes);
  if (stages > kWrMaxStages) stages = kWrMaxStages;
  if (stages < 2) return 1;
  p.stages = stages;
  p.dw = a->dw; p.ldw = a->ldw;
  p.noflush = env_int("DDPM_WGRAD_NOFLUSH", 0);

  const long long out_tiles = 3LL * p.cin_tiles * p.cout_tiles;
  int splits = a->splits;
  if (splits <= 0) {
    long long want = kNumSMs / out_tiles;          // one CTA per SM (TMEM: 512 columns each)
    if (want < 1) want = 1;
    long long cap = p.total_segs / 4;              // >= 4 segments per CTA so the red.add epilogue is amortised
    if (cap < 1) cap = 1;
    splits = static_cast<int>(want < cap ? want : cap);
  }
  if (splits > p.total_segs) splits = p.total_segs;
  p.segs_per_split = (p.total_segs + splits - 1) / splits;
  splits = (p.total_segs + p.segs_per_split - 1) / p.segs_per_split;

  CUtensorMap my, mx0, mx1;
  if (int e = make_act_map(&my, a->dy, a->cout, a->ldy, a->n, a->h, W, kw, hb, 1)) return e;
  if (int e = make_act_map(&mx0, a->x0, a->c0, a->ld0, a->n, a->h, W, kw + 2, hb, 1)) return e;
  if (a->c1 > 0) {
    if (!a->x1) { set_last_error("ddpm_conv_wgrad: c1>0 but x1 is null"); return DDPM_ERR_INVALID; }
    if (int e = make_act_map(&mx1, a->x1, a->c1, a->ld1, a->n, a->h, W, kw + 2, hb, 1)) return e;
  } else {
    mx1 = mx0;
  }
  const size_t smem = static_cast<size_t>(stages) * p.stage_bytes + 256 + 1024;
  static size_t configured = 0;
  if (smem > configured) {
    DDPM_CUDA(cudaFuncSetAttribute(conv_wgrad_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid(3 * p.cin_tiles, p.cout_tiles, splits);
  conv_wgrad_row_kernel<<<grid, kWrThreads, smem, stream>>>(my, mx0, mx1, p);
  return check_launch("conv_wgrad_row_kernel");
}

}  // namespace ddpm
