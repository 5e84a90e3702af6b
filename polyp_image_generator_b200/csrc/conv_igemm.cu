// Implicit-GEMM convolution / linear kernels for sm_100a: TMA -> shared memory -> tcgen05.mma -> TMEM.
//
// Replaces what diffusers' nn.Conv2d / nn.Linear layers dispatch to cuDNN / cuBLAS on the reference path
// (UNet2DModel built at /root/reference/generator_model/PolypGeneratorModel.py:25-48, run at
//  /root/reference/generator_model/train_from_scratch.py:100,103) -- SURVEY.md §2.3 rows K1/K2/K3/K7.
//
// Data layout: activations NHWC bf16 (channel stride given, so channel-slices of wider tensors work),
// weights bf16 [Cout][tap][Cin] ("K-major": the reduction index is contiguous), accumulation fp32 in TMEM.
//
//  * conv_gemm_kernel (fprop and dgrad):  OUT[pixel, co] = sum_{tap, ci} X[pixel + tap_offset, ci] * Wt[co][tap][ci]
//      A tile = 128 output pixels x 64 channels, fetched per tap as ONE shifted 4-D TMA box over (C, W, H, N);
//      out-of-bounds rows/columns are zero-filled by TMA, which *is* the conv padding.  The channel dimension
//      may be split over two source tensors (torch.cat([h, skip], 1) is never materialised).  Taps carry an
//      arbitrary (dn, dh, dw) offset, so stride-2 convs (space-to-depth phases stacked along N), dgrad
//      (flipped taps, transposed weights) and plain linears (1 tap, H = 1) all use the same kernel.
//      Epilogue fuses bias, the per-(sample, channel) time-embedding add and the residual add.
//  * conv_wgrad_kernel:  dW[co][tap][ci] += sum_{pixel} dY[pixel, co] * X[pixel + tap_offset, ci]
//      both operands are "MN-major" straight out of NHWC (pixels are the reduction index), split-K over
//      pixel tiles with fp32 red.global.add into the gradient buffer.
#include "common.cuh"
#include "conv_epilogue.cuh"

#include <mutex>
#include <unordered_map>
#include <string>
#include <cstring>
#include <cstdlib>

#include "../../include/ddpm_b200.h"

namespace ddpm {

constexpr int kBlockM = 128;   // output pixels per tile (UMMA M)
constexpr int kBlockK = 64;    // bf16 channels per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kMaxTaps = DDPM_MAX_TAPS;
constexpr int kGemmThreads = 192;  // warp0: TMA producer, warp1: MMA issuer + TMEM owner, warps 2-5: epilogue

struct TapTable {
  int ntaps;
  int dn[kMaxTaps], dh[kMaxTaps], dw[kMaxTaps], wk[kMaxTaps];
};

struct GemmParams {
  int N, H, W;            // pixel grid
  int kb0, kb1;           // 64-channel blocks from source 0 / 1
  int wb, hb, nb;         // TMA box (pixels) = M tile shape
  int tiles_w, tiles_h, tiles_n;
  uint32_t a_bytes;       // bytes one A box delivers (wb*hb*nb*128)
  int iters_per_split;    // split-K over blockIdx.z: (tap, k-block) iterations per CTA; 0 = whole reduction
  float* splitk_ws;       // fp32 [pixels][Cout] partial-sum workspace (zero-filled) when splitting
  EpiParams epi;
  TapTable taps;
};

template <int BLOCK_N, int STAGES>
struct GemmSmem {
  static constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KB
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarOffset + 256 + 1024;  // barriers + alignment slack
};

// ---------------------------------------------------------------------------------------------
// fprop / dgrad / linear
// ---------------------------------------------------------------------------------------------
template <int BLOCK_N, int STAGES, bool GN>
__global__ void __launch_bounds__(kGemmThreads, (BLOCK_N == 128 ? 2 : 1))
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ GemmParams p) {
  using L = GemmSmem<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates: blockIdx.x = cout tile (fastest, so CTAs sharing an A tile run together), y = pixel tile
  const int nt = blockIdx.x;
  int mt = blockIdx.y;
  const int tw = mt % p.tiles_w;
  mt /= p.tiles_w;
  const int th = mt % p.tiles_h;
  const int tn = mt / p.tiles_h;
  const int w0 = tw * p.wb, h0 = th * p.hb, n0 = tn * p.nb;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.kb1 > 0) tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_base_slot, BLOCK_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  const int kbt = p.kb0 + p.kb1;
  const int all_iters = p.taps.ntaps * kbt;
  // split-K: this CTA reduces the (tap, k-block) iterations [g_begin, g_begin + total_iters)
  const int g_begin = p.iters_per_split > 0 ? static_cast<int>(blockIdx.z) * p.iters_per_split : 0;
  const int total_iters = p.iters_per_split > 0 ? max(0, min(all_iters - g_begin, p.iters_per_split)) : all_iters;

  // Single-warp roles are issue-bound unless their loops are lean (profiles/r1_halo_issue.md): ring slot / phase and
  // the (tap, k-block) pair advance incrementally -- no per-iteration division or modulo.
  if (warp == 0) {
    int t = g_begin / kbt, kb = g_begin - t * kbt;
    uint32_t s = 0, ph = 1;     // ph: parity to wait for on empty_bar (starts "free")
    for (int it = 0; it < total_iters; ++it) {
      const int cw = w0 + p.taps.dw[t], ch = h0 + p.taps.dh[t], cn = n0 + p.taps.dn[t], wk = p.taps.wk[t];
      mbar_wait(&empty_bar[s], ph);
      if (elect_one()) {
        uint8_t* sa = smem + s * L::kStageBytes;
        uint8_t* sb = sa + L::kABytes;
        mbar_expect_tx(&full_bar[s], p.a_bytes + L::kBBytes);
        if (kb < p.kb0)
          tma_load_4d(sa, &tmA0, &full_bar[s], kb * kBlockK, cw, ch, cn);
        else
          tma_load_4d(sa, &tmA1, &full_bar[s], (kb - p.kb0) * kBlockK, cw, ch, cn);
        tma_load_2d(sb, &tmB, &full_bar[s], wk + kb * kBlockK, nt * BLOCK_N);
      }
      __syncwarp();
      if (++kb == kbt) { kb = 0; ++t; }
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // whole warp runs the loop (warp-uniform descriptors stay in uniform registers); one elected lane issues
    constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N, false, false);
    const uint64_t da_base = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    uint64_t da0 = da_base;
    uint32_t s = 0, ph = 0;
    for (int it = 0; it < total_iters; ++it) {
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      const uint64_t db0 = da0 + (L::kABytes >> 4);
      const uint32_t first = it != 0 ? 1u : 0u;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k)
          umma_bf16(tmem_base, da0 + static_cast<uint64_t>((k * kUmmaK * 2) >> 4),
                    db0 + static_cast<uint64_t>((k * kUmmaK * 2) >> 4), idesc, k == 0 ? first : 1u);
        umma_commit(&empty_bar[s]);
      }
      __syncwarp();
      da0 += L::kStageBytes >> 4;
      if (++s == STAGES) { s = 0; ph ^= 1; da0 = da_base; }
    }
    if (elect_one()) umma_commit(tmem_full_bar);
    __syncwarp();
  } else {
    // epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31; one thread per output pixel (tile row)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int rows_per_img = p.wb * p.hb;
    const int ni = row / rows_per_img;
    const int rem = row - ni * rows_per_img;
    const int hi = rem / p.wb;
    const int wi = rem - hi * p.wb;
    const int n = n0 + ni, h = h0 + hi, w = w0 + wi;
    const bool valid = (ni < p.nb) && (n < p.N) && (h < p.H) && (w < p.W);
    const long long pix = (static_cast<long long>(n) * p.H + h) * p.W + w;
    const int col0 = nt * BLOCK_N;

    EpiX xcur, xnext;
    const bool pref = GN || (p.epi.res != nullptr && p.splitk_ws == nullptr);   // side input: GroupNorm x or residual
    if (pref) epi_load_x(p.epi, valid, pix, col0, xcur);   // in flight while the mainloop runs
    else {
#pragma unroll
      for (int j = 0; j < 16; ++j) xcur.w[j] = xnext.w[j] = 0u;
    }
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int n_warp = (GN && p.epi.gsums) ? epi_warp_sample(valid, n) : -1;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 32; ++c) {
      if (pref && c + 1 < BLOCK_N / 32) epi_load_x(p.epi, valid, pix, col0 + (c + 1) * 32, xnext);
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, r);
      tmem_ld_wait();
      if (!GN && p.splitk_ws != nullptr) {
        // split-K: add the raw partial sums to the fp32 workspace; bias / residual / bf16 happen in the finalize pass
        if (valid && col0 + c * 32 < p.epi.Cout && total_iters > 0) {
          float* dst = p.splitk_ws + pix * p.epi.Cout + col0 + c * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(r[j])),
                         "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3]))
                         : "memory");
        }
        continue;
      }
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      float t1 = 0.f, t2 = 0.f;
      epi_chunk<GN, true>(p.epi, v, valid, n, pix, col0 + c * 32, lane, t1, t2, xcur);
      if (GN) epi_flush_sums(p.epi, n_warp, col0 + c * 32, lane, t1, t2);
      xcur = xnext;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BLOCK_N);
}

// ---------------------------------------------------------------------------------------------
// fprop / dgrad / linear, PERSISTENT variant (problems with many output tiles)
// ---------------------------------------------------------------------------------------------
// The kernel above gives every output tile its own CTA: barrier init, TMEM allocation, descriptor prefetch and the
// pipeline fill are paid per tile, and the epilogue of a tile overlaps nothing of its own CTA.  For the short-K GEMMs of
// the UNet -- the 1x1 shortcut convs (2-6 k-blocks), conv_in / conv_out's dgrad (ONE k-block), the attention projections
// -- that overhead IS the kernel: they are memory-bound by bytes (0.12 ms for 256->128 1x1 @128^2, batch 64) and ran at
// 0.19 ms.  Here one CTA per SM walks a static round-robin of tiles with the TMA ring running ahead across tile
// boundaries and two TMEM accumulators, so tile i's epilogue (8 warps) runs under tile i+1's loads and MMAs.
constexpr int kPersistThreads = 320;   // warp0: TMA producer, warp1: MMA issuer + TMEM owner, warps 2-9: epilogue
constexpr int kPersistMaxStages = 8;

struct PersistParams {
  GemmParams g;
  int n_tiles, total_tiles, stages;
};

template <int BLOCK_N, bool GN>
__global__ void __launch_bounds__(kPersistThreads, 1)
conv_gemm_persist_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                         const __grid_constant__ CUtensorMap tmB, const __grid_constant__ PersistParams pp) {
  const GemmParams& p = pp.g;
  constexpr int kABytes = kBlockM * kBlockK * 2;
  constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  constexpr int kStageBytes = kABytes + kBBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = pp.stages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * kStageBytes);
  uint64_t* empty_bar = full_bar + kPersistMaxStages;
  uint64_t* tmem_full = empty_bar + kPersistMaxStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;                  // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.kb1 > 0) tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);     // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_base_slot, 2 * BLOCK_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  const int kbt = p.kb0 + p.kb1;
  const int tiles_hw = p.tiles_w * p.tiles_h;

  if (warp == 0) {
    uint32_t s = 0, ph = 1;
    for (int tile = blockIdx.x; tile < pp.total_tiles; tile += gridDim.x) {
      const int nt = tile % pp.n_tiles;
      const int mt = tile / pp.n_tiles;
      const int tn = mt / tiles_hw;
      const int r2 = mt - tn * tiles_hw;
      const int th = r2 / p.tiles_w;
      const int w0 = (r2 - th * p.tiles_w) * p.wb, h0 = th * p.hb, n0 = tn * p.nb;
      for (int t = 0; t < p.taps.ntaps; ++t) {
        const int cw = w0 + p.taps.dw[t], ch = h0 + p.taps.dh[t], cn = n0 + p.taps.dn[t], wk = p.taps.wk[t];
        for (int kb = 0; kb < kbt; ++kb) {
          mbar_wait(&empty_bar[s], ph);
          if (elect_one()) {
            uint8_t* sa = smem + s * kStageBytes;
            uint8_t* sb = sa + kABytes;
            mbar_expect_tx(&full_bar[s], p.a_bytes + kBBytes);
            if (kb < p.kb0)
              tma_load_4d(sa, &tmA0, &full_bar[s], kb * kBlockK, cw, ch, cn);
            else
              tma_load_4d(sa, &tmA1, &full_bar[s], (kb - p.kb0) * kBlockK, cw, ch, cn);
            tma_load_2d(sb, &tmB, &full_bar[s], wk + kb * kBlockK, nt * BLOCK_N);
          }
          __syncwarp();
          if (++s == static_cast<uint32_t>(S)) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N, false, false);
    const uint64_t da_base = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const int iters = p.taps.ntaps * kbt;
    uint64_t da0 = da_base;
    uint32_t s = 0, ph = 0, acc = 0, aph = 0;
    for (int tile = blockIdx.x; tile < pp.total_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], aph ^ 1);
      tc_fence_after();
      const uint32_t d0 = tmem_base + acc * BLOCK_N;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint64_t db0 = da0 + (kABytes >> 4);
        const uint32_t first = it != 0 ? 1u : 0u;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            umma_bf16(d0, da0 + static_cast<uint64_t>((k * kUmmaK * 2) >> 4),
                      db0 + static_cast<uint64_t>((k * kUmmaK * 2) >> 4), idesc, k == 0 ? first : 1u);
          umma_commit(&empty_bar[s]);
        }
        __syncwarp();
        da0 += kStageBytes >> 4;
        if (++s == static_cast<uint32_t>(S)) { s = 0; ph ^= 1; da0 = da_base; }
      }
      if (elect_one()) umma_commit(&tmem_full[acc]);
      __syncwarp();
      acc ^= 1;
      aph ^= (acc == 0);
    }
  } else {
    // epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31; two warps per lane quarter split the columns
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int kChunks = BLOCK_N / 64;          // 32-column chunks per warp
    const int row = q * 32 + lane;
    const int rows_per_img = p.wb * p.hb;
    const int ni = row / rows_per_img;
    const int rem = row - ni * rows_per_img;
    const int hi = rem / p.wb;
    const int wi = rem - hi * p.wb;
    uint32_t acc = 0, aph = 0;
    for (int tile = blockIdx.x; tile < pp.total_tiles; tile += gridDim.x) {
      const int nt = tile % pp.n_tiles;
      const int mt = tile / pp.n_tiles;
      const int tn = mt / tiles_hw;
      const int r2 = mt - tn * tiles_hw;
      const int th = r2 / p.tiles_w;
      const int n = tn * p.nb + ni, h = th * p.hb + hi, w = (r2 - th * p.tiles_w) * p.wb + wi;
      const bool valid = (ni < p.nb) && (n < p.N) && (h < p.H) && (w < p.W);
      const long long pix = (static_cast<long long>(n) * p.H + h) * p.W + w;
      const int col0 = nt * BLOCK_N + half * (BLOCK_N / 2);
      EpiX xcur, xnext;
      const bool pref = GN || p.epi.res != nullptr;
      if (pref) epi_load_x(p.epi, valid, pix, col0, xcur);
      else {
#pragma unroll
        for (int j = 0; j < 16; ++j) xcur.w[j] = xnext.w[j] = 0u;
      }
      mbar_wait(&tmem_full[acc], aph);
      tc_fence_after();
      const int n_warp = (GN && p.epi.gsums) ? epi_warp_sample(valid, n) : -1;
#pragma unroll 1
      for (int c = 0; c < kChunks; ++c) {
        if (pref && c + 1 < kChunks) epi_load_x(p.epi, valid, pix, col0 + (c + 1) * 32, xnext);
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N + half * (BLOCK_N / 2) + c * 32, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        float t1 = 0.f, t2 = 0.f;
        epi_chunk<GN, true>(p.epi, v, valid, n, pix, col0 + c * 32, lane, t1, t2, xcur);
        if (GN) epi_flush_sums(p.epi, n_warp, col0 + c * 32, lane, t1, t2);
        xcur = xnext;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      aph ^= (acc == 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * BLOCK_N);
}

// ---------------------------------------------------------------------------------------------
// wgrad
// ---------------------------------------------------------------------------------------------
struct WgradParams {
  int N, H, W;
  int kb0, kb1;           // 64-channel blocks of X source 0 / 1
  int Cout;               // multiple of 64
  int wb, hb, nb, tiles_w, tiles_h, tiles_n;
  uint32_t box_bytes;     // bytes of one 64-channel pixel box
  int box_rows;           // wb*hb*nb
  int ptiles;             // total pixel tiles
  int ptiles_per_split;
  int cin_tiles;          // ceil((kb0+kb1)*64 / BLOCK_N)
  float* dw;              // fp32 [Cout][ldw]
  long long ldw;
  int atomic;             // 1: red.add (split-K or accumulate), 0: plain store
  uint32_t lbo, sbo;      // MN-major descriptor strides (runtime so the bring-up test can probe them)
  TapTable taps;
};

template <int BLOCK_N, int STAGES>
struct WgradSmem {
  static constexpr int kABytes = 2 * kBlockM * 128;             // two 64-cout boxes of 128 pixel rows
  static constexpr int kBBytes = (BLOCK_N / 64) * kBlockM * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarOffset + 256 + 1024;
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX0,
                  const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ WgradParams p) {
  using L = WgradSmem<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tap = blockIdx.x / p.cin_tiles;
  const int cit = blockIdx.x - tap * p.cin_tiles;
  const int cot = blockIdx.y;
  const int pt_begin = blockIdx.z * p.ptiles_per_split;
  const int pt_end = min(p.ptiles, pt_begin + p.ptiles_per_split);
  const int kbt = p.kb0 + p.kb1;
  const int cb0 = cit * (BLOCK_N / 64);                 // first 64-channel block of this cin tile
  const int ncb = min(BLOCK_N / 64, kbt - cb0);         // valid 64-channel blocks in this tile

  // Pixel rows the TMA boxes never write (box_rows < 128) or channel blocks past Cin must read as zero.
  if (p.box_rows < kBlockM || ncb < BLOCK_N / 64 || (cot * 2 + 1) * 64 >= p.Cout) {
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < STAGES * L::kStageBytes / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
  }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmX0);
    if (p.kb1 > 0) tma_prefetch_desc(&tmX1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_base_slot, BLOCK_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  const int n_iters = pt_end - pt_begin;
  const int nyb = ((cot * 2 + 1) * 64 < p.Cout) ? 2 : 1;  // 64-cout boxes actually present

  if (warp == 0) {
    const int dn = p.taps.dn[tap], dh = p.taps.dh[tap], dw = p.taps.dw[tap];
    // pixel-tile coordinates advance incrementally (no per-iteration division: the producer warp is issue-bound)
    int tw = pt_begin % p.tiles_w;
    int th = (pt_begin / p.tiles_w) % p.tiles_h;
    int tn = pt_begin / (p.tiles_w * p.tiles_h);
    uint32_t s = 0, ph = 1;
    for (int it = 0; it < n_iters; ++it) {
      const int w0 = tw * p.wb, h0 = th * p.hb, n0 = tn * p.nb;
      mbar_wait(&empty_bar[s], ph);
      if (elect_one()) {
        uint8_t* sa = smem + s * L::kStageBytes;
        uint8_t* sb = sa + L::kABytes;
        mbar_expect_tx(&full_bar[s], p.box_bytes * (nyb + ncb));
        for (int b = 0; b < nyb; ++b)
          tma_load_4d(sa + b * (kBlockM * 128), &tmY, &full_bar[s], (cot * 2 + b) * 64, w0, h0, n0);
        for (int b = 0; b < ncb; ++b) {
          const int cb = cb0 + b;
          if (cb < p.kb0)
            tma_load_4d(sb + b * (kBlockM * 128), &tmX0, &full_bar[s], cb * 64, w0 + dw, h0 + dh, n0 + dn);
          else
            tma_load_4d(sb + b * (kBlockM * 128), &tmX1, &full_bar[s], (cb - p.kb0) * 64, w0 + dw, h0 + dh,
                        n0 + dn);
        }
      }
      __syncwarp();
      if (++tw == p.tiles_w) {
        tw = 0;
        if (++th == p.tiles_h) { th = 0; ++tn; }
      }
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N, true, true);
    const uint64_t da_base = make_smem_desc_sw128(smem_u32(smem), p.lbo, p.sbo);
    uint64_t da0 = da_base;
    uint32_t s = 0, ph = 0;
    for (int it = 0; it < n_iters; ++it) {
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      const uint64_t db0 = da0 + (L::kABytes >> 4);
      const uint32_t first = it != 0 ? 1u : 0u;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kBlockM / kUmmaK; ++k)   // 128 pixel rows = 8 MMAs of K=16
          umma_bf16(tmem_base, da0 + static_cast<uint64_t>((k * kUmmaK * 128) >> 4),
                    db0 + static_cast<uint64_t>((k * kUmmaK * 128) >> 4), idesc, k == 0 ? first : 1u);
        umma_commit(&empty_bar[s]);
      }
      __syncwarp();
      da0 += L::kStageBytes >> 4;
      if (++s == STAGES) { s = 0; ph ^= 1; da0 = da_base; }
    }
    if (elect_one()) umma_commit(tmem_full_bar);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;         // cout index within tile
    const int co = cot * kBlockM + row;
    const int cin_total = kbt * 64;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, r);
      tmem_ld_wait();
      const int ci = cit * BLOCK_N + c * 32;
      if (co < p.Cout && ci < cin_total && n_iters > 0) {
        float* dst = p.dw + static_cast<long long>(co) * p.ldw + p.taps.wk[tap] + ci;
        if (p.atomic) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(r[j])),
                         "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])),
                         "f"(__uint_as_float(r[j + 3]))
                         : "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                              __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BLOCK_N);
}

// ---------------------------------------------------------------------------------------------
// host side: tensor-map construction (cached) and launchers
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  uint64_t d[4];
  uint64_t ld;
  uint64_t st[3];
  uint32_t box[4];
  bool operator==(const MapKey& o) const { return std::memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return static_cast<size_t>(h);
  }
};
static std::mutex g_map_mu;
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

// NHWC bf16 activation view [n][h][w][c] with pixel stride ld (elements); box = (64, wb, hb, nb)
int make_act_map(CUtensorMap* out, const void* ptr, int c, long long ld, int n, int h, int w, int wb, int hb,
                        int nb) {
  MapKey k;
  std::memset(&k, 0, sizeof(k));
  k.ptr = ptr;
  k.d[0] = c; k.d[1] = w; k.d[2] = h; k.d[3] = n;
  k.ld = ld;
  k.box[0] = 64; k.box[1] = wb; k.box[2] = hb; k.box[3] = nb;
  {
    std::lock_guard<std::mutex> g(g_map_mu);
    auto it = g_maps.find(k);
    if (it != g_maps.end()) { *out = it->second; return DDPM_OK; }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_last_error("cuTensorMapEncodeTiled entry point unavailable"); return DDPM_ERR_CUDA; }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld % 8) || (c % 8)) {
    set_last_error("activation view not 16-byte aligned (ptr=%p ld=%lld c=%d)", ptr, ld, c);
    return DDPM_ERR_INVALID;
  }
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * w, (cuuint64_t)ld * 2 * w * h};
  cuuint32_t box[4] = {64, (cuuint32_t)wb, (cuuint32_t)hb, (cuuint32_t)nb};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(act c=%d w=%d h=%d n=%d ld=%lld box=%d,%d,%d) failed: %d", c, w, h, n, ld,
                   wb, hb, nb, (int)r);
    return DDPM_ERR_CUDA;
  }
  std::lock_guard<std::mutex> g(g_map_mu);
  if (g_maps.size() > 65536) g_maps.clear();
  g_maps.emplace(k, *out);
  return DDPM_OK;
}

// K-major weight matrix [rows][ld] bf16; box = (64, box_rows)
int make_wgt_map(CUtensorMap* out, const void* ptr, long long k_total, long long ld, int rows, int box_rows) {
  MapKey k;
  std::memset(&k, 0, sizeof(k));
  k.ptr = ptr;
  k.d[0] = k_total; k.d[1] = rows; k.d[2] = 0; k.d[3] = 0xB;
  k.ld = ld;
  k.box[0] = 64; k.box[1] = box_rows;
  {
    std::lock_guard<std::mutex> g(g_map_mu);
    auto it = g_maps.find(k);
    if (it != g_maps.end()) { *out = it->second; return DDPM_OK; }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_last_error("cuTensorMapEncodeTiled entry point unavailable"); return DDPM_ERR_CUDA; }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld % 8)) {
    set_last_error("weight matrix not 16-byte aligned (ptr=%p ld=%lld)", ptr, ld);
    return DDPM_ERR_INVALID;
  }
  cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(wgt k=%lld rows=%d ld=%lld) failed: %d", k_total, rows, ld, (int)r);
    return DDPM_ERR_CUDA;
  }
  std::lock_guard<std::mutex> g(g_map_mu);
  g_maps.emplace(k, *out);
  return DDPM_OK;
}

// generic 4-D bf16 view: dims[0] contiguous, dims[1..3] with element strides st[0..2]; SWIZZLE_128B, box[0] = 64
int make_map4(CUtensorMap* out, const void* ptr, const long long dims_[4], const long long st_[3], const int box_[4]) {
  MapKey k;
  std::memset(&k, 0, sizeof(k));
  k.ptr = ptr;
  k.ld = 0x4D41503400000000ull;   // tag: generic map
  for (int i = 0; i < 4; ++i) { k.d[i] = dims_[i]; k.box[i] = box_[i]; }
  for (int i = 0; i < 3; ++i) k.st[i] = st_[i];
  {
    std::lock_guard<std::mutex> g(g_map_mu);
    auto it = g_maps.find(k);
    if (it != g_maps.end()) { *out = it->second; return DDPM_OK; }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_last_error("cuTensorMapEncodeTiled entry point unavailable"); return DDPM_ERR_CUDA; }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (st_[0] % 8) || (st_[1] % 8) || (st_[2] % 8)) {
    set_last_error("matrix view not 16-byte aligned (ptr=%p strides=%lld,%lld,%lld)", ptr, st_[0], st_[1], st_[2]);
    return DDPM_ERR_INVALID;
  }
  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4], es[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) { dims[i] = (cuuint64_t)dims_[i]; box[i] = (cuuint32_t)box_[i]; }
  for (int i = 0; i < 3; ++i) strides[i] = (cuuint64_t)st_[i] * 2;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(dims=%lld,%lld,%lld,%lld strides=%lld,%lld,%lld box=%d,%d,%d,%d) failed: %d",
                   dims_[0], dims_[1], dims_[2], dims_[3], st_[0], st_[1], st_[2], box_[0], box_[1], box_[2], box_[3],
                   (int)r);
    return DDPM_ERR_CUDA;
  }
  std::lock_guard<std::mutex> g(g_map_mu);
  if (g_maps.size() > 65536) g_maps.clear();
  g_maps.emplace(k, *out);
  return DDPM_OK;
}

static void choose_box(int n, int h, int w, int* wb, int* hb, int* nb) {
  *wb = w < kBlockM ? w : kBlockM;
  int rem = kBlockM / *wb;
  if (rem < 1) rem = 1;
  *hb = h < rem ? h : rem;
  rem = kBlockM / (*wb * *hb);
  if (rem < 1) rem = 1;
  *nb = n < rem ? n : rem;
}

static int fill_taps(TapTable* t, const ddpm_conv_args* a) {
  if (a->ntaps < 1 || a->ntaps > kMaxTaps) { set_last_error("ntaps=%d out of range", a->ntaps); return DDPM_ERR_INVALID; }
  t->ntaps = a->ntaps;
  for (int i = 0; i < a->ntaps; ++i) {
    t->dn[i] = a->tap_dn[i]; t->dh[i] = a->tap_dh[i]; t->dw[i] = a->tap_dw[i]; t->wk[i] = a->tap_wk[i];
  }
  return DDPM_OK;
}

int fill_epilogue(EpiParams* e, const ::ddpm_conv_args* a) {
  std::memset(e, 0, sizeof(*e));
  e->out = static_cast<__nv_bfloat16*>(a->out);
  e->out_f32 = static_cast<float*>(a->out_f32);
  e->ldo = a->ldo;
  e->bias = a->bias; e->temb = a->temb; e->ld_temb = a->ld_temb;
  e->res = static_cast<const __nv_bfloat16*>(a->res); e->ldr = a->ldr;
  e->Cout = a->cout;
  if (a->gn_sums != nullptr) {
    DDPM_REQUIRE(a->gn_x0 && a->gn_coef, "ddpm_conv_gemm: incomplete GroupNorm-backward fusion arguments");
    DDPM_REQUIRE(a->gn_c0 % 32 == 0 && a->gn_c0 > 0 && a->gn_c0 <= a->cout && (a->gn_c0 == a->cout || a->gn_x1) &&
                     a->gn_ld0 % 8 == 0 && a->gn_ld1 % 8 == 0 && a->out_f32 == nullptr,
                 "ddpm_conv_gemm: unsupported GroupNorm-backward fusion layout (gn_c0=%d cout=%d)", a->gn_c0, a->cout);
    e->gx0 = static_cast<const __nv_bfloat16*>(a->gn_x0);
    e->gx1 = static_cast<const __nv_bfloat16*>(a->gn_x1);
    e->gld0 = a->gn_ld0; e->gld1 = a->gn_ld1; e->gc0 = a->gn_c0;
    e->gcoef = a->gn_coef;
    e->gsums = a->gn_sums;
    e->gsilu = a->gn_silu;
  }
  if (a->out_csum != nullptr) {
    DDPM_REQUIRE(a->gn_sums == nullptr && a->out_f32 == nullptr,
                 "ddpm_conv_gemm: out_csum excludes gn_sums and fp32 output");
    e->gsums = a->out_csum;
    e->gstats = 1;
  }
  if (a->split_io) {
    DDPM_REQUIRE(a->gn_sums == nullptr && a->out_csum == nullptr && a->out_f32 == nullptr && a->out != nullptr &&
                     a->ldo >= 2LL * a->cout && (a->res == nullptr || a->ldr >= 2LL * a->cout) && a->cout % 32 == 0,
                 "ddpm_conv_gemm: split_io needs a bf16 split output (ldo >= 2*cout) and excludes the GroupNorm fusions");
    e->split = 1;
  }
  auto ok32 = [](const void* ptr, long long ld_elems) {
    return ptr == nullptr || ((reinterpret_cast<uintptr_t>(ptr) & 31) == 0 && (ld_elems * 2) % 32 == 0);
  };
  e->wide = (ok32(e->out, e->ldo) && ok32(e->res, e->ldr) && ok32(e->gx0, e->gld0) && ok32(e->gx1, e->gld1)) ? 1 : 0;
  if (env_int("DDPM_EPI_WIDE", 1) == 0) e->wide = 0;
  return DDPM_OK;
}

// split-K finalize: out[pix][c] = bf16(ws[pix][c] + bias[c] + temb[n][c] + res[pix][c]); 8 channels per thread
__global__ void __launch_bounds__(256)
splitk_finalize_kernel(const float* __restrict__ ws, EpiParams e, long long pixels, long long pix_per_img) {
  const int V = e.Cout / 8;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= pixels * V) return;
  const long long pix = i / V;
  const int col = static_cast<int>(i - pix * V) * 8;
  const float4 a = *reinterpret_cast<const float4*>(ws + pix * e.Cout + col);
  const float4 b = *reinterpret_cast<const float4*>(ws + pix * e.Cout + col + 4);
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  if (e.bias) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += e.bias[col + j];
  }
  if (e.temb) {
    const float* t = e.temb + (pix / pix_per_img) * e.ld_temb + col;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += t[j];
  }
  if (e.res) {
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(e.res + pix * e.ldr + col), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += f[j];
    if (e.split) {
      unpack8(*reinterpret_cast<const bf16x8*>(e.res + pix * e.ldr + e.Cout + col), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += f[j];
    }
  }
  const bf16x8 hi = pack8(v);
  *reinterpret_cast<bf16x8*>(e.out + pix * e.ldo + col) = hi;
  if (e.split) {
    float h[8];
    unpack8(hi, h);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] -= h[j];
    *reinterpret_cast<bf16x8*>(e.out + pix * e.ldo + e.Cout + col) = pack8(v);
  }
}

// Split-K plan for a problem with `ctas` output tiles and `iters` (tap, k-block) iterations: engage idle SMs when the
// tile grid covers less than half the chip (low-resolution layers: 4x4 .. 16x16 maps at the bottom of the UNet).
static int splitk_plan(long long ctas, int iters) {
  if (env_int("DDPM_SPLITK", 1) == 0 || ctas > kNumSMs / 2 || iters < 16) return 1;
  long long s = kNumSMs / ctas;
  if (s > iters / 4) s = iters / 4;
  if (s > 16) s = 16;
  return s < 2 ? 1 : static_cast<int>(s);
}

template <int BLOCK_N, int STAGES, bool GN>
static int launch_gemm(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const GemmParams& p,
                       cudaStream_t stream) {
  using L = GemmSmem<BLOCK_N, STAGES>;
  static bool configured = false;
  auto kern = conv_gemm_kernel<BLOCK_N, STAGES, GN>;
  if (!configured) {
    DDPM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured = true;
  }
  const int all_iters = p.taps.ntaps * (p.kb0 + p.kb1);
  const int splits = p.iters_per_split > 0 ? (all_iters + p.iters_per_split - 1) / p.iters_per_split : 1;
  dim3 grid((p.epi.Cout + BLOCK_N - 1) / BLOCK_N, p.tiles_w * p.tiles_h * p.tiles_n, splits);
  kern<<<grid, kGemmThreads, L::kTotal, stream>>>(a0, a1, b, p);
  return check_launch("conv_gemm_kernel");
}

template <int BLOCK_N, bool GN>
static int launch_gemm_persist(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const GemmParams& p,
                               cudaStream_t stream) {
  PersistParams pp;
  pp.g = p;
  pp.n_tiles = (p.epi.Cout + BLOCK_N - 1) / BLOCK_N;
  pp.total_tiles = pp.n_tiles * p.tiles_w * p.tiles_h * p.tiles_n;
  const int stage_bytes = kBlockM * kBlockK * 2 + BLOCK_N * kBlockK * 2;
  int stages = (227 * 1024 - 1024 - 512) / stage_bytes;
  if (stages > kPersistMaxStages) stages = kPersistMaxStages;
  pp.stages = stages;
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + 512 + 1024;
  auto kern = conv_gemm_persist_kernel<BLOCK_N, GN>;
  static bool configured = false;
  if (!configured) {
    DDPM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int grid = pp.total_tiles < kNumSMs ? pp.total_tiles : kNumSMs;
  kern<<<grid, kPersistThreads, smem, stream>>>(a0, a1, b, pp);
  return check_launch("conv_gemm_persist_kernel");
}

template <int BLOCK_N, int STAGES>
static int launch_wgrad(const CUtensorMap& y, const CUtensorMap& x0, const CUtensorMap& x1, const WgradParams& p,
                        int splits, cudaStream_t stream) {
  using L = WgradSmem<BLOCK_N, STAGES>;
  static bool configured = false;
  auto kern = conv_wgrad_kernel<BLOCK_N, STAGES>;
  if (!configured) {
    DDPM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured = true;
  }
  dim3 grid(p.cin_tiles * p.taps.ntaps, (p.Cout + kBlockM - 1) / kBlockM, splits);
  kern<<<grid, kGemmThreads, L::kTotal, stream>>>(y, x0, x1, p);
  return check_launch("conv_wgrad_kernel");
}

int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

}  // namespace ddpm

using namespace ddpm;

extern "C" int ddpm_conv_gemm(const ddpm_conv_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DDPM_REQUIRE(a && a->x0 && a->wgt && (a->out || a->out_f32), "ddpm_conv_gemm: null pointer argument");
  DDPM_REQUIRE(a->c0 > 0 && a->c0 % 64 == 0 && a->c1 >= 0 && a->c1 % 64 == 0,
               "ddpm_conv_gemm: channel counts must be multiples of 64 (c0=%d c1=%d)", a->c0, a->c1);
  DDPM_REQUIRE(a->c1 == 0 || a->x1, "ddpm_conv_gemm: c1>0 but x1 is null");
  DDPM_REQUIRE(a->cout > 0 && a->cout % 32 == 0, "ddpm_conv_gemm: cout=%d must be a multiple of 32", a->cout);
  DDPM_REQUIRE(a->n > 0 && a->h > 0 && a->w > 0, "ddpm_conv_gemm: empty pixel grid");
  DDPM_REQUIRE(a->ldo % 8 == 0 && (a->res == nullptr || a->ldr % 8 == 0), "ddpm_conv_gemm: ldo/ldr must be multiples of 8");
  {
    const int hr = launch_conv_halo(a, stream);   // 3x3 stride-1 convs at high resolution: halo-resident kernel
    if (hr <= 0) return hr;
  }
  GemmParams p;
  std::memset(&p, 0, sizeof(p));
  if (int e = fill_taps(&p.taps, a)) return e;
  p.N = a->n; p.H = a->h; p.W = a->w;
  p.kb0 = a->c0 / 64; p.kb1 = a->c1 / 64;
  choose_box(a->n, a->h, a->w, &p.wb, &p.hb, &p.nb);
  p.tiles_w = (a->w + p.wb - 1) / p.wb;
  p.tiles_h = (a->h + p.hb - 1) / p.hb;
  p.tiles_n = (a->n + p.nb - 1) / p.nb;
  p.a_bytes = static_cast<uint32_t>(p.wb * p.hb * p.nb) * 128u;
  if (int e = fill_epilogue(&p.epi, a)) return e;
  if (p.epi.gsums) {
    // the fused GroupNorm sums are reduced per warp (32 consecutive tile rows): those must share one sample
    DDPM_REQUIRE(p.nb == 1 || (p.wb * p.hb) % 32 == 0,
                 "ddpm_conv_gemm: GroupNorm-backward fusion needs H*W >= 128 or H*W %% 32 == 0 (h=%d w=%d)", a->h, a->w);
  }
  const int src_n = a->src_n > 0 ? a->src_n : a->n;

  CUtensorMap ma0, ma1, mb;
  if (int e = make_act_map(&ma0, a->x0, a->c0, a->ld0, src_n, a->h, a->w, p.wb, p.hb, p.nb)) return e;
  if (a->c1 > 0) {
    if (int e = make_act_map(&ma1, a->x1, a->c1, a->ld1, src_n, a->h, a->w, p.wb, p.hb, p.nb)) return e;
  } else {
    ma1 = ma0;
  }
  int block_n = env_int("DDPM_BLOCK_N", 0);
  if (block_n == 0) {
    const long long mtiles = static_cast<long long>(p.tiles_w) * p.tiles_h * p.tiles_n;
    block_n = (a->cout % 256 == 0 && mtiles * (a->cout / 256) >= 2 * kNumSMs) ? 256 : 128;
    // short reductions (1x1 shortcuts: 2-6 k-blocks) are memory- and epilogue-bound: the 256-wide tile runs one CTA
    // per SM with its 8-chunk epilogue fully exposed; two co-resident 128-wide CTAs overlap each other instead
    if (a->ntaps * (p.kb0 + p.kb1) <= 8) block_n = 128;
  }
  long long k_total = a->k_total > 0 ? a->k_total : a->ldw;
  {
    // many tiles: the persistent kernel (tile i's epilogue under tile i+1's mainloop, no per-tile CTA setup)
    const long long mtiles = static_cast<long long>(p.tiles_w) * p.tiles_h * p.tiles_n;
    const int pbn = (block_n == 256) ? 256 : 128;
    const long long ptiles = mtiles * ((a->cout + pbn - 1) / pbn);
    if (env_int("DDPM_PERSIST", 1) != 0 && ptiles >= env_int("DDPM_PERSIST_MIN_TILES", 2 * kNumSMs)) {
      if (int e = make_wgt_map(&mb, a->wgt, k_total, a->ldw, a->cout, pbn)) return e;
      if (pbn == 256)
        return p.epi.gsums ? launch_gemm_persist<256, true>(ma0, ma1, mb, p, stream)
                           : launch_gemm_persist<256, false>(ma0, ma1, mb, p, stream);
      return p.epi.gsums ? launch_gemm_persist<128, true>(ma0, ma1, mb, p, stream)
                         : launch_gemm_persist<128, false>(ma0, ma1, mb, p, stream);
    }
  }
  if (block_n == 256) {
    if (int e = make_wgt_map(&mb, a->wgt, k_total, a->ldw, a->cout, 256)) return e;
    return p.epi.gsums ? launch_gemm<256, 4, true>(ma0, ma1, mb, p, stream)
                       : launch_gemm<256, 4, false>(ma0, ma1, mb, p, stream);
  }
  if (int e = make_wgt_map(&mb, a->wgt, k_total, a->ldw, a->cout, 128)) return e;
  if (p.epi.gsums) return launch_gemm<128, 3, true>(ma0, ma1, mb, p, stream);
  {
    const long long ctas0 = static_cast<long long>(p.tiles_w) * p.tiles_h * p.tiles_n * ((a->cout + 127) / 128);
    const int iters = a->ntaps * (p.kb0 + p.kb1);
    const int splits = (a->splitk_ws && !a->out_f32 && a->cout % 8 == 0) ? splitk_plan(ctas0, iters) : 1;
    if (splits > 1) {
      const long long pixels = static_cast<long long>(a->n) * a->h * a->w;
      DDPM_REQUIRE(a->splitk_ws_elems >= pixels * a->cout, "ddpm_conv_gemm: split-K workspace too small");
      DDPM_CUDA(cudaMemsetAsync(a->splitk_ws, 0, sizeof(float) * pixels * a->cout, stream));
      p.iters_per_split = (iters + splits - 1) / splits;
      p.splitk_ws = a->splitk_ws;
      if (int e = launch_gemm<128, 6, false>(ma0, ma1, mb, p, stream)) return e;
      const long long items = pixels * (a->cout / 8);
      splitk_finalize_kernel<<<static_cast<unsigned>((items + 255) / 256), 256, 0, stream>>>(
          a->splitk_ws, p.epi, pixels, static_cast<long long>(a->h) * a->w);
      return check_launch("splitk_finalize_kernel");
    }
  }
  // Few tiles (low-resolution layers): at most ~one CTA per SM is resident anyway, so a 3-stage ring keeps only two
  // 32 KB loads in flight per SM and the CTA is latency-bound on L2 (8x8 512->512: 3.5x the MMA time).  Spend the
  // whole shared memory on one CTA's ring instead.
  const long long ctas = static_cast<long long>(p.tiles_w) * p.tiles_h * p.tiles_n * ((a->cout + 127) / 128);
  if (ctas <= kNumSMs + kNumSMs / 2 && env_int("DDPM_DEEP_RING", 1))
    return launch_gemm<128, 6, false>(ma0, ma1, mb, p, stream);
  return launch_gemm<128, 3, false>(ma0, ma1, mb, p, stream);
}

extern "C" int ddpm_conv_wgrad(const ddpm_wgrad_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DDPM_REQUIRE(a && a->dy && a->x0 && a->dw, "ddpm_conv_wgrad: null pointer argument");
  DDPM_REQUIRE(a->c0 > 0 && a->c0 % 64 == 0 && a->c1 >= 0 && a->c1 % 64 == 0,
               "ddpm_conv_wgrad: channel counts must be multiples of 64 (c0=%d c1=%d)", a->c0, a->c1);
  DDPM_REQUIRE(a->cout > 0 && a->cout % 64 == 0, "ddpm_conv_wgrad: cout=%d must be a multiple of 64", a->cout);
  DDPM_REQUIRE(a->ldw % 4 == 0, "ddpm_conv_wgrad: ldw must be a multiple of 4");
  {
    const int rr = launch_wgrad_row(a, stream);   // 3x3 stride-1 convs at high resolution: row-resident kernel
    if (rr < 0) return rr;
    if (rr == 0)   // (a ones-column MMA for the bias was tried: it re-reads the dY tile and cost more than this pass)
      return a->dbias ? ddpm_reduce_hw(a->dy, a->ldy, a->n, a->h * a->w, a->cout, nullptr, 0, a->dbias, stream_) : DDPM_OK;
  }
  WgradParams p;
  std::memset(&p, 0, sizeof(p));
  if (a->ntaps < 1 || a->ntaps > kMaxTaps) { set_last_error("ntaps=%d out of range", a->ntaps); return DDPM_ERR_INVALID; }
  p.taps.ntaps = a->ntaps;
  for (int i = 0; i < a->ntaps; ++i) {
    p.taps.dn[i] = a->tap_dn[i]; p.taps.dh[i] = a->tap_dh[i]; p.taps.dw[i] = a->tap_dw[i]; p.taps.wk[i] = a->tap_wk[i];
  }
  p.N = a->n; p.H = a->h; p.W = a->w;
  p.kb0 = a->c0 / 64; p.kb1 = a->c1 / 64;
  p.Cout = a->cout;
  choose_box(a->n, a->h, a->w, &p.wb, &p.hb, &p.nb);
  p.tiles_w = (a->w + p.wb - 1) / p.wb;
  p.tiles_h = (a->h + p.hb - 1) / p.hb;
  p.tiles_n = (a->n + p.nb - 1) / p.nb;
  p.box_rows = p.wb * p.hb * p.nb;
  p.box_bytes = static_cast<uint32_t>(p.box_rows) * 128u;
  p.ptiles = p.tiles_w * p.tiles_h * p.tiles_n;
  p.dw = a->dw; p.ldw = a->ldw;
  p.lbo = static_cast<uint32_t>(env_int("DDPM_WGRAD_LBO", kBlockM * 128));
  p.sbo = static_cast<uint32_t>(env_int("DDPM_WGRAD_SBO", 1024));
  const int src_n = a->src_n > 0 ? a->src_n : a->n;
  const int kbt = p.kb0 + p.kb1;

  int block_n = (kbt % 2 == 0 && kbt >= 4) ? 128 : (kbt >= 2 ? 128 : 64);
  block_n = env_int("DDPM_WGRAD_BLOCK_N", block_n);
  p.cin_tiles = (kbt * 64 + block_n - 1) / block_n;
  const long long out_tiles = static_cast<long long>(p.cin_tiles) * a->ntaps * ((a->cout + kBlockM - 1) / kBlockM);
  int splits = a->splits;
  if (splits <= 0) {
    // aim for ~2 waves of CTAs, but keep >= 8 pixel tiles per CTA so the red.add epilogue is amortised
    long long want = (2LL * kNumSMs + out_tiles - 1) / out_tiles;
    long long cap = p.ptiles / 8;
    if (cap < 1) cap = 1;
    splits = static_cast<int>(want < cap ? want : cap);
    if (splits < 1) splits = 1;
  }
  if (splits > p.ptiles) splits = p.ptiles;
  p.ptiles_per_split = (p.ptiles + splits - 1) / splits;
  splits = (p.ptiles + p.ptiles_per_split - 1) / p.ptiles_per_split;
  p.atomic = (splits > 1 || a->accumulate) ? 1 : 0;

  CUtensorMap my, mx0, mx1;
  if (int e = make_act_map(&my, a->dy, a->cout, a->ldy, a->n, a->h, a->w, p.wb, p.hb, p.nb)) return e;
  if (int e = make_act_map(&mx0, a->x0, a->c0, a->ld0, src_n, a->h, a->w, p.wb, p.hb, p.nb)) return e;
  if (a->c1 > 0) {
    DDPM_REQUIRE(a->x1, "ddpm_conv_wgrad: c1>0 but x1 is null");
    if (int e = make_act_map(&mx1, a->x1, a->c1, a->ld1, src_n, a->h, a->w, p.wb, p.hb, p.nb)) return e;
  } else {
    mx1 = mx0;
  }
  int rc;
  if (block_n == 64) rc = launch_wgrad<64, 4>(my, mx0, mx1, p, splits, stream);
  else if (block_n == 128) rc = launch_wgrad<128, 3>(my, mx0, mx1, p, splits, stream);
  else { set_last_error("ddpm_conv_wgrad: unsupported block_n=%d", block_n); return DDPM_ERR_UNSUPPORTED; }
  if (rc != DDPM_OK || a->dbias == nullptr) return rc;
  // bias gradient: a separate column reduction over dY
  return ddpm_reduce_hw(a->dy, a->ldy, a->n, a->h * a->w, a->cout, nullptr, 0, a->dbias, stream_);
}

extern "C" long long ddpm_conv_gemm_workspace_elems(const ddpm_conv_args* a) {
  // fp32 elements of split-K workspace ddpm_conv_gemm would use for this problem (0: no split)
  if (!a || a->out_f32 || a->gn_sums || a->out_csum || a->cout % 8) return 0;
  if (a->n <= 0 || a->h <= 0 || a->w <= 0 || a->cout <= 0 || a->ntaps <= 0) return 0;   // ddpm_conv_gemm rejects these
  if (conv_uses_halo(a)) return 0;                                             // halo-resident kernel
  int wb, hb, nb;
  choose_box(a->n, a->h, a->w, &wb, &hb, &nb);
  const long long mt = static_cast<long long>((a->w + wb - 1) / wb) * ((a->h + hb - 1) / hb) * ((a->n + nb - 1) / nb);
  if (a->cout % 256 == 0 && mt * (a->cout / 256) >= 2 * kNumSMs) return 0;    // BLOCK_N = 256 path
  const int iters = a->ntaps * ((a->c0 + a->c1) / 64);
  return splitk_plan(mt * ((a->cout + 127) / 128), iters) > 1 ? static_cast<long long>(a->n) * a->h * a->w * a->cout : 0;
}
