// Halo-resident 3x3 stride-1 convolution (fprop and dgrad) on tcgen05 -- the high-resolution workhorse.
//
// Why: the generic tap-by-tap kernel (conv_igemm.cu) re-fetches every input pixel nine times from L2 and every
// weight tile once per 128 output pixels; ncu shows it pinned on L2->SM bandwidth (lts throughput ~70 %, 8.7 TB/s)
// at 34 % of tensor peak on the 128->128 @128^2 layers that carry 57 % of the UNet's FLOPs (profiles/).
//
// How: address each image in a *padded slot space* -- slot q = h * Wp + (w + 1), Wp = W + 1, where the one extra
// slot per row is zero and serves as the right pad of row h and the left pad of row h + 1.  A 3x3 tap (dh, dw) then
// is a pure shift by dh * Wp + dw slots.  A CTA owns 256 consecutive output slots (two M = 128 UMMA tiles, two TMEM
// accumulators).  Per 64-channel chunk it loads the R padded rows that cover the tile plus its one-row halo ONCE
// (one 4-D TMA box starting at w = -1: out-of-bounds rows/columns are zero-filled = the conv padding) and issues
// all nine taps from that resident halo by shifting the start address of the SWIZZLE_128B operand descriptor by a
// whole number of 128-byte rows (profiles/probe_shift: the hardware swizzles on absolute address bits, so any
// 128-byte-aligned start is valid).  L2->SM traffic per FLOP drops ~2.6x.  Slots that fall on a pad position
// produce garbage accumulator rows that are simply not stored (W / (W + 1) efficiency).
//
// Wide images (W > ~140: the 224^2 default of the reference, the 256^2 LoRA configs) do not fit a double-buffered halo of
// full rows in shared memory, so each row is cut into S column STRIPS of Ws = W / S pixels and a (image, strip) pair
// becomes the unit that is tiled: the slot space of a unit has pitch Wp = Ws + 2 -- one halo column on EACH side, real
// pixels of the neighbouring strip or TMA zero fill at the image border -- the box starts at w = strip * Ws - 1, and
// the slots on the two halo columns are the garbage rows (Ws / (Ws + 2) efficiency: 98.5 % at 256^2).  S = 1 keeps the
// shared pad slot (pitch W + 1).
//
// Structure: persistent CTAs (one per SM), static round-robin tile schedule, warp-specialised:
//   warps 0-7  epilogue (TMEM -> registers -> bias / time-embedding / residual -> bf16 NHWC), overlapped with the
//              next tile's mainloop through double-buffered accumulators (2 x 256 TMEM columns)
//   warp  8    halo producer (TMA, double-buffered halo)       warp 9   weight-tile producer (TMA, ring)
//   warp 10    MMA issuer + TMEM owner
#include "common.cuh"
#include "conv_epilogue.cuh"

#include <atomic>
#include <cstring>

#include "../../include/ddpm_b200.h"

namespace ddpm {

constexpr int kHaloThreads = 352;
constexpr int kHaloPairThreads = 384;   // pair kernel: 8 epilogue warps + 3 role warps + 1 idle warp = 3 full warpgroups
constexpr int kHaloMaxBStages = 8;   // weight-tile ring depth is chosen at launch from the shared memory left over
constexpr int kHaloMaxPairStages = 16;   // CTA-pair variant: half tiles (8 KB), twice the depth
constexpr int kTileSlots = 256;
constexpr int kHaloBBytes = 128 * 128;   // one weight tile: 128 cout rows x 64 ch

struct HaloParams {
  int N, H, W, Wp;
  int S, Ws;             // column strips per row, pixels per strip (S * Ws == W); units = N * S
  int units;
  int kb0, kb1;
  int Cout, Cin_total;
  int R;                 // padded rows per halo box
  int tiles_per_img, n_tiles, total_tiles;
  uint32_t halo_bytes;   // R * Wp * 128
  uint32_t halo_stride;  // halo_bytes rounded up to 1024
  int b_stages;          // weight-tile ring depth
  int mma_n;             // UMMA N: 128, or Cout rounded up to 16 for narrow outputs (conv_out: 3 -> 32 columns)
  unsigned int* sched_ctr;   // pair kernel: global tile counter of the dynamic schedule (nullptr: static round-robin)
  EpiParams epi;
};

// ---- dynamic tile schedule of the CTA-pair kernel ---------------------------------------------------------------
// A static round-robin makes the kernel as slow as its LAST-STARTED pair: when some SMs are still held by a kernel of
// another stream (the weight-gradient GEMMs of the second stream, NCCL's all-reduce CTAs under DDP -- neither can
// co-reside with a 200 KB / 59 K-register CTA) those pairs start late and still own a full share of the tiles.  With
// the dynamic schedule the first tile of a pair is static (tile = pair index, no latency) and every further tile is
// drawn from a global counter by the otherwise idle 12th warp of the leader CTA; it hands the tile index to the role
// warps of BOTH CTAs through a 4-deep ring in the LEADER's shared memory (plain store, then a release arrive on a
// barrier in each CTA; readers acquire at cluster scope and load the word over its shared::cluster address), running up
// to 4 tiles ahead of the slowest consumer, so the ~1 us atomic never sits on the critical path.  The counter resets
// itself: the draws of one launch number exactly total_tiles (every pair draws until it gets a value past the end), so
// the draw that returns total_tiles - 1 is the last one.
constexpr int kSchedDepth = 4;
constexpr uint32_t kSchedConsumers = 21;   // role warps that read a tile index: 8 epilogue + halo + weight producer in
                                           // both CTAs, + the leader's MMA warp

struct TileFeed {
  uint64_t* full;        // [kSchedDepth], this CTA
  uint32_t tiles0;       // shared::cluster address of the LEADER's tile ring
  uint32_t empty0;       // shared::cluster address of the LEADER's empty[0]
  uint32_t s, ph;
  int pt, step, total;   // static schedule (dyn == false)
  bool dyn;

  // next tile of this pair, or -1; called by all lanes of a converged warp
  __device__ __forceinline__ int next() {
    if (!dyn) {
      const int t = pt;
      pt += step;
      return t < total ? t : -1;
    }
    mbar_wait_acquire_cluster(&full[s], ph);
    const int t = static_cast<int>(ld_shared_cluster_u32(tiles0 + s * 4u));
    __syncwarp();
    // the slot may be refilled once every consumer warp of the pair has read it; the (always true) test on the loaded
    // value keeps the arrive behind the completion of the load
    if (elect_one() && t >= -1) mbar_arrive_cluster(empty0 + s * 8u);
    if (++s == static_cast<uint32_t>(kSchedDepth)) { s = 0; ph ^= 1; }
    return t;
  }
  // the tile after the one next() returned last, without consuming it (-1: none)
  __device__ __forceinline__ int peek() {
    if (!dyn) return pt < total ? pt : -1;
    mbar_wait_acquire_cluster(&full[s], ph);
    return static_cast<int>(ld_shared_cluster_u32(tiles0 + s * 4u));
  }
};

constexpr unsigned int kTileCtrPool = 256;
__device__ unsigned int g_halo_tile_ctr[kTileCtrPool];   // zero at load; every launch leaves its counter at zero

__device__ __forceinline__ int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

__global__ void __launch_bounds__(kHaloThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* halo = smem;                                   // 2 x halo_stride
  uint8_t* bsm = smem + 2 * p.halo_stride;                // b_stages x 16 KB
  const int kHaloBStages = p.b_stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bsm + kHaloBStages * kHaloBBytes);
  uint64_t* halo_full = bars;            // [2]
  uint64_t* halo_empty = bars + 2;       // [2]
  uint64_t* b_full = bars + 4;           // [kHaloBStages]
  uint64_t* b_empty = b_full + kHaloMaxBStages;
  uint64_t* tmem_full = b_empty + kHaloMaxBStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kbt = p.kb0 + p.kb1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.kb1 > 0) tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&halo_full[i], 1);
      mbar_init(&halo_empty[i], 1);
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);     // one arrival per epilogue warp
    }
    for (int i = 0; i < kHaloBStages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 10) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // The three single-warp roles below are ISSUE-bound if written naively (ncu source view, profiles/r1_halo_issue.md:
  // the MMA warp never waits on a barrier, it spends ~1000 clk per tap executing ~170 SASS instructions -- a runtime
  // modulo for the ring slot, tap / 3, descriptor construction, R2UR moves -- against 512 clk of tensor work).  So:
  // ring slots and phases advance incrementally, taps are fully unrolled, and descriptors are one 64-bit add away from
  // a per-chunk base.
  if (warp == 8) {
    // ---------------- halo producer (warp-uniform loop, one elected lane issues) ----------------
    uint32_t hb = 0, hph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int mt = tile / p.n_tiles;
      const int unit = mt / p.tiles_per_img;
      const int q0 = (mt - unit * p.tiles_per_img) * kTileSlots;
      const int img = unit / p.S;
      const int x0 = (unit - img * p.S) * p.Ws - 1;      // first column of the box: the left halo column of the strip
      const int row_lo = floor_div(q0 - p.Wp - 1, p.Wp);
      for (int c = 0; c < kbt; ++c) {
        mbar_wait(&halo_empty[hb], hph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&halo_full[hb], p.halo_bytes);
          if (c < p.kb0)
            tma_load_4d(halo + hb * p.halo_stride, &tmA0, &halo_full[hb], c * 64, x0, row_lo, img);
          else
            tma_load_4d(halo + hb * p.halo_stride, &tmA1, &halo_full[hb], (c - p.kb0) * 64, x0, row_lo, img);
        }
        __syncwarp();
        hb ^= 1;
        hph ^= (hb == 0);
      }
    }
  } else if (warp == 9) {
    // ---------------- weight-tile producer ----------------
    uint32_t bs = 0, bph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles;
      for (int c = 0; c < kbt; ++c) {
        int kcol = c * 64;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap, kcol += p.Cin_total) {
          mbar_wait(&b_empty[bs], bph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&b_full[bs], kHaloBBytes);
            tma_load_2d(bsm + bs * kHaloBBytes, &tmB, &b_full[bs], kcol, nt * 128);
          }
          __syncwarp();
          if (++bs == static_cast<uint32_t>(kHaloBStages)) { bs = 0; bph ^= 1; }
        }
      }
    }
  } else if (warp == 10) {
    // ---------------- MMA issuer (whole warp runs the loop; one elected lane issues) ----------------
    const uint32_t idesc = make_idesc_bf16(128, p.mma_n, false, false);
    const uint64_t db_base = make_smem_desc_sw128(smem_u32(bsm), 16, 1024);
    const uint32_t row_step = static_cast<uint32_t>(p.Wp) * 8u;     // one padded image row, in 16-byte descriptor units
    uint32_t bs = 0, bph = 0, hb = 0, hph = 0, acc = 0, aph = 0;
    uint64_t db = db_base;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int mt = tile / p.n_tiles;
      const int unit = mt / p.tiles_per_img;
      const int q0 = (mt - unit * p.tiles_per_img) * kTileSlots;
      const int row_lo = floor_div(q0 - p.Wp - 1, p.Wp);
      const int rel00 = q0 - row_lo * p.Wp - p.Wp - 1;   // slot of tap (-1, -1) of output slot q0 inside the halo buffer
      mbar_wait(&tmem_empty[acc], aph ^ 1);
      tc_fence_after();
      const uint32_t d0 = tmem_base + acc * 256;
      for (int c = 0; c < kbt; ++c) {
        mbar_wait(&halo_full[hb], hph);
        const uint32_t h_addr = smem_u32(halo + hb * p.halo_stride);
        // descriptors differ only in the 14-bit start-address field: build one per chunk, then add (bytes >> 4)
        const uint64_t da_c = make_smem_desc_sw128(h_addr + static_cast<uint32_t>(rel00) * 128u, 16, 1024);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(&b_full[bs], bph);
          tc_fence_after();
          const uint64_t da = da_c + static_cast<uint64_t>((tap / 3) * row_step + (tap % 3) * 8u);
          const uint32_t first = (tap != 0 || c != 0) ? 1u : 0u;
          if (elect_one()) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d0 + u * 128, da + static_cast<uint64_t>(u * 1024 + k * 2), db + static_cast<uint64_t>(k * 2),
                          idesc, k == 0 ? first : 1u);
            }
            umma_commit(&b_empty[bs]);
          }
          __syncwarp();
          db += kHaloBBytes >> 4;
          if (++bs == static_cast<uint32_t>(kHaloBStages)) { bs = 0; bph ^= 1; db = db_base; }
        }
        if (elect_one()) umma_commit(&halo_empty[hb]);
        __syncwarp();
        hb ^= 1;
        hph ^= (hb == 0);
      }
      if (elect_one()) umma_commit(&tmem_full[acc]);
      __syncwarp();
      acc ^= 1;
      aph ^= (acc == 0);
    }
  } else {
    // ---------------- epilogue (8 warps) ----------------
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int half = warp >> 2;       // column half: chunks {0,1} or {2,3}
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
      const int nt = tile % p.n_tiles;
      const int mt = tile / p.n_tiles;
      const int unit = mt / p.tiles_per_img;
      const int q0 = (mt - unit * p.tiles_per_img) * kTileSlots;
      const int img = unit / p.S;
      const int xs = (unit - img * p.S) * p.Ws;          // first image column of this strip
      const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
      // step i = (cc, u): column chunk half*2 + i/2, row block i%2; the GroupNorm input of step i+1 is loaded while
      // step i is processed, and that of step 0 while the tile's mainloop is still running
      auto geom = [&](int i, int& col, bool& valid, long long& pix, int& u) {
        u = i & 1;
        col = nt * 128 + (half * 2 + (i >> 1)) * 32;
        const int slot = q0 + u * 128 + q * 32 + lane;
        const int h = slot / p.Wp;
        const int w = slot - h * p.Wp - 1;
        valid = (w >= 0) && (w < p.Ws) && (h < p.H);
        pix = (static_cast<long long>(img) * p.H + h) * p.W + xs + w;
      };
      EpiX xcur, xnext;
      {
        int col, u; bool valid; long long pix;
        geom(0, col, valid, pix, u);
        epi_load_x(p.epi, valid, pix, col, xcur);
      }
      if (p.epi.res != nullptr || epi_slot_is_gn(p.epi)) {
        // side input of the NEXT tile of this CTA -> L2, one whole tile ahead of its first use
        const int tn = tile + gridDim.x;
        if (tn < p.total_tiles) {
          const int nt2 = tn % p.n_tiles, mt2 = tn / p.n_tiles;
          const int unit2 = mt2 / p.tiles_per_img;
          const int q02 = (mt2 - unit2 * p.tiles_per_img) * kTileSlots;
          const int img2 = unit2 / p.S, xs2 = (unit2 - img2 * p.S) * p.Ws;
#pragma unroll
          for (int u2 = 0; u2 < 2; ++u2) {
            const int slot = q02 + u2 * 128 + q * 32 + lane;
            const int h2 = slot / p.Wp, w2 = slot - h2 * p.Wp - 1;
            epi_prefetch_side(p.epi, (w2 >= 0) && (w2 < p.Ws) && (h2 < p.H),
                              (static_cast<long long>(img2) * p.H + h2) * p.W + xs2 + w2, nt2 * 128 + half * 64, 64);
          }
        }
      }
      mbar_wait(&tmem_full[acc], aph);
      tc_fence_after();
      float t1 = 0.f, t2 = 0.f;
#pragma unroll 1
      for (int i = 0; i < 4; ++i) {
        int col, u; bool valid; long long pix;
        if (i + 1 < 4) {
          geom(i + 1, col, valid, pix, u);
          epi_load_x(p.epi, valid, pix, col, xnext);
        }
        geom(i, col, valid, pix, u);
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256 + u * 128 + (half * 2 + (i >> 1)) * 32, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        epi_chunk(p.epi, v, valid, img, pix, col, lane, t1, t2, xcur);
        if (u == 1) {   // both row blocks of this column chunk done; a tile lies within one image
          epi_flush_sums(p.epi, img, col, lane, t1, t2);
          t1 = t2 = 0.f;
        }
        xcur = xnext;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 10) tmem_dealloc(tmem_base, 512);
}

// =====================================================================================================
// CTA-pair variant (thread-block cluster of 2, tcgen05 cta_group::2).
// The two CTAs of a pair work on the SAME tile position of two DIFFERENT images (img = 2*pair + rank), so their halo
// buffers have identical geometry and one shared-memory descriptor addresses both.  The leader's MMA warp issues
// M = 256 instructions: rows 0-127 come from its own halo and land in its own TMEM, rows 128-255 from the peer's halo
// into the peer's TMEM; each CTA stages only HALF of every weight tile (64 of the 128 cout rows).  Per FLOP this halves
// the issue cost (one issuing warp drives two tensor cores), the weight-tile L2 traffic and the weight shared-memory
// footprint (the ring doubles in depth).  Barriers: "full" barriers live in the leader and count one arrival per CTA
// (the peer's TMA signals them through their shared::cluster address); "empty" barriers live in both CTAs and are
// released by multicast tcgen05.commit; the accumulator-empty barrier of the leader collects both epilogues.
// =====================================================================================================
template <int MODE>      // epilogue specialisation (conv_epilogue.cuh::epi_specialize); kEpiAny = all variants at run time
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kHaloPairThreads, 1)
conv_halo_pair_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                      const __grid_constant__ CUtensorMap tmB, const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  // identical carve-up in both CTAs: offsets (not generic addresses) must match across the pair
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* halo = smem;                                   // 2 x halo_stride
  uint8_t* bsm = smem + 2 * p.halo_stride;                // b_stages x 8 KB (this CTA's half of each weight tile)
  const int kStages = p.b_stages;
  constexpr int kHalfB = kHaloBBytes / 2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bsm + kStages * kHalfB);
  uint64_t* halo_full = bars;            // [2]   (used in the leader)
  uint64_t* halo_empty = bars + 2;       // [2]   (both CTAs)
  uint64_t* b_full = bars + 4;           // [kHaloMaxPairStages] (leader)
  uint64_t* b_empty = b_full + kHaloMaxPairStages;
  uint64_t* tmem_full = b_empty + kHaloMaxPairStages;   // [2] (both CTAs)
  uint64_t* tmem_empty = tmem_full + 2;                 // [2] (leader; 16 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* sched_full = tmem_empty + 3;                // [kSchedDepth] (both CTAs)
  uint64_t* sched_empty = sched_full + kSchedDepth;     // [kSchedDepth] (leader)
  int* sched_tiles = reinterpret_cast<int*>(sched_empty + kSchedDepth);   // [kSchedDepth] (both CTAs)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kbt = p.kb0 + p.kb1;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const bool dyn = p.sched_ctr != nullptr && p.total_tiles > n_pairs;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.kb1 > 0) tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&halo_full[i], 2);      // one arrival per CTA of the pair
      mbar_init(&halo_empty[i], 1);
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 16);    // 8 epilogue warps x 2 CTAs
    }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&b_full[i], 2);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < kSchedDepth; ++i) {
      mbar_init(&sched_full[i], 1);                 // the scheduler's release arrive
      mbar_init(&sched_empty[i], kSchedConsumers);
    }
    fence_barrier_init();
  }
  if (warp == 10) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Register split (ncu, profiles/r2_epilogue_stalls.md: the fused epilogues stalled on the first use of their side-input
  // loads, one 32-column chunk of prefetch was all that 168 registers allowed).  The four role warps (one issuing lane
  // each) shrink to 56 registers, the eight epilogue warps grow to 224 and hold the side input of ALL FOUR chunks of a
  // tile, loaded before they wait for the accumulator.
  // (the setmaxnreg instructions sit at the top of the two role regions below so that ptxas allocates each region
  // against its own budget)


  // pair tile -> (image and first column of this CTA's strip, first slot, cout tile); the pair works on two
  // consecutive (image, strip) units
  auto decode = [&](int pt, int& img, int& xs, bool& img_ok, int& q0, int& nt) {
    nt = pt % p.n_tiles;
    const int mt = pt / p.n_tiles;
    const int ip = mt / p.tiles_per_img;
    q0 = (mt - ip * p.tiles_per_img) * kTileSlots;
    int unit = 2 * ip + static_cast<int>(rank);
    img_ok = unit < p.units;
    if (!img_ok) unit = p.units - 1;     // odd unit count: the last peer recomputes the last unit and stores nothing
    img = unit / p.S;
    xs = (unit - img * p.S) * p.Ws;
  };

  TileFeed feed;
  feed.full = sched_full;
  feed.tiles0 = mapa_u32(smem_u32(&sched_tiles[0]), 0);
  feed.empty0 = mapa_u32(smem_u32(&sched_empty[0]), 0);
  feed.s = 0; feed.ph = 0;
  feed.pt = pair_id; feed.step = n_pairs; feed.total = p.total_tiles;
  feed.dyn = dyn;

  if (warp >= 8) {
  setmaxnreg_dec<56>();
  if (warp == 8) {
    // ---------------- halo producer (both CTAs; completion is signalled on the LEADER's barrier) ----------------
    uint32_t hb = 0, hph = 0;
    const uint32_t full0 = mapa_u32(smem_u32(&halo_full[0]), 0), full1 = mapa_u32(smem_u32(&halo_full[1]), 0);
    for (int pt = feed.next(); pt >= 0; pt = feed.next()) {
      int img, xs, q0, nt; bool ok;
      decode(pt, img, xs, ok, q0, nt);
      const int row_lo = floor_div(q0 - p.Wp - 1, p.Wp);
      for (int c = 0; c < kbt; ++c) {
        mbar_wait(&halo_empty[hb], hph ^ 1);
        if (elect_one()) {
          const uint32_t fb = hb ? full1 : full0;
          if (c < p.kb0)
            tma_load_4d_pair(halo + hb * p.halo_stride, &tmA0, fb, c * 64, xs - 1, row_lo, img);
          else
            tma_load_4d_pair(halo + hb * p.halo_stride, &tmA1, fb, (c - p.kb0) * 64, xs - 1, row_lo, img);
          if (leader) mbar_expect_tx(&halo_full[hb], 2u * p.halo_bytes);   // arrive (1 of 2) + both CTAs' bytes
          else mbar_arrive_cluster(fb);                                     // arrive (2 of 2)
        }
        __syncwarp();
        hb ^= 1;
        hph ^= (hb == 0);
      }
    }
  } else if (warp == 9) {
    // ---------------- weight-tile producer: this CTA's 64 cout rows of every tile ----------------
    uint32_t bs = 0, bph = 0;
    const uint32_t bfull0 = mapa_u32(smem_u32(&b_full[0]), 0);
    for (int pt = feed.next(); pt >= 0; pt = feed.next()) {
      const int nt = pt % p.n_tiles;
      const int row0 = nt * 128 + static_cast<int>(rank) * (p.mma_n / 2);   // this CTA's half of the N columns
      for (int c = 0; c < kbt; ++c) {
        int kcol = c * 64;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap, kcol += p.Cin_total) {
          mbar_wait(&b_empty[bs], bph ^ 1);
          if (elect_one()) {
            const uint32_t fb = bfull0 + bs * 8u;
            tma_load_2d_pair(bsm + bs * kHalfB, &tmB, fb, kcol, row0);
            if (leader) mbar_expect_tx(&b_full[bs], static_cast<uint32_t>(kHaloBBytes));
            else mbar_arrive_cluster(fb);
          }
          __syncwarp();
          if (++bs == static_cast<uint32_t>(kStages)) { bs = 0; bph ^= 1; }
        }
      }
    }
  } else if (warp == 10) {
    if (leader) {
      // ---------------- MMA issuer (leader only): M = 256 across the pair ----------------
      const uint32_t idesc = make_idesc_bf16(256, p.mma_n, false, false);
      const uint64_t db_base = make_smem_desc_sw128(smem_u32(bsm), 16, 1024);
      const uint32_t row_step = static_cast<uint32_t>(p.Wp) * 8u;
      uint32_t bs = 0, bph = 0, hb = 0, hph = 0, acc = 0, aph = 0;
      uint64_t db = db_base;
      for (int pt = feed.next(); pt >= 0; pt = feed.next()) {
        int img, xs, q0, nt; bool ok;
        decode(pt, img, xs, ok, q0, nt);
        const int row_lo = floor_div(q0 - p.Wp - 1, p.Wp);
        const int rel00 = q0 - row_lo * p.Wp - p.Wp - 1;
        mbar_wait(&tmem_empty[acc], aph ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + acc * 256;
        for (int c = 0; c < kbt; ++c) {
          mbar_wait(&halo_full[hb], hph);
          const uint32_t h_addr = smem_u32(halo + hb * p.halo_stride);
          const uint64_t da_c = make_smem_desc_sw128(h_addr + static_cast<uint32_t>(rel00) * 128u, 16, 1024);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_full[bs], bph);
            tc_fence_after();
            const uint64_t da = da_c + static_cast<uint64_t>((tap / 3) * row_step + (tap % 3) * 8u);
            const uint32_t first = (tap != 0 || c != 0) ? 1u : 0u;
            if (elect_one()) {
#pragma unroll
              for (int u = 0; u < 2; ++u) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16_pair(d0 + u * 128, da + static_cast<uint64_t>(u * 1024 + k * 2),
                                 db + static_cast<uint64_t>(k * 2), idesc, k == 0 ? first : 1u);
              }
              umma_commit_pair(&b_empty[bs]);
            }
            __syncwarp();
            db += kHalfB >> 4;
            if (++bs == static_cast<uint32_t>(kStages)) { bs = 0; bph ^= 1; db = db_base; }
          }
          if (elect_one()) umma_commit_pair(&halo_empty[hb]);
          __syncwarp();
          hb ^= 1;
          hph ^= (hb == 0);
        }
        if (elect_one()) umma_commit_pair(&tmem_full[acc]);
        __syncwarp();
        acc ^= 1;
        aph ^= (acc == 0);
      }
    }
  } else if (dyn && leader) {
    // ---------------- tile scheduler (warp 11 of the leader) ----------------
    // The WHOLE warp walks the loop (lane 0 draws and publishes): the kernel ends in an .aligned cluster barrier, which a
    // warp must reach converged -- a one-lane loop let lanes 1-31 arrive there early and lane 0 a second time.
    uint32_t s = 0, ph = 0;
    const uint32_t full_p = mapa_u32(smem_u32(&sched_full[0]), 1);
    const unsigned int total = static_cast<unsigned int>(p.total_tiles);
    bool first = true;
    for (;;) {
      mbar_wait(&sched_empty[s], ph ^ 1);
      uint32_t done = 0;
      if (lane == 0) {
        unsigned int t;
        if (first) {
          t = static_cast<unsigned int>(pair_id);
        } else {
          const unsigned int old = atomicAdd(p.sched_ctr, 1u);
          if (old == total - 1u) *reinterpret_cast<volatile unsigned int*>(p.sched_ctr) = 0u;   // the launch's last draw
          t = static_cast<unsigned int>(n_pairs) + old;
        }
        done = t >= total ? 1u : 0u;
        *reinterpret_cast<volatile int*>(&sched_tiles[s]) = done ? -1 : static_cast<int>(t);
        mbar_arrive_release_cluster(full_p + s * 8u);     // peer's consumers
        mbar_arrive(&sched_full[s]);                      // this CTA's consumers (release.cta)
      }
      first = false;
      done = __shfl_sync(0xffffffffu, done, 0);
      if (done) break;
      if (++s == static_cast<uint32_t>(kSchedDepth)) { s = 0; ph ^= 1; }
    }
    __syncwarp();
  }
  } else {
    setmaxnreg_inc<224>();
    // ---------------- epilogue (8 warps per CTA; each CTA drains its own TMEM = its own image) ----------------
    const EpiParams epi = epi_specialize<MODE>(p.epi);
    const int q = warp & 3;
    const int half = warp >> 2;
    uint32_t acc = 0, aph = 0;
    const uint32_t empty0 = mapa_u32(smem_u32(&tmem_empty[0]), 0);
    for (int pt = feed.next(); pt >= 0; pt = feed.next()) {
      int img, xs, q0, nt; bool img_ok;
      decode(pt, img, xs, img_ok, q0, nt);
      auto geom = [&](int i, int& col, bool& valid, long long& pix, int& u) {
        u = i & 1;
        col = nt * 128 + (half * 2 + (i >> 1)) * 32;
        const int slot = q0 + u * 128 + q * 32 + lane;
        const int h = slot / p.Wp;
        const int w = slot - h * p.Wp - 1;
        valid = img_ok && (w >= 0) && (w < p.Ws) && (h < p.H);
        pix = (static_cast<long long>(img) * p.H + h) * p.W + xs + w;
      };
      // side input (GroupNorm x or residual) of all four chunks, in flight while the tile's mainloop still runs
      EpiX x0, x1, x2, x3;
      {
        int col, u; bool valid; long long pix;
        geom(0, col, valid, pix, u);
        epi_load_x(epi, valid, pix, col, x0);
        geom(1, col, valid, pix, u);
        epi_load_x(epi, valid, pix, col, x1);
        geom(2, col, valid, pix, u);
        epi_load_x(epi, valid, pix, col, x2);
        geom(3, col, valid, pix, u);
        epi_load_x(epi, valid, pix, col, x3);
      }
      const int pt_next = (epi.res != nullptr || epi_slot_is_gn(epi)) ? feed.peek() : -1;
      if (pt_next >= 0) {
        // side input of the NEXT tile of this CTA -> L2, one whole tile ahead of its first use
        int img2, xs2, q02, nt2; bool ok2;
        decode(pt_next, img2, xs2, ok2, q02, nt2);
#pragma unroll
        for (int u2 = 0; u2 < 2; ++u2) {
          const int slot = q02 + u2 * 128 + q * 32 + lane;
          const int h2 = slot / p.Wp, w2 = slot - h2 * p.Wp - 1;
          epi_prefetch_side(epi, ok2 && (w2 >= 0) && (w2 < p.Ws) && (h2 < p.H),
                            (static_cast<long long>(img2) * p.H + h2) * p.W + xs2 + w2, nt2 * 128 + half * 64, 64);
        }
      }
      mbar_wait(&tmem_full[acc], aph);
      tc_fence_after();
      // two column chunks per warp; the two row blocks (u = 0, 1) of a chunk share its columns, so their per-column
      // partials of the fused reductions are added in registers and cross the warp ONCE
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        EpiAcc racc;
        epi_acc_zero(racc);
        int col = 0;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          int uu; bool valid; long long pix;
          geom(2 * cc + u, col, valid, pix, uu);
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256 + u * 128 + (half * 2 + cc) * 32, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          epi_math<true, true>(epi, v, valid, img, pix, col, u == 0 ? x0 : x1, racc);
          epi_store(epi, v, valid, pix, col);
        }
        float t1 = 0.f, t2 = 0.f;
        epi_reduce<true>(epi, lane, racc, t1, t2);
        epi_flush_sums(epi, img_ok ? img : -1, col, lane, t1, t2);
        x0 = x2;
        x1 = x3;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(empty0 + acc * 8u);     // the LEADER's accumulator-empty barrier
      acc ^= 1;
      aph ^= (acc == 0);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 10) tmem_dealloc_pair(tmem_base, 512);
}

// Strip plan for an image width: the smallest number of column strips S (S | W) whose double-buffered halo leaves room
// for a weight-tile ring of >= 3 full (or 6 half) tiles.  Returns 0 when the width is not served by this kernel.
struct HaloGeom {
  int S, Ws, Wp, R;
  uint32_t halo_bytes, halo_stride;
  size_t fixed;
};
static bool halo_geometry(int W, HaloGeom* g) {
  // 32x32 maps included: measured on the 128^2 training step, 35.80 -> 35.39 ms against the generic kernel (L2-bound
  // tap-by-tap refetch); below that a 256-slot tile wastes too many slots on the last partial tile of an image
  const int min_w = env_int("DDPM_HALO_MIN_W", 32);
  if (W < min_w) return false;
  for (int S = 1; S <= 8; ++S) {
    if (W % S) continue;
    const int Ws = W / S;
    if (Ws < 32 || (S > 1 && Ws < min_w)) break;
    const int Wp = (S == 1) ? W + 1 : Ws + 2;
    if (Wp > 256) continue;                                  // TMA box extent limit
    // rows covering [q0 - Wp - 1, q0 + 255 + Wp + 1] for any q0
    const int R = (kTileSlots + 2 * Wp + 1) / Wp + 2;
    if (R > 256) continue;
    const uint32_t hbytes = static_cast<uint32_t>(R) * Wp * 128u;
    const uint32_t hstride = (hbytes + 1023u) & ~1023u;
    const size_t fixed = 2ull * hstride + 512 + 1024;
    if (fixed + 3 * kHaloBBytes > 227 * 1024) continue;
    g->S = S; g->Ws = Ws; g->Wp = Wp; g->R = R;
    g->halo_bytes = hbytes; g->halo_stride = hstride; g->fixed = fixed;
    return true;
  }
  return false;
}

int conv_halo_strips(int W) {
  HaloGeom g;
  if (env_int("DDPM_HALO", 1) == 0 || !halo_geometry(W, &g)) return 0;
  return g.S;
}

static bool halo_taps_ok(const ::ddpm_conv_args* a) {
  if (a->ntaps != 9 || (a->src_n != 0 && a->src_n != a->n)) return false;
  const int cin_total = a->c0 + a->c1;
  for (int t = 0; t < 9; ++t) {
    if (a->tap_dn[t] != 0 || a->tap_dh[t] != t / 3 - 1 || a->tap_dw[t] != t % 3 - 1 || a->tap_wk[t] != t * cin_total)
      return false;
  }
  return true;
}

// would ddpm_conv_gemm run this problem on the halo-resident kernel?
bool conv_uses_halo(const ::ddpm_conv_args* a) {
  return halo_taps_ok(a) && conv_halo_strips(a->w) > 0;
}

int launch_conv_halo(const ::ddpm_conv_args* a, cudaStream_t stream) {
  if (env_int("DDPM_HALO", 1) == 0) return 1;
  if (!halo_taps_ok(a) || (a->out == nullptr && a->out_f32 == nullptr)) return 1;
  const int cin_total = a->c0 + a->c1;
  const int W = a->w, H = a->h;
  HaloGeom g;
  if (!halo_geometry(W, &g)) return 1;
  HaloParams p;
  std::memset(&p, 0, sizeof(p));
  p.N = a->n; p.H = H; p.W = W; p.Wp = g.Wp;
  p.S = g.S; p.Ws = g.Ws; p.units = a->n * g.S;
  p.kb0 = a->c0 / 64; p.kb1 = a->c1 / 64;
  p.Cout = a->cout; p.Cin_total = cin_total;
  p.R = g.R;
  p.halo_bytes = g.halo_bytes;
  p.halo_stride = g.halo_stride;
  const size_t fixed = g.fixed;
  int bst = static_cast<int>((227 * 1024 - fixed) / kHaloBBytes);
  if (bst > kHaloMaxBStages) bst = kHaloMaxBStages;
  bst = env_int("DDPM_HALO_BSTAGES", bst) < bst ? env_int("DDPM_HALO_BSTAGES", bst) : bst;
  if (bst < 2) return 1;
  p.b_stages = bst;
  p.mma_n = a->cout >= 128 ? 128 : ((a->cout + 15) / 16) * 16;   // narrow outputs: do not multiply the zero rows
  const size_t smem = fixed + static_cast<size_t>(bst) * kHaloBBytes;
  p.tiles_per_img = (H * g.Wp + kTileSlots - 1) / kTileSlots;     // tiles per (image, strip) unit
  p.n_tiles = (a->cout + 127) / 128;
  p.total_tiles = p.units * p.tiles_per_img * p.n_tiles;
  if (int e = fill_epilogue(&p.epi, a)) return e;

  CUtensorMap ma0, ma1, mb;
  if (int e = make_act_map(&ma0, a->x0, a->c0, a->ld0, a->n, H, W, g.Wp, g.R, 1)) return e;
  if (a->c1 > 0) {
    if (int e = make_act_map(&ma1, a->x1, a->c1, a->ld1, a->n, H, W, g.Wp, g.R, 1)) return e;
  } else {
    ma1 = ma0;
  }
  const long long k_total = a->k_total > 0 ? a->k_total : a->ldw;
  if (env_int("DDPM_HALO_PAIR", 1) != 0 && p.units >= 2) {
    // CTA-pair variant: half weight tiles (8 KB per stage), pair tiles over pairs of (image, strip) units
    HaloParams pp = p;
    int pst = static_cast<int>((227 * 1024 - fixed) / (kHaloBBytes / 2));
    if (pst > kHaloMaxPairStages) pst = kHaloMaxPairStages;
    pp.b_stages = pst;
    pp.total_tiles = ((p.units + 1) / 2) * p.tiles_per_img * p.n_tiles;
    const size_t psmem = fixed + static_cast<size_t>(pst) * (kHaloBBytes / 2);
    if (int e = make_wgt_map(&mb, a->wgt, k_total, a->ldw, a->cout, 64)) return e;
    // one instantiation per epilogue variant (DDPM_HALO_SPECIALIZE=0: the all-variants kernel for every launch)
    using PairFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const HaloParams);
    static const PairFn fns[5] = {conv_halo_pair_kernel<kEpiAny>, conv_halo_pair_kernel<kEpiLean>,
                                  conv_halo_pair_kernel<kEpiRes>, conv_halo_pair_kernel<kEpiStats>,
                                  conv_halo_pair_kernel<kEpiGnBwd>};
    const int fi = env_int("DDPM_HALO_SPECIALIZE", 1) != 0 ? 1 + epi_mode_of(pp.epi) : 0;
    static size_t pconfigured[5] = {0, 0, 0, 0, 0};
    if (psmem > pconfigured[fi]) {
      DDPM_CUDA(cudaFuncSetAttribute(fns[fi], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
      pconfigured[fi] = psmem;
    }
    int pgrid = 2 * pp.total_tiles < kNumSMs ? 2 * pp.total_tiles : (kNumSMs / 2) * 2;
    pp.sched_ctr = nullptr;
    if (env_int("DDPM_HALO_DYNAMIC", 1) != 0 && pp.total_tiles > pgrid / 2) {
      // one self-resetting counter per launch, round-robin over a small pool: launches on ONE stream never overlap,
      // and a captured graph keeps the slot it was captured with
      static unsigned int* pool = nullptr;
      static std::atomic<unsigned int> seq{0};
      if (pool == nullptr) DDPM_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&pool), g_halo_tile_ctr));
      pp.sched_ctr = pool + (seq.fetch_add(1) % kTileCtrPool);
    }
    fns[fi]<<<pgrid, kHaloPairThreads, psmem, stream>>>(ma0, ma1, mb, pp);
    return check_launch("conv_halo_pair_kernel");
  }
  if (int e = make_wgt_map(&mb, a->wgt, k_total, a->ldw, a->cout, 128)) return e;
  static size_t configured = 0;
  if (smem > configured) {
    DDPM_CUDA(cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
  conv_halo_kernel<<<grid, kHaloThreads, smem, stream>>>(ma0, ma1, mb, p);
  return check_launch("conv_halo_kernel");
}

}  // namespace ddpm

extern "C" int ddpm_conv_halo_strips(int w) { return ddpm::conv_halo_strips(w); }
