// Shared epilogue of the tcgen05 conv kernels (conv_igemm.cu, conv_halo.cu): one thread owns one output pixel
// (accumulator row) and walks 32-column chunks read from TMEM.
//
// Fused here, in the order diffusers applies them (ResnetBlock2D.forward, App. A.2 of SURVEY.md):
//   + bias[co]  + temb[n, co] (time_emb_proj output)  + residual[pix, co]
// and, for the dgrad GEMM that produces the gradient w.r.t. a GroupNorm(+SiLU) OUTPUT, the first half of the
// GroupNorm backward:  dz = dy * act'(xhat*gamma + beta)  is written INSTEAD of dy, and the per-(sample, channel)
// sums  S1 = sum dz,  S2 = sum dz * x  are reduced over the tile (butterfly over the 32 rows of a warp) and added to
// gsums[n][c][2].  The epilogue warps are idle for most of a tile's mainloop, so the SiLU-derivative (1 MUFU + ~10
// FP32 ops per element) costs nothing there, and the remaining GroupNorm backward is a single transcendental-free
// streaming pass (gn_bwd_apply in groupnorm.cu): 6 B/elem instead of 10.
#pragma once
#include "common.cuh"

namespace ddpm {

struct EpiParams {
  __nv_bfloat16* out;
  long long ldo;
  float* out_f32;          // optional fp32 output instead of bf16 (same indexing, ldo)
  const float* bias;
  const float* temb;
  int ld_temb;
  const __nv_bfloat16* res;
  long long ldr;
  int Cout;
  // ---- GroupNorm-backward fusion (all NULL/0 when unused) ----
  const __nv_bfloat16* gx0;   // GroupNorm INPUT x (channels [0, gc0)) ...
  const __nv_bfloat16* gx1;   // ... and [gc0, Cout) when the input was a channel concat
  long long gld0, gld1;
  int gc0;
  const float* gcoef;         // [N][Cout/2][4] = (ka0, ka1, kb0, kb1): z = x*ka + kb  (written by ddpm_gn_fwd)
  float* gsums;               // [N][Cout][2] += (sum dz, sum dz*x)   (gstats: [N][Cout/4][2] += (sum out, sum out^2))
  int gsilu;
  int gstats;                 // 1: gsums receives the per-(sample, channel) moments of the STORED output -- the
                              // statistics of the GroupNorm that consumes this tensor (no gx / gcoef involved)
  int wide;                   // every pointer / stride above allows 32-byte (256-bit) row accesses
  int split;                  // fp32-faithful mode: out / res rows are [hi (Cout) | lo (Cout)], value = hi + lo
};

// 256-bit global accesses (sm_100: LDG/STG.E.256).  One thread owns one pixel row here, so a warp-wide access touches
// 32 different 128-byte lines whatever its width; moving 32 instead of 16 bytes per lane halves the number of L1
// wavefronts per tile, which is what bounded the fused epilogue (ncu: LSU data-pipe 65 %, profiles/r1_halo_fused.md).
__device__ __forceinline__ void ld256(const void* p, uint32_t* r) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void st256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// 64 bytes (32 bf16) from / to a 16-byte-aligned address; `wide` = the address is known to be 32-byte aligned
__device__ __forceinline__ void ld64B(const void* p, uint32_t* w, bool wide) {
  if (wide) {
    ld256(p, w);
    ld256(static_cast<const uint8_t*>(p) + 32, w + 8);
  } else {
    const uint4* q = static_cast<const uint4*>(p);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 t = q[j];
      w[4 * j] = t.x; w[4 * j + 1] = t.y; w[4 * j + 2] = t.z; w[4 * j + 3] = t.w;
    }
  }
}
__device__ __forceinline__ void st64B(void* p, const uint32_t* w, bool wide) {
  if (wide) {
    st256(p, w);
    st256(static_cast<uint8_t*>(p) + 32, w + 8);
  } else {
    uint4* q = static_cast<uint4*>(p);
#pragma unroll
    for (int j = 0; j < 4; ++j) q[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
  }
}

__device__ __forceinline__ float epi_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Column totals over the 32 lanes of a warp: a[j] (j = column) in, lane L returns the total of column L.
__device__ __forceinline__ float warp_column_sums(float (&a)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? a[i] : a[i + s];
      const float keep = up ? a[i + s] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return a[0];
}

// Totals over the 32 lanes of a warp of EIGHT values per lane (4-channel granules of a 32-column chunk): three
// transposing butterfly steps (8 -> 4 -> 2 -> 1 values, 7 shuffles) then two plain ones.  Every lane returns the total of
// granule epi_granule_of(lane); the four lanes that share a granule hold the same value.  9 shuffles instead of the 31
// of warp_column_sums -- GroupNorm groups are multiples of 4 channels in every configuration of this UNet (4 ... 32), so
// per-granule moments are all the consuming GroupNorm needs.
__device__ __forceinline__ float warp_granule_sums(float (&a)[8], int lane) {
#pragma unroll
  for (int s = 16, m = 4; s >= 4; s >>= 1, m >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < m; ++i) {
      const float send = up ? a[i] : a[i + m];
      const float keep = up ? a[i + m] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  float v = a[0];
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
__device__ __forceinline__ int epi_granule_of(int lane) { return (lane >> 2) & 7; }

// Build-time switches of the epilogue arithmetic.  All four cut the static instruction count of a chunk row (902 -> 805
// SASS instructions with all of them on), and none of them paid: timed against each other in one process per library
// (profiles/epi_variants.sh, profiles/r2_epilogue_variants.md) they move the fused variants of the 128->128 conv by
// -5 ... +15 %, mostly in branches they do not touch -- the kernel is ONE function whose variants share a register
// allocation (168 registers, 13-22 spilled), so an edit in one branch reshuffles the spills of the others.  They stay
// off; what did pay is the two-element conversion in the forward-statistics branch below (-17 / -22 us).
#ifndef EPI_PACKED_ADDS
#define EPI_PACKED_ADDS 0      // bias / temb / residual adds as packed fp32x2 instructions
#endif
#ifndef EPI_PACKED_GNACC
#define EPI_PACKED_GNACC 0     // GroupNorm-backward column partials with packed fp32x2 instructions
#endif
#ifndef EPI_CSUM_PAIRS
#define EPI_CSUM_PAIRS 0       // forward-statistics partials kept as pairs until the per-chunk reduction
#endif
#ifndef EPI_SILU_FORM2
#define EPI_SILU_FORM2 0       // silu'(z) = (1 + t + (z/2)(1 - t^2)) / 2 instead of the sigmoid form
#endif

// (a, b) += (x, y) as ONE packed fp32x2 add
__device__ __forceinline__ void epi_add2(float& a, float& b, float x, float y) {
#if EPI_PACKED_ADDS
  const float2 r = __fadd2_rn(make_float2(a, b), make_float2(x, y));
  a = r.x;
  b = r.y;
#else
  a += x;
  b += y;
#endif
}

// Epilogue specialisation of a kernel instantiation: the fused variants (statistics, GroupNorm backward, residual) are
// runtime branches of ONE epilogue, and a kernel that carries all of them shares one register allocation between them
// (168 registers, 13+ spilled in the CTA-pair halo conv).  Forcing the flags a launch cannot have to constants lets
// the compiler drop the other variants' code and registers from that instantiation.
enum : int { kEpiAny = -1, kEpiLean = 0, kEpiRes = 1, kEpiStats = 2, kEpiGnBwd = 3 };
template <int MODE>
__device__ __forceinline__ EpiParams epi_specialize(EpiParams e) {
  if (MODE == kEpiLean || MODE == kEpiRes) {      // no fused reduction
    e.gsums = nullptr; e.gstats = 0; e.gx0 = nullptr; e.gx1 = nullptr; e.gcoef = nullptr; e.gsilu = 0;
  }
  if (MODE == kEpiLean) { e.res = nullptr; e.split = 0; }
  if (MODE == kEpiStats) { e.gstats = 1; e.gx0 = nullptr; e.gx1 = nullptr; e.gcoef = nullptr; e.gsilu = 0; }
  if (MODE == kEpiGnBwd) e.gstats = 0;
  return e;
}
// host side: the specialisation a launch with these parameters may use
inline int epi_mode_of(const EpiParams& e) {
  if (e.gsums == nullptr) return (e.res != nullptr || e.split) ? kEpiRes : kEpiLean;
  return e.gstats ? kEpiStats : kEpiGnBwd;
}

struct EpiX {
  uint32_t w[16];   // 32 bf16 channels at this thread's pixel: the GroupNorm input (backward fusion) or the residual
};
// does the prefetch slot carry the GroupNorm input (true) or the residual (false)?
__device__ __forceinline__ bool epi_slot_is_gn(const EpiParams& e) { return e.gsums != nullptr && !e.gstats; }
// Issue the loads of the per-pixel side input of (pix, col .. col+31) -- the GroupNorm input of the backward fusion, or
// else the residual -- EARLY (before waiting on the accumulator / while the previous chunk is processed) so that the
// DRAM latency is off the epilogue's critical path.  (A residual loaded inside epi_chunk cost +0.055 ms on a 0.21 ms
// 128->128 conv at 128^2: four exposed ~1 us round trips per tile against a 4.8 us mainloop.)
__device__ __forceinline__ void epi_load_x(const EpiParams& e, bool valid, long long pix, int col, EpiX& x) {
  if (epi_slot_is_gn(e) && valid && col < e.Cout) {
    const __nv_bfloat16* xp = (col < e.gc0) ? e.gx0 + pix * e.gld0 + col : e.gx1 + pix * e.gld1 + (col - e.gc0);
    ld64B(xp, x.w, e.wide != 0);
  } else if (!epi_slot_is_gn(e) && e.res != nullptr && valid && col < e.Cout) {
    ld64B(e.res + pix * e.ldr + col, x.w, e.wide != 0);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) x.w[j] = 0u;
  }
}

// L2 prefetch of the side input (GroupNorm x or residual) of one pixel row, `ncols` channels from `col`: no register
// destination, so it can run a whole tile ahead.  ncu (profiles/r2_epilogue_stalls.md): with the register prefetch one
// 32-column chunk ahead the epilogue warps still sat in long-scoreboard stalls on the first use of these loads -- a chunk
// takes ~700 clk, a DRAM round trip under load ~2000 -- and the fused epilogues cost +25 ... +60 % of the kernel.
__device__ __forceinline__ void epi_prefetch_side(const EpiParams& e, bool valid, long long pix, int col, int ncols) {
  if (!valid || col >= e.Cout) return;
  const __nv_bfloat16* p0 = nullptr;
  if (epi_slot_is_gn(e))
    p0 = (col < e.gc0) ? e.gx0 + pix * e.gld0 + col : e.gx1 + pix * e.gld1 + (col - e.gc0);
  else if (e.res != nullptr)
    p0 = e.res + pix * e.ldr + col;
  if (p0 == nullptr) return;
  for (int b = 0; b < ncols * 2; b += 128)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const uint8_t*>(p0) + b));
}

// Per-thread column partials of the fused reductions, accumulated over the row blocks that share a 32-column chunk and
// reduced across the warp ONCE (epi_reduce): GroupNorm-backward fusion a[j] = dz, b[j] = dz * x per column;
// forward statistics a[2k] + a[2k+1] = sum, b[2k] + b[2k+1] = sum of squares per 4-channel granule (k < 8).
struct EpiAcc {
  float a[32];
  float b[32];
};
__device__ __forceinline__ void epi_acc_zero(EpiAcc& acc) {
#pragma unroll
  for (int j = 0; j < 32; ++j) acc.a[j] = acc.b[j] = 0.f;
}

// v[32]: fp32 accumulators of (row = this thread's pixel, columns col .. col+31).  `n`, `pix` describe the pixel;
// rows with valid == false are not stored and contribute zero to the sums.  Applies bias / temb / residual and, with
// GroupNorm fusion, turns v into dz; the fused reductions' per-column partials are ADDED to acc.
// PREF: xin was filled by epi_load_x for this (pix, col) -- the residual, when there is one and the slot is not taken by
// the GroupNorm input, comes from there instead of a load issued here.
template <bool GN, bool PREF>
__device__ __forceinline__ void epi_math(const EpiParams& e, float (&v)[32], bool valid, int n, long long pix, int col,
                                         const EpiX& xin, EpiAcc& acc) {
  const bool col_ok = col < e.Cout;
  if (valid && col_ok) {
    if (e.bias) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = *reinterpret_cast<const float4*>(e.bias + col + j);
        epi_add2(v[j], v[j + 1], b.x, b.y);
        epi_add2(v[j + 2], v[j + 3], b.z, b.w);
      }
    }
    if (e.temb) {
      const float* t = e.temb + static_cast<long long>(n) * e.ld_temb + col;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = *reinterpret_cast<const float4*>(t + j);
        epi_add2(v[j], v[j + 1], b.x, b.y);
        epi_add2(v[j + 2], v[j + 3], b.z, b.w);
      }
    }
    if (e.res) {
      if (PREF && !epi_slot_is_gn(e)) {
#pragma unroll
        for (int j = 0; j < 16; ++j) epi_add2(v[2 * j], v[2 * j + 1], bf16lo_f(xin.w[j]), bf16hi_f(xin.w[j]));
      } else {
        uint32_t rw[16];
        ld64B(e.res + pix * e.ldr + col, rw, e.wide != 0);
#pragma unroll
        for (int j = 0; j < 16; ++j) epi_add2(v[2 * j], v[2 * j + 1], bf16lo_f(rw[j]), bf16hi_f(rw[j]));
      }
      if (e.split) {      // low halves of the split residual
        uint32_t rw[16];
        ld64B(e.res + pix * e.ldr + e.Cout + col, rw, e.wide != 0);
#pragma unroll
        for (int j = 0; j < 16; ++j) epi_add2(v[2 * j], v[2 * j + 1], bf16lo_f(rw[j]), bf16hi_f(rw[j]));
      }
    }
  }
  if (GN && e.gsums != nullptr && e.gstats) {
    // ---- statistics of the consuming GroupNorm: moments of the bf16 values that are stored, per 4-channel granule
    // (in-thread fold of 4 columns; the 9-shuffle granule butterfly runs once per chunk in epi_reduce) ----
    // Two-element conversions, then packed fp32x2 adds / FMAs on the unpacked words: ~110 instead of ~160 instructions
    // per chunk row.  (This epilogue has no side input to wait for -- it is issue-bound: +57 us on a 221 us 128->128
    // conv at 128^2.)  v is left holding the rounded values, which epi_store packs again exactly.
    if (valid && col_ok) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t w0 = pack2_bf16_(v[4 * k], v[4 * k + 1]), w1 = pack2_bf16_(v[4 * k + 2], v[4 * k + 3]);
        const float2 p0 = make_float2(bf16lo_f(w0), bf16hi_f(w0));
        const float2 p1 = make_float2(bf16lo_f(w1), bf16hi_f(w1));
        v[4 * k] = p0.x; v[4 * k + 1] = p0.y; v[4 * k + 2] = p1.x; v[4 * k + 3] = p1.y;
#if EPI_CSUM_PAIRS
        // per-granule partials stay PAIRS (a[2k], a[2k+1]) until epi_reduce folds them once per chunk
        const float2 sa = __fadd2_rn(__fadd2_rn(p0, p1), make_float2(acc.a[2 * k], acc.a[2 * k + 1]));
        const float2 sq = __ffma2_rn(p1, p1, __ffma2_rn(p0, p0, make_float2(acc.b[2 * k], acc.b[2 * k + 1])));
        acc.a[2 * k] = sa.x; acc.a[2 * k + 1] = sa.y;
        acc.b[2 * k] = sq.x; acc.b[2 * k + 1] = sq.y;
#else
        const float2 sa = __fadd2_rn(p0, p1);
        const float2 sq = __ffma2_rn(p1, p1, __fmul2_rn(p0, p0));
        acc.a[2 * k] += sa.x + sa.y;
        acc.b[2 * k] += sq.x + sq.y;
#endif
      }
    }
  } else if (GN && e.gsums != nullptr) {
    // ---- GroupNorm backward, part 1 (packed fp32x2 math) ----
    if (valid && col_ok) {
      float x[32];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        x[2 * j] = bf16lo_f(xin.w[j]);
        x[2 * j + 1] = bf16hi_f(xin.w[j]);
      }
      if (e.gsilu) {
        const float4* cf = reinterpret_cast<const float4*>(e.gcoef) + (static_cast<long long>(n) * e.Cout + col) / 2;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float4 c4 = cf[j / 2];   // (ka_j, ka_j+1, kb_j, kb_j+1), warp-uniform address
          // silu'(z) = s + z s (1 - s), s = (1 + t) / 2, t = tanh(z / 2)  =>  silu'(z) = (1 + t + (z/2)(1 - t^2)) / 2:
          // 6 packed instructions + 2 MUFU per channel pair (the sigmoid form took 7 + 2)
          const float2 z = __ffma2_rn(make_float2(x[j], x[j + 1]), make_float2(c4.x, c4.y), make_float2(c4.z, c4.w));
          const float2 h = __fmul2_rn(z, make_float2(0.5f, 0.5f));
#if EPI_SILU_FORM2
          const float2 t = make_float2(epi_tanh(h.x), epi_tanh(h.y));
          const float2 w = __ffma2_rn(make_float2(-t.x, -t.y), t, make_float2(1.f, 1.f));
          const float2 r = __ffma2_rn(h, w, t);
          const float2 vh = __fmul2_rn(make_float2(v[j], v[j + 1]), make_float2(0.5f, 0.5f));
          const float2 dz = __ffma2_rn(vh, r, vh);
#else
          const float2 sg = __ffma2_rn(make_float2(epi_tanh(h.x), epi_tanh(h.y)), make_float2(0.5f, 0.5f),
                                       make_float2(0.5f, 0.5f));
          const float2 om = __ffma2_rn(sg, make_float2(-1.f, -1.f), make_float2(1.f, 1.f));
          const float2 u = __ffma2_rn(z, om, make_float2(1.f, 1.f));
          const float2 dz = __fmul2_rn(__fmul2_rn(make_float2(v[j], v[j + 1]), sg), u);
#endif
          v[j] = dz.x;
          v[j + 1] = dz.y;
        }
      }
#if EPI_PACKED_GNACC
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float2 ra = __fadd2_rn(make_float2(acc.a[j], acc.a[j + 1]), make_float2(v[j], v[j + 1]));
        acc.a[j] = ra.x;
        acc.a[j + 1] = ra.y;
        const float2 r = __ffma2_rn(make_float2(v[j], v[j + 1]), make_float2(x[j], x[j + 1]),
                                    make_float2(acc.b[j], acc.b[j + 1]));
        acc.b[j] = r.x;
        acc.b[j + 1] = r.y;
      }
#else
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        acc.a[j] += v[j];
        acc.b[j] = fmaf(v[j], x[j], acc.b[j]);
      }
#endif
    }
  }
}

// stores v (after epi_math) as bf16 (or fp32 / split-bf16) at (pix, col .. col+31)
__device__ __forceinline__ void epi_store(const EpiParams& e, const float (&v)[32], bool valid, long long pix, int col) {
  if (!(valid && col < e.Cout)) return;
  if (e.out_f32) {
    float4* op = reinterpret_cast<float4*>(e.out_f32 + pix * e.ldo + col);
#pragma unroll
    for (int j = 0; j < 8; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
    uint32_t ow[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) ow[j] = pack2_bf16_(v[2 * j], v[2 * j + 1]);
    st64B(e.out + pix * e.ldo + col, ow, e.wide != 0);
    if (e.split) {      // lo = bf16(v - hi): together 16 mantissa bits of the fp32 accumulator
      uint32_t lw[16];
#pragma unroll
      for (int j = 0; j < 16; ++j)
        lw[j] = pack2_bf16_(v[2 * j] - bf16lo_f(ow[j]), v[2 * j + 1] - bf16hi_f(ow[j]));
      st64B(e.out + pix * e.ldo + e.Cout + col, lw, e.wide != 0);
    }
  }
}

// Cross-lane reduction of the per-column partials: this lane's totals are ADDED to (t1, t2) (the layout epi_flush_sums
// expects).  All 32 lanes of the warp must call this together, and the warp's valid rows must belong to ONE sample.
template <bool GN>
__device__ __forceinline__ void epi_reduce(const EpiParams& e, int lane, EpiAcc& acc, float& t1, float& t2) {
  if (!GN || e.gsums == nullptr) return;
  if (e.gstats) {
    float g1[8], g2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { g1[k] = acc.a[2 * k] + acc.a[2 * k + 1]; g2[k] = acc.b[2 * k] + acc.b[2 * k + 1]; }
    t1 += warp_granule_sums(g1, lane);
    t2 += warp_granule_sums(g2, lane);
  } else {
    t1 += warp_column_sums(acc.a, lane);
    t2 += warp_column_sums(acc.b, lane);
  }
}

// One chunk, one row block: math + reduction + store (kernels whose row blocks do not share a chunk's columns).
template <bool GN = true, bool PREF = GN>
__device__ __forceinline__ void epi_chunk(const EpiParams& e, float (&v)[32], bool valid, int n, long long pix, int col,
                                          int lane, float& t1, float& t2, const EpiX& xin) {
  EpiAcc acc;
  if (GN) epi_acc_zero(acc);
  epi_math<GN, PREF>(e, v, valid, n, pix, col, xin, acc);
  epi_reduce<GN>(e, lane, acc, t1, t2);
  epi_store(e, v, valid, pix, col);
}

// Adds this lane's column totals to gsums[n][col + lane][0..1].  `n_valid` < 0: the warp had no valid row.
__device__ __forceinline__ void epi_flush_sums(const EpiParams& e, int n_valid, int col, int lane, float t1, float t2) {
  if (e.gsums == nullptr || n_valid < 0 || col >= e.Cout) return;
  if (e.gstats) {     // forward statistics: [n][Cout / 4][2], one lane per 4-channel granule
    if ((lane & 3) != 0) return;
    float* dst = e.gsums + (static_cast<long long>(n_valid) * (e.Cout >> 2) + (col >> 2) + epi_granule_of(lane)) * 2;
    atomicAdd(dst, t1);
    atomicAdd(dst + 1, t2);
    return;
  }
  float* dst = e.gsums + (static_cast<long long>(n_valid) * e.Cout + col + lane) * 2;
  atomicAdd(dst, t1);
  atomicAdd(dst + 1, t2);
}
// sample index of the warp's valid rows (host guarantees they agree), or -1
__device__ __forceinline__ int epi_warp_sample(bool valid, int n) {
  const unsigned vm = __ballot_sync(0xffffffffu, valid);
  if (vm == 0u) return -1;
  return __shfl_sync(0xffffffffu, n, __ffs(vm) - 1);
}

}  // namespace ddpm
