// Shared helpers for the sm_100a DDPM kernels: error plumbing for the C-ABI and
// thin inline-PTX wrappers (mbarrier, TMA, tcgen05/TMEM).  No torch types here.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

// ---------------------------------------------------------------------------------------------
// C-ABI error plumbing (thread-local last error, see include/ddpm_b200.h)
// ---------------------------------------------------------------------------------------------
namespace ddpm {
void set_last_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError -> 0 / DDPM_ERR_CUDA
}  // namespace ddpm

#define DDPM_OK 0
#define DDPM_ERR_INVALID (-1)
#define DDPM_ERR_CUDA (-2)
#define DDPM_ERR_UNSUPPORTED (-3)

#define DDPM_REQUIRE(cond, ...)                \
  do {                                         \
    if (!(cond)) {                             \
      ddpm::set_last_error(__VA_ARGS__);       \
      return DDPM_ERR_INVALID;                 \
    }                                          \
  } while (0)

#define DDPM_CUDA(call)                                                                  \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      ddpm::set_last_error("%s failed: %s", #call, cudaGetErrorString(e__));             \
      return DDPM_ERR_CUDA;                                                              \
    }                                                                                    \
  } while (0)

struct ddpm_conv_args;

namespace ddpm {

constexpr int kNumSMs = 148;

// host-side helpers shared by the tcgen05 kernels (defined in conv_igemm.cu / conv_halo.cu)
int make_act_map(CUtensorMap* out, const void* ptr, int c, long long ld, int n, int h, int w, int wb, int hb, int nb);
int make_wgt_map(CUtensorMap* out, const void* ptr, long long k_total, long long ld, int rows, int box_rows);
int make_map4(CUtensorMap* out, const void* ptr, const long long dims[4], const long long strides[3], const int box[4]);
int env_int(const char* name, int dflt);
// halo-resident 3x3 conv: returns 1 when the problem is not eligible (caller falls back to the generic kernel)
int launch_conv_halo(const ::ddpm_conv_args* a, cudaStream_t stream);
bool conv_uses_halo(const ::ddpm_conv_args* a);   // would ddpm_conv_gemm take the halo-resident kernel for this problem?
int conv_halo_strips(int w);                      // column strips per image row at this width (0: not served)
}
struct ddpm_wgrad_args;
namespace ddpm {
// row-resident 3x3 wgrad: returns 1 when not eligible
int launch_wgrad_row(const ::ddpm_wgrad_args* a, cudaStream_t stream);
struct EpiParams;
int fill_epilogue(EpiParams* e, const ::ddpm_conv_args* a);

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
// d/dx silu(x) = s + x*s*(1-s), s = sigmoid(x)
__device__ __forceinline__ float silu_grad_f(float x) {
  float s = 1.0f / (1.0f + __expf(-x));
  return s * (1.0f + x * (1.0f - s));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Eight bf16 values moved as ONE 128-bit access.  (A struct of four __nv_bfloat162 is copied member-wise by nvcc --
// four 32-bit LDG/STG per vector, 4x the L1/L2 wavefronts -- so the carrier type is the builtin uint4.)
using bf16x8 = uint4;
using bf16x4 = uint2;

__device__ __forceinline__ float bf16lo_f(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi_f(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack2_bf16_(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
  f[0] = bf16lo_f(p.x); f[1] = bf16hi_f(p.x);
  f[2] = bf16lo_f(p.y); f[3] = bf16hi_f(p.y);
  f[4] = bf16lo_f(p.z); f[5] = bf16hi_f(p.z);
  f[6] = bf16lo_f(p.w); f[7] = bf16hi_f(p.w);
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  return make_uint4(pack2_bf16_(f[0], f[1]), pack2_bf16_(f[2], f[3]), pack2_bf16_(f[4], f[5]), pack2_bf16_(f[6], f[7]));
}
__device__ __forceinline__ void unpack4(const bf16x4& p, float* f) {
  f[0] = bf16lo_f(p.x); f[1] = bf16hi_f(p.x);
  f[2] = bf16lo_f(p.y); f[3] = bf16hi_f(p.y);
}
__device__ __forceinline__ bf16x4 pack4(const float* f) {
  return make_uint2(pack2_bf16_(f[0], f[1]), pack2_bf16_(f[2], f[3]));
}
__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) { return pack2_bf16_(a, b); }

// ---------------------------------------------------------------------------------------------
// PTX: shared-memory addresses, mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a mis-programmed pipeline traps (-> CUDA error at the next sync) instead of hanging
// the GPU box.  try_wait suspends in hardware, so the loop count stays small on the good path.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) {
      printf("ddpm_b200: mbarrier wait timed out (block %d,%d thread %d, barrier @smem 0x%x, parity %u)\n", blockIdx.x,
             blockIdx.y, threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// PTX: TMA (cp.async.bulk.tensor), tiled mode, mbarrier completion
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// PTX: tcgen05 / TMEM (cta_group::1)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// One lane of a CONVERGED warp.  The MMA / TMA issuing warps run their loops warp-uniformly and predicate only the
// issuing instruction with this: inside an `if (lane == 0)` region the compiler cannot keep descriptors in uniform
// registers and emits an ELECT / R2UR.BROADCAST retry loop (~17 SASS instructions) around every tcgen05.mma, which made
// the single issuing thread -- not the tensor pipe -- the limiter of the conv kernels (~60 % of peak).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// Register re-partitioning between warpgroups (4 consecutive warps; every warp of the group must execute it): the
// epilogue warps take what the single-lane TMA / MMA issuing warps do not need.
template <uint32_t kRegs>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs));
}
template <uint32_t kRegs>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs));
}

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate.  One thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// PTX: CTA pairs (thread-block cluster of 2, tcgen05 cta_group::2)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// default semantics (.release.cta): a .cluster-scope release costs a full memory barrier per arrival -- measured: the
// peer's producer warp spent ~700 clk per tap in it and became the pipeline's bottleneck (profiles/r1_halo_pair.md)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Publishing a word to the other CTA of the pair through (distributed) shared memory: the writer stores it, then
// arrives with RELEASE at CLUSTER scope on a barrier in the reader's CTA; the reader waits with ACQUIRE at cluster scope
// and loads the word through its shared::cluster address.  (One fence per published word on the writer -- affordable
// once per tile, not once per tap: see mbar_arrive_cluster above.)
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_acquire_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_acquire_cluster(bar, parity)) {
    if (++spins > (1u << 24)) {
      printf("ddpm_b200: mbarrier wait timed out (block %d,%d thread %d, barrier @smem 0x%x, parity %u, cluster scope)\n",
             blockIdx.x, blockIdx.y, threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}
__device__ __forceinline__ uint32_t ld_shared_cluster_u32(uint32_t cluster_addr) {
  uint32_t v;
  asm volatile("ld.volatile.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(cluster_addr) : "memory");
  return v;
}
// TMA loads of a CTA pair: data lands in THIS CTA's shared memory, the transaction bytes are signalled on a barrier
// that may live in the peer CTA (shared::cluster address)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                 int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D (M = 256: rows 0-127 in the leader's TMEM, 128-255 in the peer's) (+)= A * B; each CTA supplies its own 128 rows of
// A and HALF of B (N/2 rows) at the same shared-memory offsets.  Issued by the leader CTA only.
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

// Shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.
//   K-major  operand: rows (M or N index) are 128 B apart, 8-row groups are `sbo` bytes apart.
//   MN-major operand: 64-element MN blocks are `lbo` bytes apart, 8-row K groups are `sbo` bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;  // SWIZZLE_128B
  return d;
}
// Instruction descriptor, kind::f16: bf16 x bf16 -> fp32, shape M x N (K = 16).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

}  // namespace ddpm
