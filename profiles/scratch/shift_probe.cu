// Hardware probe (bring-up tool, not on the product path): does a tcgen05 K-major SWIZZLE_128B operand descriptor
// whose start address is shifted by a whole number of 128-byte rows (i.e. NOT aligned to the 1024-byte swizzle
// repeat) read the rows TMA wrote, and which `base_offset` encoding does it need?  The halo-resident 3x3 conv
// kernel applies its taps exactly this way.  D[128 x 64] = A[shift : shift+128, 0:64] * B[64 x 64]^T.
#include "common.cuh"

#include "../../include/ddpm_b200.h"

namespace ddpm {

__global__ void __launch_bounds__(128, 1)
shift_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   float* __restrict__ out, int shift, int base_offset_mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 256 rows x 128 B
  uint8_t* sB = smem + 256 * 128;     // 64 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 64 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], 256 * 128 + 64 * 128);
    tma_load_2d(sA, &tmA, &bars[0], 0, 0);
    tma_load_2d(sB, &tmB, &bars[0], 0, 0);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, false, false);
    const uint32_t a_addr = smem_u32(sA) + shift * 128;
    const uint32_t b_addr = smem_u32(sB);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
      if (base_offset_mode == 1) da |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
      const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
      umma_bf16(tmem_base, da, db, idesc, k != 0 ? 1u : 0u);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[row * 64 + c * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace ddpm

using namespace ddpm;

extern "C" int ddpm_debug_shift_probe(const void* a /*bf16 [256][64]*/, const void* b /*bf16 [64][64]*/,
                                      float* out /*[128][64]*/, int shift, int base_offset_mode, void* stream) {
  DDPM_REQUIRE(a && b && out && shift >= 0 && shift <= 128, "ddpm_debug_shift_probe: bad argument");
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  DDPM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
  DDPM_REQUIRE(f && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled unavailable");
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(f);
  CUtensorMap ma, mb;
  cuuint64_t dims_a[2] = {64, 256}, dims_b[2] = {64, 64}, strides[1] = {128};
  cuuint32_t box_a[2] = {64, 256}, box_b[2] = {64, 64}, es[2] = {1, 1};
  CUresult r1 = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a), dims_a, strides, box_a, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(b), dims_b, strides, box_b, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DDPM_REQUIRE(r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS, "tensor map encode failed (%d, %d)", (int)r1, (int)r2);
  const int smem = 256 * 128 + 64 * 128 + 64 + 1024;
  DDPM_CUDA(cudaFuncSetAttribute(shift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  shift_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(ma, mb, out, shift, base_offset_mode);
  return check_launch("shift_probe_kernel");
}
