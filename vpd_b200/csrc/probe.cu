// Hardware-behaviour probe (test-only): can a K-major SWIZZLE_128B UMMA operand
// start at an arbitrary 128-byte row of a swizzled shared-memory patch, with an
// arbitrary stride between 8-row groups? This decides whether the 3x3 taps of a
// convolution can all be read from ONE halo patch in shared memory.
#include "common.cuh"
#include "tma_host.h"

namespace vpd {

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __nv_bfloat16* __restrict__ src, int rows, int row_start, int sbo_bytes,
                  int base_offset_mode, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;                          // rows x 128 B, swizzled like TMA would
  uint8_t* sB = smem + ((rows * 128 + 1023) & ~1023);   // 64 x 128 B identity
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x;
  // address-based 128B swizzle: 16-byte chunk index ^= (row % 8), base 1024-aligned
  for (int i = tid; i < rows * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    const uint4 v = reinterpret_cast<const uint4*>(src)[r * 8 + c];
    *reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
  for (int i = tid; i < 64 * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    __nv_bfloat16 v[8];
    for (int j = 0; j < 8; ++j) v[j] = __float2bfloat16_rn((c * 8 + j) == r ? 1.f : 0.f);
    *reinterpret_cast<uint4*>(sB + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<uint4*>(v);
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (tid < 32) {
    tmem_alloc(&tmem_ptr, 64);
    tmem_relinquish();
  }
  fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (tid == 0) {
    const uint32_t a_addr = smem_u32(sA) + row_start * 128;
    uint64_t adesc = make_smem_desc(a_addr, 16, sbo_bytes);
    if (base_offset_mode & 1) adesc |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
    const uint64_t bdesc = make_smem_desc(smem_u32(sB), 16, 1024);
    // base_offset_mode bit 1: the A rows hold fp16 (a_format field = 0) while B stays bf16 - does
    // kind::f16 take a different 16-bit format per operand?
    const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0) & ~((base_offset_mode & 2) ? (1u << 7) : 0u);
    for (int k = 0; k < 4; ++k) umma_bf16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = tid >> 5, lane = tid & 31;
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 64);
}

int umma_probe(const __nv_bfloat16* src, int rows, int row_start, int sbo_bytes,
               int base_offset_mode, float* out, cudaStream_t stream) {
  const int smem = rows * 128 + 64 * 128 + 4096;
  VPD_REQUIRE(smem <= 200 * 1024, "probe: too many rows");
  VPD_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      smem));
  VPD_CHECK_CUDA(launch_kernel(umma_probe_kernel, dim3(1), dim3(128), smem, stream, src, rows,
                               row_start, sbo_bytes, base_offset_mode, out));
  VPD_LAUNCHED(1);
  return 0;
}

}  // namespace vpd
