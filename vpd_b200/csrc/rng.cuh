// Counter-based random numbers for the augmentation kernels (assemble.cu, augment.cu).
#pragma once
#include <cuda_runtime.h>

namespace vpd {

// Counter-based normal generator for the noise augmentation: Philox4x32-10 keyed by the
// seed, counter = (element index, frame), Box-Muller on the first two words.
__device__ __forceinline__ float philox_normal(unsigned long long seed, unsigned int idx,
                                               unsigned int frame) {
  unsigned int c0 = idx, c1 = frame, c2 = 0x1234u, c3 = 0u;
  unsigned int k0 = static_cast<unsigned int>(seed), k1 = static_cast<unsigned int>(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  const float u1 = (static_cast<float>(c0 >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0, 1)
  const float u2 = (static_cast<float>(c1 >> 8) + 0.5f) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * __logf(u1)) * __cosf(6.28318530718f * u2);
}

}  // namespace vpd
