// Counter-based uniform random numbers for the two noise sources of a training step
// (utils/network.py:4516 latent noise, :610 / :679 kernel noise at q = 1), drawn INSIDE the kernels
// that consume them: Philox4x32-10 (Salmon et al., SC'11; the generator behind torch.rand on CUDA),
// keyed by a host seed, indexed by (element, stream id, device-resident step counter).  No noise
// tensor, no generator kernels and no seed/offset bookkeeping kernels in the captured step; forward
// and backward kernels of one step regenerate identical values from the same counter.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nvf {
namespace rng {

enum Stream : uint32_t { kLatentNoise = 1, kKernelNoise = 2 };

__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
  c[1] = (uint32_t)p1;
  c[3] = (uint32_t)p0;
  c[0] = n0;
  c[2] = n2;
}

// U[0,1) with 24 random bits: value number `elem` of stream `stream` at step `step`
__host__ __device__ __forceinline__ float uniform(uint64_t seed, uint64_t step, uint32_t stream, uint64_t elem) {
  uint32_t c[4] = {(uint32_t)(elem >> 2), (uint32_t)(elem >> 34) ^ (stream << 24), (uint32_t)step, (uint32_t)(step >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return (float)(c[elem & 3] >> 8) * (1.0f / 16777216.0f);
}

}  // namespace rng
}  // namespace nvf
