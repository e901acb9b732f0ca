// Philox4x32-10 counter-based generator (Salmon et al., SC'11) -> one FP64 uniform in [0,1) per counter.
// key = (seed_lo, seed_hi), counter = (idx_lo, idx_hi, 0, 0); u = (((x0 << 32) | x1) >> 11) * 2^-53.
// tests/philox_ref.py holds the numpy twin used to check accept decisions bit-for-bit.
#pragma once
#include <cstdint>

namespace cgs {

__host__ __device__ inline double philox_uniform_f64(uint64_t seed, uint64_t index) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32), c2 = 0u, c3 = 0u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0;
    const uint64_t p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  const uint64_t bits = (((uint64_t)c0 << 32) | (uint64_t)c1) >> 11;
  return (double)bits * (1.0 / 9007199254740992.0);
}

}  // namespace cgs
