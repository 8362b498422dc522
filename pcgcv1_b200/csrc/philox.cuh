// Counter-based noise of the "noise" quantisation mode, shared by entropy.cu and train.cu.
#pragma once
#include <stdint.h>

namespace pcgc {

// "noise" quantisation (entropy_model.py:105-107, conditional_entropy_model.py:62-64): x + U(-1/2, 1/2).  Counter-based
// Philox4x32-10 keyed by the seed, counter = element index / 4, so the draw of an element does not depend on the launch
// shape and the oracle (oracle/entropy.py:philox_uniform) reproduces it bit for bit.
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  uint32_t c[4] = {c0, c1, 0u, 0u};
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
// 4 uniforms in [-1/2, 1/2) for the vector of elements 4*v .. 4*v+3: (top 24 bits) * 2^-24 - 1/2 (exact in float)
__host__ __device__ __forceinline__ void noise4(uint64_t seed, uint64_t v, float u[4]) {
  uint32_t r[4];
  philox4x32_10((uint32_t)v, (uint32_t)(v >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), r);
#pragma unroll
  for (int k = 0; k < 4; ++k) u[k] = (float)(r[k] >> 8) * 5.9604644775390625e-8f - 0.5f;
}

}  // namespace pcgc
