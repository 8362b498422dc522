// Device helpers for the "PM" split-bf16 plane-major activation format (see umma_conv.cuh).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace pcgc {

// 8 float32 channel values -> hi cell + lo cell (16 bytes each): hi = bf16(v), lo = bf16(v - hi).
__device__ __forceinline__ void split_store(__nv_bfloat16* hi_cell, __nv_bfloat16* lo_cell, const float* v) {
  __align__(16) __nv_bfloat16 h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    h[i] = __float2bfloat16_rn(v[i]);
    l[i] = __float2bfloat16_rn(v[i] - __bfloat162float(h[i]));
  }
  *reinterpret_cast<uint4*>(hi_cell) = *reinterpret_cast<const uint4*>(h);
  *reinterpret_cast<uint4*>(lo_cell) = *reinterpret_cast<const uint4*>(l);
}

__device__ __forceinline__ void load_cell_sum(const __nv_bfloat16* hi_cell, const __nv_bfloat16* lo_cell, float* v) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(hi_cell));
  const uint4 b = __ldg(reinterpret_cast<const uint4*>(lo_cell));
  const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&a);
  const __nv_bfloat16* l = reinterpret_cast<const __nv_bfloat16*>(&b);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __bfloat162float(h[i]) + __bfloat162float(l[i]);
}

// hi cell + lo cell already in registers -> 8 float32 values
__device__ __forceinline__ void cell_sum(const uint4& a, const uint4& b, float* v) {
  const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&a);
  const __nv_bfloat16* l = reinterpret_cast<const __nv_bfloat16*>(&b);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __bfloat162float(h[i]) + __bfloat162float(l[i]);
}

// half a cell (4 channels)
__device__ __forceinline__ void load_half_cell_sum(const __nv_bfloat16* hi4, const __nv_bfloat16* lo4, float* v) {
  const uint2 a = __ldg(reinterpret_cast<const uint2*>(hi4));
  const uint2 b = __ldg(reinterpret_cast<const uint2*>(lo4));
  const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&a);
  const __nv_bfloat16* l = reinterpret_cast<const __nv_bfloat16*>(&b);
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __bfloat162float(h[i]) + __bfloat162float(l[i]);
}

}  // namespace pcgc
