// GPU-side range coding of the per-cube latent strings: the byte streams of
// SymmetricConditional.compress / decompress (models/conditional_entropy_model.py:126-201) are written and read on the
// device, so neither the per-element CDF rows (65 536 x N uint16 per cube) nor the per-element intervals cross PCIe and no
// host thread pool sits between the transforms.  Cubes are independent strings (transform.py:157-168): the coder is
// sequential inside a string and parallel over cubes.
//
// The state machines are range_coder.h's -- the SAME source as the host coder (coder.cpp) -- so a string written here is
// byte-identical to pcgc_range_encode_intervals' and decodes with pcgc_range_decode_rows (tests/test_gpu_coder.py).
//   encode: one THREAD per cube (the encoder is ~20 integer instructions per symbol with one data-dependent branch);
//           CPW cubes per warp keeps the divergence of that branch small and spreads the cubes over the SMs.
//   decode: one WARP per cube.  The symbol search is lane-parallel: lane k holds cdf[k] of the current symbol's row and
//           tests size*cdf[k] <= offset; a ballot gives the symbol without a division or a serial scan.  Rows stream
//           through a double-buffered shared-memory window (cp.async), the next 16-bit word of the string is prefetched.
#include <stdint.h>

#include "common.cuh"
#include "range_coder.h"

namespace pcgc {

namespace {

constexpr int ENC_CPW = 8;            // cubes (= active lanes) per encoder warp

__global__ void __launch_bounds__(32)
range_encode_intervals_kernel(const uint32_t* __restrict__ iv, int B, int64_t E, int precision, uint8_t* __restrict__ out,
                              int64_t stride, int64_t* __restrict__ lens, int* __restrict__ err) {
  const int lane = threadIdx.x;
  const int b = blockIdx.x * ENC_CPW + lane;
  if (lane >= ENC_CPW || b >= B) return;
  RangeEncoder e(out + (size_t)b * stride, stride, precision);
  const uint32_t* src = iv + (size_t)b * E;
  int64_t i = 0;
  for (; i + 8 <= E; i += 8) {
    // the loads do not depend on the coder state: issue a batch, then run the serial chain over it
    const uint4 w0 = __ldg(reinterpret_cast<const uint4*>(src + i)), w1 = __ldg(reinterpret_cast<const uint4*>(src + i + 4));
    const uint32_t w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) { const uint32_t lower = w[k] & 0xFFFF; e.encode(lower, lower + (w[k] >> 16) + 1); }
  }
  for (; i < E; ++i) { const uint32_t w = __ldg(src + i), lower = w & 0xFFFF; e.encode(lower, lower + (w >> 16) + 1); }
  const int64_t m = e.finish();
  if (m < 0) { atomicExch(err, PCGC_ERR_OVERFLOW); lens[b] = 0; } else lens[b] = m;
}

// Concatenates the B strings: offsets[b] = sum of lens[0..b), offsets[B] = total; bytes beyond cap are dropped (error flag).
__global__ void __launch_bounds__(256)
pack_strings_kernel(const uint8_t* __restrict__ in, int64_t stride, const int64_t* __restrict__ lens, int B,
                    uint8_t* __restrict__ packed, int64_t cap, int64_t* __restrict__ offsets, int* __restrict__ err) {
  __shared__ int64_t s_part[256];
  const int b = blockIdx.x;
  int64_t acc = 0;
  for (int j = threadIdx.x; j < b; j += 256) acc += lens[j];
  s_part[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) s_part[threadIdx.x] += s_part[threadIdx.x + o]; __syncthreads(); }
  const int64_t off = s_part[0], n = lens[b];
  if (threadIdx.x == 0) { offsets[b] = off; if (b == B - 1) offsets[B] = off + n; }
  if (off + n > cap) { if (threadIdx.x == 0) atomicExch(err, PCGC_ERR_OVERFLOW); return; }
  const uint8_t* src = in + (size_t)b * stride;
  for (int64_t i = threadIdx.x; i < n; i += 256) packed[off + i] = src[i];
}

constexpr int DEC_G = 32;                       // symbols per shared-memory window
constexpr int DEC_MAXN = 64;                    // lane-parallel search: lane k tests entries k and k + 32
constexpr int DEC_WARPS = 1;                    // cubes per block: a 32-thread block with a few KB of shared memory fits beside the
                                                // persistent conv CTAs of another stream

__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}

__global__ void __launch_bounds__(32 * DEC_WARPS)
range_decode_rows_kernel(const uint8_t* __restrict__ packed, const int64_t* __restrict__ offsets, int B, int64_t E,
                         const uint16_t* __restrict__ rows, const int64_t* __restrict__ row_offset,
                         const int32_t* __restrict__ minmax, int precision, float* __restrict__ y_hat, int* __restrict__ err,
                         int win_elems) {
  extern __shared__ __align__(16) uint16_t s_rows[];              // [DEC_WARPS][2][win_elems], win_elems >= DEC_G * max N of the launch
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * DEC_WARPS + warp;
  if (b >= B) return;
  const int min_v = minmax[2 * b], N = minmax[2 * b + 1] - min_v + 1;
  if (N < 1 || N > DEC_MAXN || DEC_G * N > win_elems) { if (lane == 0) atomicExch(err, PCGC_ERR_BAD_RANGE); return; }
  const uint8_t* str = packed + offsets[b];
  const int64_t nbytes = offsets[b + 1] - offsets[b];
  const uint32_t* rsrc = reinterpret_cast<const uint32_t*>(rows + row_offset[b]);    // E*N is even: 4-byte aligned windows
  const int words = DEC_G * N / 2;                                                   // uint32 words per window
  uint16_t* const wbuf[2] = {s_rows + (size_t)(2 * warp) * win_elems, s_rows + (size_t)(2 * warp + 1) * win_elems};
  const uint32_t sbuf[2] = {(uint32_t)__cvta_generic_to_shared(wbuf[0]), (uint32_t)__cvta_generic_to_shared(wbuf[1])};
  const int64_t groups = E / DEC_G;                                                  // E % 32 == 0 (checked by the host)
  auto fetch = [&](int64_t g, int buf) {
    const uint32_t* src = rsrc + (size_t)g * words;
    for (int w = lane; w < words; w += 32) cp_async4(sbuf[buf] + 4u * w, src + w);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // decoder state, replicated in every lane (RangeDecoder of range_coder.h, unrolled for the warp)
  int64_t pos = 0;
  auto byte_at = [&](int64_t p) -> uint32_t { return p < nbytes ? (uint32_t)__ldg(str + p) : 0u; };
  uint32_t value = (byte_at(0) << 24) | (byte_at(1) << 16) | (byte_at(2) << 8) | byte_at(3);
  pos = 4;
  uint32_t nxt = (byte_at(pos) << 8) | byte_at(pos + 1);                             // prefetched next word
  uint32_t base = 0, size_minus1 = 0xFFFFFFFFu;
  const uint32_t top = 1u << precision;
  const bool wide = N > 32;                                                          // warp-uniform
  fetch(0, 0);
  for (int64_t g = 0; g < groups; ++g) {
    const int buf = (int)(g & 1);
    if (g + 1 < groups) { fetch(g + 1, buf ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    const uint16_t* win = wbuf[buf];
    int mine = 0;
#pragma unroll 4
    for (int j = 0; j < DEC_G; ++j) {
      const uint16_t* row = win + j * N;
      const uint32_t c_lo = lane < N ? (uint32_t)row[lane] : top;                    // cdf[lane]
      const uint32_t c_hi = lane + 1 < N ? (uint32_t)row[lane + 1] : top;            // cdf[lane + 1] (cdf[N] = 2^precision)
      const uint64_t offset = (((uint64_t)(uint32_t)(value - base) + 1) << precision) - 1;
      const uint64_t p_lo = (uint64_t)size_minus1 * c_lo + c_lo;                     // size * cdf[lane]
      const uint64_t p_hi = (uint64_t)size_minus1 * c_hi + c_hi;
      const unsigned m = __ballot_sync(0xffffffffu, lane < N && (lane == 0 || p_lo <= offset));
      int s = 31 - __clz((int)m);                                                    // last entry with size*cdf <= offset
      uint32_t a_l = (uint32_t)(p_lo >> precision), n_l = (uint32_t)(p_hi >> precision) - 1u - a_l;
      if (wide) {                                                                    // entries 32 .. N-1
        const uint32_t d_lo = lane + 32 < N ? (uint32_t)row[lane + 32] : top, d_hi = lane + 33 < N ? (uint32_t)row[lane + 33] : top;
        const uint64_t q_lo = (uint64_t)size_minus1 * d_lo + d_lo, q_hi = (uint64_t)size_minus1 * d_hi + d_hi;
        const unsigned m2 = __ballot_sync(0xffffffffu, lane + 32 < N && q_lo <= offset);
        if (m2) {
          s = 63 - __clz((int)m2);
          a_l = (uint32_t)(q_lo >> precision); n_l = (uint32_t)(q_hi >> precision) - 1u - a_l;
        }
      }
      const uint32_t a = __shfl_sync(0xffffffffu, a_l, s & 31);
      size_minus1 = __shfl_sync(0xffffffffu, n_l, s & 31);
      base += a;
      if ((size_minus1 >> 16) == 0) {
        base <<= 16;
        size_minus1 = (size_minus1 << 16) | 0xFFFF;
        value = (value << 16) | nxt;
        pos += 2;
        nxt = (byte_at(pos) << 8) | byte_at(pos + 1);
      }
      if (lane == j) mine = s;
    }
    y_hat[(size_t)b * E + g * DEC_G + lane] = (float)(mine + min_v);
    __syncwarp();                                                                     // window `buf` is refilled two iterations later
  }
}

}  // namespace

cudaError_t launch_range_encode_intervals(const uint32_t* iv, int B, int64_t E, int precision, uint8_t* scratch, int64_t stride,
                                          int64_t* lens, uint8_t* packed, int64_t cap, int64_t* offsets, int* err,
                                          cudaStream_t s, int64_t* launches) {
  if (B <= 0) return cudaSuccess;
  range_encode_intervals_kernel<<<(B + ENC_CPW - 1) / ENC_CPW, 32, 0, s>>>(iv, B, E, precision, scratch, stride, lens, err);
  pack_strings_kernel<<<B, 256, 0, s>>>(scratch, stride, lens, B, packed, cap, offsets, err);
  if (launches) *launches += 2;
  return cudaGetLastError();
}

cudaError_t launch_range_decode_rows(const uint8_t* packed, const int64_t* offsets, int B, int64_t E, const uint16_t* rows,
                                     const int64_t* row_offset, const int32_t* minmax, int max_n, int precision, float* y_hat, int* err,
                                     cudaStream_t s, int64_t* launches) {
  if (B <= 0) return cudaSuccess;
  if (E % DEC_G || max_n < 1 || max_n > DEC_MAXN) return cudaErrorInvalidValue;
  const int win_elems = DEC_G * ((max_n + 7) / 8 * 8);              // 16-byte multiple per window
  const size_t smem = (size_t)DEC_WARPS * 2 * win_elems * sizeof(uint16_t);
  range_decode_rows_kernel<<<(B + DEC_WARPS - 1) / DEC_WARPS, 32 * DEC_WARPS, smem, s>>>(packed, offsets, B, E, rows, row_offset, minmax,
                                                                                         precision, y_hat, err, win_elems);
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace pcgc
