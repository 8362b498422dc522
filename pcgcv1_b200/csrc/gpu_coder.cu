// GPU-side range coding of the per-cube latent strings: the byte streams of
// SymmetricConditional.compress / decompress (models/conditional_entropy_model.py:126-201) are written and read on the
// device, so neither the per-element CDF rows (65 536 x N uint16 per cube) nor the per-element intervals cross PCIe and no
// host thread pool sits between the transforms.  Cubes are independent strings (transform.py:157-168): the coder is
// sequential inside a string and parallel over cubes.
//
// The state machines are range_coder.h's -- the SAME source as the host coder (coder.cpp) -- so a string written here is
// byte-identical to pcgc_range_encode_intervals' and decodes with pcgc_range_decode_rows (tests/test_gpu_coder.py).
//   encode: one WARP per cube with the state replicated in every lane: coalesced interval loads handed round by shuffles, a
//           warp-uniform renormalisation branch, lane 0 stores the words.
//   decode: one WARP per cube.  The symbol search is lane-parallel: lane k holds cdf[k] of the current symbol's row and
//           tests size*cdf[k] <= offset; a ballot gives the symbol without a division or a serial scan.  Rows stream
//           through a double-buffered shared-memory window (cp.async), the next 16-bit word of the string is prefetched.
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"
#include "range_coder.h"

namespace pcgc {

namespace {

// One WARP per cube, precision 16, state replicated in every lane (RangeEncoder16 of range_coder.h, unrolled for the warp):
// the lanes fetch 32 intervals with one coalesced load and hand them round by shuffles, the renormalisation branch is
// warp-uniform (no divergence), lane 0 stores the emitted 16-bit words.  A thread-per-cube form with 4-8 cubes per warp was
// measured at ~260 cycles per symbol (divergent renormalisation, 64-bit output bookkeeping on the serial chain).
// Every symbol renormalises at most once and every renormalisation emits at most one word on average, so 2*E + 8 bytes always
// suffice: the launcher checks the stride once instead of the kernel checking every store.
__global__ void __launch_bounds__(32)
range_encode_intervals_kernel(const uint32_t* __restrict__ iv, int B, int64_t E, uint8_t* __restrict__ out,
                              int64_t stride, int64_t* __restrict__ lens) {
  const int lane = threadIdx.x;
  const int b = blockIdx.x;
  const unsigned FULL = 0xffffffffu;
  const uint32_t* src = iv + (size_t)b * E;
  uint16_t* o16 = reinterpret_cast<uint16_t*>(out + (size_t)b * stride);          // stride and the buffer are 2-byte aligned
  uint32_t base = 0, carry = 0, sm1 = 0xFFFFFFFFu, cache = 0, have = 0, pending = 0, n16 = 0;
  auto emit = [&](uint32_t w) {                                                     // big-endian 16-bit word
    if (lane == 0) o16[n16] = (uint16_t)(((w & 0xFF) << 8) | ((w >> 8) & 0xFF));
    ++n16;
  };
  const int64_t groups = E / 32;                                                    // E % 32 == 0 (checked by the launcher)
  // Every lane reads the SAME 32 intervals (8 broadcast 16-byte loads) one group ahead into registers: nothing inside the
  // 32-symbol loop touches memory or another lane (a shuffle per symbol sat on the serial chain: 186 cycles per symbol).
  const uint4* src4 = reinterpret_cast<const uint4*>(src);
  uint4 cur[8], nxt[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) cur[q] = __ldg(src4 + q);
  for (int64_t g = 0; g < groups; ++g) {
    if (g + 1 < groups) {
#pragma unroll
      for (int q = 0; q < 8; ++q) nxt[q] = __ldg(src4 + (g + 1) * 8 + q);
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const uint4 qv = cur[j >> 2];
      const uint32_t w = (j & 3) == 0 ? qv.x : ((j & 3) == 1 ? qv.y : ((j & 3) == 2 ? qv.z : qv.w));
      const uint32_t lower = w & 0xFFFFu, upper = lower + (w >> 16) + 1u;
      const uint32_t a = (uint32_t)(((uint64_t)sm1 * lower + lower) >> 16);
      const uint32_t bq = (uint32_t)(((uint64_t)sm1 * upper + upper) >> 16) - 1u;
      const uint32_t nb = base + a;
      carry |= (uint32_t)(nb < a);
      base = nb;
      const uint32_t t = bq - a;
      const bool renorm = t < 0x10000u;                                             // warp-uniform
      sm1 = renorm ? ((t << 16) | 0xFFFFu) : t;                                     // the size chain does not wait for the emission below
      // (a straight-line predicated form of the block below -- one basic block per 32 symbols -- was measured slower: 4.48 vs
      // 3.84 ms per string; the renormalisation happens for about one symbol in three and the branch skips ~15 instructions)
      if (renorm) {
        if (base < 0xFFFF0000u || carry) {
          if (have) emit((cache + carry) & 0xFFFF);
#pragma unroll 1
          for (; pending > 0; --pending) emit((0xFFFF + carry) & 0xFFFF);
          cache = base >> 16;
          have = 1;
        } else {
          ++pending;
        }
        base <<= 16;
        carry = 0;
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) cur[q] = nxt[q];
  }
  // finish(): the multiple of 2^16 inside the interval, then drop trailing zero bytes
  const uint64_t v = (((uint64_t)carry << 32) + base + 0xFFFF) >> 16;
  const uint32_t c = (uint32_t)(v >> 16), word = (uint32_t)(v & 0xFFFF);
  if (have) emit((cache + c) & 0xFFFF);
#pragma unroll 1
  for (; pending > 0; --pending) emit((0xFFFF + c) & 0xFFFF);
  emit(word);
  if (lane == 0) {
    __threadfence_block();
    const uint8_t* o8 = out + (size_t)b * stride;
    int64_t n = 2 * (int64_t)n16;
    while (n > 0 && o8[n - 1] == 0) --n;
    lens[b] = n;
  }
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
constexpr int DEC_MAXN = 64;                    // lane-parallel search: lane k tests entry k (and k + 32 in the WIDE form)

__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}

// One warp per cube, precision 16.  Per symbol the serial chain is: size -> one 32x32->64 multiply per lane -> compare with
// value - base -> two warp reductions (REDUX max / min) -> new base / size.  Derivation (range_coder.h RangeDecoder):
//   size*cdf[k] <= offset = ((d+1) << 16) - 1  with d = value - base   <=>   a_k := (size*cdf[k]) >> 16 <= d,
//   so the decoded symbol s is the LAST lane whose a_k <= d, its interval starts at a_s = max{a_k : a_k <= d} (a_k is
//   monotone) and ends at a_{s+1} - 1 = min{a_k - 1 : a_k > d}; lanes >= N hold cdf = 2^16, i.e. a_k - 1 = size - 1.
// Nothing inside the 32-symbol loop touches memory: the lane's CDF entries are read into registers first, the next 16-bit
// words of the string are prefetched one per lane and handed out by shuffles.
template <bool WIDE>
__global__ void __launch_bounds__(256)
range_decode_rows_kernel(const uint8_t* __restrict__ packed, const int64_t* __restrict__ offsets, int B, int64_t E,
                         const uint16_t* __restrict__ rows, const int64_t* __restrict__ row_offset,
                         const int32_t* __restrict__ minmax, float* __restrict__ y_hat, int* __restrict__ err, int win_elems) {
  extern __shared__ __align__(16) uint16_t s_rows_all[];          // [cubes per block][2][win_elems], win_elems >= DEC_G * max N of the launch
  const int lane = threadIdx.x;
  const int b = blockIdx.x * blockDim.y + threadIdx.y;            // blockDim = (32, cubes per block): one warp per cube
  if (b >= B) return;
  uint16_t* const s_rows = s_rows_all + (size_t)threadIdx.y * 2 * win_elems;
  const unsigned FULL = 0xffffffffu;
  const int min_v = minmax[2 * b], N = minmax[2 * b + 1] - min_v + 1;
  if (N < 1 || N > (WIDE ? DEC_MAXN : 32) || DEC_G * N > win_elems) { if (lane == 0) atomicExch(err, PCGC_ERR_BAD_RANGE); return; }
  const uint8_t* str = packed + offsets[b];
  const int64_t nbytes = offsets[b + 1] - offsets[b];
  const uint32_t* rsrc = reinterpret_cast<const uint32_t*>(rows + row_offset[b]);    // E*N is even: 4-byte aligned windows
  const int words = DEC_G * N / 2;                                                   // uint32 words per window
  uint16_t* const wbuf[2] = {s_rows, s_rows + win_elems};
  const uint32_t sbuf[2] = {(uint32_t)__cvta_generic_to_shared(wbuf[0]), (uint32_t)__cvta_generic_to_shared(wbuf[1])};
  const int64_t groups = E / DEC_G;                                                  // E % 32 == 0 (checked by the host)
  auto fetch = [&](int64_t g, int buf) {
    const uint32_t* src = rsrc + (size_t)g * words;
    for (int w = lane; w < words; w += 32) cp_async4(sbuf[buf] + 4u * w, src + w);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // big-endian 16-bit word `i` of the string (zero beyond its end), i counted from the start of the string
  auto word_at = [&](int64_t i) -> uint32_t {
    const int64_t p = 2 * i;
    const uint32_t hi = p < nbytes ? (uint32_t)__ldg(str + p) : 0u, lo = p + 1 < nbytes ? (uint32_t)__ldg(str + p + 1) : 0u;
    return (hi << 8) | lo;
  };
  // decoder state, identical in every lane
  uint32_t value = (word_at(0) << 16) | word_at(1);
  int64_t wpos = 2;                                   // index of the next unread word
  uint32_t w0 = word_at(wpos + lane), w1 = word_at(wpos + 32 + lane);      // words wpos + [0, 64) spread over the lanes
  uint32_t base = 0, sm1 = 0xFFFFFFFFu;
  fetch(0, 0);
  for (int64_t g = 0; g < groups; ++g) {
    const int buf = (int)(g & 1);
    if (g + 1 < groups) { fetch(g + 1, buf ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    const uint32_t wb = sbuf[buf];
    uint32_t c[DEC_G], c2[WIDE ? DEC_G : 1];
#pragma unroll
    for (int j = 0; j < DEC_G; ++j) {
      uint32_t v = 0x10000u;
      if (lane < N) { uint16_t t; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(t) : "r"(wb + 2u * (uint32_t)(j * N + lane))); v = t; }
      c[j] = v;
      if (WIDE) {
        uint32_t v2 = 0x10000u;
        if (lane + 32 < N) { uint16_t t; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(t) : "r"(wb + 2u * (uint32_t)(j * N + lane + 32))); v2 = t; }
        c2[j] = v2;
      }
    }
    uint32_t r = 0;                                   // renormalisations (= words consumed) in this group
    uint32_t nxt = __shfl_sync(FULL, w0, 0);
    int mine = 0;
#pragma unroll
    for (int j = 0; j < DEC_G; ++j) {
      const uint32_t d = value - base;
      const uint64_t q = ((uint64_t)sm1 * c[j] + c[j]) >> 16;                        // (size * cdf[lane]) >> 16
      const uint32_t a_l = (uint32_t)q;
      const bool pred = a_l <= d && c[j] < 0x10000u;
      uint32_t va = pred ? a_l : 0u, vu = pred ? 0xFFFFFFFFu : a_l - 1u;
      unsigned cnt = __ballot_sync(FULL, pred);
      int s = __popc(cnt) - 1;
      if (WIDE) {
        const uint32_t a2 = (uint32_t)(((uint64_t)sm1 * c2[j] + c2[j]) >> 16);
        const bool pred2 = a2 <= d && c2[j] < 0x10000u;
        va = pred2 ? a2 : va;
        vu = pred2 ? vu : min(vu, a2 - 1u);
        s += __popc(__ballot_sync(FULL, pred2));
      }
      const uint32_t a = __reduce_max_sync(FULL, va);
      const uint32_t bm1 = min(__reduce_min_sync(FULL, vu), sm1);
      base += a;
      const uint32_t t = bm1 - a;
      const bool renorm = t < 0x10000u;                                             // warp-uniform; selects instead of a branch keep the
      sm1 = renorm ? ((t << 16) | 0xFFFFu) : t;                                     // 32 symbols one basic block for the scheduler
      base = renorm ? (base << 16) : base;
      value = renorm ? ((value << 16) | nxt) : value;
      r += renorm ? 1u : 0u;
      nxt = __shfl_sync(FULL, w0, r & 31);
      if (lane == j) mine = s;
    }
    y_hat[(size_t)b * E + g * DEC_G + lane] = (float)(mine + min_v);
    // advance the word window by r (<= 32): new w0 = words [r, r + 32) of the old 64, new w1 is fetched (latency hidden by the
    // next group)
    {
      const uint32_t t = r + (uint32_t)lane;
      const uint32_t x0 = __shfl_sync(FULL, w0, t & 31), x1 = __shfl_sync(FULL, w1, t & 31);
      w0 = t < 32 ? x0 : x1;
      wpos += r;
      // w1 must become words [wpos + 32, wpos + 64): those below old wpos + 64 are already in w1 (shifted), the rest are new
      const uint32_t y1 = __shfl_sync(FULL, w1, t & 31);
      w1 = t < 32 ? y1 : word_at(wpos + 32 + lane);
    }
    __syncwarp();                                                                     // window `buf` is refilled two iterations later
  }
}

}  // namespace

cudaError_t launch_range_encode_intervals(const uint32_t* iv, int B, int64_t E, int precision, uint8_t* scratch, int64_t stride,
                                          int64_t* lens, uint8_t* packed, int64_t cap, int64_t* offsets, int* err,
                                          cudaStream_t s, int64_t* launches) {
  if (B <= 0) return cudaSuccess;
  if (precision != 16 || E % 32 || stride < 2 * E + 8 || (stride & 1)) return cudaErrorInvalidValue;
  PCGC_CARVEOUT_ONCE(range_encode_intervals_kernel);
  PCGC_CARVEOUT_ONCE(pack_strings_kernel);
  range_encode_intervals_kernel<<<B, 32, 0, s>>>(iv, B, E, scratch, stride, lens);
  pack_strings_kernel<<<B, 256, 0, s>>>(scratch, stride, lens, B, packed, cap, offsets, err);
  if (launches) *launches += 2;
  return cudaGetLastError();
}

cudaError_t launch_range_decode_rows(const uint8_t* packed, const int64_t* offsets, int B, int64_t E, const uint16_t* rows,
                                     const int64_t* row_offset, const int32_t* minmax, int max_n, int precision, float* y_hat, int* err,
                                     cudaStream_t s, int64_t* launches) {
  if (B <= 0) return cudaSuccess;
  if (E % DEC_G || max_n < 1 || max_n > DEC_MAXN || precision != 16) return cudaErrorInvalidValue;
  const int win_elems = DEC_G * ((max_n + 7) / 8 * 8);              // 16-byte multiple per window
  // cubes (= warps) per block.  1 spreads the decoder warps over all SMs; more packs them onto few SMs (experiments: PCGC_DEC_CPB)
  static const int cpb = [] { const char* e = getenv("PCGC_DEC_CPB"); const int v = e ? atoi(e) : 1; return v < 1 ? 1 : (v > 8 ? 8 : v); }();
  const size_t smem = (size_t)cpb * 2 * win_elems * sizeof(uint16_t);
  PCGC_CARVEOUT_ONCE(range_decode_rows_kernel<true>);
  PCGC_CARVEOUT_ONCE(range_decode_rows_kernel<false>);
  const dim3 block(32, cpb), grid((B + cpb - 1) / cpb);
  if (max_n > 32)
    range_decode_rows_kernel<true><<<grid, block, smem, s>>>(packed, offsets, B, E, rows, row_offset, minmax, y_hat, err, win_elems);
  else
    range_decode_rows_kernel<false><<<grid, block, smem, s>>>(packed, offsets, B, E, rows, row_offset, minmax, y_hat, err, win_elems);
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace pcgc
