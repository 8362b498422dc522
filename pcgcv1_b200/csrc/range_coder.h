// The range coder's state machines, ONE source for the host coder (coder.cpp) and the GPU coder (gpu_coder.cu), so the
// byte streams of the two are identical by construction.  Replaces coder_ops.range_encode / range_decode
// (models/entropy_model.py:258-259,298-299; models/conditional_entropy_model.py:161,195).  The upstream C++
// (tensorflow-gpu==1.13.1, tensorflow/contrib/coder/kernels/range_coder.cc) is not vendored; this is a fresh
// carry-propagating 32-bit range coder with 16-bit renormalisation built to the published contract: interval update
// a=(size*lower)>>p, b=((size*upper)>>p)-1; big-endian 16-bit words; finalisation by the upstream Finalize rule (see finish()).
// oracle/coder.py holds a literal restatement of the upstream delay-based encoder; this carry-propagating form produces the same
// bytes (tests/test_host_coder.py, tests/test_gpu_coder.py).  Byte identity with a TF BINARY stays unpinned (no TF wheel, no
// golden stream in the reference).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PCGC_RC __host__ __device__ __forceinline__
#else
#define PCGC_RC inline
#endif

namespace pcgc {

struct RangeEncoder {
  uint64_t base = 0;            // bit 32 holds a carry that has not been propagated yet
  uint32_t size_minus1 = 0xFFFFFFFFu;
  bool have_cache = false;
  uint32_t cache = 0;           // delayed 16-bit word
  int64_t pending = 0;          // delayed 0xFFFF words following `cache`
  uint8_t* out;
  int64_t n = 0, cap;
  bool overflow = false;
  int precision;

  PCGC_RC RangeEncoder(uint8_t* o, int64_t c, int p) : out(o), cap(c), precision(p) {}

  PCGC_RC void emit16(uint32_t w) {
    if (n + 2 > cap) { overflow = true; return; }
    out[n++] = (uint8_t)(w >> 8);
    out[n++] = (uint8_t)w;
  }
  PCGC_RC void shift() {
    const uint32_t carry = (uint32_t)(base >> 32);
    const uint32_t low32 = (uint32_t)base;
    if (low32 < 0xFFFF0000u || carry) {
      if (have_cache) emit16((cache + carry) & 0xFFFF);
      for (; pending > 0; --pending) emit16((0xFFFF + carry) & 0xFFFF);
      cache = (low32 >> 16) & 0xFFFF;
      have_cache = true;
    } else {
      ++pending;
    }
    base = (uint64_t)(low32 & 0xFFFF) << 16;
  }
  PCGC_RC void encode(uint32_t lower, uint32_t upper) {
    const uint64_t size = (uint64_t)size_minus1 + 1;
    const uint32_t a = (uint32_t)((size * lower) >> precision);
    const uint32_t b = (uint32_t)(((size * upper) >> precision) - 1);
    base += a;
    size_minus1 = b - a;
    if ((size_minus1 >> 16) == 0) {
      shift();
      size_minus1 = (size_minus1 << 16) | 0xFFFF;
    }
  }
  // RangeEncoder::Finalize of the upstream coder in this state machine's terms (oracle/coder.py restates the upstream one literally;
  // tests compare the two byte for byte).  The interval is [base, base + size_minus1], base exact (33 bits).  If it still holds a
  // multiple of 2^32 above base (upstream: delay_ != 0) that multiple is written: the delayed word + 1, the zeros after it left
  // out.  Otherwise base is rounded up to a multiple of 2^16: earlier words in full, of the last word the low byte only if it is
  // not zero, and nothing when the low 32 bits of base are zero.
  PCGC_RC void emit8(uint32_t b) {
    if (n + 1 > cap) { overflow = true; return; }
    out[n++] = (uint8_t)b;
  }
  PCGC_RC int64_t finish() {
    const uint32_t carry = (uint32_t)(base >> 32), low32 = (uint32_t)base;
    if (!carry && (uint32_t)(low32 + size_minus1) < low32) {
      const uint32_t w = (cache + 1) & 0xFFFF;
      emit8(w >> 8);
      if (w & 0xFF) emit8(w & 0xFF);
    } else {
      if (have_cache) emit16((cache + carry) & 0xFFFF);
      for (; pending > 0; --pending) emit16((0xFFFF + carry) & 0xFFFF);
      if (low32 != 0) {
        const uint32_t mid = ((low32 - 1) >> 16) + 1;
        emit8(mid >> 8);
        if (mid & 0xFF) emit8(mid & 0xFF);
      }
    }
    return overflow ? -1 : n;
  }
};

// RangeEncoder specialised for precision 16 with the 33-bit base kept as 32 bits + a carry flag and no 64-bit shifts: the form
// the GPU encoder runs (one thread per string; every 64-bit operation of the generic form costs several SASS instructions on
// the serial chain).  Same decisions, same bytes as RangeEncoder (tests/test_host_coder.py compares them on random intervals).
struct RangeEncoder16 {
  uint32_t base = 0, carry = 0;          // base + (carry << 32) = RangeEncoder::base
  uint32_t size_minus1 = 0xFFFFFFFFu;
  uint32_t cache = 0, have_cache = 0;
  uint32_t pending = 0;
  uint8_t* out;
  int64_t n = 0, cap;
  bool overflow = false;

  PCGC_RC RangeEncoder16(uint8_t* o, int64_t c) : out(o), cap(c) {}

  PCGC_RC void emit16(uint32_t w) {
    if (n + 2 > cap) { overflow = true; return; }
    out[n] = (uint8_t)(w >> 8);
    out[n + 1] = (uint8_t)w;
    n += 2;
  }
  PCGC_RC void shift() {
    if (base < 0xFFFF0000u || carry) {
      if (have_cache) emit16((cache + carry) & 0xFFFF);
      for (; pending > 0; --pending) emit16((0xFFFF + carry) & 0xFFFF);
      cache = base >> 16;
      have_cache = 1;
    } else {
      ++pending;
    }
    base <<= 16;
    carry = 0;
  }
  // interval word of pcgc_laplace_intervals: lower | (upper - lower - 1) << 16
  PCGC_RC void encode_word(uint32_t w) {
    const uint32_t lower = w & 0xFFFFu, upper = lower + (w >> 16) + 1u;
    const uint32_t a = (uint32_t)(((uint64_t)size_minus1 * lower + lower) >> 16);
    const uint32_t b = (uint32_t)(((uint64_t)size_minus1 * upper + upper) >> 16) - 1u;     // (size*upper >> 16) <= 2^32: wraps correctly
    const uint32_t nb = base + a;
    carry |= (uint32_t)(nb < a);
    base = nb;
    size_minus1 = b - a;
    if (size_minus1 < 0x10000u) {
      shift();
      size_minus1 = (size_minus1 << 16) | 0xFFFFu;
    }
  }
  PCGC_RC void emit8(uint32_t b) {
    if (n + 1 > cap) { overflow = true; return; }
    out[n++] = (uint8_t)b;
  }
  PCGC_RC int64_t finish() {                      // the same rule as RangeEncoder::finish
    if (!carry && (uint32_t)(base + size_minus1) < base) {
      const uint32_t w = (cache + 1) & 0xFFFF;
      emit8(w >> 8);
      if (w & 0xFF) emit8(w & 0xFF);
    } else {
      if (have_cache) emit16((cache + carry) & 0xFFFF);
      for (; pending > 0; --pending) emit16((0xFFFF + carry) & 0xFFFF);
      if (base != 0) {
        const uint32_t mid = ((base - 1) >> 16) + 1;
        emit8(mid >> 8);
        if (mid & 0xFF) emit8(mid & 0xFF);
      }
    }
    return overflow ? -1 : n;
  }
};

struct RangeDecoder {
  const uint8_t* p;
  int64_t nbytes, pos = 0;
  uint32_t base = 0, size_minus1 = 0xFFFFFFFFu, value;
  int precision;

  PCGC_RC RangeDecoder(const uint8_t* d, int64_t n, int prec) : p(d), nbytes(n), precision(prec) {
    value = read16() << 16;
    value |= read16();
  }
  PCGC_RC uint32_t read16() {
    uint32_t v = 0;
    for (int k = 0; k < 2; ++k) { v <<= 8; if (pos < nbytes) v |= p[pos++]; }
    return v;
  }
  // cdf(i) for i in [0, N]; returns the symbol.
  template <typename CdfAt>
  PCGC_RC int decode(int N, CdfAt cdf) {
    const uint64_t size = (uint64_t)size_minus1 + 1;
    const uint64_t offset = (((uint64_t)(uint32_t)(value - base) + 1) << precision) - 1;
    int lo = 1, hi = N;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (size * (uint64_t)cdf(mid) > offset) hi = mid; else lo = mid + 1;
    }
    return narrow(lo - 1, size, cdf);
  }
  // Same search through one division: size*c > offset  <=>  c > floor(offset/size); find(t) returns the symbol s with
  // cdf(s) <= t < cdf(s+1).
  template <typename Find, typename CdfAt>
  PCGC_RC int decode_at(Find find, CdfAt cdf) {
    const uint64_t size = (uint64_t)size_minus1 + 1;
    const uint64_t offset = (((uint64_t)(uint32_t)(value - base) + 1) << precision) - 1;
    uint64_t t = offset / size;
    const uint64_t top = ((uint64_t)1 << precision) - 1;
    if (t > top) t = top;                         // only a corrupt stream gets here
    return narrow(find((uint32_t)t), size, cdf);
  }
  template <typename CdfAt>
  PCGC_RC int narrow(const int s, const uint64_t size, CdfAt cdf) {
    const uint32_t a = (uint32_t)((size * (uint64_t)cdf(s)) >> precision);
    const uint32_t b = (uint32_t)(((size * (uint64_t)cdf(s + 1)) >> precision) - 1);
    base += a;
    size_minus1 = b - a;
    if ((size_minus1 >> 16) == 0) {
      base <<= 16;
      size_minus1 = (size_minus1 << 16) | 0xFFFF;
      value = (value << 16) | read16();
    }
    return s;
  }
};

}  // namespace pcgc
