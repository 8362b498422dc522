// Host range coder of libpcgc_b200.so -- the reference's sequential tail, kept on the host as
// BASELINE.json's north_star asks.  Replaces coder_ops.range_encode / range_decode /
// pmf_to_quantized_cdf (models/entropy_model.py:218,258-259,298-299;
// models/conditional_entropy_model.py:122,161,195).  The upstream C++ (tensorflow-gpu==1.13.1,
// tensorflow/contrib/coder/kernels/range_coder.cc) is not vendored; this is a fresh
// carry-propagating 32-bit range coder with 16-bit renormalisation built to the published
// contract: interval update a=(size*lower)>>p, b=((size*upper)>>p)-1; big-endian 16-bit words;
// finalisation by the upstream Finalize rule (range_coder.h finish()).
// Byte-identical to the literal restatement of the upstream delay-based RangeEncoder in oracle/coder.py (UpstreamRangeEncoder;
// tests/test_host_coder.py); NOT verified against a TF binary: the reference holds no golden stream to pin either.
// Per-cube strings are independent, so the batch entry points fan out over a thread pool.  The state machines live in
// range_coder.h and are shared with the GPU coder (gpu_coder.cu).
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <atomic>
#include <thread>
#include <vector>

#include "../../include/pcgc_b200.h"
#include "cdf_norm.h"
#include "det_math.h"
#include "range_coder.h"

namespace {

using Encoder = pcgc::RangeEncoder;      // range_coder.h: shared with the GPU coder (gpu_coder.cu)
using Decoder = pcgc::RangeDecoder;

template <typename F>
void parallel_for(int n, int threads, F f) {
  int hw = (int)std::thread::hardware_concurrency();
  if (hw <= 0) hw = 1;
  if (threads <= 0 || threads > hw) threads = hw;
  if (threads > n) threads = n;
  if (threads <= 1) { for (int i = 0; i < n; ++i) f(i); return; }
  std::atomic<int> next(0);
  std::vector<std::thread> pool;
  pool.reserve(threads);
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&] { for (int i; (i = next.fetch_add(1)) < n;) f(i); });
  for (auto& th : pool) th.join();
}

// One long string (the hyper latents of a whole cloud, the factorized latents of model_simple: millions of symbols on ONE thread, the
// serial tail of the sharded codec) without the unpredictable renormalisation branch.  The number a range encoder writes is a SUM of
// the interval offsets, so every renormalised word can be stored at once and a later carry simply walks back through the bytes already
// written (rare, well predicted); what is left per symbol is the size recurrence with selects and one unconditional 16-bit store whose
// position advances by 0 or 2.  Same bytes as Encoder (range_coder.h) incl. the upstream Finalize rule: tests/test_host_coder.py.
inline void carry_back(uint8_t* out, int64_t n) {
  int64_t i = n - 1;
  while (i >= 0 && out[i] == 0xFF) out[i--] = 0;
  if (i >= 0) ++out[i];
}

int encode_string_branch_free(const int16_t* sym, int64_t n, const int32_t* cdf, int cdf_rows, int N, int precision, uint8_t* out,
                              int64_t* len) {
  uint32_t low = 0, sm1 = 0xFFFFFFFFu;
  int64_t pos = 0;                                   // bytes written; out has room for 2 n + 8
  int r = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int s = sym[i];
    if (s < 0 || s >= N) return PCGC_ERR_BAD_RANGE;
    const int32_t* row = cdf + (int64_t)r * (N + 1);
    const uint32_t lower = (uint32_t)row[s], upper = (uint32_t)row[s + 1];
    if (!(lower < upper)) return PCGC_ERR_BAD_ARG;
    const uint64_t size = (uint64_t)sm1 + 1;
    const uint32_t a = (uint32_t)((size * lower) >> precision);
    const uint32_t t = (uint32_t)(((size * upper) >> precision) - 1) - a;
    low += a;
    if (low < a) carry_back(out, pos);
    const bool renorm = t < 0x10000u;
    out[pos] = (uint8_t)(low >> 24);
    out[pos + 1] = (uint8_t)(low >> 16);
    pos += renorm ? 2 : 0;
    low = renorm ? low << 16 : low;
    sm1 = renorm ? ((t << 16) | 0xFFFFu) : t;
    if (++r == cdf_rows) r = 0;
  }
  if ((uint32_t)(low + sm1) < low) {                 // upstream Finalize, delayed state: the value 2^32
    if (pos == 0) return PCGC_ERR_NOT_READY;         // cannot happen (a wrapped interval needs a renormalisation); use the generic coder
    carry_back(out, pos);
    while (pos > 0 && out[pos - 1] == 0) --pos;      // the delayed word + 1 is not zero: this stops inside it
  } else if (low != 0) {
    const uint32_t mid = ((low - 1) >> 16) + 1;
    out[pos++] = (uint8_t)(mid >> 8);
    if (mid & 0xFF) out[pos++] = (uint8_t)mid;
  }
  *len = pos;
  return PCGC_OK;
}

// K independent per-cube streams decoded in ONE loop: the serial dependency chain of a range decoder (division -> search ->
// interval update -> renormalise, ~60 cycles per symbol) leaves an out-of-order core mostly idle; interleaving K cubes lets it
// overlap their chains.  Same arithmetic per stream, so the symbols are identical to the one-at-a-time decoder's.
template <int K, typename OutT, typename Conv>
void decode_rows_interleaved(const uint8_t* const* data, const int64_t* nbytes, const int* cube, int64_t E, const uint16_t* rows,
                             const int64_t* row_offset, const int32_t* minmax, int precision, OutT* out, Conv conv) {
  Decoder* d[K];
  alignas(Decoder) unsigned char store[K][sizeof(Decoder)];
  const uint16_t* base[K];
  OutT* o[K];
  int N[K], mn[K];
  for (int j = 0; j < K; ++j) {
    const int b = cube[j];
    d[j] = new (store[j]) Decoder(data[b], nbytes[b], precision);
    mn[j] = minmax[2 * b]; N[j] = minmax[2 * b + 1] - mn[j] + 1;
    base[j] = rows + row_offset[b];
    o[j] = out + (int64_t)b * E;
  }
  const uint32_t top = 1u << precision;
  for (int64_t i = 0; i < E; ++i) {
#pragma GCC unroll 4
    for (int j = 0; j < K; ++j) {
      const uint16_t* row = base[j] + i * N[j];
      const int n = N[j];
      const int s = d[j]->decode_at([row, n](uint32_t t) {
        int c = 0;
        for (int k = 1; k < n; ++k) c += (uint32_t)row[k] <= t;
        return c;
      }, [row, n, top](int k) { return k == n ? top : (uint32_t)row[k]; });
      o[j][i] = conv(s, mn[j]);
    }
  }
}

// The encoder's chain (two multiplies, carry-propagating renormalisation) interleaves the same way.
template <int K>
int encode_intervals_interleaved(const uint32_t* iv, const int* cube, int64_t E, int precision, uint8_t* out, int64_t stride, int64_t* lens) {
  alignas(Encoder) unsigned char store[K][sizeof(Encoder)];
  Encoder* e[K];
  const uint32_t* src[K];
  for (int j = 0; j < K; ++j) {
    e[j] = new (store[j]) Encoder(out + (int64_t)cube[j] * stride, stride, precision);
    src[j] = iv + (int64_t)cube[j] * E;
  }
  for (int64_t i = 0; i < E; ++i) {
#pragma GCC unroll 4
    for (int j = 0; j < K; ++j) {
      const uint32_t w = src[j][i], lower = w & 0xFFFF;
      e[j]->encode(lower, lower + (w >> 16) + 1);
    }
  }
  int rc = PCGC_OK;
  for (int j = 0; j < K; ++j) {
    const int64_t m = e[j]->finish();
    if (m < 0) rc = PCGC_ERR_OVERFLOW; else lens[cube[j]] = m;
  }
  return rc;
}

template <typename OutT, typename Conv>
int decode_rows_batch_impl(const uint8_t* const* data, const int64_t* nbytes, int B, int64_t E, const uint16_t* rows,
                           const int64_t* row_offset, const int32_t* minmax, int precision, OutT* out, int threads, Conv conv) {
  // interleave width: as wide as leaves every host thread a group (64 cubes on 16 threads -> 4; 16 cubes -> 1)
  int hw = (int)std::thread::hardware_concurrency();
  if (hw <= 0) hw = 1;
  const int T = (threads <= 0 || threads > hw) ? hw : threads;
  const int K = std::max(1, std::min(4, B / std::max(1, T)));
  const int groups = (B + K - 1) / K;
  parallel_for(groups, threads, [&](int g) {
    int cube[4];
    const int n = std::min(K, B - g * K);
    for (int j = 0; j < n; ++j) cube[j] = g * K + j;
    switch (n) {
      case 4: decode_rows_interleaved<4>(data, nbytes, cube, E, rows, row_offset, minmax, precision, out, conv); break;
      case 3: decode_rows_interleaved<3>(data, nbytes, cube, E, rows, row_offset, minmax, precision, out, conv); break;
      case 2: decode_rows_interleaved<2>(data, nbytes, cube, E, rows, row_offset, minmax, precision, out, conv); break;
      default: decode_rows_interleaved<1>(data, nbytes, cube, E, rows, row_offset, minmax, precision, out, conv); break;
    }
  });
  return PCGC_OK;
}

}  // namespace

extern "C" {

int pcgc_abi_version(void) { return PCGC_B200_ABI_VERSION; }

int pcgc_pmf_to_quantized_cdf(const float* pmf, int64_t rows, int N, int precision, int32_t* cdf) {
  if (!pmf || !cdf || rows < 0 || precision < 1 || precision > 16) return PCGC_ERR_BAD_ARG;
  if (N < 2) return PCGC_ERR_BAD_RANGE;   // upstream: "`pmf` size should be at least 2 in the last axis"
  std::vector<int32_t> v(N);
  std::vector<float> g(N);
  for (int64_t r = 0; r < rows; ++r) {
    if (pcgc::quantize_pmf_row(pmf + r * N, N, precision, v.data(), g.data()) != 0) return PCGC_ERR_BAD_RANGE;
    int32_t* row = cdf + r * (N + 1);
    int32_t acc = 0;
    row[0] = 0;
    for (int i = 0; i < N; ++i) { acc += v[i]; row[i + 1] = acc; }
  }
  return PCGC_OK;
}

int pcgc_range_encode(const int16_t* sym, int64_t n, const int32_t* cdf, int cdf_rows, int N, int precision,
                      uint8_t* out, int64_t cap, int64_t* len) {
  if (!sym || !cdf || !out || !len || cdf_rows < 1 || N < 1) return PCGC_ERR_BAD_ARG;
  if (precision >= 1 && precision <= 16 && cap >= 2 * n + 8) {
    const int rc = encode_string_branch_free(sym, n, cdf, cdf_rows, N, precision, out, len);
    if (rc != PCGC_ERR_NOT_READY) return rc;
  }
  Encoder e(out, cap, precision);
  int r = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int s = sym[i];
    if (s < 0 || s >= N) return PCGC_ERR_BAD_RANGE;
    const int32_t* row = cdf + (int64_t)r * (N + 1);
    if (!(row[s] < row[s + 1])) return PCGC_ERR_BAD_ARG;
    e.encode((uint32_t)row[s], (uint32_t)row[s + 1]);
    if (++r == cdf_rows) r = 0;
  }
  const int64_t m = e.finish();
  if (m < 0) return PCGC_ERR_OVERFLOW;
  *len = m;
  return PCGC_OK;
}

int pcgc_range_decode(const uint8_t* data, int64_t nbytes, int64_t n, const int32_t* cdf, int cdf_rows, int N,
                      int precision, int16_t* sym) {
  if ((!data && nbytes) || !cdf || !sym || cdf_rows < 1 || N < 1 || precision < 1 || precision > 16) return PCGC_ERR_BAD_ARG;
  Decoder d(data, nbytes, precision);
  // The rows are shared by n / cdf_rows symbols each, so a coarse table per row pays: start[r][t >> sh] is the symbol whose
  // interval holds the first value of the bucket; a short, well-predicted scan finishes the search.
  const int sh = precision > 8 ? precision - 8 : 0;
  std::vector<uint16_t> start((size_t)cdf_rows * 256);
  for (int r = 0; r < cdf_rows; ++r) {
    const int32_t* row = cdf + (int64_t)r * (N + 1);
    int s = 0;
    for (int q = 0; q < 256; ++q) {
      const int32_t t = q << sh;
      while (s + 1 < N && row[s + 1] <= t) ++s;
      start[(size_t)r * 256 + q] = (uint16_t)s;
    }
  }
  int r = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int32_t* row = cdf + (int64_t)r * (N + 1);
    const uint16_t* st = start.data() + (size_t)r * 256;
    sym[i] = (int16_t)d.decode_at([row, st, sh, N](uint32_t t) {
      int s = st[t >> sh];
      while (s + 1 < N && (uint32_t)row[s + 1] <= t) ++s;
      return s;
    }, [row](int k) { return (uint32_t)row[k]; });
    if (++r == cdf_rows) r = 0;
  }
  return PCGC_OK;
}

/* pcgc_range_decode that publishes its position: after every `step` symbols (and at the end) the number of symbols written
 * to sym is stored to *progress with release order, so another thread can consume the head of one long string (the hyper
 * latents of the first cubes) while the tail is still being decoded. */
int pcgc_range_decode_progress(const uint8_t* data, int64_t nbytes, int64_t n, const int32_t* cdf, int cdf_rows, int N,
                               int precision, int16_t* sym, int64_t* progress, int64_t step) {
  if ((!data && nbytes) || !cdf || !sym || !progress || cdf_rows < 1 || N < 1 || precision < 1 || precision > 16 || step < 1) {
    if (progress) __atomic_store_n(progress, (int64_t)-1, __ATOMIC_RELEASE);
    return PCGC_ERR_BAD_ARG;
  }
  Decoder d(data, nbytes, precision);
  const int sh = precision > 8 ? precision - 8 : 0;
  std::vector<uint16_t> start((size_t)cdf_rows * 256);
  for (int r = 0; r < cdf_rows; ++r) {
    const int32_t* row = cdf + (int64_t)r * (N + 1);
    int s = 0;
    for (int q = 0; q < 256; ++q) {
      const int32_t t = q << sh;
      while (s + 1 < N && row[s + 1] <= t) ++s;
      start[(size_t)r * 256 + q] = (uint16_t)s;
    }
  }
  int r = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int32_t* row = cdf + (int64_t)r * (N + 1);
    const uint16_t* st = start.data() + (size_t)r * 256;
    sym[i] = (int16_t)d.decode_at([row, st, sh, N](uint32_t t) {
      int s = st[t >> sh];
      while (s + 1 < N && (uint32_t)row[s + 1] <= t) ++s;
      return s;
    }, [row](int k) { return (uint32_t)row[k]; });
    if (++r == cdf_rows) r = 0;
    if ((i + 1) % step == 0) __atomic_store_n(progress, i + 1, __ATOMIC_RELEASE);
  }
  __atomic_store_n(progress, n, __ATOMIC_RELEASE);
  return PCGC_OK;
}

int pcgc_range_encode_intervals(const uint32_t* iv, int64_t n, int precision, uint8_t* out, int64_t cap,
                                int64_t* len) {
  if (!iv || !out || !len) return PCGC_ERR_BAD_ARG;
  if (precision == 16) {
    // the 32-bit state machine the GPU encoder runs (range_coder.h RangeEncoder16); the batch entry point below runs the generic
    // RangeEncoder -- tests/test_host_coder.py holds the two to the same bytes
    pcgc::RangeEncoder16 e16(out, cap);
    for (int64_t i = 0; i < n; ++i) e16.encode_word(iv[i]);
    const int64_t m16 = e16.finish();
    if (m16 < 0) return PCGC_ERR_OVERFLOW;
    *len = m16;
    return PCGC_OK;
  }
  Encoder e(out, cap, precision);
  for (int64_t i = 0; i < n; ++i) {
    const uint32_t lower = iv[i] & 0xFFFF;
    const uint32_t upper = lower + (iv[i] >> 16) + 1;
    e.encode(lower, upper);
  }
  const int64_t m = e.finish();
  if (m < 0) return PCGC_ERR_OVERFLOW;
  *len = m;
  return PCGC_OK;
}

int pcgc_range_decode_rows(const uint8_t* data, int64_t nbytes, int64_t n, const uint16_t* rows, int N,
                           int precision, int16_t* sym) {
  if ((!data && nbytes) || !rows || !sym || N < 1) return PCGC_ERR_BAD_ARG;
  Decoder d(data, nbytes, precision);
  const uint32_t top = 1u << precision;
  for (int64_t i = 0; i < n; ++i) {
    const uint16_t* row = rows + i * N;
    // symbol = number of interior boundaries row[1..N-1] that are <= t (branch-free; the rows are short)
    sym[i] = (int16_t)d.decode_at([row, N](uint32_t t) {
      int s = 0;
      for (int k = 1; k < N; ++k) s += (uint32_t)row[k] <= t;
      return s;
    }, [row, N, top](int k) { return k == N ? top : (uint32_t)row[k]; });
  }
  return PCGC_OK;
}

int pcgc_range_encode_intervals_batch(const uint32_t* iv, int B, int64_t E, int precision, uint8_t* out,
                                      int64_t stride, int64_t* lens, int threads) {
  if (!iv || !out || !lens || B < 0) return PCGC_ERR_BAD_ARG;
  std::atomic<int> rc(PCGC_OK);
  int hw = (int)std::thread::hardware_concurrency();
  if (hw <= 0) hw = 1;
  const int T = (threads <= 0 || threads > hw) ? hw : threads;
  const int K = std::max(1, std::min(4, B / std::max(1, T)));
  const int groups = (B + K - 1) / K;
  parallel_for(groups, threads, [&](int g) {
    int cube[4];
    const int n = std::min(K, B - g * K);
    for (int j = 0; j < n; ++j) cube[j] = g * K + j;
    int r;
    switch (n) {
      case 4: r = encode_intervals_interleaved<4>(iv, cube, E, precision, out, stride, lens); break;
      case 3: r = encode_intervals_interleaved<3>(iv, cube, E, precision, out, stride, lens); break;
      case 2: r = encode_intervals_interleaved<2>(iv, cube, E, precision, out, stride, lens); break;
      default: r = encode_intervals_interleaved<1>(iv, cube, E, precision, out, stride, lens); break;
    }
    if (r != PCGC_OK) rc.store(r);
  });
  return rc.load();
}

int pcgc_range_decode_rows_batch(const uint8_t* const* data, const int64_t* nbytes, int B, int64_t E,
                                 const uint16_t* rows, const int64_t* row_offset, const int32_t* minmax,
                                 int precision, int16_t* sym, int threads) {
  if (!data || !nbytes || !rows || !row_offset || !minmax || !sym || B < 0) return PCGC_ERR_BAD_ARG;
  for (int b = 0; b < B; ++b) if (minmax[2 * b + 1] - minmax[2 * b] + 1 < 1) return PCGC_ERR_BAD_ARG;
  return decode_rows_batch_impl(data, nbytes, B, E, rows, row_offset, minmax, precision, sym, threads,
                                [](int s, int) { return (int16_t)s; });
}

/* As pcgc_range_decode_rows_batch, but writes y_hat = symbol + min_v as float32 (what
 * SymmetricConditional.decompress returns, conditional_entropy_model.py:196-199) so the host layer
 * can upload it without another pass over the data. */
int pcgc_range_decode_rows_batch_f32(const uint8_t* const* data, const int64_t* nbytes, int B, int64_t E,
                                     const uint16_t* rows, const int64_t* row_offset, const int32_t* minmax,
                                     int precision, float* y_hat, int threads) {
  if (!data || !nbytes || !rows || !row_offset || !minmax || !y_hat || B < 0) return PCGC_ERR_BAD_ARG;
  for (int b = 0; b < B; ++b) if (minmax[2 * b + 1] - minmax[2 * b] + 1 < 1) return PCGC_ERR_BAD_ARG;
  return decode_rows_batch_impl(data, nbytes, B, E, rows, row_offset, minmax, precision, y_hat, threads,
                                [](int s, int mn) { return (float)(s + mn); });
}

/* Host twin of pcgc_laplace_cdf (SymmetricConditional._get_cdf, conditional_entropy_model.py:95-124): the same det_math.h
 * likelihood and the same cdf_norm.h normaliser as laplace_cdf_kernel, so the rows are bit-identical to the GPU's and a
 * CPU can decode a GPU-written stream given (loc, scale).  Test / interoperability aid, not on the product path. */
int pcgc_host_laplace_cdf(const float* loc, const float* scale, int B, int64_t E, const int32_t* minmax, float likelihood_bound,
                          int precision, const int64_t* row_offset, uint16_t* rows, int threads) {
  if (!loc || !scale || !minmax || !row_offset || !rows || B < 0 || E < 0 || precision < 1 || precision > 16) return PCGC_ERR_BAD_ARG;
  for (int b = 0; b < B; ++b) {
    const int N = minmax[2 * b + 1] - minmax[2 * b] + 1;
    if (N < 2 || N > PCGC_MAX_SYMBOLS) return PCGC_ERR_BAD_RANGE;
  }
  std::atomic<int> rc(PCGC_OK);
  const int64_t CH = 4096;                                   // elements per job
  const int64_t per = (E + CH - 1) / CH;
  parallel_for((int)(B * per), threads, [&](int job) {
    const int b = (int)(job / per);
    const int64_t e0 = (job % per) * CH, e1 = std::min(E, e0 + CH);
    const int min_v = minmax[2 * b], N = minmax[2 * b + 1] - min_v + 1;
    float pmf[PCGC_MAX_SYMBOLS], g[PCGC_MAX_SYMBOLS];
    int32_t v[PCGC_MAX_SYMBOLS];
    for (int64_t e = e0; e < e1; ++e) {
      const float l = loc[(int64_t)b * E + e], s = scale[(int64_t)b * E + e];
      pcgc::det_laplace_pmf_row(min_v, N, l, s, likelihood_bound, pmf);
      if (pcgc::quantize_pmf_row(pmf, N, precision, v, g) != 0) { rc.store(PCGC_ERR_BAD_RANGE); continue; }
      uint16_t* row = rows + row_offset[b] + e * N;
      uint32_t acc = 0;
      for (int k = 0; k < N; ++k) { row[k] = (uint16_t)acc; acc += (uint32_t)v[k]; }
    }
  });
  return rc.load();
}

}  // extern "C"
