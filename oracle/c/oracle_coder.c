/* Oracle restatement in C of tensorflow.contrib.coder's pmf_to_quantized_cdf and range coder.
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into the product library.  **parity unpinned**
 * (see oracle/coder.py): the upstream op sources (tensorflow-gpu==1.13.1,
 * tensorflow/contrib/coder/kernels/) are not in the reference tree; this follows the published
 * algorithm and is cross-checked against the pure-Python definition in oracle/coder.py.
 * Reference call sites: models/entropy_model.py:218,258,298;
 * models/conditional_entropy_model.py:122,161,195.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- pmf -> quantised cdf */
static int pmf_row(const float* pmf, int n, int precision, int32_t* cdf) {
  const int64_t target = (int64_t)1 << precision;
  int64_t v[4096];
  double mass[4096];
  if (n < 2 || n > 4096) return -1;
  int64_t s = 0;
  for (int i = 0; i < n; ++i) {
    mass[i] = (double)pmf[i];
    int64_t q = (int64_t)rintf(pmf[i] * (float)target);
    v[i] = q < 1 ? 1 : q;
    s += v[i];
  }
  while (s > target) {            /* smallest penalty first, lowest index on ties */
    int best = -1;
    double bp = INFINITY;
    for (int i = 0; i < n; ++i) {
      if (v[i] <= 1) continue;
      double p = mass[i] * (log2((double)v[i]) - log2((double)(v[i] - 1)));
      if (p < bp) { bp = p; best = i; }
    }
    if (best < 0) return -2;
    v[best] -= 1; s -= 1;
  }
  while (s < target) {            /* largest gain first, lowest index on ties */
    int best = 0;
    double bg = -INFINITY;
    for (int i = 0; i < n; ++i) {
      double g = mass[i] * (log2((double)(v[i] + 1)) - log2((double)v[i]));
      if (g > bg) { bg = g; best = i; }
    }
    v[best] += 1; s += 1;
  }
  int64_t acc = 0;
  cdf[0] = 0;
  for (int i = 0; i < n; ++i) { acc += v[i]; cdf[i + 1] = (int32_t)acc; }
  return 0;
}

int orc_pmf_to_quantized_cdf(const float* pmf, int64_t rows, int n, int precision, int32_t* cdf) {
  for (int64_t r = 0; r < rows; ++r) {
    int rc = pmf_row(pmf + r * n, n, precision, cdf + r * (n + 1));
    if (rc) return rc;
  }
  return 0;
}

/* ---------------------------------------------------------------- range encoder */
typedef struct {
  uint64_t base;          /* bit 32 = carry not yet propagated */
  uint32_t size_minus1;
  int have_cache;
  uint32_t cache;
  int64_t pending;
  uint8_t* out;
  int64_t n, cap;
  int overflow;
} enc_t;

static void emit16(enc_t* e, uint32_t w) {
  if (e->n + 2 > e->cap) { e->overflow = 1; return; }
  e->out[e->n++] = (uint8_t)(w >> 8);
  e->out[e->n++] = (uint8_t)w;
}

static void shift(enc_t* e) {
  uint32_t carry = (uint32_t)(e->base >> 32);
  uint32_t low32 = (uint32_t)e->base;
  if (low32 < 0xFFFF0000u || carry) {
    if (e->have_cache) emit16(e, (e->cache + carry) & 0xFFFF);
    for (int64_t i = 0; i < e->pending; ++i) emit16(e, (0xFFFF + carry) & 0xFFFF);
    e->pending = 0;
    e->cache = (low32 >> 16) & 0xFFFF;
    e->have_cache = 1;
  } else {
    e->pending += 1;
  }
  e->base = (uint64_t)(low32 & 0xFFFF) << 16;
}

int64_t orc_range_encode(const int16_t* sym, int64_t count, const int32_t* cdf, int cdf_cols,
                         const int32_t* cdf_index, int precision, uint8_t* out, int64_t cap) {
  enc_t e;
  memset(&e, 0, sizeof e);
  e.size_minus1 = 0xFFFFFFFFu;
  e.out = out; e.cap = cap;
  for (int64_t i = 0; i < count; ++i) {
    int s = sym[i];
    if (s < 0 || s >= cdf_cols - 1) return -1;
    const int32_t* row = cdf + (int64_t)cdf_index[i] * cdf_cols;
    uint64_t lower = (uint64_t)row[s], upper = (uint64_t)row[s + 1];
    if (!(lower < upper)) return -2;
    uint64_t size = (uint64_t)e.size_minus1 + 1;
    uint32_t a = (uint32_t)((size * lower) >> precision);
    uint32_t b = (uint32_t)(((size * upper) >> precision) - 1);
    e.base += a;
    e.size_minus1 = b - a;
    if ((e.size_minus1 >> 16) == 0) {
      shift(&e);
      e.size_minus1 = (e.size_minus1 << 16) | 0xFFFF;
    }
  }
  /* RangeEncoder::Finalize of the upstream coder (see oracle/coder.py _Encoder.finish): an interval that still holds a multiple
   * of 2^32 above base is closed with that multiple (delayed word + 1, zeros after it left out); otherwise base is rounded up to a
   * multiple of 2^16, earlier words go out in full, the last word loses a zero low byte and is omitted when base's low 32 bits are 0. */
  uint32_t carry = (uint32_t)(e.base >> 32), low32 = (uint32_t)e.base;
  if (e.n + 2 * (e.pending + 2) > e.cap) return -3;
  if (!carry && (uint32_t)(low32 + e.size_minus1) < low32) {
    uint32_t w = (e.cache + 1) & 0xFFFF;
    out[e.n++] = (uint8_t)(w >> 8);
    if (w & 0xFF) out[e.n++] = (uint8_t)w;
  } else {
    if (e.have_cache) emit16(&e, (e.cache + carry) & 0xFFFF);
    for (int64_t i = 0; i < e.pending; ++i) emit16(&e, (0xFFFF + carry) & 0xFFFF);
    if (low32 != 0) {
      uint32_t mid = ((low32 - 1) >> 16) + 1;
      out[e.n++] = (uint8_t)(mid >> 8);
      if (mid & 0xFF) out[e.n++] = (uint8_t)mid;
    }
  }
  if (e.overflow) return -3;
  return e.n;
}

/* ---------------------------------------------------------------- range decoder */
typedef struct { const uint8_t* p; int64_t n, pos; } src_t;
static uint32_t read16(src_t* s) {
  uint32_t v = 0;
  for (int k = 0; k < 2; ++k) { v <<= 8; if (s->pos < s->n) v |= s->p[s->pos++]; }
  return v;
}

int orc_range_decode(const uint8_t* data, int64_t nbytes, int64_t count, const int32_t* cdf, int cdf_cols,
                     const int32_t* cdf_index, int precision, int16_t* out) {
  src_t src = {data, nbytes, 0};
  uint32_t base = 0, size_minus1 = 0xFFFFFFFFu;
  uint32_t value = read16(&src) << 16;
  value |= read16(&src);
  for (int64_t i = 0; i < count; ++i) {
    const int32_t* row = cdf + (int64_t)cdf_index[i] * cdf_cols;
    uint64_t size = (uint64_t)size_minus1 + 1;
    uint64_t offset = (((uint64_t)(uint32_t)(value - base) + 1) << precision) - 1;
    int lo = 1, hi = cdf_cols - 1;
    while (lo < hi) {
      int mid = (lo + hi) / 2;
      if (size * (uint64_t)row[mid] > offset) hi = mid; else lo = mid + 1;
    }
    int s = lo - 1;
    uint32_t a = (uint32_t)((size * (uint64_t)row[s]) >> precision);
    uint32_t b = (uint32_t)(((size * (uint64_t)row[s + 1]) >> precision) - 1);
    base += a;
    size_minus1 = b - a;
    if ((size_minus1 >> 16) == 0) {
      base <<= 16;
      size_minus1 = (size_minus1 << 16) | 0xFFFF;
      value = (value << 16) | read16(&src);
    }
    out[i] = (int16_t)s;
  }
  return 0;
}
