// Equivalence check of the two forms of the 16-bit CDF normaliser and of the shared-edge likelihood row (host build of the same
// headers the CUDA kernels compile): quantize_pmf_row_reg<8|16|32> == quantize_pmf_row and det_laplace_pmf_row[_reg] ==
// per-symbol det_laplace_likelihood, bit for bit, on random Laplace rows over all symbol counts and on generic / unnormalised
// rows (both directions, large deficits).  g++ -O2 -ffp-contract=off -I pcgcv1_b200/csrc tools/cdf_check.cpp; exit code 0 = equal.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#include <cmath>
#include <random>
#define PCGC_MAX_SYMBOLS 64
#include "cdf_norm.h"
#include "det_math.h"
using namespace pcgc;
static long long n_rows = 0, n_bad = 0;
template <int NMAX> static void one(const float* pmf, int N, int min_v, float l, float s, bool from_ls) {
  float g[64]; int32_t a[64];
  float pr[NMAX]; int32_t b[NMAX];
  float pm[64];
  if (from_ls) { det_laplace_pmf_row(min_v, N, l, s, 1e-9f, pm);
    for (int k = 0; k < N; ++k) { const float ref = fmaxf(det_laplace_likelihood((float)(min_v + k), l, s), 1e-9f); if (memcmp(&ref, &pm[k], 4)) { ++n_bad; printf("LIKELIHOOD MISMATCH\n"); } } det_laplace_pmf_row_reg<NMAX>(min_v, N, l, s, 1e-9f, pr); if (memcmp(pm, pr, 4 * N)) { ++n_bad; printf("PMF MISMATCH\n"); } }
  else { for (int k = 0; k < NMAX; ++k) pr[k] = k < N ? pmf[k] : 0.f; memcpy(pm, pmf, 4 * N); }
  int ra = quantize_pmf_row(pm, N, 16, a, g), rb = quantize_pmf_row_reg<NMAX>(pr, N, 16, b);
  ++n_rows;
  if (ra != rb || (ra == 0 && memcmp(a, b, 4 * N))) { if (n_bad++ < 5) printf("MISMATCH N=%d NMAX=%d rc %d %d\n", N, NMAX, ra, rb); }
}
static void row(const float* pmf, int N, int min_v = 0, float l = 0, float s = 0, bool from_ls = false) {
  if (N <= 8) one<8>(pmf, N, min_v, l, s, from_ls);
  if (N <= 16) one<16>(pmf, N, min_v, l, s, from_ls);
  if (N <= 32) one<32>(pmf, N, min_v, l, s, from_ls);
}
int main() {
  FILE* f = nullptr;
  if (f) {
    int64_t hdr[2]; if (fread(hdr, 8, 2, f) != 2) return 1;
    int B = hdr[0]; int64_t E = hdr[1];
    std::vector<int32_t> mm(2 * B); if (fread(mm.data(), 4, 2 * B, f) != (size_t)2 * B) return 1;
    std::vector<float> loc(B * E), sc(B * E);
    if (fread(loc.data(), 4, B * E, f) != (size_t)(B * E) || fread(sc.data(), 4, B * E, f) != (size_t)(B * E)) return 1;
    for (int b = 0; b < B; ++b) { int min_v = mm[2 * b], N = mm[2 * b + 1] - min_v + 1; for (int64_t e = 0; e < E; ++e) row(nullptr, N, min_v, loc[b * E + e], sc[b * E + e], true); }
    printf("workload rows: %lld checks, %lld mismatches\n", n_rows, n_bad); fflush(stdout);
  }
  std::mt19937_64 rng(99); std::uniform_real_distribution<double> U(0, 1);
  for (int it = 0; it < 300000; ++it) {
    int min_v = -(int)(U(rng) * 16), max_v = (int)(U(rng) * 16); if (max_v == min_v) max_v++;
    int N = max_v - min_v + 1; if (N > 32) continue;
    float l = (float)((U(rng) - 0.5) * 12), s = (float)std::exp((U(rng) - 0.6) * 7);
    row(nullptr, N, min_v, l, s, true);
  }
  printf("random Laplace rows: %lld checks, %lld mismatches\n", n_rows, n_bad); fflush(stdout);
  for (int it = 0; it < 60000; ++it) {
    int N = 2 + (int)(U(rng) * 31); if (N > 32) N = 32;
    float pmf[32]; double tot = 0; int kind = it % 5;
    for (int k = 0; k < N; ++k) { double x = U(rng); if (kind == 1) x = std::pow(x, 8); else if (kind == 2) x = (k % 3 == 0) ? 0.3 : 1e-7 * x; else if (kind == 3) x = 1.0; else if (kind == 4) x = std::exp(-12 * x); pmf[k] = (float)x; tot += x; }
    double scale = ((it % 50 == 0 ? 0.6 : 0.995) + (it % 50 == 0 ? 0.8 : 0.01) * U(rng)) / tot;
    for (int k = 0; k < N; ++k) pmf[k] = fmaxf((float)(pmf[k] * scale), 1e-9f);
    row(pmf, N);
  }
  printf("generic rows: %lld checks, %lld mismatches\n", n_rows, n_bad);
  return n_bad != 0;
}
