// pmf -> 16-bit quantised CDF normaliser shared by the host coder (coder.cpp) and the CDF kernels
// (entropy.cu).  Replaces coder_ops.pmf_to_quantized_cdf (call sites models/entropy_model.py:218,
// models/conditional_entropy_model.py:122; upstream tensorflow/contrib/coder/kernels/pmf_to_cdf_op.cc,
// not vendored -- see oracle/coder.py for the restated contract).
//
// Contract (identical to the oracle's greedy definition):
//   v_i = max(rint(pmf_i * 2^precision), 1)
//   sum > target: repeatedly decrement the entry with the smallest penalty
//                 pmf_i*(log2 v_i - log2(v_i-1)), entries at 1 excluded, lowest index on ties;
//   sum < target: repeatedly increment the entry with the largest gain
//                 pmf_i*(log2(v_i+1) - log2 v_i), lowest index on ties.
// The deficit case can need thousands of steps (Laplace tails cut at min_v/max_v), so it is done
// as an exact water-filling: every increment whose gain is >= a threshold lambda that provably
// admits at most `deficit` increments is granted at once (counts verified with exact gain
// evaluations), and only the remainder runs the step-by-step greedy.  Because per-entry gains
// are strictly decreasing, the result equals the step-by-step greedy's.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PCGC_HD __host__ __device__ __forceinline__
#else
#define PCGC_HD inline
#endif

namespace pcgc {

PCGC_HD double cdf_gain(double m, int v) { return m * (log2((double)(v + 1)) - log2((double)v)); }

// pmf[n] -> v[n] (counts, sum == 2^precision).  g: scratch double[n].  Returns 0, or -2 if the row
// cannot be shrunk (all ones and n > target).
PCGC_HD int quantize_pmf_row(const float* pmf, int n, int precision, int32_t* v, double* g) {
  const int target = 1 << precision;
  const float scale = (float)target;
  long long sum = 0;
  for (int i = 0; i < n; ++i) {
    int q = (int)rintf(pmf[i] * scale);
    q = q < 1 ? 1 : q;
    v[i] = q;
    sum += q;
  }
  if (sum > target) {
    long long surplus = sum - target;
    for (int i = 0; i < n; ++i) g[i] = v[i] > 1 ? cdf_gain((double)pmf[i], v[i] - 1) : INFINITY;
    while (surplus > 0) {
      int best = -1;
      double bp = INFINITY;
      for (int i = 0; i < n; ++i)
        if (g[i] < bp) { bp = g[i]; best = i; }
      if (best < 0) return -2;
      v[best] -= 1;
      g[best] = v[best] > 1 ? cdf_gain((double)pmf[best], v[best] - 1) : INFINITY;
      --surplus;
    }
    return 0;
  }
  long long deficit = target - sum;
  if (deficit == 0) return 0;

  bool have_gains = false;
  if (deficit > 2 * n + 8) {
    // ---- water-filling: continuous solution of gain_i(v) ~ m_i*log2(e)/(v+0.5) == lambda ----
    const double L = 1.4426950408889634;
    // active set: entries that grow at the solution; iterate a few times.
    double inv_lambda = 0.0;                 // 1/lambda
    {
      double t_act = (double)target, m_act = 0.0;
      int n_act = n;
      for (int i = 0; i < n; ++i) m_act += (double)pmf[i];
      for (int it = 0; it < 4; ++it) {
        inv_lambda = m_act > 0.0 ? (t_act + 0.5 * n_act) / (L * m_act) : 0.0;
        double t2 = (double)target, m2 = 0.0;
        int n2 = 0;
        for (int i = 0; i < n; ++i) {
          const double want = (double)pmf[i] * L * inv_lambda - 0.5;
          if (want >= (double)v[i]) { m2 += (double)pmf[i]; ++n2; } else { t2 -= (double)v[i]; }
        }
        if (n2 == n_act && m2 == m_act) break;
        t_act = t2; m_act = m2; n_act = n2;
        if (n2 == 0) break;
      }
    }
    double lambda = inv_lambda > 0.0 ? 1.0 / inv_lambda : INFINITY;
    for (int attempt = 0; attempt < 64 && lambda < INFINITY; ++attempt) {
      // exact count c_i = #{k >= 0 : gain_i(v_i + k) >= lambda}, gains cached for the final greedy
      long long granted = 0;
      bool ok = true;
      for (int i = 0; i < n && ok; ++i) {
        const double m = (double)pmf[i];
        double approx = m * L / lambda - 0.5;                 // largest v with gain >= lambda (approx)
        long long c = approx >= (double)v[i] ? (long long)(approx - (double)v[i]) + 1 : 0;
        if (c > deficit) c = deficit + 1;
        while (c > 0 && cdf_gain(m, (int)(v[i] + c - 1)) < lambda) --c;
        double nxt = cdf_gain(m, (int)(v[i] + c));
        while (nxt >= lambda) {
          ++c;
          if (c > deficit) break;
          nxt = cdf_gain(m, (int)(v[i] + c));
        }
        g[i] = nxt;
        granted += c;
        v[i] += (int32_t)c;                                   // undone below if infeasible
        if (granted > deficit) ok = false;
      }
      if (ok) { deficit -= granted; have_gains = true; break; }
      // infeasible: restore v from pmf and raise lambda slightly
      for (int i = 0; i < n; ++i) { int q = (int)rintf(pmf[i] * scale); v[i] = q < 1 ? 1 : q; }
      lambda *= 1.0 + ldexp(1.0, -14 + attempt / 2);
    }
  }
  if (!have_gains)
    for (int i = 0; i < n; ++i) g[i] = cdf_gain((double)pmf[i], v[i]);
  while (deficit > 0) {
    int best = 0;
    double bg = -INFINITY;
    for (int i = 0; i < n; ++i)
      if (g[i] > bg) { bg = g[i]; best = i; }
    v[best] += 1;
    g[best] = cdf_gain((double)pmf[best], v[best]);
    --deficit;
  }
  return 0;
}

}  // namespace pcgc
