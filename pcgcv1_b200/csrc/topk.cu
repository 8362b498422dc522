// Top-k occupancy classification on the GPU (dataprocess/inout_points.py:147-179).
//
// select_voxels sorts every cube's 262144 logits on the host (values.sort(); thres = values[-num]).
// Here one CTA per cube finds the k-th largest logit with a deterministic 3-pass radix select
// (11/11/10-bit digits of the order-preserving uint32 image of the float, histograms in shared
// memory), then writes mask = (logit >= thres) exactly like np.greater_equal -- ties included, so
// the set matches the reference's bit for bit given the same logits.  The reference first filters
// vol > -2.0 and falls back to all voxels if fewer than k remain; both branches yield the k-th
// largest of the whole cube, which is what is computed.  k == 0 reproduces the values[-0] quirk
// (threshold = smallest logit > -2.0); k > V or an empty candidate set raise PCGC_ERR_BAD_RANGE
// where the reference raises IndexError.
#include <limits.h>

#include "common.cuh"

namespace pcgc {

constexpr int TK_THREADS = 1024;
constexpr int TK_BINS = 2048;

__device__ __forceinline__ uint32_t f2key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// Finds, among hist[0..nbins), the highest bin i with sum_{j>=i} hist[j] >= k.  Returns bin and the
// number of elements in higher bins through shared outputs.  All threads must call.
__device__ void select_bin(const uint32_t* hist, int nbins, uint32_t k, uint32_t* s_scan, uint32_t* s_bin,
                           uint32_t* s_above) {
  // thread t owns reversed bins r = 2t, 2t+1  (r = nbins-1-bin)
  const int t = threadIdx.x;
  uint32_t a = 0, b = 0;
  if (2 * t < nbins) a = hist[nbins - 1 - 2 * t];
  if (2 * t + 1 < nbins) b = hist[nbins - 2 - 2 * t];
  uint32_t sum = a + b, incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
    if ((t & 31) >= o) incl += n;
  }
  if ((t & 31) == 31) s_scan[t >> 5] = incl;
  __syncthreads();
  if (t < 32) {
    uint32_t w = s_scan[t], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t n = __shfl_up_sync(0xffffffffu, wi, o);
      if (t >= o) wi += n;
    }
    s_scan[t] = wi - w;                      // exclusive warp offsets
  }
  __syncthreads();
  const uint32_t before = s_scan[t >> 5] + incl - sum;   // elements in bins above reversed bin 2t
  if (before < k && before + a >= k) { *s_bin = (uint32_t)(nbins - 1 - 2 * t); *s_above = before; }
  else if (before + a < k && before + sum >= k) { *s_bin = (uint32_t)(nbins - 2 - 2 * t); *s_above = before + a; }
  __syncthreads();
}

__global__ void __launch_bounds__(TK_THREADS)
topk_kernel(const float* __restrict__ logits, int64_t V, const int32_t* __restrict__ ks, uint8_t* __restrict__ mask,
            float* __restrict__ thres_out, int32_t* __restrict__ count_out, int* __restrict__ err) {
  __shared__ uint32_t hist[TK_BINS];
  __shared__ uint32_t s_scan[32];
  __shared__ uint32_t s_bin, s_above;
  __shared__ float s_red[32];
  __shared__ int s_cnt[32];
  const int b = blockIdx.x;
  const float* v = logits + (size_t)b * V;
  const float4* v4 = reinterpret_cast<const float4*>(v);
  const int64_t n4 = V / 4;
  const int k_in = ks[b];
  float thres;

  if (k_in < 0 || (int64_t)k_in > V) {
    if (threadIdx.x == 0) atomicExch(err, PCGC_ERR_BAD_RANGE);
    thres = __int_as_float(0x7f800000);        // +inf: empty mask
  } else if (k_in == 0) {
    // values[-0] == values[0]: the minimum of {v > -2.0}
    float m = __int_as_float(0x7f800000);
    for (int64_t i = threadIdx.x; i < V; i += TK_THREADS) { const float f = v[i]; if (f > -2.0f) m = fminf(m, f); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < TK_THREADS / 32; ++i) m = fminf(m, s_red[i]);
      s_red[0] = m;
      if (m == __int_as_float(0x7f800000)) atomicExch(err, PCGC_ERR_BAD_RANGE);   // IndexError in the reference
    }
    __syncthreads();
    thres = s_red[0];
  } else {
    uint32_t k = (uint32_t)k_in, prefix = 0;
    // digit 0: bits 31..21, digit 1: bits 20..10, digit 2: bits 9..0
    const int shifts[3] = {21, 10, 0};
    const int nb[3] = {2048, 2048, 1024};
    for (int pass = 0; pass < 3; ++pass) {
      for (int i = threadIdx.x; i < TK_BINS; i += TK_THREADS) hist[i] = 0;
      __syncthreads();
      const int sh = shifts[pass];
      const uint32_t dm = (uint32_t)nb[pass] - 1;
      const uint32_t pmask = pass == 0 ? 0u : (pass == 1 ? 0xFFE00000u : 0xFFFFFC00u);
      for (int64_t i = threadIdx.x; i < n4; i += TK_THREADS) {
        const float4 f = __ldg(v4 + i);
        const uint32_t k0 = f2key(f.x), k1 = f2key(f.y), k2 = f2key(f.z), k3 = f2key(f.w);
        if ((k0 & pmask) == prefix) atomicAdd(&hist[(k0 >> sh) & dm], 1u);
        if ((k1 & pmask) == prefix) atomicAdd(&hist[(k1 >> sh) & dm], 1u);
        if ((k2 & pmask) == prefix) atomicAdd(&hist[(k2 >> sh) & dm], 1u);
        if ((k3 & pmask) == prefix) atomicAdd(&hist[(k3 >> sh) & dm], 1u);
      }
      __syncthreads();
      select_bin(hist, nb[pass], k, s_scan, &s_bin, &s_above);
      prefix |= s_bin << sh;
      k -= s_above;
      __syncthreads();
    }
    thres = key2f(prefix);
  }

  // mask = logits >= thres (np.greater_equal)
  int cnt = 0;
  uchar4* m4 = reinterpret_cast<uchar4*>(mask + (size_t)b * V);
  for (int64_t i = threadIdx.x; i < n4; i += TK_THREADS) {
    const float4 f = __ldg(v4 + i);
    uchar4 m;
    m.x = f.x >= thres; m.y = f.y >= thres; m.z = f.z >= thres; m.w = f.w >= thres;
    cnt += m.x + m.y + m.z + m.w;
    m4[i] = m;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int c = 0;
    for (int i = 0; i < TK_THREADS / 32; ++i) c += s_cnt[i];
    if (count_out) count_out[b] = c;
    if (thres_out) thres_out[b] = thres;
  }
}

cudaError_t launch_topk(const float* logits, int B, int64_t V, const int32_t* ks, uint8_t* mask, float* thres,
                        int32_t* count, int* err_flag, cudaStream_t s, int64_t* launches) {
  if (V % 4 != 0) return cudaErrorInvalidValue;
  PCGC_CARVEOUT_ONCE(topk_kernel);
  topk_kernel<<<B, TK_THREADS, 0, s>>>(logits, V, ks, mask, thres, count, err_flag);
  if (launches) ++*launches;
  return cudaGetLastError();
}

__global__ void threshold_kernel(const float* __restrict__ logits, int64_t V, float thres, uint8_t* __restrict__ mask,
                                 int32_t* __restrict__ count) {
  __shared__ int s_cnt[8];
  const int b = blockIdx.y;
  const float4* v4 = reinterpret_cast<const float4*>(logits + (size_t)b * V);
  uchar4* m4 = reinterpret_cast<uchar4*>(mask + (size_t)b * V);
  int cnt = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < V / 4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 f = __ldg(v4 + i);
    uchar4 m;
    m.x = f.x >= thres; m.y = f.y >= thres; m.z = f.z >= thres; m.w = f.w >= thres;
    cnt += m.x + m.y + m.z + m.w;
    m4[i] = m;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0 && count) {
    int c = 0;
    for (int i = 0; i < (int)blockDim.x / 32; ++i) c += s_cnt[i];
    atomicAdd(count + b, c);
  }
}

cudaError_t launch_threshold(const float* logits, int B, int64_t V, float thres, uint8_t* mask, int32_t* count,
                             cudaStream_t s, int64_t* launches) {
  if (V % 4 != 0) return cudaErrorInvalidValue;
  if (count) { cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int32_t) * B, s); if (e != cudaSuccess) return e; }
  dim3 grid(32, (unsigned)B);
  threshold_kernel<<<grid, 256, 0, s>>>(logits, V, thres, mask, count);
  if (launches) ++*launches;
  return cudaGetLastError();
}

}  // namespace pcgc
