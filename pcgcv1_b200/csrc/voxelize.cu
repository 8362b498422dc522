// Points <-> occupancy cubes on the GPU (dataprocess/inout_points.py:116-143).
//
// points2voxels builds a float64 [B,S,S,S,1] tensor on the host (2 MiB per cube, then copied to the GPU);
// voxels2points pulls the float32 mask back and runs np.where per cube.  Here the partitioned points (6 B each) go up,
// a scatter kernel sets the uint8 occupancy bytes, and on the way back an ordered compaction turns the top-k mask into
// the coordinate list directly -- lexicographic (d,h,w) order per cube, cubes in order, exactly np.where's order -- so
// only 6 B per point cross PCIe instead of 256 KiB per cube.  Both kernels are byte/index work bound by HBM.
#include "common.cuh"

namespace pcgc {

constexpr int VX_THREADS = 256;
constexpr int VX_PER_THREAD = 16;                    // one 16-byte load
constexpr int VX_CHUNK = VX_THREADS * VX_PER_THREAD; // voxels per block

__global__ void voxelize_kernel(const int16_t* __restrict__ local, const int64_t* __restrict__ offsets, int B, int S,
                                int64_t n, uint8_t* __restrict__ cubes, int* err) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int lo = 0, hi = B;                                // cube b with offsets[b] <= i < offsets[b+1]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (offsets[mid] <= i) lo = mid; else hi = mid;
  }
  const int x = local[3 * i], y = local[3 * i + 1], z = local[3 * i + 2];
  if ((unsigned)x >= (unsigned)S || (unsigned)y >= (unsigned)S || (unsigned)z >= (unsigned)S) { atomicExch(err, -201); return; }
  cubes[(((int64_t)lo * S + x) * S + y) * S + z] = 1;   // duplicates write the same byte
}

__device__ __forceinline__ int count16(const uint4 v) {
  // number of non-zero bytes among 16
  auto nz = [](uint32_t w) { return ((w & 0xFFu) != 0) + ((w & 0xFF00u) != 0) + ((w & 0xFF0000u) != 0) + ((w & 0xFF000000u) != 0); };
  return nz(v.x) + nz(v.y) + nz(v.z) + nz(v.w);
}

__global__ void __launch_bounds__(VX_THREADS) extract_count_kernel(const uint8_t* __restrict__ mask, int64_t* __restrict__ chunk_count) {
  const int64_t chunk = blockIdx.x;
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(mask + chunk * VX_CHUNK) + threadIdx.x);
  int c = count16(v);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  __shared__ int s[VX_THREADS / 32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < VX_THREADS / 32; ++w) t += s[w];
    chunk_count[chunk] = t;
  }
}

// Exclusive scan of the chunk counts in place (one block walks the array in fixed order: deterministic), plus the
// per-cube totals and the grand total.
__global__ void __launch_bounds__(1024) extract_scan_kernel(int64_t* __restrict__ chunk, int64_t n_chunks, int chunks_per_cube, int B,
                                                            int32_t* __restrict__ counts, int64_t* __restrict__ total) {
  __shared__ int64_t s_warp[32];
  __shared__ int64_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n_chunks; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int64_t v = i < n_chunks ? chunk[i] : 0;
    int64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      const int64_t w = s_warp[threadIdx.x];
      int64_t wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, wi, o);
        if (threadIdx.x >= o) wi += t;
      }
      s_warp[threadIdx.x] = wi - w;
    }
    __syncthreads();
    const int64_t excl = s_carry + s_warp[threadIdx.x >> 5] + incl - v;
    if (i < n_chunks) chunk[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = excl + v;
    __syncthreads();
  }
  // per-cube counts from the exclusive offsets
  for (int b = threadIdx.x; b < B; b += 1024) {
    const int64_t a = chunk[(int64_t)b * chunks_per_cube];
    const int64_t e = b + 1 < B ? chunk[(int64_t)(b + 1) * chunks_per_cube] : s_carry;
    counts[b] = (int32_t)(e - a);
  }
  if (threadIdx.x == 0) *total = s_carry;
}

__global__ void __launch_bounds__(VX_THREADS) extract_write_kernel(const uint8_t* __restrict__ mask, const int64_t* __restrict__ chunk_off,
                                                                    int S, int chunks_per_cube, int16_t* __restrict__ points, int64_t cap) {
  const int64_t chunk = blockIdx.x;
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(mask + chunk * VX_CHUNK) + threadIdx.x);
  const int c = count16(v);
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += t;
  }
  __shared__ int s[VX_THREADS / 32];
  if ((threadIdx.x & 31) == 31) s[threadIdx.x >> 5] = incl;
  __syncthreads();
  int before = incl - c;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += s[w];
  if (c == 0) return;
  int64_t o = chunk_off[chunk] + before;
  const int64_t vox0 = (chunk % chunks_per_cube) * VX_CHUNK + (int64_t)threadIdx.x * VX_PER_THREAD;   // voxel index inside the cube
  const uint32_t words[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < VX_PER_THREAD; ++j) {
    if ((words[j >> 2] >> (8 * (j & 3))) & 0xFFu) {
      const int64_t vi = vox0 + j;
      if (o < cap) {
        points[3 * o] = (int16_t)(vi / ((int64_t)S * S));
        points[3 * o + 1] = (int16_t)((vi / S) % S);
        points[3 * o + 2] = (int16_t)(vi % S);
      }
      ++o;
    }
  }
}

cudaError_t launch_voxelize(const int16_t* local, const int64_t* offsets, int B, int S, int64_t n, uint8_t* cubes, int* err,
                            cudaStream_t s, int64_t* launches) {
  cudaError_t e = cudaMemsetAsync(cubes, 0, (size_t)B * S * S * S, s);
  if (e != cudaSuccess || n == 0) return e;
  voxelize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(local, offsets, B, S, n, cubes, err);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_extract_points(const uint8_t* mask, int B, int S, int64_t* chunk_ws, int32_t* counts, int16_t* points, int64_t cap,
                                  int64_t* total, cudaStream_t s, int64_t* launches) {
  const int64_t V = (int64_t)S * S * S;
  const int cpc = (int)(V / VX_CHUNK);
  const int64_t n_chunks = (int64_t)B * cpc;
  extract_count_kernel<<<(unsigned)n_chunks, VX_THREADS, 0, s>>>(mask, chunk_ws);
  extract_scan_kernel<<<1, 1024, 0, s>>>(chunk_ws, n_chunks, cpc, B, counts, total);
  extract_write_kernel<<<(unsigned)n_chunks, VX_THREADS, 0, s>>>(mask, chunk_ws, S, cpc, points, cap);
  *launches += 3;
  return cudaGetLastError();
}

}  // namespace pcgc
