// tcgen05 window-GEMM kernel for the large-kernel stride-2 layers of model_simple (models/model_simple.py:21-42,58-86:
// Conv3D 9^3 s2 1->32, 5^3 s2 32->32; Conv3DTranspose 5^3 s2 32->32, 9^3 s2 32->1), 96 % of that model's MACs.
//
// Every one of those layers is a STRIDE-1 window GEMM once the stride is folded into channels or columns:
//   * stride-2 conv, kernel k: on the space-to-depth input X'[t][p*C + c] = x[2t + p][c] the taps k = 2*cell + parity - before
//     become a ceil(k/2)+... window of whole cells: 9^3 s2 -> 5^3 cells x 8 parities, 5^3 s2 -> 3^3 cells x 8 parities
//     (cell/parity pairs without a tap carry zero weights; all-zero K blocks are never issued);
//   * stride-2 transposed conv, kernel k: output voxel 2t + r gathers x[t + c] W[r + before - 2c]: a 3^3 (k = 5) or 5^3 (k = 9)
//     window over the INPUT grid whose column blocks are the 8 output-parity classes r (gather form, no atomics, deterministic).
// GEMM view per CTA: M = 128 rows = 8 (x) x 16 (y) voxels of one z slice, zt slices per tile, N = np columns, K = window x Cin.
// K is walked in CHUNKS = (window z cell, 16 input channels): one 5-D TMA box {8+wx-1 cells, 16+wy-1 lines, zt slices, 4 planes}
// lands the chunk's haloed brick in the UMMA no-swizzle K-major layout, the (ty, tx) taps of the chunk are descriptor
// start-address offsets into it (umma_conv.cu's scheme), read from a small table so that one kernel serves every window shape
// and zero structure.  Cin = 8 (the 8 parities of the 1-channel occupancy cube) pairs two taps per K = 16.
// Split-bf16 operands as everywhere in the engine: D = x_hi*[w_hi | w_lo] + x_lo*w_hi, FP32 accumulation in TMEM.
// Pipeline: one elected thread (warp 8) issues TMA + MMAs over a 2-stage ring (chunk g+1 loads while chunk g multiplies: the load
// is issued after the MMAs of g are queued, when chunk g-1 -- the stage's previous user -- retires); two TMEM accumulator sets
// alternate between tiles so the 8 epilogue warps drain tile i while tile i+1 multiplies.  Persistent, one CTA per SM.
// 8^3 grids (smaller than a 16-line M tile) use zp = 2: the tensor map lists z before y, so the brick is [y][2 slices][x] and the 16
// row groups (y, z) of an 8 x 8 x 2 tile are again equally spaced (SBO = one x run).
// Deterministic: fixed chunk / tap order, no split-K, no atomics.
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "pm_format.cuh"
#include "umma_dev.cuh"
#include "umma_win.cuh"

namespace pcgc {

namespace {

constexpr int WIN_EPI_WARPS = 8;
constexpr int WIN_THREADS = 32 * (WIN_EPI_WARPS + 1);

struct WinArgs {
  int n, zt, batch;
  int exc, ey;                 // brick cells along x, lines along y
  int ox, oy;                  // brick origin relative to the tile origin
  int paired;                  // cin == 8: K = 16 spans two taps (LBO per table entry)
  int zp;                      // 1: M rows = 8 x * 16 y of one slice; 2: 8 x * 8 y * 2 z (8^3 grids: brick laid out [y][z][x])
  int n_chunks, n_entries;
  int plane_bytes, a_bytes, tile_bytes, stage_bytes;
  const __nv_bfloat16* wpacked;
  const int4* chunks;          // {dz, plane0, first entry, entries}
  const uint32_t* entries;     // A-descriptor increment
  const float* bias;
  int n_real, flags;
  float* out_f32; int out_cs, out_co;
  __nv_bfloat16* out_pm; int out_planes, out_s2d;
  int up_ncls, up_cout, up_cls[8];
  int nsets, set_cols, tmem_cols;
  int* err;
};

template <int NP, int EPI>
__global__ void __launch_bounds__(WIN_THREADS, 1) conv_umma_win_kernel(const __grid_constant__ CUtensorMap tmap, const WinArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + 2 * (size_t)a.stage_bytes);   // full[2], free[2], done[2], accfree[2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 8);
  float* s_bias = reinterpret_cast<float*>(s_bar + 10);                             // 16-byte aligned
  int4* s_chunks = reinterpret_cast<int4*>(s_bias + NP);
  uint32_t* s_entries = reinterpret_cast<uint32_t*>(s_chunks + a.n_chunks);
  const uint32_t bar_full = smem_u32(s_bar), bar_free = smem_u32(s_bar + 2), bar_done = smem_u32(s_bar + 4), bar_accfree = smem_u32(s_bar + 6);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tx_n = a.n / 8, ty_n = a.zp == 2 ? a.n / 8 : a.n / 16, tz_n = a.n / (a.zp == 2 ? 2 : a.zt);
  const int total_tiles = tx_n * ty_n * tz_n * a.batch;
  const int my_tiles = (int)blockIdx.x < total_tiles ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  for (int i = tid; i < NP; i += WIN_THREADS) s_bias[i] = a.bias ? a.bias[i] : 0.f;
  for (int i = tid; i < a.n_chunks; i += WIN_THREADS) s_chunks[i] = a.chunks[i];
  for (int i = tid; i < a.n_entries; i += WIN_THREADS) s_entries[i] = a.entries[i];
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_full + 8 * i, 1); mbar_init(bar_free + 8 * i, 1);
      mbar_init(bar_done + 8 * i, 1); mbar_init(bar_accfree + 8 * i, WIN_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WIN_EPI_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(a.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *s_tmem;

  auto tile_origin = [&](int it, int& b, int& x0, int& y0, int& z0) {
    int r = (int)blockIdx.x + it * (int)gridDim.x;
    x0 = (r % tx_n) * 8; r /= tx_n;
    y0 = (r % ty_n) * (a.zp == 2 ? 8 : 16); r /= ty_n;
    z0 = (r % tz_n) * (a.zp == 2 ? 2 : a.zt); r /= tz_n;
    b = r;
  };

  if (warp == WIN_EPI_WARPS) {
    if (elect_one() && my_tiles > 0) {
      // ------------------------------ TMA producer + MMA issuer (one thread) ------------------------------
      const uint32_t stage0 = smem_u32(smem);
      const uint32_t PL = (uint32_t)a.plane_bytes, SBO = (uint32_t)a.exc * 16u;
      const uint32_t klbo = a.paired ? 0u : 2u * PL;
      const uint64_t z_step = (uint64_t)((a.ey * a.exc * 16) >> 4);
      constexpr uint64_t b_step = (uint64_t)((2 * NP * 32) >> 4);
      constexpr uint32_t idesc_full = make_idesc(128, 2 * NP), idesc_half = make_idesc(128, NP);
      const long total_chunks = (long)my_tiles * a.n_chunks;
      auto load = [&](int it, int c, int st) {
        int b, x0, y0, z0;
        tile_origin(it, b, x0, y0, z0);
        const int4 ck = s_chunks[c];
        const uint32_t dst = stage0 + (uint32_t)st * (uint32_t)a.stage_bytes;
        const uint32_t bytes_b = (uint32_t)ck.w * (uint32_t)a.tile_bytes;
        mbar_expect_tx(bar_full + 8 * st, (uint32_t)a.a_bytes + bytes_b);
        if (a.zp == 2) tma_load_5d(dst, &tmap, bar_full + 8 * st, (x0 + a.ox) * 8, z0 + ck.x, y0 + a.oy, ck.y, b);   // dims (x, z, y, plane, cube)
        else tma_load_5d(dst, &tmap, bar_full + 8 * st, (x0 + a.ox) * 8, y0 + a.oy, z0 + ck.x, ck.y, b);
        bulk_load(dst + (uint32_t)a.a_bytes, reinterpret_cast<const uint8_t*>(a.wpacked) + (size_t)ck.z * a.tile_bytes, bytes_b, bar_full + 8 * st);
      };
      load(0, 0, 0);
      bool alive = true;
      int it = 0, c = 0;
      for (long g = 0; g < total_chunks && alive; ++g) {
        const int st = (int)(g & 1);
        const int set = it % a.nsets, suse = it / a.nsets;
        if (c == 0 && suse >= 1) { alive = mbar_wait(bar_accfree + 8 * set, (uint32_t)(suse - 1) & 1u, a.err, -134); if (!alive) break; }
        alive = mbar_wait(bar_full + 8 * st, (uint32_t)(g >> 1) & 1u, a.err, -131);
        if (!alive) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int4 ck = s_chunks[c];
        const uint32_t brick = stage0 + (uint32_t)st * (uint32_t)a.stage_bytes;
        const uint64_t a_hi0 = make_desc(brick, klbo, SBO), a_lo0 = make_desc(brick + PL, klbo, SBO);
        const uint64_t b0 = make_desc(brick + (uint32_t)a.a_bytes, 2 * NP * 16, 128);
        for (int zi = 0; zi < a.zt; ++zi) {
          const uint32_t d = tmem_base + (uint32_t)(set * a.set_cols + zi * 2 * NP);
          const uint64_t ah = a_hi0 + zi * z_step, al = a_lo0 + zi * z_step;
#pragma unroll 4
          for (int e = 0; e < ck.w; ++e) {
            const uint64_t add = (uint64_t)s_entries[ck.z + e];
            const uint64_t bd = b0 + (uint64_t)e * b_step;
            umma_f16(d, ah + add, bd, idesc_full, (c == 0 && e == 0) ? 0u : 1u);    // x_hi * [w_hi | w_lo]
            umma_f16(d, al + add, bd, idesc_half, 1u);                              // x_lo * w_hi
          }
        }
        umma_commit(bar_free + 8 * st);                       // the stage is reusable once these MMAs retire
        const bool last = c + 1 == a.n_chunks;
        if (last) umma_commit(bar_done + 8 * set);            // accumulators of the tile final
        int nit = it, nc = c + 1;
        if (last) { nit = it + 1; nc = 0; }
        if (g + 1 < total_chunks) {
          // chunk g+1 goes into the stage chunk g-1 used: wait for g-1's MMAs (they retire while g's are running)
          if (g >= 1) { alive = mbar_wait(bar_free + 8 * (st ^ 1), (uint32_t)((g - 1) >> 1) & 1u, a.err, -132); if (!alive) break; }
          load(nit, nc, st ^ 1);
        }
        it = nit; c = nc;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ epilogue: 8 warps; warp w reads TMEM lanes 32*(w&3).., z slices of parity w>>2 ------------------------------
    const int row = (warp & 3) * 32 + lane;
    const size_t plane_elems = (size_t)a.n * a.n * a.n * 8;
    for (int it = 0; it < my_tiles; ++it) {
      int b, x0, y0, z0;
      tile_origin(it, b, x0, y0, z0);
      const int set = it % a.nsets, suse = it / a.nsets;
      const int vx = x0 + (row & 7), vy = y0 + (a.zp == 2 ? (row >> 4) : (row >> 3));
      if (!mbar_wait(bar_done + 8 * set, (uint32_t)suse & 1u, a.err, -133)) break;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int zi = (warp >> 2); zi < a.zt; zi += 2) {
        const int vz = z0 + (a.zp == 2 ? ((row >> 3) & 1) : zi);
        const uint32_t lane_base = tmem_base + (uint32_t)(set * a.set_cols + zi * 2 * NP) + ((uint32_t)((warp & 3) * 32) << 16);
        if (EPI == WEPI_UP_PM) {
          const int on = 2 * a.n;
          const size_t out_plane = (size_t)on * on * on * 8;
          __nv_bfloat16* ob = a.out_pm + (size_t)b * a.out_planes * out_plane;
#pragma unroll 1
          for (int cls = 0; cls < a.up_ncls; ++cls) {
            const int gc = a.up_cls[cls];
            const int oz = 2 * vz + ((gc >> 2) & 1), oy = 2 * vy + ((gc >> 1) & 1), ox = 2 * vx + (gc & 1);
            __nv_bfloat16* oc = ob + (((size_t)oz * on + oy) * on + ox) * 8;
            for (int j = 0; j < a.up_cout / 16; ++j) {
              const int col = cls * a.up_cout + j * 16;
              float d1[16], d2[16], t[16];
              tmem_ld16(lane_base + (uint32_t)col, d1);
              tmem_ld16(lane_base + (uint32_t)(NP + col), d2);
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float v = (d1[i] + d2[i]) + s_bias[col + i];
                t[i] = (a.flags & EPI_RELU) ? fmaxf(v, 0.f) : v;
              }
              split_store(oc + (size_t)(4 * j) * out_plane, oc + (size_t)(4 * j + 1) * out_plane, t);
              split_store(oc + (size_t)(4 * j + 2) * out_plane, oc + (size_t)(4 * j + 3) * out_plane, t + 8);
            }
          }
        } else if (EPI == WEPI_UP_F32) {
          // one output channel per class: columns 0..up_ncls-1
          float d1[16], d2[16];
          tmem_ld16(lane_base, d1);
          tmem_ld16(lane_base + (uint32_t)NP, d2);
          const int on = 2 * a.n;
          float* ob = a.out_f32 + (size_t)b * on * on * on * a.out_cs + a.out_co;
#pragma unroll
          for (int cls = 0; cls < 8; ++cls) {
            if (cls < a.up_ncls) {
              const int gc = a.up_cls[cls];
              const int oz = 2 * vz + ((gc >> 2) & 1), oy = 2 * vy + ((gc >> 1) & 1), ox = 2 * vx + (gc & 1);
              float v = (d1[cls] + d2[cls]) + s_bias[cls];
              if (a.flags & EPI_RELU) v = fmaxf(v, 0.f);
              ob[(((size_t)oz * on + oy) * on + ox) * a.out_cs] = v;
            }
          }
        } else {
          float v[NP];
#pragma unroll
          for (int j = 0; j < NP / 16; ++j) {
            float d1[16], d2[16];
            tmem_ld16(lane_base + (uint32_t)(j * 16), d1);
            tmem_ld16(lane_base + (uint32_t)(NP + j * 16), d2);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float t = (d1[i] + d2[i]) + s_bias[j * 16 + i];
              v[j * 16 + i] = (a.flags & EPI_RELU) ? fmaxf(t, 0.f) : t;
            }
          }
          const size_t vox = ((size_t)vz * a.n + vy) * a.n + vx;
          if (EPI == WEPI_F32) {
            float* op = a.out_f32 + ((size_t)b * a.n * a.n * a.n + vox) * a.out_cs + a.out_co;
#pragma unroll
            for (int i = 0; i < NP; ++i) if (i < a.n_real) op[i] = v[i];
          } else {
            // PM on the GEMM grid, or space-to-depth: voxel (z,y,x) channel c -> voxel (z/2,y/2,x/2) of the n/2 grid, channel
            // parity * C + c (the tensor then has 8x the planes of 1/8 the size)
            size_t ops = plane_elems;
            __nv_bfloat16* ob = a.out_pm + (size_t)b * a.out_planes * plane_elems + vox * 8;
            if (a.out_s2d) {
              const int h = a.n >> 1;
              ops = plane_elems >> 3;
              const int par = ((vz & 1) << 2) | ((vy & 1) << 1) | (vx & 1);
              ob = a.out_pm + ((size_t)b * a.out_planes * 8 + (size_t)par * a.out_planes) * ops + ((((size_t)(vz >> 1) * h + (vy >> 1)) * h + (vx >> 1)) * 8);
            }
#pragma unroll
            for (int c8 = 0; c8 < NP / 8; ++c8)
              if (c8 * 8 < a.n_real) split_store(ob + (size_t)(2 * c8) * ops, ob + (size_t)(2 * c8 + 1) * ops, v + c8 * 8);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_accfree + 8 * set) : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == WIN_EPI_WARPS) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
  }
}

// occupancy cube -> space-to-depth PM: cell (z,y,x) of the 32^3 grid holds the 8 voxels 2(z,y,x) + (pz,py,px), channel = pz*4+py*2+px
template <typename T>
__global__ void __launch_bounds__(256) cubes_to_s2d_pm_kernel(const T* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t total) {
  constexpr int N = 64, H = 32;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % H), y = (int)((i / H) % H), z = (int)((i / (H * H)) % H);
    const size_t b = i / ((size_t)H * H * H);
    const T* ib = in + b * (size_t)N * N * N;
    float v[8];
#pragma unroll
    for (int p = 0; p < 8; ++p)
      v[p] = (float)ib[((size_t)(2 * z + (p >> 2)) * N + (2 * y + ((p >> 1) & 1))) * N + 2 * x + (p & 1)];
    const size_t pe = (size_t)H * H * H * 8;
    __nv_bfloat16* ob = out + b * 2 * pe + (i % ((size_t)H * H * H)) * 8;
    split_store(ob, ob + pe, v);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn win_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

cudaError_t win_tmap(const PmTensor& t, int exc, int ey, int ez, int ppc, int zp, CUtensorMap* out) {
  EncodeTiledFn fn = win_encode_fn();
  if (!fn) return cudaErrorNotSupported;
  const cuuint64_t n = (cuuint64_t)t.n, planes = (cuuint64_t)(2 * t.c / 8);
  cuuint64_t gdim[5] = {n * 8, n, n, planes, (cuuint64_t)t.B};
  cuuint64_t gstride[4] = {n * 16, n * n * 16, n * n * n * 16, planes * n * n * n * 16};
  cuuint32_t box[5] = {(cuuint32_t)(exc * 8), (cuuint32_t)ey, (cuuint32_t)ez, (cuuint32_t)ppc, 1};
  if (zp == 2) {                       // dimension order (x, z, y, plane, cube): two z slices interleaved per y line
    gstride[0] = n * n * 16; gstride[1] = n * 16;
    box[1] = 2; box[2] = (cuuint32_t)ey;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)t.p, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <int NP, int EPI>
cudaError_t win_launch(const CUtensorMap& tm, const WinArgs& a, int grid, size_t smem, cudaStream_t s) {
  PCGC_CARVEOUT_ONCE((conv_umma_win_kernel<NP, EPI>));
  cudaError_t e = cudaFuncSetAttribute(conv_umma_win_kernel<NP, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  conv_umma_win_kernel<NP, EPI><<<grid, WIN_THREADS, smem, s>>>(tm, a);
  return cudaGetLastError();
}

}  // namespace

cudaError_t pack_win_layer(int cin, int wz, int wy, int wx, int oz, int oy, int ox, int n_cols,
                           const std::function<float(int, int, int, int, int)>& weight, const float* bias, WinLayer& out, int zp) {
  free_win_layer(out);
  if (zp != 1 && zp != 2) return cudaErrorInvalidValue;
  if (!(cin == 8 || (cin % 16 == 0 && cin <= 256)) || n_cols < 1 || n_cols > 64) return cudaErrorNotSupported;
  const int np = n_cols <= 16 ? 16 : (n_cols <= 32 ? 32 : 64);
  const bool paired = cin == 8;
  const int kch = paired ? 1 : cin / 16, exc = 8 + wx - 1, pitch = zp * exc;      // cells between consecutive y lines of the brick
  const size_t tile = (size_t)2 * np * 16;
  std::vector<__nv_bfloat16> packed;
  std::vector<int4> chunks;
  std::vector<uint32_t> entries;
  double macs = 0;
  for (int tz = 0; tz < wz; ++tz)
    for (int ty = 0; ty < wy; ++ty)
      for (int tx = 0; tx < wx; ++tx)
        for (int ci = 0; ci < cin; ++ci)
          for (int co = 0; co < n_cols; ++co) macs += weight(tz, ty, tx, ci, co) != 0.f;
  int max_entries = 0;
  for (int tz = 0; tz < wz; ++tz)
    for (int cc = 0; cc < kch; ++cc) {
      const int e0 = (int)entries.size();
      const int ntap = wy * wx;
      for (int m = 0; m < (paired ? (ntap + 1) / 2 : ntap); ++m) {
        const int t0 = paired ? 2 * m : m, t1 = paired ? (2 * m + 1 < ntap ? 2 * m + 1 : -1) : -1;
        std::vector<__nv_bfloat16> tl(tile, __float2bfloat16(0.f));
        bool nz = false;
        for (int k = 0; k < 16; ++k) {
          const int t = paired ? (k < 8 ? t0 : t1) : t0;
          if (t < 0) continue;
          const int ci = paired ? k % 8 : cc * 16 + k;
          for (int co = 0; co < n_cols; ++co) {
            const float w = weight(tz, t / wx, t % wx, ci, co);
            if (w == 0.f) continue;
            const __nv_bfloat16 hi = __float2bfloat16_rn(w);
            const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
            auto at = [&](int row) { return (size_t)(k / 8) * (2 * np * 8) + (size_t)(row / 8) * 64 + (row % 8) * 8 + (k % 8); };
            tl[at(co)] = hi;
            tl[at(np + co)] = lo;
            nz = true;
          }
        }
        if (!nz) continue;
        const int off0 = ((t0 / wx) * pitch + t0 % wx) * 16;
        uint32_t add = (uint32_t)(off0 >> 4);
        if (paired && t1 >= 0) add |= (uint32_t)(((((t1 / wx) * pitch + t1 % wx) * 16) - off0) >> 4) << 16;
        entries.push_back(add);
        packed.insert(packed.end(), tl.begin(), tl.end());
      }
      const int ne = (int)entries.size() - e0;
      if (ne > 0) {
        chunks.push_back(make_int4(oz + tz, cc * (paired ? 2 : 4), e0, ne));
        max_entries = std::max(max_entries, ne);
      }
    }
  if (chunks.empty()) return cudaErrorInvalidValue;
  std::vector<float> bz(np, 0.f);
  if (bias) for (int i = 0; i < n_cols; ++i) bz[i] = bias[i];
  cudaError_t e;
  if ((e = cudaMalloc(&out.packed, packed.size() * sizeof(__nv_bfloat16))) != cudaSuccess) return e;
  if ((e = cudaMemcpy(out.packed, packed.data(), packed.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&out.chunks, chunks.size() * sizeof(int4))) != cudaSuccess) return e;
  if ((e = cudaMemcpy(out.chunks, chunks.data(), chunks.size() * sizeof(int4), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&out.entries, entries.size() * sizeof(uint32_t))) != cudaSuccess) return e;
  if ((e = cudaMemcpy(out.entries, entries.data(), entries.size() * sizeof(uint32_t), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
  if ((e = cudaMalloc((void**)&out.bias, np * sizeof(float))) != cudaSuccess) return e;
  if ((e = cudaMemcpy(out.bias, bz.data(), np * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
  out.cin = cin; out.wz = wz; out.wy = wy; out.wx = wx; out.oz = oz; out.oy = oy; out.ox = ox;
  out.n_cols = n_cols; out.np = np; out.n_chunks = (int)chunks.size(); out.n_entries = (int)entries.size(); out.max_entries = max_entries;
  out.ppc = paired ? 2 : 4; out.macs_per_row = macs; out.zp = zp; out.ok = true;
  return cudaSuccess;
}

void free_win_layer(WinLayer& w) {
  if (w.packed) cudaFree(w.packed);
  if (w.chunks) cudaFree(w.chunks);
  if (w.entries) cudaFree(w.entries);
  if (w.bias) cudaFree(w.bias);
  w = WinLayer();
}

cudaError_t launch_conv_umma_win(const WinCall& c, const WinLayer& w, cudaStream_t s, int64_t* launches) {
  const int n = c.in.n;
  if (!w.ok || c.in.c != w.cin || (w.zp == 1 ? n % 16 != 0 : n % 8 != 0)) return cudaErrorNotSupported;
  WinArgs a;
  a.n = n; a.batch = c.in.B; a.zp = w.zp;
  a.exc = 8 + w.wx - 1; a.ey = (w.zp == 2 ? 8 : 16) + w.wy - 1; a.ox = w.ox; a.oy = w.oy;
  a.paired = w.cin == 8;
  a.n_chunks = w.n_chunks; a.n_entries = w.n_entries;
  a.tile_bytes = 2 * w.np * 32;
  const size_t fixed = 10 * 8 + (size_t)w.np * 4 + (size_t)w.n_chunks * 16 + (size_t)w.n_entries * 4 + 64;
  int zt = w.zp == 2 ? 1 : 8;                                 // zp = 2: one MMA row set = two slices; the brick holds exactly those
  const int zmul = w.zp == 2 ? 2 : 1;
  auto stage_bytes = [&](int z) { return (w.ppc * z * zmul * a.ey * a.exc * 16 + w.max_entries * a.tile_bytes + 127) / 128 * 128; };
  while (zt > 1 && (zt > n || zt * 2 * w.np > 256 || 2 * (size_t)stage_bytes(zt) + fixed > (size_t)200 * 1024)) zt /= 2;
  if (2 * (size_t)stage_bytes(zt) + fixed > (size_t)226 * 1024) return cudaErrorNotSupported;
  a.zt = zt;
  a.plane_bytes = zt * zmul * a.ey * a.exc * 16;
  a.a_bytes = w.ppc * a.plane_bytes;
  a.stage_bytes = stage_bytes(zt);
  a.wpacked = (const __nv_bfloat16*)w.packed; a.chunks = (const int4*)w.chunks; a.entries = (const uint32_t*)w.entries; a.bias = w.bias;
  a.n_real = w.n_cols; a.flags = c.flags;
  a.out_f32 = c.out_f32; a.out_cs = c.out_cs; a.out_co = c.out_co;
  a.out_pm = c.out.p; a.out_s2d = c.out_s2d;
  a.out_planes = 2 * (c.out_s2d ? c.out.c / 8 : c.out.c) / 8;
  a.up_ncls = w.up_ncls; a.up_cout = w.up_cout;
  for (int i = 0; i < 8; ++i) a.up_cls[i] = w.up_cls[i];
  a.err = c.err;
  auto pow2 = [](int v) { int c2 = 32; while (c2 < v) c2 *= 2; return c2; };
  a.set_cols = zt * 2 * w.np;
  a.nsets = 2 * a.set_cols <= 512 ? 2 : 1;
  a.tmem_cols = pow2(a.nsets * a.set_cols);
  if (c.epi == WEPI_PM && (w.n_cols % 8 != 0 || c.out.c != (c.out_s2d ? 8 : 1) * w.n_cols || c.out.n != (c.out_s2d ? n / 2 : n))) return cudaErrorInvalidValue;
  if (c.epi == WEPI_UP_PM && (w.up_ncls * w.up_cout != w.n_cols || w.up_cout % 16 != 0 || c.out.c != w.up_cout || c.out.n != 2 * n)) return cudaErrorInvalidValue;
  if (c.epi == WEPI_UP_F32 && (w.up_cout != 1 || w.up_ncls != w.n_cols || w.n_cols > 8 || !c.out_f32)) return cudaErrorInvalidValue;
  if (c.epi == WEPI_F32 && !c.out_f32) return cudaErrorInvalidValue;
  CUtensorMap tm;
  cudaError_t e = win_tmap(c.in, a.exc, a.ey, zt, w.ppc, w.zp, &tm);
  if (e != cudaSuccess) return e;
  const int sms = conv_sm_count();
  const int tiles = w.zp == 2 ? (n / 8) * (n / 8) * (n / 2) * c.in.B : (n / 8) * (n / 16) * (n / zt) * c.in.B;
  const int grid = std::min(tiles, sms);
  const size_t smem = 2 * (size_t)a.stage_bytes + fixed;
  if (launches) ++*launches;
  switch (w.np * 10 + c.epi) {
    case 160 + WEPI_F32: return win_launch<16, WEPI_F32>(tm, a, grid, smem, s);
    case 160 + WEPI_PM: return win_launch<16, WEPI_PM>(tm, a, grid, smem, s);
    case 160 + WEPI_UP_F32: return win_launch<16, WEPI_UP_F32>(tm, a, grid, smem, s);
    case 320 + WEPI_F32: return win_launch<32, WEPI_F32>(tm, a, grid, smem, s);
    case 320 + WEPI_PM: return win_launch<32, WEPI_PM>(tm, a, grid, smem, s);
    case 640 + WEPI_UP_PM: return win_launch<64, WEPI_UP_PM>(tm, a, grid, smem, s);
  }
  return cudaErrorNotSupported;
}

cudaError_t launch_cubes_to_s2d_pm(const void* cubes, int dtype, const PmTensor& out, cudaStream_t s, int64_t* launches) {
  if (out.n != 32 || out.c != 8) return cudaErrorInvalidValue;
  const size_t total = (size_t)out.B * 32 * 32 * 32;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  if (launches) ++*launches;
  if (dtype == PCGC_DTYPE_U8) { PCGC_CARVEOUT_ONCE(cubes_to_s2d_pm_kernel<uint8_t>); cubes_to_s2d_pm_kernel<uint8_t><<<blocks, 256, 0, s>>>((const uint8_t*)cubes, out.p, total); }
  else if (dtype == PCGC_DTYPE_F32) { PCGC_CARVEOUT_ONCE(cubes_to_s2d_pm_kernel<float>); cubes_to_s2d_pm_kernel<float><<<blocks, 256, 0, s>>>((const float*)cubes, out.p, total); }
  else { PCGC_CARVEOUT_ONCE(cubes_to_s2d_pm_kernel<double>); cubes_to_s2d_pm_kernel<double><<<blocks, 256, 0, s>>>((const double*)cubes, out.p, total); }
  return cudaGetLastError();
}

}  // namespace pcgc
