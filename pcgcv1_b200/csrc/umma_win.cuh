// tcgen05 window-GEMM kernel for the large-kernel stride-2 layers of model_simple (models/model_simple.py:21-42,58-86):
// interface.  See umma_win.cu for the design.
#pragma once
#include <functional>

#include "umma_conv.cuh"

namespace pcgc {

enum WinEpilogue : int {
  WEPI_F32 = 0,      // +bias, [relu] -> float32 NDHWC
  WEPI_PM = 1,       // +bias, [relu] -> PM, on the GEMM grid or space-to-depth (for a following stride-2 layer)
  WEPI_UP_PM = 2,    // stride-2 transposed conv: column block `cls` is output voxel 2t + r(cls), +bias, [relu] -> PM on the 2n grid
  WEPI_UP_F32 = 3    // the same with ONE output channel per class -> float32 [2n]^3 (the logits of model_simple)
};

// One layer in its window-GEMM form: rows = voxels t of an n^3 grid, K = (window cell, input channel), columns = n_cols.
struct WinLayer {
  bool ok = false;
  int cin = 0;                 // channels of the input PM tensor as the kernel reads it (8 = one cell, paired taps; else a multiple of 16)
  int wz = 0, wy = 0, wx = 0;  // window extents in cells (3 or 5)
  int oz = 0, oy = 0, ox = 0;  // cell offset of the first window cell relative to t (e.g. -2)
  int n_cols = 0, np = 0;      // real / padded (multiple of 16) output columns
  int n_chunks = 0, n_entries = 0, max_entries = 0;
  int ppc = 0;                 // planes per chunk: 2 (cin 8) or 4
  int zp = 1;                  // 2: the 8 x 8 x 2-voxel tile form for 8^3 grids (umma_win.cu)
  void* packed = nullptr;      // device bf16: [entry][2*np x 16] canonical no-swizzle K-major tiles (hi rows, then lo rows)
  void* chunks = nullptr;      // device int4 {dz, plane0, first entry, entries}
  void* entries = nullptr;     // device uint32: A-descriptor increment (start offset >> 4 | LBO >> 4 << 16)
  float* bias = nullptr;       // device [np]
  int up_cout = 0;             // WEPI_UP_*: channels per class (columns = classes x up_cout)
  int up_cls[8] = {0};         // WEPI_UP_*: class id (rz<<2 | ry<<1 | rx) of every column block
  int up_ncls = 0;
  double macs_per_row = 0;     // algorithmic MACs per GEMM row (reference taps only: zero padding does not count)
};

// weight(tz, ty, tx, ci, col): value of window cell (tz, ty, tx), input channel ci, output column col (0 where the reference
// layer has no tap).  bias: host [n_cols] or null.
cudaError_t pack_win_layer(int cin, int wz, int wy, int wx, int oz, int oy, int ox, int n_cols,
                           const std::function<float(int, int, int, int, int)>& weight, const float* bias, WinLayer& out, int zp = 1);
void free_win_layer(WinLayer& w);

struct WinCall {
  PmTensor in;
  int epi = WEPI_F32;
  int flags = 0;
  float* out_f32 = nullptr; int out_cs = 0, out_co = 0;
  PmTensor out; int out_s2d = 0;
  int* err = nullptr;
};
cudaError_t launch_conv_umma_win(const WinCall& c, const WinLayer& w, cudaStream_t s, int64_t* launches);

// occupancy cube (uint8 / float32 / float64, [B,64,64,64,1]) -> space-to-depth PM tensor [B][2 planes][32^3][8 parities]
cudaError_t launch_cubes_to_s2d_pm(const void* cubes, int dtype, const PmTensor& out, cudaStream_t s, int64_t* launches);

}  // namespace pcgc
