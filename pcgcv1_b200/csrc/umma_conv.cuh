// tcgen05 (UMMA) implicit-GEMM convolution engine: interface.  See umma_conv.cu.
#pragma once
#include "common.cuh"

namespace pcgc {

struct UmmaWeights {
  bool ok = false;          // layer qualifies and the packed weights are resident
  int cin = 0, cout = 0;
  int cin_pad = 0;          // Cin padded to a multiple of 8 (one 16-byte bf16 cell)
  int n_pad = 0;            // MMA N (couts incl. hi/lo columns, padded to 16)
  void* packed = nullptr;   // device: bf16 B-operand tiles, see umma_conv.cu
};

// kernel: HOST float32 [3,3,3,Cin,Cout] (Keras layout).  cudaErrorNotSupported if the shape does not qualify.
cudaError_t pack_umma_weights(const float* kernel, int cin, int cout, UmmaWeights& out);
void free_umma_weights(UmmaWeights& w);
// 3x3x3 stride-1 SAME conv.  cudaErrorNotSupported if this call cannot be taken (caller falls back).
cudaError_t launch_conv_umma(const ConvCall& c, const UmmaWeights& w, cudaStream_t s, int64_t* launches);

}  // namespace pcgc
