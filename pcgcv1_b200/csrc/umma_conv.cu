// tcgen05 (UMMA) implicit-GEMM convolution engine -- placeholder until the kernel lands.
#include "umma_conv.cuh"

namespace pcgc {

cudaError_t pack_umma_weights(const float*, int, int, UmmaWeights& out) { out.ok = false; return cudaErrorNotSupported; }
void free_umma_weights(UmmaWeights& w) { if (w.packed) cudaFree(w.packed); w.packed = nullptr; w.ok = false; }
cudaError_t launch_conv_umma(const ConvCall&, const UmmaWeights&, cudaStream_t, int64_t*) { return cudaErrorNotSupported; }

}  // namespace pcgc
