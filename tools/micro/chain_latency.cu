// Dev tool: dependent-issue latency (cycles) of the instructions on the range coder's serial chain, one warp on one SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 4096
template <int OP>
__global__ void k(uint32_t* out, uint32_t seed, uint32_t m, long long* cyc) {
  uint32_t x = seed + threadIdx.x, y = m;
  const unsigned F = 0xffffffffu;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = x * y + y;                                             // IMAD
    else if (OP == 1) x = __umulhi(x, y) + y;                               // IMAD.HI
    else if (OP == 2) x = __reduce_max_sync(F, x) + 1;                      // REDUX.MAX (+IADD)
    else if (OP == 3) x = __reduce_or_sync(F, x) + 1;                       // REDUX.OR (+IADD)
    else if (OP == 4) x = __shfl_sync(F, x, x & 31) + 1;                    // SHFL with a data-dependent lane (+IADD)
    else if (OP == 5) x = __ballot_sync(F, x & 1) + 1 + x;                  // VOTE (+IADD)
    else if (OP == 6) x = (x < 0x10000u ? (x << 16 | 0xffffu) : x) + y;     // ISETP + SEL (+LEA) + IADD
    else if (OP == 7) x = (x ^ y) + (x >> 3);                               // LOP3 + SHF + IADD
    else if (OP == 8) { uint32_t b = __ballot_sync(F, ((x >> (threadIdx.x & 7)) & 1) != 0); x = __shfl_sync(F, x, (__ffs(b | 0x80000000u) - 1) & 31) + 1; }   // VOTE + FLO + SHFL
    else if (OP == 9) x = __reduce_add_sync(F, x) + 1;                      // REDUX.ADD
    else if (OP == 10) x = min(x, y) + 1;                                   // IMNMX
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  uint32_t* o; long long* c; cudaMalloc(&o, 128); cudaMalloc(&c, 8);
  const char* names[] = {"IMAD", "IMAD.HI+add", "REDUX.MAX+IADD", "REDUX.OR+IADD", "SHFL+IADD", "VOTE+IADD", "ISETP+SEL+IADD", "LOP3+SHF+IADD", "VOTE+FLO+SHFL+IADD", "REDUX.ADD+IADD", "IMNMX+IADD"};
#define RUN(OP) { k<OP><<<1, 32>>>(o, 1, 3, c); k<OP><<<1, 32>>>(o, 1, 3, c); long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-22s %.1f cycles per iteration\n", names[OP], (double)h / N); }
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10)
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
