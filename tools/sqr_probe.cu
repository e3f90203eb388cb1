// Micro-probe (not part of libsvk): Montgomery squaring vs multiplication rate, inlined and out-of-line, at
// several occupancies.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17
//   --expt-relaxed-constexpr -I snark_verifier_axiom_b200/csrc -o tools/sqr_probe.bin tools/sqr_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "field.cuh"

template <int MODE>
__global__ void k(u32* out, int iters) {
  Fq a, b;
  u32 t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; i++) { a.v[i] = t * 2654435761u + i; b.v[i] = (t ^ (0x9e3779b9u * (i + 1))); if (i == 7) { a.v[i] &= 0x0fffffff; b.v[i] &= 0x0fffffff; } }
  for (int k2 = 0; k2 < iters; k2++) {
    if (MODE == 0) a = Fq::mul_inline(a, b);
    if (MODE == 1) a = Fq::sqr_inline(a);
    if (MODE == 2) a = a * b;
    if (MODE == 3) a = a.sqr();
    if (MODE == 4) a = Fq::mul_inline(a, a);
  }
  u32 x = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) x ^= a.v[i];
  if (x == 0x12345678u) out[0] = x;
}

template <int MODE>
void run(const char* name, int blocks, int threads, int iters, u32* d) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(d, 8);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(d, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)blocks * threads * iters;
  printf("%-12s warps/SMSP=%5.2f  %8.3f ms  %7.2f Gop/s  ns/op/thread=%.1f\n", name, (double)blocks * threads / 32 / (148 * 4), ms, ops / ms / 1e6, ms * 1e6 / iters);
}

int main() {
  u32* d; cudaMalloc(&d, 256);
  int iters = 4000;
  int cfg[][2] = {{148, 128}, {148 * 2, 128}, {148 * 2, 256}, {148 * 4, 256}, {148 * 8, 256}};
  for (auto& c : cfg) {
    run<0>("mul_inline", c[0], c[1], iters, d);
    run<4>("mul(a,a)", c[0], c[1], iters, d);
    run<1>("sqr_inline", c[0], c[1], iters, d);
    run<2>("mul_call", c[0], c[1], iters, d);
    run<3>("sqr_call", c[0], c[1], iters, d);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
