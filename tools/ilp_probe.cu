// Micro-probe (not part of libsvk): Montgomery-mul rate vs ILP (independent chains per thread) and
// warps per SM sub-partition.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17
//   --expt-relaxed-constexpr -I snark_verifier_axiom_b200/csrc -o gpurun_out/ilp_probe tools/ilp_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "field.cuh"

template <int ILP>
__global__ void k(u32* out, int iters) {
  Fq a[ILP], b[ILP];
  u32 t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int j = 0; j < ILP; j++)
#pragma unroll
    for (int i = 0; i < 8; i++) { a[j].v[i] = t * 2654435761u + i + j; b[j].v[i] = (t ^ (0x9e3779b9u * (i + 1))) + j; if (i == 7) { a[j].v[i] &= 0x0fffffff; b[j].v[i] &= 0x0fffffff; } }
  for (int k2 = 0; k2 < iters; k2++) {
#pragma unroll
    for (int j = 0; j < ILP; j++) a[j] = Fq::mul_inline(a[j], b[j]);
  }
  u32 x = 0;
#pragma unroll
  for (int j = 0; j < ILP; j++)
#pragma unroll
    for (int i = 0; i < 8; i++) x ^= a[j].v[i];
  if (x == 0x12345678u) out[0] = x;
}

template <int ILP>
void run(int blocks, int threads, int iters, u32* d) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP><<<blocks, threads>>>(d, 8);
  cudaEventRecord(e0);
  k<ILP><<<blocks, threads>>>(d, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double muls = (double)blocks * threads * ILP * iters;
  int sm = 148;
  double warps_per_smsp = (double)blocks * threads / 32 / (sm * 4);
  printf("ILP=%d blocks=%d thr=%d warps/SMSP=%.2f  %.3f ms  %.2f Gmodmul/s  ns/mul/thread=%.1f\n", ILP, blocks, threads, warps_per_smsp, ms, muls / ms / 1e6, ms * 1e6 / (ILP * iters));
}

int main() {
  u32* d; cudaMalloc(&d, 256);
  int iters = 2000;
  int cfg[][2] = {{148, 32}, {148, 128}, {148, 256}, {148 * 2, 256}, {148 * 4, 256}, {148 * 8, 256}};
  for (auto& c : cfg) {
    run<1>(c[0], c[1], iters, d);
    run<2>(c[0], c[1], iters, d);
    run<3>(c[0], c[1], iters, d);
    run<4>(c[0], c[1], iters, d);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
