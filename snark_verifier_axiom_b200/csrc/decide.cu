// K5/K6 kernels: KzgAs::decide (snark-verifier/src/pcs/kzg/decider.rs:60-81), two schedules (svk_ctx.h decide_coop_max):
//   k_decide       one accumulator per thread (throughput form: batches of thousands); the G2 line tables and Frobenius
//                  constants are read uniformly by all threads (broadcast loads, L1/L2 resident: 2 x 102 x 128 B)
//   k_decide_coop  one accumulator per BLOCK of 128 threads (latency form: the single pairing of a folded batch), the
//                  block-cooperative program of coop_pairing.cuh
#include "svk_ctx.h"

__device__ __forceinline__ G1Affine load_g1_canon(const uint8_t* p) {
  G1Affine a;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3];
  a.x.v[0] = w0.x; a.x.v[1] = w0.y; a.x.v[2] = w0.z; a.x.v[3] = w0.w;
  a.x.v[4] = w1.x; a.x.v[5] = w1.y; a.x.v[6] = w1.z; a.x.v[7] = w1.w;
  a.y.v[0] = w2.x; a.y.v[1] = w2.y; a.y.v[2] = w2.z; a.y.v[3] = w2.w;
  a.y.v[4] = w3.x; a.y.v[5] = w3.y; a.y.v[6] = w3.z; a.y.v[7] = w3.w;
  return a;
}

// status: bit0 = accept.  A point that is not a canonical on-curve encoding makes the reference's
// `G1Affine` unconstructible; such accumulators are reported as reject.
#ifndef SVK_DECIDE_MINBLOCKS
#define SVK_DECIDE_MINBLOCKS 1
#endif
__global__ void __launch_bounds__(64, SVK_DECIDE_MINBLOCKS) k_decide(size_t n, const uint8_t* accs, size_t acc_stride, uint8_t* out_ok, size_t ok_stride,
                                               const G2Line* t_g2, const G2Line* t_neg_sg2, const PairingConsts* consts) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine lhs = load_g1_canon(accs + i * acc_stride);
  G1Affine rhs = load_g1_canon(accs + i * acc_stride + 64);
  bool ok = Fq::is_canonical(lhs.x.v) && Fq::is_canonical(lhs.y.v) && Fq::is_canonical(rhs.x.v) && Fq::is_canonical(rhs.y.v);
  if (!lhs.is_identity()) { lhs.x = lhs.x.to_mont(); lhs.y = lhs.y.to_mont(); }
  if (!rhs.is_identity()) { rhs.x = rhs.x.to_mont(); rhs.y = rhs.y.to_mont(); }
  ok = ok && g1_on_curve(lhs) && g1_on_curve(rhs);
  bool acc = false;
  if (ok) acc = kzg_decide(lhs, rhs, t_g2, t_neg_sg2, *consts);
  out_ok[i * ok_stride] = acc ? 1 : 0;
}

// Latency form (coop_pairing.cuh): one accumulator per block of COOP_THREADS threads.  Used when there are too few
// accumulators to fill the machine with one thread each (the single pairing that ends a batch, the per-rank or per-batch
// pairings of a sharded / multi-batch call).
#define COOP_THREADS 128
struct DevExec {
  static constexpr bool kWarp12 = true;  // a 12-task phase runs on lanes 0..11 of warp 0
  template <class F>
  __device__ __forceinline__ void par(int n_tasks, F f) {
    for (int i = threadIdx.x; i < n_tasks; i += COOP_THREADS) f(i);
    __syncthreads();
  }
};

__global__ void __launch_bounds__(COOP_THREADS) k_decide_coop(size_t n, const uint8_t* accs, size_t acc_stride, uint8_t* out_ok, size_t ok_stride,
                                                              const G2LineX* t_g2, const G2LineX* t_neg_sg2, const PairingConsts* consts, Fq* line_scratch) {
  __shared__ CoopMem m;
  __shared__ G1Affine pts[2];
  __shared__ int pt_ok[2];
  size_t i = blockIdx.x;
  if (threadIdx.x < 2) {
    G1Affine p = load_g1_canon(accs + i * acc_stride + 64 * threadIdx.x);
    bool ok = Fq::is_canonical(p.x.v) && Fq::is_canonical(p.y.v);
    if (!p.is_identity()) { p.x = p.x.to_mont(); p.y = p.y.to_mont(); }
    ok = ok && g1_on_curve(p);
    pts[threadIdx.x] = p;
    pt_ok[threadIdx.x] = ok ? 1 : 0;
  }
  __syncthreads();
  if (!(pt_ok[0] && pt_ok[1])) {  // not a `G1Affine` at all: reject (uniform across the block)
    if (threadIdx.x == 0) out_ok[i * ok_stride] = 0;
    return;
  }
  DevExec ex;
  bool acc = coop_kzg_decide(ex, pts[0], pts[1], t_g2, t_neg_sg2, *consts, line_scratch + i * (size_t)COOP_SCRATCH_FQ, m);
  if (threadIdx.x == 0) out_ok[i * ok_stride] = acc ? 1 : 0;
}

int svk_decide_launch_strided(svk_ctx* ctx, int dk, size_t n, const void* d_accs, size_t acc_stride, void* d_ok, size_t ok_stride);
int svk_decide_launch(svk_ctx* ctx, int dk, size_t n, const void* d_accs, void* d_ok) {
  return svk_decide_launch_strided(ctx, dk, n, d_accs, 128, d_ok, 1);
}

int svk_decide_launch_strided(svk_ctx* ctx, int dk, size_t n, const void* d_accs, size_t acc_stride, void* d_ok, size_t ok_stride) {
  if (dk < 0 || dk >= (int)ctx->dks.size()) return svk_fail(ctx, "bad deciding-key id %d", dk);
  if (n == 0) return 0;
  const DkDevice& k = ctx->dks[dk];
  if (n <= ctx->decide_coop_max) {
    Fq* d_lines;
    if (svk_scratch(ctx, 19, n * (size_t)COOP_SCRATCH_FQ * sizeof(Fq), (void**)&d_lines)) return -1;
    SVK_LAUNCH(ctx, "k_decide_coop",
               k_decide_coop<<<(unsigned)n, COOP_THREADS, 0, ctx->stream>>>(n, (const uint8_t*)d_accs, acc_stride, (uint8_t*)d_ok, ok_stride, k.d_linesx_g2,
                                                                           k.d_linesx_neg_sg2, ctx->d_pairing_consts, d_lines));
    SVK_CUDA(ctx, cudaGetLastError());
    return 0;
  }
  unsigned block = 64;
  unsigned grid = (unsigned)((n + block - 1) / block);
  SVK_LAUNCH(ctx, "k_decide",
             k_decide<<<grid, block, 0, ctx->stream>>>(n, (const uint8_t*)d_accs, acc_stride, (uint8_t*)d_ok, ok_stride, k.d_lines_g2, k.d_lines_neg_sg2,
                                                       ctx->d_pairing_consts));
  SVK_CUDA(ctx, cudaGetLastError());
  return 0;
}

// ---- integer-multiply roofline micro-benchmark ------------------------------------------------
__global__ void __launch_bounds__(256) k_modmul_peak(u32* out, int iters) {
  // 2 independent dependent-chains per thread (ILP 2), 8 warps per SM sub-partition
  Fq a, b, c, d;
  u32 t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; i++) { a.v[i] = t * 2654435761u + i; b.v[i] = t ^ (0x9e3779b9u * (i + 1)); c.v[i] = t + 77 * i; d.v[i] = ~t - i; }
  a.v[7] &= 0x0fffffffu; b.v[7] &= 0x0fffffffu; c.v[7] &= 0x0fffffffu; d.v[7] &= 0x0fffffffu;
  for (int k = 0; k < iters; k++) {
    a = Fq::mul_inline(a, b);
    c = Fq::mul_inline(c, d);
    b = Fq::mul_inline(b, a);
    d = Fq::mul_inline(d, c);
  }
  Fq r = a + b + c + d;
  u32 x = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) x ^= r.v[i];
  if (x == 0x12345678u) out[0] = x;  // keep the chain alive
}

int svk_modmul_peak_launch(svk_ctx* ctx, int iters, double* out_rate, double* out_ms) {
  void* d_out;
  if (svk_scratch(ctx, 7, 256, &d_out)) return -1;
  int blocks = ctx->sm_count * 8, threads = 256;
  cudaEvent_t e0, e1;
  SVK_CUDA(ctx, cudaEventCreate(&e0));
  SVK_CUDA(ctx, cudaEventCreate(&e1));
  k_modmul_peak<<<blocks, threads, 0, ctx->stream>>>((u32*)d_out, 16);  // warm-up
  SVK_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
  k_modmul_peak<<<blocks, threads, 0, ctx->stream>>>((u32*)d_out, iters);
  SVK_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
  ctx->launches += 2;
  SVK_CUDA(ctx, cudaEventSynchronize(e1));
  float ms = 0;
  SVK_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  double muls = (double)blocks * threads * 4.0 * iters;
  *out_rate = muls / (ms * 1e-3);
  *out_ms = ms;
  return 0;
}
