// IPA accumulation decider on the Pasta curves.
//
// Replaces `<IpaAs<C, MOS> as AccumulationDecider<C, NativeLoader>>::{decide, decide_all}`
// (snark-verifier/src/pcs/ipa/decider.rs:33-67):
//     let h = h_coeffs(&xi, C::Scalar::one());                          pcs/ipa.rs:379-395
//     (u == multi_scalar_multiplication(&h, &dk.g).to_affine())         util/msm.rs:238-317
// `IpaDecidingKey { svk, g }` (decider.rs:5-16) crosses the ABI as the 2^k committing-key points; an
// `IpaAccumulator { xi, u }` (pcs/ipa/accumulator.rs:5-25) as k scalars + one affine point.
//
//   k_h_coeffs     one coefficient per thread: h[j] = prod_{i : bit i of j set} xi[k-1-i]  -- the closed form of the
//                  reference's doubling loop (block i of length 2^i is the copy of blocks < i times xi.rev()[i])
//   msm (msm.cu)   the K4 Pippenger pipeline instantiated over the curve (CurvePallas / CurveVesta / CurveBn254)
//   k_ipa_verdict  U == commit(G, h)  ->  status 0 / SVK_ASSERTION_FAILURE
#include "pasta.cuh"
#include "svk_ctx.h"

int svk_msm_curve_launch(svk_ctx* ctx, int curve, size_t n, const uint8_t* d_scalars, const uint8_t* d_points, uint8_t* d_out, int* d_status);

template <class Fs>
__global__ void __launch_bounds__(128) k_h_coeffs(u32 k, const uint8_t* xi, uint8_t* h, int* bad) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ((size_t)1 << k)) return;
  Fs acc = Fs::one();
  for (u32 i = 0; i < k; i++) {
    if (!((j >> i) & 1)) continue;
    Fs x;
    const uint4* q = reinterpret_cast<const uint4*>(xi + (size_t)(k - 1 - i) * 32);
    uint4 lo = q[0], hi = q[1];
    x.v[0] = lo.x; x.v[1] = lo.y; x.v[2] = lo.z; x.v[3] = lo.w; x.v[4] = hi.x; x.v[5] = hi.y; x.v[6] = hi.z; x.v[7] = hi.w;
    if (!Fs::is_canonical(x.v)) { *bad = 2; x = Fs::zero(); }
    acc = acc * x.to_mont();
  }
  acc = acc.from_mont();
  uint4* o = reinterpret_cast<uint4*>(h + j * 32);
  o[0] = make_uint4(acc.v[0], acc.v[1], acc.v[2], acc.v[3]);
  o[1] = make_uint4(acc.v[4], acc.v[5], acc.v[6], acc.v[7]);
}

// status: 0 ok; 3 = Error::AssertionFailure("U == commit(G, h)") (decider.rs:52-54).  Inputs no halo2curves value can hold
// (xi >= the scalar modulus, a committing-key point off the curve) also fail the assertion; `*invalid` tells them apart.
__global__ void k_ipa_verdict(const uint8_t* commit, const uint8_t* u, const int* msm_status, const int* h_bad, int32_t* out_status, int32_t* invalid) {
  u32 diff = 0;
  for (int i = 0; i < 16; i++) diff |= reinterpret_cast<const u32*>(commit)[i] ^ reinterpret_cast<const u32*>(u)[i];
  int inv = (*msm_status != 0) || (*h_bad != 0);
  *out_status = (diff == 0 && !inv) ? 0 : 3;
  if (inv) *invalid = 1;
}

// d_g: 2^k x 64 B affine points; d_xi: n x k x 32 B; d_u: n x 64 B; d_out_status: n x int32; d_invalid: one int32 (sticky)
int svk_ipa_decide_launch(svk_ctx* ctx, int curve, u32 k, const uint8_t* d_g, size_t n, const uint8_t* d_xi, const uint8_t* d_u,
                          int32_t* d_out_status, int32_t* d_invalid) {
  if (k == 0 || k > 26) return svk_fail(ctx, "ipa_decide: k out of range (the reference asserts !xi.is_empty())");
  if (curve < 0 || curve > 2) return svk_fail(ctx, "ipa_decide: unknown curve id");
  size_t m = (size_t)1 << k;
  uint8_t* d_h;
  uint8_t* d_tmp;
  if (svk_scratch(ctx, 19, m * 32, (void**)&d_h)) return -1;
  if (svk_scratch(ctx, 20, 256, (void**)&d_tmp)) return -1;
  cudaStream_t s = ctx->stream;
  int* d_msm_status = (int*)(d_tmp + 64);
  int* d_h_bad = (int*)(d_tmp + 128);
  for (size_t i = 0; i < n; i++) {
    SVK_CUDA(ctx, cudaMemsetAsync(d_h_bad, 0, 4, s));
    unsigned gb = (unsigned)((m + 127) / 128);
    const uint8_t* xi = d_xi + i * (size_t)k * 32;
    switch (curve) {
      case 0: SVK_LAUNCH(ctx, "k_h_coeffs", k_h_coeffs<CurveBn254::Scalar><<<gb, 128, 0, s>>>(k, xi, d_h, d_h_bad)); break;
      case 1: SVK_LAUNCH(ctx, "k_h_coeffs", k_h_coeffs<CurvePallas::Scalar><<<gb, 128, 0, s>>>(k, xi, d_h, d_h_bad)); break;
      default: SVK_LAUNCH(ctx, "k_h_coeffs", k_h_coeffs<CurveVesta::Scalar><<<gb, 128, 0, s>>>(k, xi, d_h, d_h_bad)); break;
    }
    if (svk_msm_curve_launch(ctx, curve, m, d_h, d_g, d_tmp, d_msm_status)) return -1;
    SVK_LAUNCH(ctx, "k_ipa_verdict", k_ipa_verdict<<<1, 1, 0, s>>>(d_tmp, d_u + i * 64, d_msm_status, d_h_bad, d_out_status + i, d_invalid));
  }
  SVK_CUDA(ctx, cudaGetLastError());
  return 0;
}
