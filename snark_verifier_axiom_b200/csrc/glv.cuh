// GLV scalar decomposition for BN254 G1 (j-invariant 0: phi(x, y) = (beta x, y) = lambda (x, y) on the r-torsion).
//
// Only group elements are observable at the boundary (`NativeLoader::multi_scalar_multiplication`,
// snark-verifier/src/loader/native.rs:61-71; util/msm.rs:238-317), so k P may be evaluated as k1 P + k2 phi(P) with
// k = k1 + k2 lambda (mod r), |k1|, |k2| < 2^128: half the doublings of every scalar multiplication.
// Constants and the rounding scheme are derived and checked by tools/glv_constants.py; the split is CORRECT for any
// (c1, c2) because (a1, b1), (a2, b2) lie in the lattice { a + b lambda = 0 mod r } -- rounding only decides the size, which
// `glv_decompose` reports (tests/test_host_arith.py::test_glv_decompose checks k1 + k2 lambda = k and the bound).
#pragma once
#include "field.cuh"

// lambda = 4407920970296243842393367215006156084916469457145843978461, beta = 2203960485148121921418603742825762020974279258880205651966
HD Fq glv_beta_mont() {
  const u32 b[8] = {0xd782e155u, 0x71930c11u, 0xffbe3323u, 0xa6bb947cu, 0xd4741444u, 0xaa303344u, 0x26594943u, 0x2c3b3f0du};
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = b[i];
  return r;
}

// out[0..no) = low `no` limbs of a[0..na) * b[0..nb) shifted right by `skip` limbs
template <int NA, int NB, int SKIP, int NO>
HD void glv_mul(const u32* a, const u32* b, u32* out) {
  u32 t[NA + NB];
#pragma unroll
  for (int i = 0; i < NA + NB; i++) t[i] = 0;
#pragma unroll
  for (int i = 0; i < NA; i++) {
    u64 c = 0;
#pragma unroll
    for (int j = 0; j < NB; j++) {
      u64 v = (u64)a[i] * b[j] + t[i + j] + c;
      t[i + j] = (u32)v;
      c = v >> 32;
    }
    t[i + NB] = (u32)c;
  }
#pragma unroll
  for (int i = 0; i < NO; i++) out[i] = (SKIP + i < NA + NB) ? t[SKIP + i] : 0;
}

// r (5 limbs, two's complement mod 2^160) = x - y
HD void glv_sub5(u32* r, const u32* x, const u32* y) {
  u64 br = 0;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    u64 v = (u64)x[i] - y[i] - br;
    r[i] = (u32)v;
    br = (v >> 32) & 1;
  }
}

// |v| and the sign of a 160-bit two's-complement value; returns false when |v| >= 2^128
HD bool glv_abs5(u32* v, u32& neg) {
  neg = v[4] >> 31;
  if (neg) {
    u64 c = 1;
#pragma unroll
    for (int i = 0; i < 5; i++) {
      u64 t = (u64)(~v[i]) + c;
      v[i] = (u32)t;
      c = t >> 32;
    }
  }
  return v[4] == 0;
}

// k (8 limbs, canonical, < r)  ->  |k1|, |k2| (4 limbs each, < 2^128) and their signs; k = (+-k1) + (+-k2) lambda (mod r)
HD bool glv_decompose(const u32* k, u32* k1, u32& neg1, u32* k2, u32& neg2) {
  const u32 G1[3] = {0xc7e0b3d7u, 0xd91d232eu, 0x00000002u};                                   // round(2^256 b2 / r)
  const u32 G2[5] = {0x391eb18eu, 0x7a7bd9d4u, 0xa773d2cfu, 0x4ccef014u, 0x00000002u};           // round(2^256 |b1| / r)
  const u32 A1[2] = {0x94d213e3u, 0x89d32568u};                                                // a1 = b2 = 9931322734385697763
  const u32 NB1[4] = {0x7d4f1128u, 0x8211bbebu, 0xeeb859fcu, 0x6f4d8248u};                       // -b1
  const u32 A2[4] = {0x1221250bu, 0x0be4e154u, 0xeeb859fdu, 0x6f4d8248u};                        // a2
  u32 c1[3], c2[5];
  glv_mul<8, 3, 8, 3>(k, G1, c1);   // c1 = (k g1) >> 256  < 2^66
  glv_mul<8, 5, 8, 5>(k, G2, c2);   // c2 = (k g2) >> 256  < 2^130
  // k1 = k - c1 a1 - c2 a2 ;  k2 = c1 (-b1) - c2 b2      (mod 2^160; both are small)
  u32 t1[5], t2[5], acc[5];
  glv_mul<3, 2, 0, 5>(c1, A1, t1);
  glv_mul<5, 4, 0, 5>(c2, A2, t2);
  glv_sub5(acc, k, t1);
  glv_sub5(acc, acc, t2);
  u32 v1[5], v2[5];
#pragma unroll
  for (int i = 0; i < 5; i++) v1[i] = acc[i];
  glv_mul<3, 4, 0, 5>(c1, NB1, t1);
  glv_mul<5, 2, 0, 5>(c2, A1, t2);  // b2 = a1
  glv_sub5(v2, t1, t2);
  bool ok = glv_abs5(v1, neg1);
  ok = glv_abs5(v2, neg2) && ok;
#pragma unroll
  for (int i = 0; i < 4; i++) { k1[i] = v1[i]; k2[i] = v2[i]; }
  return ok;
}
