// K5/K6: optimal-ate Miller loop with precomputed G2 line coefficients + final exponentiation.
//
// Replaces `M::multi_miller_loop(&[(lhs, g2), (rhs, -s_g2)]).final_exponentiation().is_identity()`
// (snark-verifier/src/pcs/kzg/decider.rs:60-68; arithmetic in halo2curves 0.3.1 `bn256::Bn256`).
// The two G2 points are fixed per `KzgDecidingKey` (decider.rs:6-13), so every line's slope is
// computed ONCE on the host (`G2LineTable`); the device only evaluates lines at the G1 points.
// Only the accept bit is observable, so lines are used unnormalised.
//
// Untwist (x', y') -> (x' w^2, y' w^3).  Line through T (slope lam) at P = (xP, yP):
//     l = yP + (-lam * xP) w + (lam * xT - yT) w^3            (w^3 = v w)
// i.e. Fq12 coefficients  c0 = (yP, 0, 0),  c1 = (-lam*xP, lam*xT - yT, 0).
#pragma once
#include "g1.cuh"
#include "tower.cuh"

#define SVK_ATE_LOOP_HI 0x1u          // 6x+2 = 0x1_9d797039_be763ba8 (65 bits)
#define SVK_ATE_LOOP_MID 0x9d797039u
#define SVK_ATE_LOOP_LO 0xbe763ba8u
#define SVK_BN_X_HI 0x44e992b4u       // x = 4965661367192848881 = 0x44e992b4_4a6909f1
#define SVK_BN_X_LO 0x4a6909f1u
#define SVK_N_LINES (64 + 36 + 2)     // doubling steps + addition steps (popcount(6x+2) - 1) + 2 Frobenius steps

struct G2Line {
  Fq2 neg_lam;  // -lambda
  Fq2 c3;       // lambda * xT - yT
};

struct G2Affine {
  Fq2 x, y;
};

struct PairingConsts {
  Fq2 g1[5];  // xi^(i (p-1)/6),   i = 1..5
  Fq g2[5];   // xi^(i (p^2-1)/6)  (in Fq)
  Fq2 g3[5];  // xi^(i (p^3-1)/6)
};

HD int ate_bit(int i) {  // bit i of 6x+2
  if (i >= 64) return (SVK_ATE_LOOP_HI >> (i - 64)) & 1;
  if (i >= 32) return (SVK_ATE_LOOP_MID >> (i - 32)) & 1;
  return (SVK_ATE_LOOP_LO >> i) & 1;
}

// ---- Frobenius on Fq12: coefficient of w^i is conj^k(c_i) * gamma_{k,i};
// w-power layout: c0.c0 = w^0, c1.c0 = w^1, c0.c1 = w^2, c1.c1 = w^3, c0.c2 = w^4, c1.c2 = w^5
HDN Fq12 fq12_frob1(const Fq12& f, const PairingConsts& k) {
  Fq12 r;
  r.c0.c0 = f.c0.c0.conj();
  r.c1.c0 = f.c1.c0.conj() * k.g1[0];
  r.c0.c1 = f.c0.c1.conj() * k.g1[1];
  r.c1.c1 = f.c1.c1.conj() * k.g1[2];
  r.c0.c2 = f.c0.c2.conj() * k.g1[3];
  r.c1.c2 = f.c1.c2.conj() * k.g1[4];
  return r;
}
HDN Fq12 fq12_frob2(const Fq12& f, const PairingConsts& k) {
  Fq12 r;
  r.c0.c0 = f.c0.c0;
  r.c1.c0 = f.c1.c0.mul_fq(k.g2[0]);
  r.c0.c1 = f.c0.c1.mul_fq(k.g2[1]);
  r.c1.c1 = f.c1.c1.mul_fq(k.g2[2]);
  r.c0.c2 = f.c0.c2.mul_fq(k.g2[3]);
  r.c1.c2 = f.c1.c2.mul_fq(k.g2[4]);
  return r;
}
HDN Fq12 fq12_frob3(const Fq12& f, const PairingConsts& k) {
  Fq12 r;
  r.c0.c0 = f.c0.c0.conj();
  r.c1.c0 = f.c1.c0.conj() * k.g3[0];
  r.c0.c1 = f.c0.c1.conj() * k.g3[1];
  r.c1.c1 = f.c1.c1.conj() * k.g3[2];
  r.c0.c2 = f.c0.c2.conj() * k.g3[3];
  r.c1.c2 = f.c1.c2.conj() * k.g3[4];
  return r;
}

// ---- Miller loop over two (G1, fixed G2) pairs sharing the squarings.
// A pair whose G1 point is the identity contributes 1 (its lines are skipped).
HDN Fq12 miller_loop_2(const G1Affine& p1, const G2Line* t1, const G1Affine& p2, const G2Line* t2) {
  Fq12 f = Fq12::one();
  bool use1 = !p1.is_identity(), use2 = !p2.is_identity();
  int li = 0;
  for (int i = 63; i >= 0; i--) {
    f = f.sqr();
    if (use1) f = f.mul_by_line(p1.y, t1[li].neg_lam.mul_fq(p1.x), t1[li].c3);
    if (use2) f = f.mul_by_line(p2.y, t2[li].neg_lam.mul_fq(p2.x), t2[li].c3);
    li++;
    if (ate_bit(i)) {
      if (use1) f = f.mul_by_line(p1.y, t1[li].neg_lam.mul_fq(p1.x), t1[li].c3);
      if (use2) f = f.mul_by_line(p2.y, t2[li].neg_lam.mul_fq(p2.x), t2[li].c3);
      li++;
    }
  }
  for (int s = 0; s < 2; s++) {
    if (use1) f = f.mul_by_line(p1.y, t1[li].neg_lam.mul_fq(p1.x), t1[li].c3);
    if (use2) f = f.mul_by_line(p2.y, t2[li].neg_lam.mul_fq(p2.x), t2[li].c3);
    li++;
  }
  return f;
}

// f^x for the BN parameter x (63 bits), f in the cyclotomic subgroup
HDN Fq12 fq12_pow_x(const Fq12& f) {
  Fq12 r = f;
  for (int i = 61; i >= 0; i--) {  // bit 62 is the MSB
    r = r.cyclotomic_sqr();
    int bit = (i >= 32) ? ((SVK_BN_X_HI >> (i - 32)) & 1) : ((SVK_BN_X_LO >> i) & 1);
    if (bit) r = r * f;
  }
  return r;
}

// f^((p^12 - 1)/r) exactly:  easy part (p^6-1)(p^2+1), then
// hard = (p^4-p^2+1)/r = p^3 + (6x^2+1) p^2 + (-36x^3-18x^2-12x+1) p + (-36x^3-30x^2-18x-2)
HDN Fq12 final_exponentiation(const Fq12& f0, const PairingConsts& k) {
  Fq12 t = f0.conj() * f0.inv();      // ^(p^6 - 1)
  Fq12 f = fq12_frob2(t, k) * t;      // ^(p^2 + 1)   -> cyclotomic subgroup: inverse == conj
  Fq12 fx = fq12_pow_x(f);
  Fq12 fx2 = fq12_pow_x(fx);
  Fq12 fx3 = fq12_pow_x(fx2);
  // small powers (everything here lives in the cyclotomic subgroup: Granger-Scott squarings)
  Fq12 a2 = fx2.cyclotomic_sqr();                // fx2^2
  Fq12 a6 = a2.cyclotomic_sqr() * a2;            // fx2^6
  Fq12 a12 = a6.cyclotomic_sqr();                // fx2^12
  Fq12 a18 = a12 * a6;                           // fx2^18
  Fq12 a30 = a18 * a12;                          // fx2^30
  Fq12 b2 = fx.cyclotomic_sqr();
  Fq12 b6 = b2.cyclotomic_sqr() * b2;            // fx^6
  Fq12 b12 = b6.cyclotomic_sqr();                // fx^12
  Fq12 b18 = b12 * b6;                           // fx^18
  Fq12 c2 = fx3.cyclotomic_sqr();
  Fq12 c4 = c2.cyclotomic_sqr();
  Fq12 c8 = c4.cyclotomic_sqr();
  Fq12 c9 = c8 * fx3;
  Fq12 c18 = c9.cyclotomic_sqr();
  Fq12 c36 = c18.cyclotomic_sqr();               // fx3^36
  Fq12 e2 = a6 * f;                                   // f^(6x^2+1)
  Fq12 e1 = (c36 * a18 * b12).conj() * f;             // f^(-36x^3-18x^2-12x+1)
  Fq12 e0 = (c36 * a30 * b18 * f.cyclotomic_sqr()).conj();  // f^(-36x^3-30x^2-18x-2)
  return fq12_frob3(f, k) * fq12_frob2(e2, k) * fq12_frob1(e1, k) * e0;
}

// KzgAs::decide (decider.rs:60-68): accept iff e(lhs, g2) * e(rhs, -s_g2) == 1
HD bool kzg_decide(const G1Affine& lhs, const G1Affine& rhs, const G2Line* t_g2, const G2Line* t_neg_sg2,
                   const PairingConsts& k) {
  Fq12 f = miller_loop_2(lhs, t_g2, rhs, t_neg_sg2);
  return final_exponentiation(f, k).is_one();
}
