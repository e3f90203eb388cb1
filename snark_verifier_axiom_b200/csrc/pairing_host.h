// Host-side, once-per-key precomputation for the pairing kernels: Frobenius constants and the
// G2 line table of a `KzgDecidingKey` (snark-verifier/src/pcs/kzg/decider.rs:6-36).  Runs the
// same limb code as the device (PTX primitives emulated, csrc/ptx_arith.cuh); it is setup work,
// like halo2curves' `G2Prepared::from` (decider.rs:64), not a verification fallback.
#pragma once
#include <vector>

#include "pairing.cuh"

namespace svk_host {

inline Fq2 fq2_pow_limbs(const Fq2& a, const u32* e, int nlimbs) {
  Fq2 r = Fq2::one();
  for (int w = nlimbs - 1; w >= 0; w--)
    for (int b = 31; b >= 0; b--) {
      r = r.sqr();
      if ((e[w] >> b) & 1) r = r * a;
    }
  return r;
}

inline Fq2 fq2_xi() {
  Fq one = Fq::one();
  Fq nine = one.dbl().dbl().dbl() + one;
  return {nine, one};
}

inline PairingConsts make_pairing_consts() {
  // e = (p - 1) / 6
  u32 e[8];
  for (int i = 0; i < 8; i++) e[i] = FqParams::mod(i);
  e[0] -= 1;
  u64 rem = 0;
  for (int i = 7; i >= 0; i--) {
    u64 cur = (rem << 32) | e[i];
    e[i] = (u32)(cur / 6);
    rem = cur % 6;
  }
  PairingConsts k;
  Fq2 g = fq2_pow_limbs(fq2_xi(), e, 8);
  Fq2 acc = g;
  for (int i = 0; i < 5; i++) {
    k.g1[i] = acc;
    // gamma_{2,i} = gamma_{1,i} * conj(gamma_{1,i})  (lies in Fq)
    Fq2 n = acc * acc.conj();
    k.g2[i] = n.c0;
    acc = acc * g;
  }
  // gamma_{3,i} = gamma_{1,i} * gamma_{2,i}^p ... = conj-twisted product; frob3 = frob1 o frob2:
  // coefficient i: conj(c_i * g2_i) * g1_i = conj(c_i) * (g2_i * g1_i)   (g2_i in Fq)
  for (int i = 0; i < 5; i++) k.g3[i] = k.g1[i].mul_fq(k.g2[i]);
  return k;
}

inline Fq2 fq2_b_twist() {  // 3 / xi
  Fq one = Fq::one();
  Fq2 three = {one + one + one, Fq::zero()};
  return three * fq2_xi().inv();
}

inline bool g2_on_curve(const G2Affine& q) { return q.y.sqr() == q.x.sqr() * q.x + fq2_b_twist(); }

// Line table for the fixed G2 point `q` (affine, Montgomery), in the exact order
// `miller_loop_2` consumes it.
inline std::vector<G2Line> make_line_table(const G2Affine& q, const PairingConsts& k) {
  std::vector<G2Line> out;
  out.reserve(SVK_N_LINES);
  Fq2 xT = q.x, yT = q.y;
  auto dbl_step = [&]() {
    Fq2 x2 = xT.sqr();
    Fq2 lam = (x2.dbl() + x2) * yT.dbl().inv();
    out.push_back({lam.neg(), lam * xT - yT});
    Fq2 x3 = lam.sqr() - xT.dbl();
    Fq2 y3 = lam * (xT - x3) - yT;
    xT = x3;
    yT = y3;
  };
  auto add_step = [&](const Fq2& xQ, const Fq2& yQ) {
    Fq2 lam = (yQ - yT) * (xQ - xT).inv();
    out.push_back({lam.neg(), lam * xT - yT});
    Fq2 x3 = lam.sqr() - xT - xQ;
    Fq2 y3 = lam * (xT - x3) - yT;
    xT = x3;
    yT = y3;
  };
  for (int i = 63; i >= 0; i--) {
    dbl_step();
    if (ate_bit(i)) add_step(q.x, q.y);
  }
  // Q1 = pi(Q) = (conj(x) g1[1], conj(y) g1[2]);  Q2 = -pi^2(Q)
  Fq2 x1 = q.x.conj() * k.g1[1], y1 = q.y.conj() * k.g1[2];
  Fq2 x2 = x1.conj() * k.g1[1], y2 = (y1.conj() * k.g1[2]).neg();
  add_step(x1, y1);
  add_step(x2, y2);
  return out;
}

}  // namespace svk_host
