// Host build of the device arithmetic headers (PTX primitives emulated bit-exactly, see
// csrc/ptx_arith.cuh) so the limb algorithms can be checked on a CPU-only box against the
// Python oracle.  Test-only; never linked into libsvk.
#include "../../snark_verifier_axiom_b200/csrc/field.cuh"
#include <cstring>

extern "C" {
// op: 0 mul, 1 add, 2 sub, 3 neg, 4 inv, 5 sqrt_candidate, 6 to_mont, 7 from_mont, 8 sqr, 9 dbl
// field: 0 Fq, 1 Fr.  a,b,out: 8 u32 limbs (raw, no conversion)
void host_fe_op(int field, int op, const u32* a, const u32* b, u32* out) {
  if (field == 0) {
    Fq x, y, r;
    memcpy(x.v, a, 32); memcpy(y.v, b, 32);
    switch (op) {
      case 0: r = x * y; break; case 1: r = x + y; break; case 2: r = x - y; break; case 3: r = x.neg(); break;
      case 4: r = x.inv(); break; case 5: r = x.sqrt_candidate(); break; case 6: r = x.to_mont(); break;
      case 7: r = x.from_mont(); break; case 8: r = x.sqr(); break; default: r = x.dbl(); break;
    }
    memcpy(out, r.v, 32);
  } else {
    Fr x, y, r;
    memcpy(x.v, a, 32); memcpy(y.v, b, 32);
    switch (op) {
      case 0: r = x * y; break; case 1: r = x + y; break; case 2: r = x - y; break; case 3: r = x.neg(); break;
      case 4: r = x.inv(); break; case 5: r = x.sqrt_candidate(); break; case 6: r = x.to_mont(); break;
      case 7: r = x.from_mont(); break; case 8: r = x.sqr(); break; default: r = x.dbl(); break;
    }
    memcpy(out, r.v, 32);
  }
}
int host_is_canonical(int field, const u32* a) { return field == 0 ? Fq::is_canonical(a) : Fr::is_canonical(a); }
}

#include "../../snark_verifier_axiom_b200/csrc/g1.cuh"
static G1Affine load_aff(const u32* xy) {  // canonical x,y (16 limbs) -> Montgomery
  G1Affine p; memcpy(p.x.v, xy, 32); memcpy(p.y.v, xy + 8, 32);
  if (p.is_identity()) return p;
  p.x = p.x.to_mont(); p.y = p.y.to_mont(); return p;
}
static void store_aff(u32* xy, const G1Affine& p) {
  Fq x = p.x.from_mont(), y = p.y.from_mont(); memcpy(xy, x.v, 32); memcpy(xy + 8, y.v, 32);
}
extern "C" {
// out = k*P + Q computed through the requested representation (0 Jacobian, 1 XYZZ incl. general add)
void host_g1_muladd(int repr, const u32* p_xy, const u32* k, const u32* q_xy, u32* out_xy) {
  G1Affine p = load_aff(p_xy), q = load_aff(q_xy);
  G1Jac kp = g1_scalar_mul(p, k);
  if (repr == 0) {
    G1Jac r = kp.add_affine(q);
    r = r.add(G1Jac::identity());
    store_aff(out_xy, r.to_affine());
  } else {
    G1Xyzz acc = G1Xyzz::identity();
    for (int w = 7; w >= 0; w--) for (int b = 31; b >= 0; b--) { acc = acc.dbl(); if ((k[w] >> b) & 1) acc = acc.add_affine(p); }
    G1Xyzz qq = G1Xyzz::from_affine(q);
    G1Xyzz r = acc.add(qq);
    G1Jac rj = r.to_jac();
    G1Affine a1 = r.to_affine(), a2 = rj.to_affine();
    if (!(a1.x == a2.x) || !(a1.y == a2.y)) { memset(out_xy, 0xff, 64); return; }
    store_aff(out_xy, a1);
  }
}
// general Jacobian add: out = k1*P + k2*P via add()
void host_g1_jac_add(const u32* p_xy, const u32* k1, const u32* k2, u32* out_xy) {
  G1Affine p = load_aff(p_xy);
  G1Jac a = g1_scalar_mul(p, k1), b = g1_scalar_mul(p, k2);
  store_aff(out_xy, a.add(b).to_affine());
}
int host_g1_decompress(const uint8_t* bytes, u32* out_xy) {
  G1Affine p; u32 xc[8], yc[8];
  int rc = g1_decompress(bytes, p, xc, yc);
  if (rc == 0) { memcpy(out_xy, xc, 32); memcpy(out_xy + 8, yc, 32); if (!g1_on_curve(p)) return 9; }
  return rc;
}
}

#include "../../snark_verifier_axiom_b200/csrc/pairing_host.h"
static Fq2 load_fq2(const u32* p) { Fq2 r; memcpy(r.c0.v, p, 32); memcpy(r.c1.v, p + 8, 32); r.c0 = r.c0.to_mont(); r.c1 = r.c1.to_mont(); return r; }
static G2Affine load_g2(const u32* p) { return {load_fq2(p), load_fq2(p + 16)}; }
extern "C" {
// gt_out: 12 x 8 canonical limbs of FE(ML(p1,q1) * ML(p2,q2)); returns decide bit; -1 if q off-curve
int host_pairing(const u32* p1_xy, const u32* q1, const u32* p2_xy, const u32* q2, u32* gt_out, u32* ml_out) {
  static PairingConsts k = svk_host::make_pairing_consts();
  G2Affine Q1 = load_g2(q1), Q2 = load_g2(q2);
  if (!svk_host::g2_on_curve(Q1) || !svk_host::g2_on_curve(Q2)) return -1;
  std::vector<G2Line> t1 = svk_host::make_line_table(Q1, k), t2 = svk_host::make_line_table(Q2, k);
  if ((int)t1.size() != SVK_N_LINES) return -2;
  G1Affine P1 = load_aff(p1_xy), P2 = load_aff(p2_xy);
  Fq12 ml = miller_loop_2(P1, t1.data(), P2, t2.data());
  Fq12 gt = final_exponentiation(ml, k);
  const Fq2* c[6] = {&gt.c0.c0, &gt.c0.c1, &gt.c0.c2, &gt.c1.c0, &gt.c1.c1, &gt.c1.c2};
  const Fq2* m[6] = {&ml.c0.c0, &ml.c0.c1, &ml.c0.c2, &ml.c1.c0, &ml.c1.c1, &ml.c1.c2};
  for (int i = 0; i < 6; i++) {
    Fq a = c[i]->c0.from_mont(), b = c[i]->c1.from_mont();
    memcpy(gt_out + 16 * i, a.v, 32); memcpy(gt_out + 16 * i + 8, b.v, 32);
    a = m[i]->c0.from_mont(); b = m[i]->c1.from_mont();
    memcpy(ml_out + 16 * i, a.v, 32); memcpy(ml_out + 16 * i + 8, b.v, 32);
  }
  return kzg_decide(P1, P2, t1.data(), t2.data(), k) ? 1 : 0;
}
}
