// BN254 G1 (y^2 = x^3 + 3 over Fq) point arithmetic.
//
// Replaces halo2curves 0.3.1 `bn256::{G1, G1Affine}` as used by
//   NativeLoader::multi_scalar_multiplication   snark-verifier/src/loader/native.rs:61-71
//   G1Affine::from_bytes / coordinates          snark-verifier/src/system/halo2/transcript/halo2.rs:216,252
// Results leave the device as canonical affine points, so any correct group law is bit-exact.
//
// Representations (all coordinates Fq in Montgomery form):
//   G1Affine  (x, y), identity encoded as (0, 0) -- (0,0) is not on the curve.
//   G1Jac     (X, Y, Z), x = X/Z^2, y = Y/Z^3, identity Z = 0       (scalar-mul accumulators)
//   G1Xyzz    (X, Y, ZZ, ZZZ), x = X/ZZ, y = Y/ZZZ, identity ZZ = 0 (Pippenger buckets: cheapest
//             mixed add, 8M + 2S, no inversion)
#pragma once
#include "field.cuh"

template <class F>
struct AffT {
  typedef AffT G1Affine;
  F x, y;
  HD bool is_identity() const { return x.is_zero() && y.is_zero(); }
  HD static G1Affine identity() { return {F::zero(), F::zero()}; }
  HD G1Affine neg() const { return is_identity() ? *this : G1Affine{x, y.neg()}; }
};

template <class F>
struct JacT {
  typedef AffT<F> G1Affine;
  typedef JacT G1Jac;
  F X, Y, Z;
  HD bool is_identity() const { return Z.is_zero(); }
  HD static G1Jac identity() { return {F::one(), F::one(), F::zero()}; }
  HD static G1Jac from_affine(const G1Affine& p) {
    if (p.is_identity()) return identity();
    return {p.x, p.y, F::one()};
  }
  // dbl-2009-l (a = 0): 2M + 5S
  HD G1Jac dbl() const {
    if (is_identity()) return *this;
    F A = X.sqr();
    F B = Y.sqr();
    F C = B.sqr();
    F t = (X + B).sqr() - A - C;
    F D = t.dbl();
    F E = A.dbl() + A;
    F EE = E.sqr();
    G1Jac r;
    r.X = EE - D.dbl();
    F C8 = C.dbl().dbl().dbl();
    r.Y = E * (D - r.X) - C8;
    r.Z = (Y * Z).dbl();
    return r;
  }
  // madd-2007-bl mixed addition: 7M + 4S; handles identity / doubling / inverse operands
  HD G1Jac add_affine(const G1Affine& q) const {
    if (q.is_identity()) return *this;
    if (is_identity()) return from_affine(q);
    F Z1Z1 = Z.sqr();
    F U2 = q.x * Z1Z1;
    F S2 = q.y * Z * Z1Z1;
    if (U2 == X) {
      if (S2 == Y) return dbl();
      return identity();
    }
    F H = U2 - X;
    F HH = H.sqr();
    F I = HH.dbl().dbl();
    F J = H * I;
    F rr = (S2 - Y).dbl();
    F V = X * I;
    G1Jac r;
    r.X = rr.sqr() - J - V.dbl();
    r.Y = F::dot2(rr, V - r.X, Y.dbl().neg_lazy(), J);  // rr (V - X3) - 2 Y J, one shared reduction
    r.Z = (Z + H).sqr() - Z1Z1 - HH;
    return r;
  }
  // add-2007-bl general addition: 11M + 5S
  HD G1Jac add(const G1Jac& q) const {
    if (q.is_identity()) return *this;
    if (is_identity()) return q;
    F Z1Z1 = Z.sqr();
    F Z2Z2 = q.Z.sqr();
    F U1 = X * Z2Z2;
    F U2 = q.X * Z1Z1;
    F S1 = Y * q.Z * Z2Z2;
    F S2 = q.Y * Z * Z1Z1;
    if (U1 == U2) {
      if (S1 == S2) return dbl();
      return identity();
    }
    F H = U2 - U1;
    F I = H.dbl().sqr();
    F J = H * I;
    F rr = (S2 - S1).dbl();
    F V = U1 * I;
    G1Jac r;
    r.X = rr.sqr() - J - V.dbl();
    r.Y = F::dot2(rr, V - r.X, S1.dbl().neg_lazy(), J);  // rr (V - X3) - 2 S1 J
    r.Z = ((Z + q.Z).sqr() - Z1Z1 - Z2Z2) * H;
    return r;
  }
  HD G1Jac neg() const { return {X, Y.neg(), Z}; }
  // to affine with a caller-supplied inverse of Z (batched inversion) or its own inversion
  HD G1Affine to_affine_with_zinv(const F& zinv) const {
    if (is_identity()) return G1Affine::identity();
    F zi2 = zinv.sqr();
    return {X * zi2, Y * zi2 * zinv};
  }
  HD G1Affine to_affine() const { return to_affine_with_zinv(Z.inv()); }
};

template <class F>
struct XyzzT {
  typedef AffT<F> G1Affine;
  typedef JacT<F> G1Jac;
  typedef XyzzT G1Xyzz;
  F X, Y, ZZ, ZZZ;
  HD bool is_identity() const { return ZZ.is_zero(); }
  HD static G1Xyzz identity() { return {F::zero(), F::zero(), F::zero(), F::zero()}; }
  HD static G1Xyzz from_affine(const G1Affine& p) {
    if (p.is_identity()) return identity();
    return {p.x, p.y, F::one(), F::one()};
  }
  // dbl-2008-s-1 (a = 0): 6M + 4S... here 2M? -> U=2Y, V=U^2, W=U*V, S=X*V, M=3X^2, X3=M^2-2S, Y3=M(S-X3)-W*Y, ZZ3=V*ZZ, ZZZ3=W*ZZZ
  HD G1Xyzz dbl() const {
    if (is_identity()) return *this;
    F U = Y.dbl();
    F V = U.sqr();
    F W = U * V;
    F S = X * V;
    F XX = X.sqr();
    F M = XX.dbl() + XX;
    G1Xyzz r;
    r.X = M.sqr() - S.dbl();
    r.Y = F::dot2(M, S - r.X, W.neg_lazy(), Y);
    r.ZZ = V * ZZ;
    r.ZZZ = W * ZZZ;
    return r;
  }
  HD static G1Xyzz dbl_affine(const G1Affine& p) {
    F U = p.y.dbl();
    F V = U.sqr();
    F W = U * V;
    F S = p.x * V;
    F XX = p.x.sqr();
    F M = XX.dbl() + XX;
    G1Xyzz r;
    r.X = M.sqr() - S.dbl();
    r.Y = F::dot2(M, S - r.X, W.neg_lazy(), p.y);
    r.ZZ = V;
    r.ZZZ = W;
    return r;
  }
  // madd-2008-s: 8M + 2S
  HD G1Xyzz add_affine(const G1Affine& q) const {
    if (q.is_identity()) return *this;
    if (is_identity()) return from_affine(q);
    F U2 = q.x * ZZ;
    F S2 = q.y * ZZZ;
    if (U2 == X) {
      if (S2 == Y) return dbl_affine(q);
      return identity();
    }
    F Pp = U2 - X;
    F Rr = S2 - Y;
    F PP = Pp.sqr();
    F PPP = Pp * PP;
    F Q = X * PP;
    G1Xyzz r;
    r.X = Rr.sqr() - PPP - Q.dbl();
    r.Y = F::dot2(Rr, Q - r.X, Y.neg_lazy(), PPP);
    r.ZZ = ZZ * PP;
    r.ZZZ = ZZZ * PPP;
    return r;
  }
  // add-2008-s: 12M + 2S
  HD G1Xyzz add(const G1Xyzz& q) const {
    if (q.is_identity()) return *this;
    if (is_identity()) return q;
    F U1 = X * q.ZZ;
    F U2 = q.X * ZZ;
    F S1 = Y * q.ZZZ;
    F S2 = q.Y * ZZZ;
    if (U1 == U2) {
      if (S1 == S2) return dbl();
      return identity();
    }
    F Pp = U2 - U1;
    F Rr = S2 - S1;
    F PP = Pp.sqr();
    F PPP = Pp * PP;
    F Q = U1 * PP;
    G1Xyzz r;
    r.X = Rr.sqr() - PPP - Q.dbl();
    r.Y = F::dot2(Rr, Q - r.X, S1.neg_lazy(), PPP);
    r.ZZ = ZZ * q.ZZ * PP;
    r.ZZZ = ZZZ * q.ZZZ * PPP;
    return r;
  }
  // x = X/ZZ, y = Y/ZZZ.  One inversion: (ZZ*ZZZ)^-1 -> 1/ZZ = inv*ZZZ, 1/ZZZ = inv*ZZ
  HD G1Affine to_affine() const {
    if (is_identity()) return G1Affine::identity();
    F inv = (ZZ * ZZZ).inv();
    return {X * (inv * ZZZ), Y * (inv * ZZ)};
  }
  HD G1Jac to_jac() const {
    // (X, Y, ZZ, ZZZ) -> Jacobian with Z = ZZZ/ZZ is not polynomial; use Z' = ZZ*ZZZ... instead:
    // x = X/ZZ = X*ZZ*ZZZ^2/(ZZ*ZZZ)^2, y = Y/ZZZ = Y*ZZ^3*ZZZ^2/(ZZ*ZZZ)^3 with Z = ZZ*ZZZ
    if (is_identity()) return G1Jac::identity();
    F Z = ZZ * ZZZ;
    F ZZZ2 = ZZZ.sqr();
    return {X * ZZ * ZZZ2, Y * ZZ.sqr() * ZZ * ZZZ2, Z};
  }
};

// BN254 G1 instantiation (the KZG path); csrc/pasta.cuh instantiates the same templates over the Pasta base fields.
typedef AffT<Fq> G1Affine;
typedef JacT<Fq> G1Jac;
typedef XyzzT<Fq> G1Xyzz;

HD Fq fq_b3() {  // curve constant b = 3 in Montgomery form
  Fq one = Fq::one();
  return one + one + one;
}

HD bool g1_on_curve(const G1Affine& p) {
  if (p.is_identity()) return true;
  return p.y.sqr() == p.x.sqr() * p.x + fq_b3();
}

// 256-bit scalar (canonical, little-endian limbs) times affine point; MSB-first double-and-add.
// Only the value is observable (SURVEY finding 1), the schedule is ours.
HD G1Jac g1_scalar_mul(const G1Affine& p, const u32* k) {
  G1Jac acc = G1Jac::identity();
  bool started = false;
  for (int w = 7; w >= 0; w--) {
    for (int b = 31; b >= 0; b--) {
      if (started) acc = acc.dbl();
      if ((k[w] >> b) & 1) {
        acc = acc.add_affine(p);
        started = true;
      }
    }
  }
  return acc;
}

// halo2curves 0.3.1 compressed G1 (32 B): x little-endian, bit 7 of byte 31 = lsb(y), all-zero =
// identity.  Returns 0 ok, 1 invalid encoding (x >= p or x^3+3 non-residue), 2 identity (decodes,
// but `common_ec_point` rejects it: transcript/halo2.rs:214-224).  `xc`/`yc` get the canonical limbs.
HD int g1_decompress(const uint8_t* bytes, G1Affine& out, u32* xc, u32* yc) {
  u32 x[8];
  fe_load_le(x, bytes);
  u32 ysign = x[7] >> 31;
  x[7] &= 0x7fffffffu;
  out = G1Affine::identity();
  if (!Fq::is_canonical(x)) return 1;
  Fq xf;
#pragma unroll
  for (int i = 0; i < 8; i++) xf.v[i] = x[i];
  if (xf.is_zero() && ysign == 0) return 2;
  Fq xm = xf.to_mont();
  Fq rhs = xm.sqr() * xm + fq_b3();
  Fq y = rhs.sqrt_candidate();
  if (y.sqr() != rhs) return 1;
  Fq yc_ = y.from_mont();
  if ((yc_.v[0] & 1) != ysign) {
    y = y.neg();
    yc_ = y.from_mont();
  }
  out.x = xm;
  out.y = y;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    xc[i] = x[i];
    yc[i] = yc_.v[i];
  }
  return 0;
}
