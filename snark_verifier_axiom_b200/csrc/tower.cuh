// BN254 extension tower: Fq2 = Fq[u]/(u^2+1), Fq6 = Fq2[v]/(v^3 - xi), Fq12 = Fq6[w]/(w^2 - v),
// xi = 9 + u.  Replaces halo2curves 0.3.1 `bn256::{Fq2,Fq6,Fq12}` underneath
// `multi_miller_loop().final_exponentiation()` (snark-verifier/src/pcs/kzg/decider.rs:64-65).
#pragma once
#include "field.cuh"

struct Fq2 {
  Fq c0, c1;
  HD static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
  HD static Fq2 one() { return {Fq::one(), Fq::zero()}; }
  HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  HD bool operator==(const Fq2& b) const { return c0 == b.c0 && c1 == b.c1; }
  HD friend Fq2 operator+(const Fq2& a, const Fq2& b) { return {a.c0 + b.c0, a.c1 + b.c1}; }
  HD friend Fq2 operator-(const Fq2& a, const Fq2& b) { return {a.c0 - b.c0, a.c1 - b.c1}; }
  HD Fq2 neg() const { return {c0.neg(), c1.neg()}; }
  HD Fq2 dbl() const { return {c0.dbl(), c1.dbl()}; }
  HD Fq2 conj() const { return {c0, c1.neg()}; }
  // Two fused dot products (field.cuh `dot2`): 4 limb products + 2 reductions = 400 wide MACs against 417 for Karatsuba's
  // three Montgomery products, and none of Karatsuba's five additions/subtractions (~125 instructions).
#if defined(__CUDA_ARCH__)
  // Device: ONE out-of-line function per Fq2 operation with both of its Fq products inlined -- two independent multiply-accumulate
  // chains for the scheduler to interleave and one call instead of two (batched k_decide: 35.9 -> 33.8 ms for 2^16 accumulators)
  static __device__ __noinline__ Fq2 mul_call(Fq2 a, Fq2 b) {
    Fq x[2] = {a.c0, a.c1.neg_lazy()}, y[2] = {b.c0, b.c1}, z[2] = {a.c0, a.c1}, w[2] = {b.c1, b.c0};
    return {Fq::template dot_inline<2>(x, y), Fq::template dot_inline<2>(z, w)};
  }
  static __device__ __noinline__ Fq2 sqr_call(Fq2 a) {
    Fq t = Fq::mul_inline(a.c0, a.c1);
    return {Fq::mul_inline(a.c0 + a.c1, a.c0 - a.c1), t.dbl()};
  }
  __device__ friend Fq2 operator*(const Fq2& a, const Fq2& b) { return mul_call(a, b); }
  __device__ Fq2 sqr() const { return sqr_call(*this); }
#else
  HD friend Fq2 operator*(const Fq2& a, const Fq2& b) {
    return {Fq::dot2(a.c0, b.c0, a.c1.neg_lazy(), b.c1), Fq::dot2(a.c0, b.c1, a.c1, b.c0)};
  }
  // complex squaring: 2 Fq mul
  HD Fq2 sqr() const {
    Fq t = c0 * c1;
    return {(c0 + c1) * (c0 - c1), t.dbl()};
  }
#endif
  HD Fq2 mul_fq(const Fq& s) const { return {c0 * s, c1 * s}; }
  // * xi = (9 + u): (9a - b) + (9b + a) u
  HD Fq2 mul_xi() const {
    Fq a8 = c0.dbl().dbl().dbl();
    Fq b8 = c1.dbl().dbl().dbl();
    return {a8 + c0 - c1, b8 + c1 + c0};
  }
  HD Fq2 inv() const {
    Fq n = (c0.sqr() + c1.sqr()).inv();
    return {c0 * n, (c1 * n).neg()};
  }
};

struct Fq6 {
  Fq2 c0, c1, c2;
  HD static Fq6 zero() { return {Fq2::zero(), Fq2::zero(), Fq2::zero()}; }
  HD static Fq6 one() { return {Fq2::one(), Fq2::zero(), Fq2::zero()}; }
  HD bool operator==(const Fq6& b) const { return c0 == b.c0 && c1 == b.c1 && c2 == b.c2; }
  HD friend Fq6 operator+(const Fq6& a, const Fq6& b) { return {a.c0 + b.c0, a.c1 + b.c1, a.c2 + b.c2}; }
  HD friend Fq6 operator-(const Fq6& a, const Fq6& b) { return {a.c0 - b.c0, a.c1 - b.c1, a.c2 - b.c2}; }
  HD Fq6 neg() const { return {c0.neg(), c1.neg(), c2.neg()}; }
  // Karatsuba / Toom-like: 6 Fq2 mul
  HD friend Fq6 operator*(const Fq6& a, const Fq6& b) {
    Fq2 t0 = a.c0 * b.c0;
    Fq2 t1 = a.c1 * b.c1;
    Fq2 t2 = a.c2 * b.c2;
    Fq2 r0 = ((a.c1 + a.c2) * (b.c1 + b.c2) - t1 - t2).mul_xi() + t0;
    Fq2 r1 = (a.c0 + a.c1) * (b.c0 + b.c1) - t0 - t1 + t2.mul_xi();
    Fq2 r2 = (a.c0 + a.c2) * (b.c0 + b.c2) - t0 - t2 + t1;
    return {r0, r1, r2};
  }
  HD Fq6 sqr() const { return (*this) * (*this); }
  // * v
  HD Fq6 mul_v() const { return {c2.mul_xi(), c0, c1}; }
  HD Fq6 mul_fq(const Fq& s) const { return {c0.mul_fq(s), c1.mul_fq(s), c2.mul_fq(s)}; }
  // * (b0 + b1 v): 5 Fq2 mul
  HD Fq6 mul_by_01(const Fq2& b0, const Fq2& b1) const {
    Fq2 t0 = c0 * b0;
    Fq2 t1 = c1 * b1;
    Fq2 r0 = (c2 * b1).mul_xi() + t0;
    Fq2 r1 = (c0 + c1) * (b0 + b1) - t0 - t1;
    Fq2 r2 = c2 * b0 + t1;
    return {r0, r1, r2};
  }
  HDN Fq6 inv() const {
    Fq2 t0 = c0.sqr() - (c1 * c2).mul_xi();
    Fq2 t1 = c2.sqr().mul_xi() - c0 * c1;
    Fq2 t2 = c1.sqr() - c0 * c2;
    Fq2 d = c0 * t0 + (c2 * t1 + c1 * t2).mul_xi();
    Fq2 di = d.inv();
    return {t0 * di, t1 * di, t2 * di};
  }
};

struct Fq12 {
  Fq6 c0, c1;
  HD static Fq12 one() { return {Fq6::one(), Fq6::zero()}; }
  HD bool operator==(const Fq12& b) const { return c0 == b.c0 && c1 == b.c1; }
  HD bool is_one() const { return *this == one(); }
  // 3 Fq6 mul = 54 Fq mul
  HDN friend Fq12 operator*(const Fq12& a, const Fq12& b) {
    Fq6 t0 = a.c0 * b.c0;
    Fq6 t1 = a.c1 * b.c1;
    Fq6 r1 = (a.c0 + a.c1) * (b.c0 + b.c1) - t0 - t1;
    return {t0 + t1.mul_v(), r1};
  }
  // complex squaring: 2 Fq6 mul = 36 Fq mul
  HDN Fq12 sqr() const {
    Fq6 ab = c0 * c1;
    Fq6 t = (c0 + c1) * (c0 + c1.mul_v()) - ab - ab.mul_v();
    return {t, ab + ab};
  }
  HD Fq12 conj() const { return {c0, c1.neg()}; }
  // Granger-Scott squaring for elements of the cyclotomic subgroup (after the easy part of the final
  // exponentiation): three Fq4 squarings = 9 Fq2 mul-equivalents = 18 Fq mul instead of 36.
  // Fq4 = Fq2[y]/(y^2 - xi) pairs: (c0.c0, c1.c1), (c1.c0, c0.c2), (c0.c1, c1.c2).
  HDN Fq12 cyclotomic_sqr() const {
    const Fq2 &z0 = c0.c0, &z4 = c0.c1, &z3 = c0.c2, &z2 = c1.c0, &z1 = c1.c1, &z5 = c1.c2;
    Fq2 tmp = z0 * z1;
    Fq2 t0 = (z0 + z1) * (z1.mul_xi() + z0) - tmp - tmp.mul_xi();
    Fq2 t1 = tmp.dbl();
    tmp = z2 * z3;
    Fq2 t2 = (z2 + z3) * (z3.mul_xi() + z2) - tmp - tmp.mul_xi();
    Fq2 t3 = tmp.dbl();
    tmp = z4 * z5;
    Fq2 t4 = (z4 + z5) * (z5.mul_xi() + z4) - tmp - tmp.mul_xi();
    Fq2 t5 = tmp.dbl();
    Fq12 r;
    // z0' = 3 t0 - 2 z0 ; z1' = 3 t1 + 2 z1 ; z2' = 3 xi t5 + 2 z2 ; z3' = 3 t4 - 2 z3 ; z4' = 3 t2 - 2 z4 ; z5' = 3 t3 + 2 z5
    r.c0.c0 = (t0 - z0).dbl() + t0;
    r.c1.c1 = (t1 + z1).dbl() + t1;
    Fq2 x5 = t5.mul_xi();
    r.c1.c0 = (x5 + z2).dbl() + x5;
    r.c0.c2 = (t4 - z3).dbl() + t4;
    r.c0.c1 = (t2 - z4).dbl() + t2;
    r.c1.c2 = (t3 + z5).dbl() + t3;
    return r;
  }
  HDN Fq12 inv() const {
    Fq6 d = (c0.sqr() - c1.sqr().mul_v()).inv();
    return {c0 * d, (c1 * d).neg()};
  }
  // * line  l = l0 + (l1 + l2 v) w   with l0 in Fq, l1, l2 in Fq2   (36 Fq mul)
  HDN Fq12 mul_by_line(const Fq& l0, const Fq2& l1, const Fq2& l2) const {
    Fq6 t0 = c0.mul_fq(l0);
    Fq6 t1 = c1.mul_by_01(l1, l2);
    Fq2 s0 = l1;
    s0.c0 = s0.c0 + l0;
    Fq6 t2 = (c0 + c1).mul_by_01(s0, l2);
    return {t0 + t1.mul_v(), t2 - t0 - t1};
  }
};
