// K0: BN254 Fq / Fr Montgomery arithmetic on 8 x 32-bit limbs (R = 2^256).
//
// Replaces halo2curves 0.3.1 `bn256::{Fq,Fr}` (un-vendored; Cargo.lock:1803-1826) at the call
// sites listed in SURVEY.md 8c.  Values are kept fully reduced in [0, p) in Montgomery form.
//
// mul(): interleaved (CIOS-style) multiply+reduce with the products split over two accumulators
// ("aligned" columns 2k,2k+1 and "offset" columns 2k+1,2k+2) so that every 32x32 partial product is
// one 64-bit multiply-accumulate with a predicate carry in and out.  ptxas turns each
// mad.lo.cc/madc.hi.cc pair into a single IMAD.WIDE.U32.X: 139 IMAD* + 37 IADD3 per modmul on
// sm_100a (cuobjdump count).  Both moduli are 254-bit, so the top accumulator pair
// never overflows.
#pragma once
#include "ptx_arith.cuh"

struct FqParams {
  HD static constexpr u32 mod(int i) {
    constexpr u32 m[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return m[i];
  }
  HD static constexpr u32 one(int i) {
    constexpr u32 m[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return m[i];
  }
  HD static constexpr u32 r2(int i) {
    constexpr u32 m[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
    return m[i];
  }
  static constexpr u32 M0 = 0xe4866389u;
};

struct FrParams {
  HD static constexpr u32 mod(int i) {
    constexpr u32 m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return m[i];
  }
  HD static constexpr u32 one(int i) {
    constexpr u32 m[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return m[i];
  }
  HD static constexpr u32 r2(int i) {
    constexpr u32 m[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
    return m[i];
  }
  static constexpr u32 M0 = 0xefffffffu;
};

template <class P>
struct Fe {
  u32 v[8];

  HD static Fe zero() {
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
  }
  HD static Fe one() {
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = P::one(i);
    return r;
  }
  HD static Fe modulus() {
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = P::mod(i);
    return r;
  }
  HD bool is_zero() const {
    u32 o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= v[i];
    return o == 0;
  }
  HD bool operator==(const Fe& b) const {
    u32 o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= v[i] ^ b.v[i];
    return o == 0;
  }
  HD bool operator!=(const Fe& b) const { return !(*this == b); }

  // r = a - p if a >= p else a     (a < 2p, possibly with a carry word `hi`)
  HD static void reduce_once(u32* a, u32 hi = 0) {
    u32 s[8];
    s[0] = ptx::sub_cc(a[0], P::mod(0));
#pragma unroll
    for (int i = 1; i < 8; i++) s[i] = ptx::subc_cc(a[i], P::mod(i));
    u32 br = ptx::subc(hi, 0);  // hi - borrow: 0xffffffff iff a < p (and hi == 0)
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = br ? a[i] : s[i];
  }

  HD friend Fe operator+(const Fe& a, const Fe& b) {
    Fe r;
    r.v[0] = ptx::add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) r.v[i] = ptx::addc_cc(a.v[i], b.v[i]);
    // 2p < 2^255: no carry out of the top limb
    reduce_once(r.v);
    return r;
  }
  HD friend Fe operator-(const Fe& a, const Fe& b) {
    Fe r;
    r.v[0] = ptx::sub_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) r.v[i] = ptx::subc_cc(a.v[i], b.v[i]);
    u32 br = ptx::subc(0, 0);  // 0xffffffff iff borrow
    u32 t[8];
    t[0] = ptx::add_cc(r.v[0], P::mod(0) & br);
#pragma unroll
    for (int i = 1; i < 7; i++) t[i] = ptx::addc_cc(r.v[i], P::mod(i) & br);
    t[7] = ptx::addc(r.v[7], P::mod(7) & br);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
    return r;
  }
  HD Fe neg() const { return zero() - *this; }
  HD Fe dbl() const { return *this + *this; }

  // ---- Montgomery multiplication -------------------------------------------------------------
  // acc pair (j, j+1) = a[j] * bi               (j = 0,2,4,6; a is a pointer to the first limb used)
  HD static void mul_row(u32* acc, const u32* a, u32 bi) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      acc[j] = ptx::mul_lo(a[j], bi);
      acc[j + 1] = ptx::mul_hi(a[j], bi);
    }
  }
  // acc pairs += a[j] * bi with one carry chain; leaves the carry-out in CC
  HD static void cmad_row(u32* acc, const u32* a, u32 bi) {
    acc[0] = ptx::mad_lo_cc(a[0], bi, acc[0]);
    acc[1] = ptx::madc_hi_cc(a[0], bi, acc[1]);
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
      acc[j] = ptx::madc_lo_cc(a[j], bi, acc[j]);
      acc[j + 1] = ptx::madc_hi_cc(a[j], bi, acc[j + 1]);
    }
  }
  HD static void cmad_row_mod(u32* acc, int off, u32 mi) {
    acc[0] = ptx::mad_lo_cc(P::mod(off), mi, acc[0]);
    acc[1] = ptx::madc_hi_cc(P::mod(off), mi, acc[1]);
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
      acc[j] = ptx::madc_lo_cc(P::mod(off + j), mi, acc[j]);
      acc[j + 1] = ptx::madc_hi_cc(P::mod(off + j), mi, acc[j + 1]);
    }
  }
  // odd[j] = odd[j+2] + (a[j] * bi) words, continuing the carry in CC (shift right by two limbs)
  HD static void madc_row_rshift(u32* odd, const u32* a, u32 bi) {
#pragma unroll
    for (int j = 0; j < 6; j += 2) {
      odd[j] = ptx::madc_lo_cc(a[j], bi, odd[j + 2]);
      odd[j + 1] = ptx::madc_hi_cc(a[j], bi, odd[j + 3]);
    }
    odd[6] = ptx::madc_lo_cc(a[6], bi, 0);
    odd[7] = ptx::madc_hi(a[6], bi, 0);
  }
  // one row: T = (T + a*bi + m*p) / 2^32.  `al` holds columns (2k,2k+1), `of` columns (2k+1,2k+2).
  // On exit the roles of the two arrays are swapped (the caller alternates them).
  HD static void mad_row_redc(u32* al, u32* of, const u32* a, u32 bi, bool first) {
    if (first) {
      mul_row(of, a + 1, bi);
      mul_row(al, a, bi);
    } else {
      al[0] = ptx::add_cc(al[0], of[1]);
      madc_row_rshift(of, a + 1, bi);
      cmad_row(al, a, bi);
      of[7] = ptx::addc(of[7], 0);
    }
    u32 mi = al[0] * P::M0;
    cmad_row_mod(of, 1, mi);
    cmad_row_mod(al, 0, mi);
    of[7] = ptx::addc(of[7], 0);
  }

  // Fully inlined multiplication (176 SASS instructions).  Device code normally goes through the
  // out-of-line copy below: with the product inlined at every call site the hot loops of the
  // latency-bound kernels (one warp per SM sub-partition) overflow the instruction caches
  // (ncu: `stalled_no_instruction` up to 2.3 per issue, profiles/r1_ncu_notes.md); a by-value call
  // costs ~16 MOVs (arguments and result travel in registers, no local memory).
  HD static Fe mul_inline(const Fe& a, const Fe& b) {
    u32 al[8], of[8];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      mad_row_redc(al, of, a.v, b.v[i], i == 0);
      mad_row_redc(of, al, a.v, b.v[i + 1], false);
    }
    // after the last row of[0] == 0 and the pending one-limb shift folds of[1..7] onto al[0..6]
    Fe r;
    r.v[0] = ptx::add_cc(al[0], of[1]);
#pragma unroll
    for (int j = 1; j < 7; j++) r.v[j] = ptx::addc_cc(al[j], of[j + 1]);
    r.v[7] = ptx::addc(al[7], 0);
    reduce_once(r.v);
    return r;
  }
#if defined(__CUDA_ARCH__)
  static __device__ __noinline__ Fe mul_call(Fe a, Fe b) { return mul_inline(a, b); }
  HD friend Fe operator*(const Fe& a, const Fe& b) { return mul_call(a, b); }
#else
  HD friend Fe operator*(const Fe& a, const Fe& b) { return mul_inline(a, b); }
#endif
  // ---- Montgomery squaring: 36 + 64 wide MACs instead of 128 ------------------------------------------------------
  // (1) the 28 off-diagonal products a_i a_j (i < j) go to two 16-limb accumulators by column parity (E: pairs at even
  //     columns, O: pairs at odd columns, O[k] = column k + 1), one carry chain per (row, parity); rows ascend, so the limb
  //     that receives a chain's carry-out holds only earlier carry-outs at that moment (no ripple);
  // (2) T = 2 (E + (O << 32)) + sum_i a_i^2 2^(64 i)   (one add chain, one funnel-shift pass, one 8-MAC chain);
  // (3) U = (T_lo + M p) / 2^256 by eight pure reduction rows (the two-limb shift of the `of` accumulator is fused into the
  //     m p MACs), result = U + T_hi < 1.19 p + 1, one conditional subtraction.
  HD static void redc_row(u32* al, u32* of, bool first) {
    if (first) {
      u32 mi = al[0] * P::M0;
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        of[j] = ptx::mul_lo(P::mod(j + 1), mi);
        of[j + 1] = ptx::mul_hi(P::mod(j + 1), mi);
      }
      cmad_row_mod(al, 0, mi);
      of[7] = ptx::addc(of[7], 0);
      } else {
      u32 mi = (al[0] + of[1]) * P::M0;
      al[0] = ptx::add_cc(al[0], of[1]);
#pragma unroll
      for (int j = 0; j < 6; j += 2) {
        of[j] = ptx::madc_lo_cc(P::mod(j + 1), mi, of[j + 2]);
        of[j + 1] = ptx::madc_hi_cc(P::mod(j + 1), mi, of[j + 3]);
      }
      of[6] = ptx::madc_lo_cc(P::mod(7), mi, 0);
      of[7] = ptx::madc_hi(P::mod(7), mi, 0);
      cmad_row_mod(al, 0, mi);
      of[7] = ptx::addc(of[7], 0);
      }
  }
  // T (16 limbs, < 2^510) -> T * 2^-256 mod p: U = (T_lo + M p) / 2^256 by eight pure reduction rows, result = U + T_hi < 2p,
  // one conditional subtraction.
  HD static Fe redc_wide(const u32* T) {
    u32 al[8], of[8];
#pragma unroll
    for (int i = 0; i < 8; i++) al[i] = T[i];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      redc_row(al, of, i == 0);
      redc_row(of, al, false);
    }
    Fe r;
    r.v[0] = ptx::add_cc(al[0], of[1]);
#pragma unroll
    for (int j = 1; j < 7; j++) r.v[j] = ptx::addc_cc(al[j], of[j + 1]);
    r.v[7] = ptx::addc(al[7], 0);
    // + T_hi
    r.v[0] = ptx::add_cc(r.v[0], T[8]);
#pragma unroll
    for (int j = 1; j < 7; j++) r.v[j] = ptx::addc_cc(r.v[j], T[8 + j]);
    r.v[7] = ptx::addc(r.v[7], T[15]);
    reduce_once(r.v);
    return r;
  }

  HD static Fe sqr_inline(const Fe& x) {
    const u32* a = x.v;
    u32 E[16], O[16];
#pragma unroll
    for (int i = 0; i < 16; i++) E[i] = O[i] = 0;
#pragma unroll
    for (int i = 0; i < 7; i++) {
      // odd columns: j = i+1, i+3, ...   ->  O[i+j-1], O[i+j]
      {
        int k = 2 * i;  // O index of column 2i+1
        O[k] = ptx::mad_lo_cc(a[i], a[i + 1], O[k]);
        O[k + 1] = ptx::madc_hi_cc(a[i], a[i + 1], O[k + 1]);
#pragma unroll
        for (int j = i + 3; j < 8; j += 2) {
          O[i + j - 1] = ptx::madc_lo_cc(a[i], a[j], O[i + j - 1]);
          O[i + j] = ptx::madc_hi_cc(a[i], a[j], O[i + j]);
        }
        int last = i + 1 + 2 * ((7 - (i + 1)) / 2);  // last j of the chain
        O[i + last + 1] = ptx::addc(O[i + last + 1], 0);
      }
      // even columns: j = i+2, i+4, ...  ->  E[i+j], E[i+j+1]
      if (i + 2 < 8) {
        E[2 * i + 2] = ptx::mad_lo_cc(a[i], a[i + 2], E[2 * i + 2]);
        E[2 * i + 3] = ptx::madc_hi_cc(a[i], a[i + 2], E[2 * i + 3]);
#pragma unroll
        for (int j = i + 4; j < 8; j += 2) {
          E[i + j] = ptx::madc_lo_cc(a[i], a[j], E[i + j]);
          E[i + j + 1] = ptx::madc_hi_cc(a[i], a[j], E[i + j + 1]);
        }
        int last = i + 2 + 2 * ((7 - (i + 2)) / 2);
        E[i + last + 2] = ptx::addc(E[i + last + 2], 0);
      }
    }
    // S = E + (O << 32); columns 0 and 15 of S are empty (a < 2^254: the off-diagonal sum is < 2^507)
    u32 T[16];
    T[0] = 0;
    T[1] = O[0];
    T[2] = ptx::add_cc(E[2], O[1]);
#pragma unroll
    for (int c = 3; c < 15; c++) T[c] = ptx::addc_cc(E[c], O[c - 1]);
    T[15] = ptx::addc(E[15], O[14]);
    // T = 2 S
#pragma unroll
    for (int c = 15; c >= 1; c--) T[c] = (T[c] << 1) | (T[c - 1] >> 31);
    // T += diagonal
    T[0] = ptx::mul_lo(a[0], a[0]);
    T[1] = ptx::mad_hi_cc(a[0], a[0], T[1]);
#pragma unroll
    for (int i = 1; i < 8; i++) {
      T[2 * i] = ptx::madc_lo_cc(a[i], a[i], T[2 * i]);
      T[2 * i + 1] = (i < 7) ? ptx::madc_hi_cc(a[i], a[i], T[2 * i + 1]) : ptx::madc_hi(a[i], a[i], T[2 * i + 1]);
    }
    return redc_wide(T);
  }
#if defined(__CUDA_ARCH__)
  static __device__ __noinline__ Fe sqr_call(Fe a) { return sqr_inline(a); }
  HD Fe sqr() const { return sqr_call(*this); }
#else
  HD Fe sqr() const { return sqr_inline(*this); }
#endif

  // ---- fused Montgomery dot products ("lazy reduction"): sum_k a_k * b_k * R^-1 mod p with ONE reduction pass.
  // N products share the 8 reduction rows: 64 N + 72 wide MACs instead of 136 N, and the additions/subtractions a
  // Karatsuba formula would need around the products disappear (a difference a*b - c*d is dot2(a, b, p - c, d)).
  // Bounds (operands <= p): the running sum stays < sum_k a_k + p < 2^32 * 2^256 (the ninth limb `of[7]` holds it) and
  // the result before the final subtraction is < (0.19 N + 1) p < 2p for N <= 4.
  template <int N>
  HD static void dot_row(u32* al, u32* of, const Fe* a, const Fe* b, int i, bool first) {
    if (first) {
      mul_row(of, a[0].v + 1, b[0].v[i]);
      mul_row(al, a[0].v, b[0].v[i]);
    } else {
      al[0] = ptx::add_cc(al[0], of[1]);
      madc_row_rshift(of, a[0].v + 1, b[0].v[i]);
      cmad_row(al, a[0].v, b[0].v[i]);
      of[7] = ptx::addc(of[7], 0);
    }
#pragma unroll
    for (int k = 1; k < N; k++) {
      cmad_row(of, a[k].v + 1, b[k].v[i]);
      cmad_row(al, a[k].v, b[k].v[i]);
      of[7] = ptx::addc(of[7], 0);
    }
    u32 mi = al[0] * P::M0;
    cmad_row_mod(of, 1, mi);
    cmad_row_mod(al, 0, mi);
    of[7] = ptx::addc(of[7], 0);
  }
  template <int N>
  HD static Fe dot_inline(const Fe* a, const Fe* b) {
    // result before the final subtraction < (N p / 2^256 + 1) p: < 2p for N <= 4 with the 254-bit BN254 moduli (p / 2^256 = 0.19),
    // for N <= 3 with the 255-bit Pasta moduli (0.25)
    static_assert(N >= 1 && N <= (P::mod(7) < 0x40000000u ? 4 : 3), "bound of the single final subtraction");
    u32 al[8], of[8];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      dot_row<N>(al, of, a, b, i, i == 0);
      dot_row<N>(of, al, a, b, i + 1, false);
    }
    Fe r;
    r.v[0] = ptx::add_cc(al[0], of[1]);
#pragma unroll
    for (int j = 1; j < 7; j++) r.v[j] = ptx::addc_cc(al[j], of[j + 1]);
    r.v[7] = ptx::addc(al[7], 0);
    reduce_once(r.v);
    return r;
  }
#if defined(__CUDA_ARCH__)
  static __device__ __noinline__ Fe dot2_call(Fe a0, Fe b0, Fe a1, Fe b1) {
    Fe a[2] = {a0, a1}, b[2] = {b0, b1};
    return dot_inline<2>(a, b);
  }
  static __device__ __noinline__ Fe dot3_call(Fe a0, Fe b0, Fe a1, Fe b1, Fe a2, Fe b2) {
    Fe a[3] = {a0, a1, a2}, b[3] = {b0, b1, b2};
    return dot_inline<3>(a, b);
  }
  HD static Fe dot2(const Fe& a0, const Fe& b0, const Fe& a1, const Fe& b1) { return dot2_call(a0, b0, a1, b1); }
  HD static Fe dot3(const Fe& a0, const Fe& b0, const Fe& a1, const Fe& b1, const Fe& a2, const Fe& b2) {
    return dot3_call(a0, b0, a1, b1, a2, b2);
  }
#else
  HD static Fe dot2(const Fe& a0, const Fe& b0, const Fe& a1, const Fe& b1) {
    Fe a[2] = {a0, a1}, b[2] = {b0, b1};
    return dot_inline<2>(a, b);
  }
  HD static Fe dot3(const Fe& a0, const Fe& b0, const Fe& a1, const Fe& b1, const Fe& a2, const Fe& b2) {
    Fe a[3] = {a0, a1, a2}, b[3] = {b0, b1, b2};
    return dot_inline<3>(a, b);
  }
#endif
  // p - a without the conditional (a <= p; 0 maps to p, which the dot products accept as an operand)
  HD Fe neg_lazy() const {
    Fe r;
    r.v[0] = ptx::sub_cc(P::mod(0), v[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = ptx::subc_cc(P::mod(i), v[i]);
    r.v[7] = ptx::subc(P::mod(7), v[7]);
    return r;
  }

  // canonical <-> Montgomery
  HD Fe to_mont() const {
    Fe r2;
#pragma unroll
    for (int i = 0; i < 8; i++) r2.v[i] = P::r2(i);
    return (*this) * r2;
  }
  // to_mont of ANY 256-bit value (value mod p, not necessarily canonical on input): R^2 goes first because the
  // multiplier's carry bound needs its FIRST operand < p; the running sum then stays < R^2-operand + p < 2p.
  HD Fe to_mont_wide() const {
    Fe r2;
#pragma unroll
    for (int i = 0; i < 8; i++) r2.v[i] = P::r2(i);
    return r2 * (*this);
  }
  HD Fe from_mont() const {
    Fe o = zero();
    o.v[0] = 1;
    return (*this) * o;
  }
  // canonical value < p ?
  HD static bool is_canonical(const u32* a) {
    ptx::sub_cc(a[0], P::mod(0));
#pragma unroll
    for (int i = 1; i < 8; i++) ptx::subc_cc(a[i], P::mod(i));
    return ptx::subc(0, 0) != 0;
  }

  // a^e for a public 256-bit exponent (left-to-right binary; the exponent is uniform across a warp)
  // a^e for a public 256-bit exponent (uniform across a warp): sliding windows of up to 5 bits over a table of the odd powers
  // a, a^3, .., a^31 (1 squaring + 15 products; the table lives in the noinline frame): ~253 squarings + ~42 + 16 products for a
  // 254-bit exponent (fixed 4-bit windows: 252 + 77).
  HDN Fe pow(const u32* e) const {
    Fe tab[16];
    tab[0] = *this;
    Fe a2 = sqr();
#pragma unroll 1
    for (int i = 1; i < 16; i++) tab[i] = tab[i - 1] * a2;
    Fe r = one();
    bool started = false;
    int i = 255;
    while (i >= 0) {
      if (!((e[i >> 5] >> (i & 31)) & 1)) {
        if (started) r = r.sqr();
        i--;
        continue;
      }
      // window [j, i] with bit j set, at most 5 bits
      int j = i - 4 < 0 ? 0 : i - 4;
      while (!((e[j >> 5] >> (j & 31)) & 1)) j++;
      u32 val = 0;
      for (int t = i; t >= j; t--) {
        val = (val << 1) | ((e[t >> 5] >> (t & 31)) & 1);
        if (started) r = r.sqr();
      }
      r = started ? r * tab[val >> 1] : tab[val >> 1];
      started = true;
      i = j - 1;
    }
    return r;
  }
  // Inversion by the binary extended Euclidean algorithm, branch-free per step so that a warp only diverges in the trip count
  // (322..388 steps for these moduli, 358 on average): invariants x1 * t = u, x2 * t = v (mod p) for the stored integer t; v stays
  // odd, every step makes u even (swap so that u >= v and subtract when u is odd) and halves it; at u = 0, v = gcd = 1 and
  // x2 = t^-1.  About 85 ALU instructions per step and NO multiplications: ~33 k instructions on the otherwise lightly loaded ALU
  // pipe instead of a 294-product Fermat chain (40 k instructions) on the integer-multiply pipe that bounds every kernel.
  // The result is the same field element.  0 -> 0 (== the `unwrap_or_else(|| value.clone())` of loader.rs:247).
  HDN Fe inv() const {
    if (is_zero()) return zero();
    u32 u[8], w[8], x1[8], x2[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { u[i] = v[i]; w[i] = P::mod(i); x1[i] = i == 0; x2[i] = 0; }
    while (true) {
      u32 nz = 0;
#pragma unroll
      for (int i = 0; i < 8; i++) nz |= u[i];
      if (!nz) break;
      u32 odd = 0u - (u[0] & 1);  // mask
      ptx::sub_cc(u[0], w[0]);
#pragma unroll
      for (int i = 1; i < 8; i++) ptx::subc_cc(u[i], w[i]);
      u32 lt = ptx::subc(0, 0);  // 0xffffffff iff u < w
      u32 sw = odd & lt;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        u32 t = (u[i] ^ w[i]) & sw;
        u[i] ^= t; w[i] ^= t;
        t = (x1[i] ^ x2[i]) & sw;
        x1[i] ^= t; x2[i] ^= t;
      }
      // u odd: u -= w (both odd, u >= w), x1 -= x2 (mod p)
      u[0] = ptx::sub_cc(u[0], w[0] & odd);
#pragma unroll
      for (int i = 1; i < 7; i++) u[i] = ptx::subc_cc(u[i], w[i] & odd);
      u[7] = ptx::subc(u[7], w[7] & odd);
      x1[0] = ptx::sub_cc(x1[0], x2[0] & odd);
#pragma unroll
      for (int i = 1; i < 8; i++) x1[i] = ptx::subc_cc(x1[i], x2[i] & odd);
      u32 br = ptx::subc(0, 0);
      x1[0] = ptx::add_cc(x1[0], P::mod(0) & br);
#pragma unroll
      for (int i = 1; i < 7; i++) x1[i] = ptx::addc_cc(x1[i], P::mod(i) & br);
      x1[7] = ptx::addc(x1[7], P::mod(7) & br);
      // u /= 2 ; x1 /= 2 (mod p)
#pragma unroll
      for (int i = 0; i < 7; i++) u[i] = (u[i] >> 1) | (u[i + 1] << 31);
      u[7] >>= 1;
      u32 om = 0u - (x1[0] & 1);
      x1[0] = ptx::add_cc(x1[0], P::mod(0) & om);
#pragma unroll
      for (int i = 1; i < 7; i++) x1[i] = ptx::addc_cc(x1[i], P::mod(i) & om);
      x1[7] = ptx::addc(x1[7], P::mod(7) & om);  // x1 + p < 2^256
#pragma unroll
      for (int i = 0; i < 7; i++) x1[i] = (x1[i] >> 1) | (x1[i + 1] << 31);
      x1[7] >>= 1;
    }
    // x2 = (a R)^-1 = a^-1 R^-1  ->  a^-1 R = x2 * R^3 * R^-1
    Fe r2, t;
#pragma unroll
    for (int i = 0; i < 8; i++) { r2.v[i] = P::r2(i); t.v[i] = x2[i]; }
    return t * (r2 * r2);
  }
  // Fermat inversion a^(p-2) (kept as the cross-check of inv() in the tests)
  HD Fe inv_fermat() const {
    u32 e[8];
    e[0] = ptx::sub_cc(P::mod(0), 2);
#pragma unroll
    for (int i = 1; i < 8; i++) e[i] = ptx::subc_cc(P::mod(i), 0);
    return pow(e);
  }
  // sqrt for p = 3 mod 4 (Fq only): candidate a^((p+1)/4); caller checks y*y == a
  HD Fe sqrt_candidate() const {
    u32 e[8];
    // (p+1)/4
    u32 c = 1;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      u64 t = (u64)P::mod(i) + c;
      e[i] = (u32)t;
      c = (u32)(t >> 32);
    }
#pragma unroll
    for (int i = 0; i < 7; i++) e[i] = (e[i] >> 2) | (e[i + 1] << 30);
    e[7] >>= 2;
    return pow(e);
  }
};

typedef Fe<FqParams> Fq;
typedef Fe<FrParams> Fr;

// 32-byte little-endian canonical <-> limbs
HD void fe_load_le(u32* v, const uint8_t* b) {
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = (u32)b[4 * i] | ((u32)b[4 * i + 1] << 8) | ((u32)b[4 * i + 2] << 16) | ((u32)b[4 * i + 3] << 24);
}
HD void fe_store_le(uint8_t* b, const u32* v) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    b[4 * i] = (uint8_t)v[i];
    b[4 * i + 1] = (uint8_t)(v[i] >> 8);
    b[4 * i + 2] = (uint8_t)(v[i] >> 16);
    b[4 * i + 3] = (uint8_t)(v[i] >> 24);
  }
}

// Fq canonical value -> Fr (integer value mod r; `fe_to_fe`, util/arithmetic.rs:256-258).
// p < 2r so one conditional subtraction suffices.  Input/outputs canonical (non-Montgomery) limbs.
HD void fq_canon_to_fr_canon(u32* out, const u32* in) {
#pragma unroll
  for (int i = 0; i < 8; i++) out[i] = in[i];
  Fr::reduce_once(out);
}
