// K5/K6, latency form: ONE accumulator decided by a whole thread block.
//
// `KzgAs::decide` (snark-verifier/src/pcs/kzg/decider.rs:60-68) ends every batch with a SINGLE pairing check; run by one
// thread (decide.cu: k_decide) it is a chain of ~16 k dependent Montgomery products = 15 ms, a third of a 4096-proof
// batch's latency.  Here the Fq12 arithmetic is spread over the lanes of a block instead: an Fq12 product is 144 Fq
// products; 48 lanes each take a fused 3-term dot product (field.cuh `dot3`), 12 lanes add the four partial sums of an
// output coefficient.  Dependent chain per Fq12 product: one dot3 + ~60 add-type instructions instead of 54 products.
//
// Representation: Fq12 over the basis { w^k u^c : k = 0..5, c = 0..1 } (w^6 = xi = 9 + u, u^2 = -1; the tower of
// tower.cuh has c0 = (w^0, w^2, w^4), c1 = (w^1, w^3, w^5)) as an "extended block" of 24 Fq values
//     blk[2k + c]      = coefficient of w^k u^c
//     blk[12 + 2k + c] = the same coefficient of xi * (f_k)       (f_k in Fq2: the wrap-around factor of w^6)
// so that every output coefficient is a plain signed sum of products blk_a[i] * blk_b[j]:
//     (a b)_k = sum_{i <= k} a_i b_{k-i} + sum_{i > k} a_i (xi b_{k-i+6}).
// Only the accept bit is observable (decider.rs:66), the schedule is ours; values are those of pairing.cuh
// (tests/test_host_arith.py::test_coop_pairing_matches_tower runs this very code lane by lane on the host).
//
// The program is written once against an executor `Ex` with `par(n_tasks, f)`: run f(task) for every task, then
// barrier.  Device: tasks strided over the block's threads + __syncthreads (decide.cu).  Host: a plain loop.
#pragma once
#include "pairing.cuh"

struct G2LineX {
  Fq2 neg_lam, c3;        // as G2Line
  Fq2 xi_neg_lam, xi_c3;  // xi * neg_lam, xi * c3 (host-precomputed per deciding key)
};

#define COOP_BLK 24          // Fq values per extended block
#define COOP_DOT_LANES 48    // dot3 tasks per Fq12 product
#define COOP_N_TMP 12        // temporaries of the final exponentiation

struct CoopMem {
  Fq f[COOP_BLK];                    // Miller accumulator / running value
  Fq tmp[COOP_N_TMP][COOP_BLK];
  Fq part[COOP_DOT_LANES];
  int flag;
};

HD Fq fq_mul9(const Fq& x) { return x.dbl().dbl().dbl() + x; }

// 9 x + y (minus = false) or 9 x - y (minus = true), both operands reduced: ONE integer pass instead of five modular additions.
// t = 9 x + w with w = y or p - y, t < 10 p < 2^258 (nine limbs); q = floor(t / p) is estimated from the top 34 bits against the
// top 30 bits of p (never too large, at most one too small), t - q p < 2p, one conditional subtraction.
HD Fq fq_mul9_addsub(const Fq& x, const Fq& y, bool minus) {
  u32 t[9];
  u64 c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    u32 w = minus ? 0 : y.v[i];
    u64 v = (u64)x.v[i] * 9 + w + c;
    t[i] = (u32)v;
    c = v >> 32;
  }
  t[8] = (u32)c;
  if (minus) {  // + (p - y)
    u64 br = 0, cy = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      u64 d = (u64)FqParams::mod(i) - y.v[i] - br;
      br = (d >> 32) & 1;
      u64 v = (u64)t[i] + (u32)d + cy;
      t[i] = (u32)v;
      cy = v >> 32;
    }
    t[8] += (u32)cy;
  }
  u64 hi = ((u64)t[8] << 32) | t[7];
  const u64 ptop = (u64)FqParams::mod(7) + 1;
  u32 q = 0;
#pragma unroll
  for (int b = 3; b >= 0; b--) {
    u64 m = ptop << b;
    if (hi >= m) { hi -= m; q |= 1u << b; }
  }
  // r = t - q p  (fits eight limbs: < 2p)
  u32 r[8];
  u64 mc = 0, br = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    u64 m = (u64)FqParams::mod(i) * q + mc;
    mc = m >> 32;
    u64 d = (u64)t[i] - (u32)m - br;
    r[i] = (u32)d;
    br = (d >> 32) & 1;
  }
  Fq::reduce_once(r);
  Fq out;
#pragma unroll
  for (int i = 0; i < 8; i++) out.v[i] = r[i];
  return out;
}

// xi-half of a block from its value half: xi (a + b u) = (9a - b) + (9b + a) u.  task t in [0, 12)
HD void coop_xi_task(Fq* blk, int t) {
  int k = t >> 1, c = t & 1;
  const Fq &a = blk[2 * k], &b = blk[2 * k + 1];
  blk[12 + t] = c == 0 ? fq_mul9_addsub(a, b, true) : fq_mul9_addsub(b, a, false);
}

// task t in [0, 48): partial dot product number (t & 3) of output coefficient (t >> 2)
HD void coop_dot_task(Fq* part, const Fq* A, const Fq* B, int t) {
  int o = t >> 2, s = t & 3, k = o >> 1, c = o & 1;
  Fq a[3], b[3];
#pragma unroll
  for (int e = 0; e < 3; e++) {
    int term = 3 * s + e, i = term >> 1, which = term & 1;
    int j = k - i;
    const Fq* bj = B + 2 * j;
    if (j < 0) bj = B + 12 + 2 * (j + 6);
    a[e] = A[2 * i + which];
    b[e] = bj[c == 0 ? which : 1 - which];
    if (c == 0 && which == 1) a[e] = a[e].neg_lazy();  // - a_i1 b_j1
  }
  part[t] = Fq::dot3(a[0], b[0], a[1], b[1], a[2], b[2]);
}

// sum of the four partials of output coefficient o (each < p, so the integer sum is < 4p < 2^256), reduced
HD Fq coop_sum4(const Fq* part, int o) {
  const Fq* p = part + 4 * o;
  u32 s[8];
  s[0] = ptx::add_cc(p[0].v[0], p[1].v[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) s[i] = ptx::addc_cc(p[0].v[i], p[1].v[i]);
  s[7] = ptx::addc(p[0].v[7], p[1].v[7]);
#pragma unroll
  for (int q = 2; q < 4; q++) {
    s[0] = ptx::add_cc(s[0], p[q].v[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) s[i] = ptx::addc_cc(s[i], p[q].v[i]);
    s[7] = ptx::addc(s[7], p[q].v[7]);
  }
  // < 4p: subtract 2p if possible, then p
  u32 d[8];
  d[0] = ptx::sub_cc(s[0], FqParams::mod(0) << 1);
#pragma unroll
  for (int i = 1; i < 8; i++) d[i] = ptx::subc_cc(s[i], (FqParams::mod(i) << 1) | (FqParams::mod(i - 1) >> 31));
  u32 br = ptx::subc(0, 0);
#pragma unroll
  for (int i = 0; i < 8; i++) s[i] = br ? s[i] : d[i];
  Fq::reduce_once(s);
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = s[i];
  return r;
}

// task t in [0, 12): output coefficient t and its xi-image (needs the partner component: summed here as well)
// `warp12`: the twelve tasks run on lanes 0..11 of one warp (the device executor): the partner's sum comes by shuffle.
HD void coop_combine_task(Fq* out, const Fq* part, int t, bool warp12 = false) {
  int c = t & 1;
  Fq mine = coop_sum4(part, t), other;
#if defined(__CUDA_ARCH__)
  if (warp12) {
#pragma unroll
    for (int i = 0; i < 8; i++) other.v[i] = __shfl_xor_sync(0xfffu, mine.v[i], 1);
  } else
#endif
    other = coop_sum4(part, t ^ 1);
  (void)warp12;
  out[t] = mine;
  out[12 + t] = fq_mul9_addsub(mine, other, c == 0);
}

// out = a * b   (blocks; out may alias a and/or b).  The block-level operations are out-of-line on the device: the pairing program
// has ~50 call sites and one warp runs it, so inlining them (0.3 MB of code) would live in instruction-cache misses.
template <class Ex>
HDN void coop_mul(Ex& ex, Fq* out, const Fq* a, const Fq* b, Fq* part) {
  ex.par(COOP_DOT_LANES, [&](int t) { coop_dot_task(part, a, b, t); });
  ex.par(12, [&](int t) { coop_combine_task(out, part, t, Ex::kWarp12); });
}

template <class Ex>
HDN void coop_copy(Ex& ex, Fq* out, const Fq* a) {
  ex.par(COOP_BLK, [&](int t) { out[t] = a[t]; });
}

// conjugation over Fq6: the odd powers of w change sign (both halves of the block)
template <class Ex>
HDN void coop_conj(Ex& ex, Fq* out, const Fq* a) {
  ex.par(COOP_BLK, [&](int t) {
    int k = (t % 12) >> 1;
    out[t] = (k & 1) ? a[t].neg() : a[t];
  });
}

// Frobenius^j (j = 1, 2, 3): coefficient k -> conj^j(f_k) * gamma_{j,k}   (pairing.cuh fq12_frob1/2/3)
template <class Ex>
HDN void coop_frob(Ex& ex, Fq* out, const Fq* a, int j, const PairingConsts& K, Fq* scratch12) {
  ex.par(12, [&](int t) {
    int k = t >> 1, c = t & 1;
    Fq a0 = a[2 * k], a1 = a[2 * k + 1];
    if (j == 2) {
      Fq v = c ? a1 : a0;
      scratch12[t] = k == 0 ? v : v * K.g2[k - 1];
      return;
    }
    if (k == 0) {
      scratch12[t] = c ? a1.neg() : a0;
      return;
    }
    const Fq2& g = j == 1 ? K.g1[k - 1] : K.g3[k - 1];
    // (a0 - a1 u)(g0 + g1 u) = (a0 g0 + a1 g1) + (a0 g1 - a1 g0) u
    scratch12[t] = c == 0 ? Fq::dot2(a0, g.c0, a1, g.c1) : Fq::dot2(a0, g.c1, a1.neg_lazy(), g.c0);
  });
  ex.par(12, [&](int t) { out[t] = scratch12[t]; });
  ex.par(12, [&](int t) { coop_xi_task(out, t); });
}

HD Fq12 coop_to_tower(const Fq* blk) {
  Fq12 f;
  f.c0.c0 = {blk[0], blk[1]};
  f.c1.c0 = {blk[2], blk[3]};
  f.c0.c1 = {blk[4], blk[5]};
  f.c1.c1 = {blk[6], blk[7]};
  f.c0.c2 = {blk[8], blk[9]};
  f.c1.c2 = {blk[10], blk[11]};
  return f;
}
HD void coop_from_tower(Fq* blk, const Fq12& f) {
  blk[0] = f.c0.c0.c0; blk[1] = f.c0.c0.c1;
  blk[2] = f.c1.c0.c0; blk[3] = f.c1.c0.c1;
  blk[4] = f.c0.c1.c0; blk[5] = f.c0.c1.c1;
  blk[6] = f.c1.c1.c0; blk[7] = f.c1.c1.c1;
  blk[8] = f.c0.c2.c0; blk[9] = f.c0.c2.c1;
  blk[10] = f.c1.c2.c0; blk[11] = f.c1.c2.c1;
}

// inversion: once per pairing, on one lane through the tower (tower.cuh)
template <class Ex>
HDN void coop_inv(Ex& ex, Fq* out, const Fq* a) {
  ex.par(1, [&](int) { coop_from_tower(out, coop_to_tower(a).inv()); });
  ex.par(12, [&](int t) { coop_xi_task(out, t); });
}

// r = x^X for the BN parameter X (pairing.cuh fq12_pow_x); x in the cyclotomic subgroup.  r must not alias x.
template <class Ex>
HDN void coop_pow_x(Ex& ex, Fq* r, const Fq* x, Fq* part) {
  coop_copy(ex, r, x);
  for (int i = 61; i >= 0; i--) {
    coop_mul(ex, r, r, r, part);
    int bit = (i >= 32) ? ((SVK_BN_X_HI >> (i - 32)) & 1) : ((SVK_BN_X_LO >> i) & 1);
    if (bit) coop_mul(ex, r, r, x, part);
  }
}

// Line blocks: lb[(pair * SVK_N_LINES + line) * 24 ..] = the line of `pair` at step `line`, evaluated at the pair's G1
// point: l = yP + (neg_lam xP) w + c3 w^3  (pairing.cuh), as an extended block.  An identity G1 point contributes 1.
// task idx in [0, 2 * SVK_N_LINES * 24)
HD void coop_line_task(Fq* lb, const G1Affine& p1, const G2LineX* t1, const G1Affine& p2, const G2LineX* t2, int idx) {
  int e = idx % COOP_BLK, li = idx / COOP_BLK;
  int pair = li / SVK_N_LINES, l = li % SVK_N_LINES;
  const G1Affine& P = pair ? p2 : p1;
  const G2LineX& L = (pair ? t2 : t1)[l];
  Fq v = Fq::zero();
  if (P.is_identity()) {
    if (e == 0 || e == 12) v = Fq::one();
    if (e == 12) v = fq_mul9(v);
    if (e == 13) v = Fq::one();
  } else {
    switch (e) {
      case 0: v = P.y; break;
      case 2: v = L.neg_lam.c0 * P.x; break;
      case 3: v = L.neg_lam.c1 * P.x; break;
      case 6: v = L.c3.c0; break;
      case 7: v = L.c3.c1; break;
      case 12: v = fq_mul9(P.y); break;  // xi * (yP + 0 u) = 9 yP + yP u   (never read by coop_dot_task: j = 0 does not wrap)
      case 13: v = P.y; break;
      case 14: v = L.xi_neg_lam.c0 * P.x; break;
      case 15: v = L.xi_neg_lam.c1 * P.x; break;
      case 18: v = L.xi_c3.c0; break;
      case 19: v = L.xi_c3.c1; break;
      default: break;
    }
  }
  lb[idx] = v;
}

// Scratch per accumulator (Fq values): the 2 x SVK_N_LINES line blocks, the partial dot products of their pairwise products and the
// SVK_N_LINES product blocks.
#define COOP_SCRATCH_FQ (2 * SVK_N_LINES * COOP_BLK + SVK_N_LINES * COOP_DOT_LANES + SVK_N_LINES * COOP_BLK)

// KzgAs::decide for one accumulator (decider.rs:60-68): e(lhs, g2) e(rhs, -s_g2) == 1.
// `lb`: COOP_SCRATCH_FQ values of scratch (global memory on the device), `m`: block-shared memory.
// The two lines of a Miller step (one per pair) are multiplied with each other FIRST -- all SVK_N_LINES products are independent
// of f and of each other, so every thread of the block works on them -- and f then takes ONE product per step instead of two.
template <class Ex>
HD bool coop_kzg_decide(Ex& ex, const G1Affine& lhs, const G1Affine& rhs, const G2LineX* t_g2, const G2LineX* t_neg_sg2,
                        const PairingConsts& K, Fq* lb, CoopMem& m, Fq* out_ml = nullptr, Fq* out_gt = nullptr) {
  ex.par(2 * SVK_N_LINES * COOP_BLK, [&](int idx) { coop_line_task(lb, lhs, t_g2, rhs, t_neg_sg2, idx); });
  Fq* f = m.f;
  Fq* part = m.part;
  ex.par(COOP_BLK, [&](int t) { f[t] = (t == 0 || t == 13) ? Fq::one() : (t == 12 ? fq_mul9(Fq::one()) : Fq::zero()); });
  const Fq* l1 = lb;
  const Fq* l2 = lb + SVK_N_LINES * COOP_BLK;
  Fq* lpart = lb + 2 * SVK_N_LINES * COOP_BLK;
  Fq* lprod = lpart + SVK_N_LINES * COOP_DOT_LANES;
  ex.par(SVK_N_LINES * COOP_DOT_LANES, [&](int idx) {
    int l = idx / COOP_DOT_LANES, t = idx % COOP_DOT_LANES;
    coop_dot_task(lpart + l * COOP_DOT_LANES, l1 + l * COOP_BLK, l2 + l * COOP_BLK, t);
  });
  ex.par(SVK_N_LINES * 12, [&](int idx) {
    int l = idx / 12, t = idx % 12;
    coop_combine_task(lprod + l * COOP_BLK, lpart + l * COOP_DOT_LANES, t);
  });
  int li = 0;
  for (int i = 63; i >= 0; i--) {
    coop_mul(ex, f, f, f, part);
    coop_mul(ex, f, f, lprod + li * COOP_BLK, part);
    li++;
    if (ate_bit(i)) {
      coop_mul(ex, f, f, lprod + li * COOP_BLK, part);
      li++;
    }
  }
  for (int s = 0; s < 2; s++) {
    coop_mul(ex, f, f, lprod + li * COOP_BLK, part);
    li++;
  }
  if (out_ml) ex.par(12, [&](int t) { out_ml[t] = f[t]; });  // test hook: the Miller value
  // ---- final exponentiation (pairing.cuh final_exponentiation, same addition chain)
  Fq *t0 = m.tmp[0], *t1 = m.tmp[1], *F = m.tmp[2], *fx = m.tmp[3], *fx2 = m.tmp[4], *fx3 = m.tmp[5];
  Fq *A = m.tmp[6], *B = m.tmp[7], *C = m.tmp[8], *D = m.tmp[9], *E = m.tmp[10], *G = m.tmp[11];
  Fq* sc = m.part;  // 12-entry scratch of coop_frob (part is free between products)
  coop_conj(ex, t0, f);
  coop_inv(ex, t1, f);
  coop_mul(ex, t0, t0, t1, part);            // f^(p^6 - 1)
  coop_frob(ex, t1, t0, 2, K, sc);
  coop_mul(ex, F, t1, t0, part);             // ^(p^2 + 1): cyclotomic subgroup, inverse == conj
  coop_pow_x(ex, fx, F, part);
  coop_pow_x(ex, fx2, fx, part);
  coop_pow_x(ex, fx3, fx2, part);
  // a6 = fx2^6, a12, a18, a30
  coop_mul(ex, A, fx2, fx2, part);           // fx2^2
  coop_mul(ex, B, A, A, part);               // fx2^4
  coop_mul(ex, A, B, A, part);               // a6 = fx2^6
  coop_mul(ex, B, A, A, part);               // a12
  coop_mul(ex, C, B, A, part);               // a18
  coop_mul(ex, D, C, B, part);               // a30
  // e2 = a6 * f
  coop_mul(ex, A, A, F, part);               // A = e2 = f^(6x^2+1)
  // b6 = fx^6, b12, b18
  coop_mul(ex, E, fx, fx, part);             // fx^2
  coop_mul(ex, G, E, E, part);               // fx^4
  coop_mul(ex, E, G, E, part);               // b6
  coop_mul(ex, G, E, E, part);               // b12
  coop_mul(ex, E, G, E, part);               // b18
  // c36 = fx3^36
  coop_mul(ex, t0, fx3, fx3, part);          // c2
  coop_mul(ex, t0, t0, t0, part);            // c4
  coop_mul(ex, t0, t0, t0, part);            // c8
  coop_mul(ex, t0, t0, fx3, part);           // c9
  coop_mul(ex, t0, t0, t0, part);            // c18
  coop_mul(ex, t0, t0, t0, part);            // c36
  // e1 = conj(c36 * a18 * b12) * f
  coop_mul(ex, t1, t0, C, part);
  coop_mul(ex, t1, t1, G, part);
  coop_conj(ex, t1, t1);
  coop_mul(ex, t1, t1, F, part);             // t1 = e1
  // e0 = conj(c36 * a30 * b18 * f^2)
  coop_mul(ex, t0, t0, D, part);
  coop_mul(ex, t0, t0, E, part);
  coop_mul(ex, B, F, F, part);
  coop_mul(ex, t0, t0, B, part);
  coop_conj(ex, t0, t0);                     // t0 = e0
  // frob3(f) * frob2(e2) * frob1(e1) * e0
  coop_frob(ex, B, F, 3, K, sc);
  coop_frob(ex, C, A, 2, K, sc);
  coop_mul(ex, B, B, C, part);
  coop_frob(ex, C, t1, 1, K, sc);
  coop_mul(ex, B, B, C, part);
  coop_mul(ex, B, B, t0, part);
  if (out_gt) ex.par(12, [&](int t) { out_gt[t] = B[t]; });
  ex.par(1, [&](int) {
    bool one = B[0] == Fq::one();
    for (int i = 1; i < 12; i++) one = one && B[i].is_zero();
    m.flag = one ? 1 : 0;
  });
  return m.flag != 0;
}
