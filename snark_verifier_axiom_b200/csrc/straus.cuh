// Interleaved (Straus) multi-scalar multiplication with signed 5-bit fixed windows, shared by the per-proof MSM
// (k_msm_var, verify.cu) and the fold group MSM (k_group_var, fold.cu) -- both stand for the reference's
// `Σ base * scalar` of NativeLoader::multi_scalar_multiplication (snark-verifier/src/loader/native.rs:61-71); only the
// resulting group element is observable, the schedule is ours.
//
// Recoding without a carry chain at read time: k' = k + C with C = Σ_{i<51} 16 * 32^i (fits 256 bits for any k < 2^255), then
//   digit_i = ((k' >> 5 i) & 31) - 16  in [-16, 15]   (i < 51),      digit_51 = k' >> 255  in {0, 1}
// and Σ digit_i 32^i = k.  Per term: a 16-entry table {1..16} P (1 + 7 doublings, 7 mixed additions, then normalised to affine
// together with all other tables of the thread) and <= 52 MIXED table additions (a digit is zero with probability 1/32); per
// thread 255 shared doublings and one inversion.  Against the unsigned
// 4-bit windows of the first version (15-entry table, 64 additions, 252 doublings): ~10 additions fewer per term.
#pragma once
#include "g1.cuh"
#include "glv.cuh"

#define STRAUS_WINDOWS 52
#define STRAUS_TABLE 16

// k (8 limbs, canonical scalar) += C
HD void straus_recode(u32* k) {
  const u32 C[8] = {0x21084210u, 0x08421084u, 0x42108421u, 0x10842108u, 0x84210842u, 0x21084210u, 0x08421084u, 0x42108421u};
  k[0] = ptx::add_cc(k[0], C[0]);
#pragma unroll
  for (int i = 1; i < 7; i++) k[i] = ptx::addc_cc(k[i], C[i]);
  k[7] = ptx::addc(k[7], C[7]);
}

// magnitude (0..16) and sign of digit w of a recoded scalar
HD u32 straus_digit(const u32* k, int w, u32& neg) {
  u32 bit = 5u * (u32)w, word = bit >> 5, sh = bit & 31;
  u32 lo = k[word], hi = word < 7 ? k[word + 1] : 0;
  u32 v = (u32)((((u64)hi << 32) | lo) >> sh) & 31;
  neg = 0;
  if (w == STRAUS_WINDOWS - 1) return v;
  int s = (int)v - 16;
  neg = s < 0;
  return (u32)(s < 0 ? -s : s);
}

// tb[(d - 1) * stride] = d * base for d = 1..16
template <class F>
HD void straus_build_table(JacT<F>* tb, size_t stride, const AffT<F>& base) {
  JacT<F> cur = JacT<F>::from_affine(base);
  tb[0] = cur;
  for (u32 d = 2; d <= STRAUS_TABLE; d++) {
    if (d & 1) cur = cur.add_affine(base);
    else cur = tb[(size_t)(d / 2 - 1) * stride].dbl();
    tb[(size_t)(d - 1) * stride] = cur;
  }
}

// Batch normalisation of a thread's tables to affine (Montgomery's trick over all its entries: one Fermat inversion per thread):
// X, Y of every entry are overwritten by the affine coordinates ((0, 0) for an identity entry), so that the main loop uses mixed
// additions (9.9 instead of 14.8 product-equivalents each).  `prefix` holds the running products of the Z's (n_entries x stride).
// Per entry 1 + 5 products and a squaring: pays off from ~3 terms per thread on (11 terms: -11 % of the kernel).
template <class F>
HD void straus_normalize(JacT<F>* tables, F* prefix, size_t stride, u32 n_entries) {
  // the first entry of every table is the base itself (Z = 1, or Z = 0 for an identity base): nothing to multiply
  F run = F::one();
  for (u32 i = 0; i < n_entries; i++) {
    F z = tables[(size_t)i * stride].Z;
    if ((i % STRAUS_TABLE) != 0 && !z.is_zero()) run = run * z;
    prefix[(size_t)i * stride] = run;
  }
  F inv = run.inv();
  for (u32 i = n_entries; i-- > 0;) {
    JacT<F>* e = tables + (size_t)i * stride;
    F z = e->Z;
    if (z.is_zero()) {
      e->X = F::zero();
      e->Y = F::zero();
      continue;
    }
    if ((i % STRAUS_TABLE) == 0) continue;  // already affine
    F zi = i ? inv * prefix[(size_t)(i - 1) * stride] : inv;
    inv = inv * z;
    F zi2 = zi.sqr();
    e->X = e->X * zi2;
    e->Y = e->Y * (zi2 * zi);
  }
}

// acc = Σ_t k[t] * base_t over `nt` recoded scalars k[t] (8 limbs each, row stride 8) whose NORMALISED tables start at
// tables[(t * 16) * stride]; entries are fetched one step ahead of their use (the tables of a launch are hundreds of MB and
// every read misses L2; a table addition is long enough to cover a DRAM round trip).
// AFFINE = false: the tables are still Jacobian (threads with one or two terms, where the inversion would not pay): full additions.
template <bool AFFINE, class F>
HD JacT<F> straus_run(const u32* k, u32 nt, const JacT<F>* tables, size_t stride) {
  typedef JacT<F> J;
  J acc = J::identity();
  if (nt == 0) return acc;
  J nxt = J::identity();
  u32 nn = 0;
  u32 dn = straus_digit(k, STRAUS_WINDOWS - 1, nn);
  if (dn) {
    const J* e = tables + (size_t)(dn - 1) * stride;
    nxt.X = e->X;
    nxt.Y = e->Y;
    if (!AFFINE) nxt.Z = e->Z;
  }
  for (int w = STRAUS_WINDOWS - 1; w >= 0; w--) {
    if (w != STRAUS_WINDOWS - 1) acc = acc.dbl().dbl().dbl().dbl().dbl();
    for (u32 t = 0; t < nt; t++) {
      J cur = nxt;
      u32 d = dn, ng = nn;
      u32 t2 = t + 1;
      int w2 = w;
      if (t2 == nt) { t2 = 0; w2 = w - 1; }
      dn = 0;
      if (w2 >= 0) {
        dn = straus_digit(k + 8 * t2, w2, nn);
        if (dn) {
          const J* e = tables + ((size_t)t2 * STRAUS_TABLE + (dn - 1)) * stride;
          nxt.X = e->X;
          nxt.Y = e->Y;
          if (!AFFINE) nxt.Z = e->Z;
        }
      }
      if (d) {
        if (ng) cur.Y = cur.Y.neg();
        if (AFFINE) acc = acc.add_affine(AffT<F>{cur.X, cur.Y});
        else acc = acc.add(cur);
      }
    }
  }
  return acc;
}


// ---- one term per thread, GLV form (latency schedule of the per-proof MSM: a lone 255-doubling chain is most of k_msm_var's time)
// k P = (+-k1) P + (+-k2) phi(P) with 128-bit halves (glv.cuh): 25 signed 5-bit windows + an unsigned top window per half, ONE
// 16-entry Jacobian table of P (phi(T) = (beta X, Y, Z) in Jacobian coordinates too), 125 doublings instead of 255.
#define STRAUS_GLV_WINDOWS 26
// k (4 limbs, < 2^128) += sum_{i<25} 16 * 32^i ; digit_i = ((k' >> 5 i) & 31) - 16 for i < 25, digit_25 = k' >> 125 (0..9)
HD void straus_glv_recode(const u32* k, u32* out5) {
  const u32 C[4] = {0x21084210u, 0x08421084u, 0x42108421u, 0x10842108u};  // sum_{i<25} 16 * 32^i  (bits 4, 9, .., 124)
  out5[0] = ptx::add_cc(k[0], C[0]);
  out5[1] = ptx::addc_cc(k[1], C[1]);
  out5[2] = ptx::addc_cc(k[2], C[2]);
  out5[3] = ptx::addc_cc(k[3], C[3] & 0x1fffffffu);
  out5[4] = ptx::addc(0, 0);
}
HD u32 straus_glv_digit(const u32* k5, int w, u32& neg) {
  u32 bit = 5u * (u32)w, word = bit >> 5, sh = bit & 31;
  u64 two = (u64)k5[word] | ((u64)(word < 4 ? k5[word + 1] : 0) << 32);
  neg = 0;
  if (w == STRAUS_GLV_WINDOWS - 1) return (u32)(two >> sh);  // k' >> 125: everything that is left
  int s = (int)((two >> sh) & 31) - 16;
  neg = s < 0;
  return (u32)(s < 0 ? -s : s);
}
template <class F>
HD JacT<F> straus_run_glv1(const u32* k1, u32 neg1, const u32* k2, u32 neg2, const JacT<F>* table, size_t stride, const F& beta) {
  typedef JacT<F> J;
  u32 r1[5], r2[5];
  straus_glv_recode(k1, r1);
  straus_glv_recode(k2, r2);
  J acc = J::identity();
  for (int w = STRAUS_GLV_WINDOWS - 1; w >= 0; w--) {
    if (w != STRAUS_GLV_WINDOWS - 1) acc = acc.dbl().dbl().dbl().dbl().dbl();
    u32 n1, n2;
    u32 d1 = straus_glv_digit(r1, w, n1), d2 = straus_glv_digit(r2, w, n2);
    if (d1) {
      J e = table[(size_t)(d1 - 1) * stride];
      if (n1 ^ neg1) e.Y = e.Y.neg();
      acc = acc.add(e);
    }
    if (d2) {
      J e = table[(size_t)(d2 - 1) * stride];
      e.X = e.X * beta;  // phi
      if (n2 ^ neg2) e.Y = e.Y.neg();
      acc = acc.add(e);
    }
  }
  return acc;
}
