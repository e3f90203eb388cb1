// Poseidon (T = 3, RATE = 2, alpha = 5) over BN254 Fr: the optimised permutation of
// snark-verifier/src/util/hash/poseidon.rs:469-501 and the sponge of :455-467.
//
// `PoseidonConsts` is the device image of `OptimizedPoseidonSpec` (poseidon.rs:59-95), built on the
// host by poseidon_host.h (Grain LFSR + sparse-MDS factorisation) for the SDK parameters
// R_F = 8, R_P = 57 (snark-verifier-sdk/src/halo2.rs:52-56).  All values in Montgomery form.
// Cost: 8 full rounds x (3 x 3 + 9) + 57 partial rounds x (3 + 5) = 600 Fr multiplications in the textbook count; every
// MDS row is ONE fused 3-term dot product (field.cuh `dot3`: 3 limb products, 1 reduction, no additions), which is
// value-identical; with the partial rounds in scaled form and taken two at a time (three dot2 + one dot3 per pair) the permutation
// costs ~450 multiplication-equivalents of integer-pipe work.
#pragma once
#include "field.cuh"

#define SVK_POSEIDON_T 3
#define SVK_POSEIDON_RATE 2
#define SVK_POSEIDON_RF 8
#define SVK_POSEIDON_RP 57

struct PoseidonConsts {
  Fr start[SVK_POSEIDON_RF / 2 + 1][3];   // constants.start: [0] pre-constants, [1..3] full rounds, [4] last
  Fr partial[SVK_POSEIDON_RP];            // constants.partial
  Fr end[SVK_POSEIDON_RF / 2 - 1][3];     // constants.end
  Fr mds[3][3];
  Fr pre_sparse_mds[3][3];
  Fr sparse_row[SVK_POSEIDON_RP][3];      // SparseMDSMatrix.row
  Fr sparse_col_hat[SVK_POSEIDON_RP][2];  // SparseMDSMatrix.col_hat
  Fr capacity;                            // 2^64  (State::default, poseidon.rs:335-342)
  // Warp-cooperative schedule (poseidon_coop.cuh): the round constant of a partial round is folded into the helpers' terms,
  //   coop_rc[r] = row_r[0] * partial[r],   coop_cc[r][w] = col_hat_r[w] * partial[r]          (derived, value-identical)
  Fr coop_rc[SVK_POSEIDON_RP];
  Fr coop_cc[SVK_POSEIDON_RP][2];
  // Scaled partial rounds: x -> x^5 commutes with scaling ((sigma t)^5 = sigma^5 t^5), so the product row_r[0] * x of the s0 chain is
  // pushed into the constants: s0 = sigma_r t_r, v_r = t_r^5, t_{r+1} = v_r + Q_r with
  //   sigma_{r+1} = row_r[0] sigma_r^5,   Q_r = (row_r[0] c_r + row_r[1] s1 + row_r[2] s2) / sigma_{r+1},   s_w += col_r[w] sigma_r^5 v_r + col_r[w] c_r
  //   sc_a[r][w] = col_hat_r[w] sigma_r^5,  sc_r[r][w] = row_r[w+1] / sigma_{r+1},  sc_k[r] = row_r[0] c_r / sigma_{r+1},  sc_end = sigma_57
  // The s0 chain of a partial round is then 3 products + 1 addition instead of 4 + 1 (value-identical: exact field arithmetic).
  Fr sc_a[SVK_POSEIDON_RP][2];
  Fr sc_r[SVK_POSEIDON_RP][2];
  Fr sc_k[SVK_POSEIDON_RP];
  Fr sc_end;
  // the same two rounds at a time for the one-thread permutation (pair p = rounds a = 2p, b = 2p + 1; derived, value-identical):
  //   [0..2]  Q_a = [0] s1 + [1] s2 + [2]                        t_b = v_a + Q_a
  //   [3..6]  Q_b = [3] v_a + [4] s1 + [5] s2 + [6]              t'  = v_b + Q_b      ([3] = r_b0 a_a0 + r_b1 a_a1)
  //   [7..9]  s1' = s1 + [7] v_a + [8] v_b + [9]                 [10..12] the same for s2
  Fr sc_pair[SVK_POSEIDON_RP / 2][13];
};

struct PoseidonState {
  Fr s[3];
};

HD Fr fr_pow5(const Fr& x) {
  Fr x2 = x.sqr();
  Fr x4 = x2.sqr();
  return x4 * x;
}

HD void poseidon_mds(PoseidonState& st, const Fr (*m)[3]) {
  Fr r0 = Fr::dot3(m[0][0], st.s[0], m[0][1], st.s[1], m[0][2], st.s[2]);
  Fr r1 = Fr::dot3(m[1][0], st.s[0], m[1][1], st.s[1], m[1][2], st.s[2]);
  Fr r2 = Fr::dot3(m[2][0], st.s[0], m[2][1], st.s[1], m[2][2], st.s[2]);
  st.s[0] = r0;
  st.s[1] = r1;
  st.s[2] = r2;
}

HD void poseidon_init(PoseidonState& st, const PoseidonConsts& k) {
  st.s[0] = k.capacity;
  st.s[1] = Fr::zero();
  st.s[2] = Fr::zero();
}

// `Poseidon::permutation(inputs)` with n_in = 0, 1 or 2 inputs (poseidon.rs:469-501);
// absorb_with_pre_constants incl. the "+1" padding on the first unused slot (:362-384).
HDN void poseidon_permute(PoseidonState& st, const PoseidonConsts& k, int n_in, const Fr& in0, const Fr& in1) {
  st.s[0] = st.s[0] + k.start[0][0];
  if (n_in >= 1) st.s[1] = st.s[1] + in0 + k.start[0][1];
  else st.s[1] = st.s[1] + k.start[0][1] + Fr::one();
  if (n_in >= 2) st.s[2] = st.s[2] + in1 + k.start[0][2];
  else if (n_in == 1) st.s[2] = st.s[2] + k.start[0][2] + Fr::one();
  else st.s[2] = st.s[2] + k.start[0][2];
  // first half of the full rounds
  for (int r = 1; r < SVK_POSEIDON_RF / 2; r++) {
    for (int i = 0; i < 3; i++) st.s[i] = fr_pow5(st.s[i]) + k.start[r][i];
    poseidon_mds(st, k.mds);
  }
  for (int i = 0; i < 3; i++) st.s[i] = fr_pow5(st.s[i]) + k.start[SVK_POSEIDON_RF / 2][i];
  poseidon_mds(st, k.pre_sparse_mds);
  // partial rounds with sparse MDS (poseidon.rs:398-410), in the SCALED form (sc_* above): t with s0 = sigma_r t, so that the
  // product row_r[0] * x of the reference's round disappears into the constants; two rounds at a time:
  //   v_a = t^5, t_b = v_a + Q_a (one dot2), v_b = t_b^5, t' = v_b + Q_b (one dot3), s_w' (one dot2 each)
  // = 2 x 381 + 200 + 264 + 2 x 200 wide MACs per pair against 2 x 381 + 264 + 328 + 2 x 200 of the unscaled pairing.
  Fr t = st.s[0];
  int r = 0;
  for (; r + 1 < SVK_POSEIDON_RP; r += 2) {
    const Fr* c = k.sc_pair[r >> 1];
    Fr va = fr_pow5(t);
    Fr tb = va + (Fr::dot2(c[0], st.s[1], c[1], st.s[2]) + c[2]);
    Fr vb = fr_pow5(tb);
    Fr qb = Fr::dot3(c[3], va, c[4], st.s[1], c[5], st.s[2]) + c[6];
    Fr n1 = Fr::dot2(c[7], va, c[8], vb) + (st.s[1] + c[9]);
    Fr n2 = Fr::dot2(c[10], va, c[11], vb) + (st.s[2] + c[12]);
    t = vb + qb;
    st.s[1] = n1;
    st.s[2] = n2;
  }
  for (; r < SVK_POSEIDON_RP; r++) {
    Fr v = fr_pow5(t);
    Fr q = Fr::dot2(k.sc_r[r][0], st.s[1], k.sc_r[r][1], st.s[2]) + k.sc_k[r];
    st.s[1] = k.sc_a[r][0] * v + (st.s[1] + k.coop_cc[r][0]);
    st.s[2] = k.sc_a[r][1] * v + (st.s[2] + k.coop_cc[r][1]);
    t = v + q;
  }
  st.s[0] = k.sc_end * t;
  // second half of the full rounds
  for (int r = 0; r < SVK_POSEIDON_RF / 2 - 1; r++) {
    for (int i = 0; i < 3; i++) st.s[i] = fr_pow5(st.s[i]) + k.end[r][i];
    poseidon_mds(st, k.mds);
  }
  for (int i = 0; i < 3; i++) st.s[i] = fr_pow5(st.s[i]);
  poseidon_mds(st, k.mds);
}
