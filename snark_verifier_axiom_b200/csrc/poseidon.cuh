// Poseidon (T = 3, RATE = 2, alpha = 5) over BN254 Fr: the optimised permutation of
// snark-verifier/src/util/hash/poseidon.rs:469-501 and the sponge of :455-467.
//
// `PoseidonConsts` is the device image of `OptimizedPoseidonSpec` (poseidon.rs:59-95), built on the
// host by poseidon_host.h (Grain LFSR + sparse-MDS factorisation) for the SDK parameters
// R_F = 8, R_P = 57 (snark-verifier-sdk/src/halo2.rs:52-56).  All values in Montgomery form.
// Cost: 8 full rounds x (3 x 3 + 9) + 57 partial rounds x (3 + 5) = 600 Fr multiplications in the textbook count; every
// MDS row is ONE fused 3-term dot product (field.cuh `dot3`: 3 limb products, 1 reduction, no additions), which is
// value-identical; with the partial rounds taken two at a time (one dot4 + two dot2 per pair) the permutation costs ~475
// multiplication-equivalents of integer-pipe work.
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
  // Two partial rounds at a time (derived from sparse_row / sparse_col_hat, value-identical): pair p = rounds (2p, 2p+1),
  //   [0..3] = row_{2p+1}[0], row_{2p+1}[1] col_{2p}[0] + row_{2p+1}[2] col_{2p}[1], row_{2p+1}[1], row_{2p+1}[2]
  //   [4..5] = col_{2p+1}[0], col_{2p}[0]      [6..7] = col_{2p+1}[1], col_{2p}[1]
  Fr pair[SVK_POSEIDON_RP / 2][8];
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

// s0 after the second round of a pair: one 4-term dot product over (y_b, y_a, s1, s2); c = pair[p]
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ Fr poseidon_pair_s0(const Fr* c, Fr yb, Fr ya, Fr s1, Fr s2) {
#else
inline Fr poseidon_pair_s0(const Fr* c, Fr yb, Fr ya, Fr s1, Fr s2) {
#endif
  Fr a[4] = {c[0], c[1], c[2], c[3]}, b[4] = {yb, ya, s1, s2};
  return Fr::dot_inline<4>(a, b);
}
// s_i after the pair: col_{r+1} y_b + col_r y_a + s_i; c = pair[p] + 4 or + 6
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ Fr poseidon_pair_si(const Fr* c, Fr yb, Fr ya, Fr si) {
#else
inline Fr poseidon_pair_si(const Fr* c, Fr yb, Fr ya, Fr si) {
#endif
  Fr a[2] = {c[0], c[1]}, b[2] = {yb, ya};
  return Fr::dot_inline<2>(a, b) + si;
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
  // partial rounds with sparse MDS (poseidon.rs:398-410)
  // Rounds are taken two at a time: with y_a, y_b the S-box outputs of the pair,
  //   s0'' = row_b[0] y_b + (row_b[1] col_a[0] + row_b[2] col_a[1]) y_a + row_b[1] s1 + row_b[2] s2     (one dot4)
  //   s_i'' = col_b[i] y_b + col_a[i] y_a + s_i                                                         (one dot2 each)
  // = 264 + 328 + 2 x 200 wide MACs per pair instead of 2 x (264 + 2 x 136); the values are those of the round-by-round form.
  int r = 0;
  for (; r + 1 < SVK_POSEIDON_RP; r += 2) {
    const Fr* c = k.pair[r >> 1];
    Fr ya = fr_pow5(st.s[0]) + k.partial[r];
    Fr t0 = Fr::dot3(k.sparse_row[r][0], ya, k.sparse_row[r][1], st.s[1], k.sparse_row[r][2], st.s[2]);
    Fr yb = fr_pow5(t0) + k.partial[r + 1];
    Fr n0 = poseidon_pair_s0(c, yb, ya, st.s[1], st.s[2]);
    Fr n1 = poseidon_pair_si(c + 4, yb, ya, st.s[1]);
    Fr n2 = poseidon_pair_si(c + 6, yb, ya, st.s[2]);
    st.s[0] = n0;
    st.s[1] = n1;
    st.s[2] = n2;
  }
  for (; r < SVK_POSEIDON_RP; r++) {
    st.s[0] = fr_pow5(st.s[0]) + k.partial[r];
    Fr n0 = Fr::dot3(k.sparse_row[r][0], st.s[0], k.sparse_row[r][1], st.s[1], k.sparse_row[r][2], st.s[2]);
    Fr n1 = k.sparse_col_hat[r][0] * st.s[0] + st.s[1];
    Fr n2 = k.sparse_col_hat[r][1] * st.s[0] + st.s[2];
    st.s[0] = n0;
    st.s[1] = n1;
    st.s[2] = n2;
  }
  // second half of the full rounds
  for (int r = 0; r < SVK_POSEIDON_RF / 2 - 1; r++) {
    for (int i = 0; i < 3; i++) st.s[i] = fr_pow5(st.s[i]) + k.end[r][i];
    poseidon_mds(st, k.mds);
  }
  for (int i = 0; i < 3; i++) st.s[i] = fr_pow5(st.s[i]);
  poseidon_mds(st, k.mds);
}
