// Poseidon (T = 3, RATE = 2, alpha = 5) over BN254 Fr: the optimised permutation of
// snark-verifier/src/util/hash/poseidon.rs:469-501 and the sponge of :455-467.
//
// `PoseidonConsts` is the device image of `OptimizedPoseidonSpec` (poseidon.rs:59-95), built on the
// host by poseidon_host.h (Grain LFSR + sparse-MDS factorisation) for the SDK parameters
// R_F = 8, R_P = 57 (snark-verifier-sdk/src/halo2.rs:52-56).  All values in Montgomery form.
// Cost: 8 full rounds x (3 x 3 + 9) + 57 partial rounds x (3 + 5) = 600 Fr multiplications in the textbook count; every
// MDS row is ONE fused 3-term dot product (field.cuh `dot3`: 3 limb products, 1 reduction, no additions), which is
// value-identical and brings the permutation to ~510 multiplication-equivalents of integer-pipe work.
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
  // partial rounds with sparse MDS (poseidon.rs:398-410)
  for (int r = 0; r < SVK_POSEIDON_RP; r++) {
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
