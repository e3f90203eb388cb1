// Host construction of the Poseidon spec (`OptimizedPoseidonSpec::new::<R_F, R_P, SECURE_MDS>()`,
// snark-verifier/src/util/hash/poseidon.rs:230-315, with `Spec::constants()` of the un-vendored
// poseidon-circuit@50015b7 restated from the Poseidon paper's Grain-LFSR procedure).  Setup work
// done once per context, mirroring the SDK's cached `POSEIDON_SPEC`
// (snark-verifier-sdk/src/halo2.rs:70).  Pinned by the reference KATs through
// svk_poseidon_permute (tests/test_gpu_poseidon.py) and the host build (tests/test_host_arith.py).
#pragma once
#include <array>
#include <vector>

#include "poseidon.cuh"

namespace svk_host {

struct Grain {
  std::vector<uint8_t> s;  // 80-bit state, s[0] oldest
  Grain(int t, int r_f, int r_p) {
    auto app = [&](u32 v, int n) {
      for (int i = n - 1; i >= 0; i--) s.push_back((v >> i) & 1);
    };
    app(1, 2);    // prime field
    app(0, 4);    // x^alpha s-box
    app(254, 12); // field size in bits
    app(t, 12);
    app(r_f, 10);
    app(r_p, 10);
    app((1u << 30) - 1, 30);
    for (int i = 0; i < 160; i++) next();
  }
  int next() {
    int b = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0];
    s.erase(s.begin());
    s.push_back((uint8_t)b);
    return b;
  }
  int bit() {  // shrinking generator
    for (;;) {
      int b1 = next(), b2 = next();
      if (b1) return b2;
    }
  }
  void raw254(u32* v) {  // 254 bits, MSB first
    for (int i = 0; i < 8; i++) v[i] = 0;
    for (int i = 253; i >= 0; i--)
      if (bit()) v[i / 32] |= 1u << (i % 32);
  }
  Fr field_element() {  // with rejection
    for (;;) {
      Fr x;
      raw254(x.v);
      if (Fr::is_canonical(x.v)) return x.to_mont();
    }
  }
  Fr field_element_without_rejection() {
    Fr x;
    raw254(x.v);
    Fr::reduce_once(x.v);  // 2^254 < 2r
    return x.to_mont();
  }
};

typedef std::vector<std::vector<Fr>> Mat;

inline Mat mat_identity(int n) {
  Mat m(n, std::vector<Fr>(n, Fr::zero()));
  for (int i = 0; i < n; i++) m[i][i] = Fr::one();
  return m;
}
inline Mat mat_mul(const Mat& a, const Mat& b) {
  int n = (int)a.size();
  Mat r(n, std::vector<Fr>(n, Fr::zero()));
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++)
      for (int k = 0; k < n; k++) r[i][j] = r[i][j] + a[i][k] * b[k][j];
  return r;
}
inline Mat mat_transpose(const Mat& a) {
  int n = (int)a.size();
  Mat r(n, std::vector<Fr>(n));
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) r[i][j] = a[j][i];
  return r;
}
inline std::vector<Fr> mat_vec(const Mat& m, const std::vector<Fr>& v) {
  int n = (int)m.size();
  std::vector<Fr> r(n, Fr::zero());
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) r[i] = r[i] + m[i][j] * v[j];
  return r;
}
inline bool mat_inverse(const Mat& m, Mat& out) {  // Gauss-Jordan over Fr
  int n = (int)m.size();
  Mat a = m;
  out = mat_identity(n);
  for (int c = 0; c < n; c++) {
    int p = -1;
    for (int i = c; i < n; i++)
      if (!a[i][c].is_zero()) { p = i; break; }
    if (p < 0) return false;
    std::swap(a[c], a[p]);
    std::swap(out[c], out[p]);
    Fr inv = a[c][c].inv();
    for (int j = 0; j < n; j++) { a[c][j] = a[c][j] * inv; out[c][j] = out[c][j] * inv; }
    for (int i = 0; i < n; i++) {
      if (i == c || a[i][c].is_zero()) continue;
      Fr f = a[i][c];
      for (int j = 0; j < n; j++) { a[i][j] = a[i][j] - f * a[c][j]; out[i][j] = out[i][j] - f * out[c][j]; }
    }
  }
  return true;
}

// returns false if the parameters cannot produce a spec
inline bool make_poseidon_consts(PoseidonConsts& k, int secure_mds = 0) {
  const int T = SVK_POSEIDON_T, RF = SVK_POSEIDON_RF, RP = SVK_POSEIDON_RP, half = RF / 2;
  Grain g(T, RF, RP);
  std::vector<std::vector<Fr>> rc(RF + RP, std::vector<Fr>(T));
  for (auto& row : rc)
    for (auto& x : row) x = g.field_element();
  Mat mds(T, std::vector<Fr>(T)), mds_inv;
  for (;;) {
    std::vector<Fr> vals;
    for (;;) {
      vals.clear();
      for (int i = 0; i < 2 * T; i++) vals.push_back(g.field_element_without_rejection());
      bool uniq = true;
      for (int i = 0; i < 2 * T; i++)
        for (int j = i + 1; j < 2 * T; j++) uniq = uniq && !(vals[i] == vals[j]);
      if (uniq) break;
    }
    if (secure_mds) { secure_mds--; continue; }
    for (int i = 0; i < T; i++)
      for (int j = 0; j < T; j++) mds[i][j] = (vals[i] + vals[T + j]).inv();
    break;
  }
  if (!mat_inverse(mds, mds_inv)) return false;

  // calculate_optimized_constants (poseidon.rs:247-297)
  auto put = [&](Fr* dst, const std::vector<Fr>& v) { for (int i = 0; i < T; i++) dst[i] = v[i]; };
  put(k.start[0], rc[0]);
  for (int i = 1; i < half; i++) put(k.start[i], mat_vec(mds_inv, rc[i]));
  std::vector<Fr> acc = rc[half + RP];
  for (int r = RP - 1; r >= 0; r--) {
    std::vector<Fr> tmp = mat_vec(mds_inv, acc);
    k.partial[r] = tmp[0];
    tmp[0] = Fr::zero();
    for (int i = 0; i < T; i++) acc[i] = tmp[i] + rc[half + r][i];
  }
  put(k.start[half], mat_vec(mds_inv, acc));
  for (int i = 0; i < half - 1; i++) put(k.end[i], mat_vec(mds_inv, rc[half + RP + 1 + i]));

  // calculate_sparse_matrices (poseidon.rs:299-315) with factorise (:172-225)
  Mat mt = mat_transpose(mds), accm = mt;
  for (int r = RP - 1; r >= 0; r--) {  // reference collects then reverses
    // factorise accm = M' * M''
    Mat m_hat(T - 1, std::vector<Fr>(T - 1)), m_hat_inv;
    std::vector<Fr> w(T - 1);
    for (int i = 0; i < T - 1; i++) {
      w[i] = accm[i + 1][0];
      for (int j = 0; j < T - 1; j++) m_hat[i][j] = accm[i + 1][j + 1];
    }
    if (!mat_inverse(m_hat, m_hat_inv)) return false;
    std::vector<Fr> w_hat = mat_vec(m_hat_inv, w);  // == Cramer's rule of :207-216
    Mat m_prime = mat_identity(T);
    for (int i = 0; i < T - 1; i++)
      for (int j = 0; j < T - 1; j++) m_prime[i + 1][j + 1] = m_hat[i][j];
    // m'' = [[m00 | m0j], [w_hat | I]];  row = first column of m'', col_hat = first row without m00
    k.sparse_row[r][0] = accm[0][0];
    for (int i = 0; i < T - 1; i++) k.sparse_row[r][i + 1] = w_hat[i];
    for (int j = 0; j < T - 1; j++) k.sparse_col_hat[r][j] = accm[0][j + 1];
    accm = mat_mul(mt, m_prime);
  }
  Mat pre = mat_transpose(accm);
  for (int i = 0; i < T; i++)
    for (int j = 0; j < T; j++) { k.mds[i][j] = mds[i][j]; k.pre_sparse_mds[i][j] = pre[i][j]; }
  for (int r = 0; r < SVK_POSEIDON_RP; r++) {
    k.coop_rc[r] = k.sparse_row[r][0] * k.partial[r];
    k.coop_cc[r][0] = k.sparse_col_hat[r][0] * k.partial[r];
    k.coop_cc[r][1] = k.sparse_col_hat[r][1] * k.partial[r];
  }
  {
    Fr sigma = Fr::one();
    for (int r = 0; r < SVK_POSEIDON_RP; r++) {
      Fr s2 = sigma.sqr(), s5 = s2.sqr() * sigma;
      k.sc_a[r][0] = k.sparse_col_hat[r][0] * s5;
      k.sc_a[r][1] = k.sparse_col_hat[r][1] * s5;
      sigma = k.sparse_row[r][0] * s5;
      if (sigma.is_zero()) return false;  // a zero first row entry: the scaled schedule does not exist for these parameters
      Fr inv = sigma.inv();
      k.sc_r[r][0] = k.sparse_row[r][1] * inv;
      k.sc_r[r][1] = k.sparse_row[r][2] * inv;
      k.sc_k[r] = k.coop_rc[r] * inv;
    }
    k.sc_end = sigma;
    for (int p = 0; p < SVK_POSEIDON_RP / 2; p++) {
      const int a = 2 * p, b = 2 * p + 1;
      Fr* c = k.sc_pair[p];
      c[0] = k.sc_r[a][0]; c[1] = k.sc_r[a][1]; c[2] = k.sc_k[a];
      c[3] = k.sc_r[b][0] * k.sc_a[a][0] + k.sc_r[b][1] * k.sc_a[a][1];
      c[4] = k.sc_r[b][0]; c[5] = k.sc_r[b][1];
      c[6] = k.sc_k[b] + k.sc_r[b][0] * k.coop_cc[a][0] + k.sc_r[b][1] * k.coop_cc[a][1];
      c[7] = k.sc_a[a][0]; c[8] = k.sc_a[b][0]; c[9] = k.coop_cc[a][0] + k.coop_cc[b][0];
      c[10] = k.sc_a[a][1]; c[11] = k.sc_a[b][1]; c[12] = k.coop_cc[a][1] + k.coop_cc[b][1];
    }
  }
  Fr cap = Fr::zero();
  cap.v[2] = 1;  // 2^64
  k.capacity = cap.to_mont();
  return true;
}

}  // namespace svk_host
