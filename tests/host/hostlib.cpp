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
      case 7: r = x.from_mont(); break; case 8: r = x.sqr(); break; case 10: r = x.to_mont_wide(); break; case 11: r = x.inv_fermat(); break; case 13: r = decltype(x)::mul_inline(x, y); break; default: r = x.dbl(); break;
    }
    memcpy(out, r.v, 32);
  } else {
    Fr x, y, r;
    memcpy(x.v, a, 32); memcpy(y.v, b, 32);
    switch (op) {
      case 0: r = x * y; break; case 1: r = x + y; break; case 2: r = x - y; break; case 3: r = x.neg(); break;
      case 4: r = x.inv(); break; case 5: r = x.sqrt_candidate(); break; case 6: r = x.to_mont(); break;
      case 7: r = x.from_mont(); break; case 8: r = x.sqr(); break; case 10: r = x.to_mont_wide(); break; case 11: r = x.inv_fermat(); break; case 13: r = decltype(x)::mul_inline(x, y); break; default: r = x.dbl(); break;
    }
    memcpy(out, r.v, 32);
  }
}
// out = sum_k a[k] * b[k] * R^-1 mod p through the fused dot product (n = 2 or 3); `neg_mask` bit k replaces a[k] by p - a[k]
void host_fe_dot(int field, int n, int neg_mask, const u32* a, const u32* b, u32* out) {
  if (field == 0) {
    Fq x[3], y[3];
    for (int k = 0; k < n; k++) { memcpy(x[k].v, a + 8 * k, 32); memcpy(y[k].v, b + 8 * k, 32); if (neg_mask >> k & 1) x[k] = x[k].neg_lazy(); }
    Fq r = n == 2 ? Fq::dot2(x[0], y[0], x[1], y[1]) : Fq::dot3(x[0], y[0], x[1], y[1], x[2], y[2]);
    memcpy(out, r.v, 32);
  } else {
    Fr x[3], y[3];
    for (int k = 0; k < n; k++) { memcpy(x[k].v, a + 8 * k, 32); memcpy(y[k].v, b + 8 * k, 32); if (neg_mask >> k & 1) x[k] = x[k].neg_lazy(); }
    Fr r = n == 2 ? Fr::dot2(x[0], y[0], x[1], y[1]) : Fr::dot3(x[0], y[0], x[1], y[1], x[2], y[2]);
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

#include "../../snark_verifier_axiom_b200/csrc/coop_pairing.cuh"
struct HostExec {
  static constexpr bool kWarp12 = false;
  template <class F>
  void par(int n, F f) { for (int i = 0; i < n; i++) f(i); }
};
extern "C" {
// The block-cooperative pairing program (coop_pairing.cuh) run lane by lane; outputs in the tower order of host_pairing
int host_pairing_coop(const u32* p1_xy, const u32* q1, const u32* p2_xy, const u32* q2, u32* gt_out, u32* ml_out) {
  static PairingConsts k = svk_host::make_pairing_consts();
  G2Affine Q1 = load_g2(q1), Q2 = load_g2(q2);
  if (!svk_host::g2_on_curve(Q1) || !svk_host::g2_on_curve(Q2)) return -1;
  std::vector<G2Line> t1 = svk_host::make_line_table(Q1, k), t2 = svk_host::make_line_table(Q2, k);
  std::vector<G2LineX> x1, x2;
  for (auto& l : t1) x1.push_back({l.neg_lam, l.c3, l.neg_lam.mul_xi(), l.c3.mul_xi()});
  for (auto& l : t2) x2.push_back({l.neg_lam, l.c3, l.neg_lam.mul_xi(), l.c3.mul_xi()});
  G1Affine P1 = load_aff(p1_xy), P2 = load_aff(p2_xy);
  std::vector<Fq> lb(COOP_SCRATCH_FQ);
  static CoopMem m;
  HostExec ex;
  Fq ml[12], gt[12];
  bool ok = coop_kzg_decide(ex, P1, P2, x1.data(), x2.data(), k, lb.data(), m, ml, gt);
  Fq12 mt = coop_to_tower(ml), gtt = coop_to_tower(gt);
  const Fq2* c[6] = {&gtt.c0.c0, &gtt.c0.c1, &gtt.c0.c2, &gtt.c1.c0, &gtt.c1.c1, &gtt.c1.c2};
  const Fq2* mm[6] = {&mt.c0.c0, &mt.c0.c1, &mt.c0.c2, &mt.c1.c0, &mt.c1.c1, &mt.c1.c2};
  for (int i = 0; i < 6; i++) {
    Fq a = c[i]->c0.from_mont(), b = c[i]->c1.from_mont();
    memcpy(gt_out + 16 * i, a.v, 32); memcpy(gt_out + 16 * i + 8, b.v, 32);
    a = mm[i]->c0.from_mont(); b = mm[i]->c1.from_mont();
    memcpy(ml_out + 16 * i, a.v, 32); memcpy(ml_out + 16 * i + 8, b.v, 32);
  }
  return ok ? 1 : 0;
}
}

#include "../../snark_verifier_axiom_b200/csrc/poseidon_host.h"
extern "C" {
// state/in: canonical limbs; n_in in {0,1,2}; returns 0 ok
int host_poseidon_permute(const u32* state_in, int n_in, const u32* in0, const u32* in1, u32* state_out) {
  static PoseidonConsts k;
  static bool init = false;
  if (!init) { if (!svk_host::make_poseidon_consts(k)) return -1; init = true; }
  PoseidonState st;
  for (int i = 0; i < 3; i++) { memcpy(st.s[i].v, state_in + 8 * i, 32); st.s[i] = st.s[i].to_mont(); }
  Fr a, b; memcpy(a.v, in0, 32); memcpy(b.v, in1, 32); a = a.to_mont(); b = b.to_mont();
  poseidon_permute(st, k, n_in, a, b);
  for (int i = 0; i < 3; i++) { Fr o = st.s[i].from_mont(); memcpy(state_out + 8 * i, o.v, 32); }
  return 0;
}
}

#include "../../snark_verifier_axiom_b200/csrc/compiler.h"
extern "C" {
// Compile + run the tape for ONE proof on the host (same code the kernels run).
// terms_out: per term 3 ints (which: 0 lhs / 1 rhs, base id, slot); returns #terms or <0.
// info_out: [n_regs, n_ops, n_perm, n_challenges, n_scalar_slots, proof_len, err_word, verify_valid, n_fr_mul, n_points]
int host_compile_run_ex(const uint8_t* blob, size_t blob_len, int mos, int transcript_kind, const uint8_t* proof, u32 proof_len,
                        const uint8_t* instances, u32 n_instances, u32* challenges_out, u32* scalars_out,
                        int* terms_out, int max_terms, long long* info_out, char* errbuf, int errbuf_len);
int host_compile_run(const uint8_t* blob, size_t blob_len, int mos, const uint8_t* proof, u32 proof_len,
                     const uint8_t* instances, u32 n_instances, u32* challenges_out, u32* scalars_out,
                     int* terms_out, int max_terms, long long* info_out, char* errbuf, int errbuf_len) {
  return host_compile_run_ex(blob, blob_len, mos, 0, proof, proof_len, instances, n_instances, challenges_out, scalars_out, terms_out,
                             max_terms, info_out, errbuf, errbuf_len);
}
int host_compile_run_ex(const uint8_t* blob, size_t blob_len, int mos, int transcript_kind, const uint8_t* proof, u32 proof_len,
                        const uint8_t* instances, u32 n_instances, u32* challenges_out, u32* scalars_out,
                        int* terms_out, int max_terms, long long* info_out, char* errbuf, int errbuf_len) {
  static PoseidonConsts pk;
  static bool init = false;
  if (!init) { svk_host::make_poseidon_consts(pk); init = true; }
  try {
    svk_host::CompiledProtocol cp = svk_host::compile_protocol(blob, blob_len, mos, transcript_kind);
    std::vector<u32> regs((size_t)cp.n_regs * 8 + 8, 0);
    u32 err = SVK_NO_ERR;
    for (auto& pr : cp.points) {
      if (transcript_kind == 1) continue;  // uncompressed points: validated by k_load_points_be on the device, absorbed from the proof bytes
      G1Affine pt; u32 xc[8], yc[8];
      Fr fx = Fr::zero(), fy = Fr::zero();
      if (pr.byte_offset + 32 > proof_len) tape_note_error(err, pr.byte_offset, SVK_T_EOF);
      else {
        int rc = g1_decompress(proof + pr.byte_offset, pt, xc, yc);
        if (rc == 1) tape_note_error(err, pr.byte_offset, SVK_T_POINT_INVALID);
        else if (rc == 2) tape_note_error(err, pr.byte_offset, SVK_T_POINT_IDENTITY);
        else {
          fq_canon_to_fr_canon(fx.v, xc); fq_canon_to_fr_canon(fy.v, yc);
          fx = fx.to_mont(); fy = fy.to_mont();
        }
      }
      memcpy(&regs[(size_t)pr.val_x * 8], fx.v, 32);
      memcpy(&regs[(size_t)pr.val_y * 8], fy.v, 32);
    }
    RegFile rf{regs.data(), 1, 0};
    TapeIo io{proof, proof_len, instances, n_instances, scalars_out, challenges_out, cp.n_challenges};
    u32 end = cp.verify_valid ? (u32)cp.ops.size() : cp.read_ops_end;
    if (transcript_kind == 1) {
      TranscriptState<true> st;
      keccak_reset(st.ks);
      tape_exec<true>(cp.ops.data(), 0, end, cp.aux.data(), cp.consts.data(), pk, rf, io, st, err);
    } else {
      TranscriptState<false> st;
      poseidon_init(st.ps, pk);
      tape_exec<false>(cp.ops.data(), 0, end, cp.aux.data(), cp.consts.data(), pk, rf, io, st, err);
    }
    int nt = 0;
    for (int which = 0; which < 2; which++)
      for (auto& t : (which ? cp.rhs : cp.lhs)) {
        if (nt >= max_terms) return -3;
        terms_out[3 * nt] = which; terms_out[3 * nt + 1] = t.base; terms_out[3 * nt + 2] = t.slot; nt++;
      }
    long long info[10] = {cp.n_regs, (long long)cp.ops.size(), cp.n_perm, cp.n_challenges, cp.n_scalar_slots, cp.proof_len, err,
                          cp.verify_valid, (long long)cp.n_fr_mul, (long long)cp.points.size()};
    memcpy(info_out, info, sizeof info);
    return nt;
  } catch (svk_host::CompileError& e) {
    snprintf(errbuf, errbuf_len, "%s", e.what());
    return e.kind == SVK_INVALID_PROTOCOL ? -2 : -1;
  }
}
}

extern "C" {
// Compiles the same protocol from the library's blob and from the reference's bincode form; returns 1 when the two compiled
// tapes are identical (ops, constants, aux tables, point schedule, Msm terms, old-accumulator indices), 0 when they differ,
// < 0 on a parse error.  *consumed / *fe_used as in svk_protocol_compile_bincode.
int host_compile_compare_bincode(const uint8_t* blob, size_t blob_len, const uint8_t* bytes, size_t len, int fe_encoding, int mos,
                                 int transcript_kind, size_t* consumed, int* fe_used, char* errbuf, int errbuf_len) {
  try {
    svk_host::CompiledProtocol a = svk_host::compile_protocol(blob, blob_len, mos, transcript_kind);
    svk_host::CompiledProtocol b = svk_host::compile_protocol_bincode(bytes, len, fe_encoding, mos, transcript_kind, consumed, fe_used);
    bool same = a.ops.size() == b.ops.size() && a.consts.size() == b.consts.size() && a.aux == b.aux && a.n_regs == b.n_regs &&
                a.proof_len == b.proof_len && a.num_instance == b.num_instance && a.n_challenges == b.n_challenges &&
                a.lhs.size() == b.lhs.size() && a.rhs.size() == b.rhs.size() && a.old_acc_idx == b.old_acc_idx &&
                a.preprocessed.size() == b.preprocessed.size() && a.points.size() == b.points.size() && a.verify_valid == b.verify_valid;
    if (!same) return 0;
    if (memcmp(a.ops.data(), b.ops.data(), a.ops.size() * sizeof(TapeOp))) return 0;
    if (memcmp(a.consts.data(), b.consts.data(), a.consts.size() * sizeof(Fr))) return 0;
    if (a.preprocessed.size() && memcmp(a.preprocessed.data(), b.preprocessed.data(), a.preprocessed.size() * sizeof(svk_g1))) return 0;
    for (size_t i = 0; i < a.lhs.size(); i++) if (a.lhs[i].base != b.lhs[i].base || a.lhs[i].slot != b.lhs[i].slot) return 0;
    for (size_t i = 0; i < a.rhs.size(); i++) if (a.rhs[i].base != b.rhs[i].base || a.rhs[i].slot != b.rhs[i].slot) return 0;
    for (size_t i = 0; i < a.points.size(); i++) if (a.points[i].byte_offset != b.points[i].byte_offset) return 0;
    return 1;
  } catch (svk_host::CompileError& e) {
    snprintf(errbuf, errbuf_len, "%s", e.what());
    return -1;
  }
}

// returns 1 if cyclotomic_sqr == sqr on an element of the cyclotomic subgroup built from random input limbs
int host_cyclotomic_check(const u32* limbs96) {
  static PairingConsts k = svk_host::make_pairing_consts();
  Fq12 f;
  Fq* p = reinterpret_cast<Fq*>(&f);
  for (int i = 0; i < 12; i++) { memcpy(p[i].v, limbs96 + 8 * i, 32); p[i].v[7] &= 0x0fffffffu; p[i] = p[i].to_mont(); }
  Fq12 t = f.conj() * f.inv();
  Fq12 g = fq12_frob2(t, k) * t;  // in the cyclotomic subgroup
  Fq12 a = g.sqr(), b = g.cyclotomic_sqr();
  return (a == b) ? 1 : 0;
}
}

#include "../../snark_verifier_axiom_b200/csrc/keccak.cuh"
extern "C" {
void host_keccak256(const uint8_t* data, u32 n, uint8_t* out) {
  KeccakSponge k; keccak_reset(k); keccak_absorb(k, data, n); keccak_finish(k, out);
}
}

// ---- Pasta: the same templates over the Pallas / Vesta base fields --------------------------
#include "../../snark_verifier_axiom_b200/csrc/pasta.cuh"
template <class F>
static void fe_op_t(int op, const u32* a, const u32* b, u32* out) {
  F x, y, r;
  memcpy(x.v, a, 32); memcpy(y.v, b, 32);
  switch (op) {
    case 0: r = x * y; break; case 1: r = x + y; break; case 2: r = x - y; break; case 3: r = x.neg(); break;
    case 4: r = x.inv(); break; case 6: r = x.to_mont(); break;
    case 7: r = x.from_mont(); break; case 8: r = x.sqr(); break; case 11: r = x.inv_fermat(); break; case 13: r = decltype(x)::mul_inline(x, y); break; default: r = x.dbl(); break;
  }
  memcpy(out, r.v, 32);
}
template <class C>
static int muladd_t(const u32* p_xy, const u32* k, const u32* q_xy, u32* out_xy) {
  typedef typename C::Base F;
  AffT<F> p, q;
  memcpy(p.x.v, p_xy, 32); memcpy(p.y.v, p_xy + 8, 32); memcpy(q.x.v, q_xy, 32); memcpy(q.y.v, q_xy + 8, 32);
  if (!p.is_identity()) { p.x = p.x.to_mont(); p.y = p.y.to_mont(); }
  if (!q.is_identity()) { q.x = q.x.to_mont(); q.y = q.y.to_mont(); }
  if (!curve_on_curve<C>(p) || !curve_on_curve<C>(q)) return 1;
  XyzzT<F> acc = XyzzT<F>::identity();
  JacT<F> jac = JacT<F>::identity();
  for (int w = 7; w >= 0; w--) for (int b = 31; b >= 0; b--) {
    acc = acc.dbl(); jac = jac.dbl();
    if ((k[w] >> b) & 1) { acc = acc.add_affine(p); jac = jac.add_affine(p); }
  }
  XyzzT<F> r = acc.add(XyzzT<F>::from_affine(q));
  JacT<F> rj = jac.add(JacT<F>::from_affine(q));
  AffT<F> a1 = r.to_affine(), a2 = rj.to_affine();
  if (!(a1.x == a2.x) || !(a1.y == a2.y)) return 2;
  if (!curve_on_curve<C>(a1)) return 3;
  F x = a1.x.from_mont(), y = a1.y.from_mont();
  memcpy(out_xy, x.v, 32); memcpy(out_xy + 8, y.v, 32);
  return 0;
}
extern "C" {
// field: 2 Pallas base (= Vesta scalar), 3 Vesta base (= Pallas scalar)
void host_pasta_fe_op(int field, int op, const u32* a, const u32* b, u32* out) {
  if (field == 2) fe_op_t<PallasFp>(op, a, b, out); else fe_op_t<VestaFp>(op, a, b, out);
}
// out = k*P + Q on curve 0 BN254 G1 / 1 Pallas / 2 Vesta through BOTH the XYZZ and the Jacobian formulas (must agree)
int host_curve_muladd(int curve, const u32* p_xy, const u32* k, const u32* q_xy, u32* out_xy) {
  if (curve == 0) return muladd_t<CurveBn254>(p_xy, k, q_xy, out_xy);
  if (curve == 1) return muladd_t<CurvePallas>(p_xy, k, q_xy, out_xy);
  return muladd_t<CurveVesta>(p_xy, k, q_xy, out_xy);
}
}
template <class F>
static void fe_dot_t(int n, int neg_mask, const u32* a, const u32* b, u32* out) {
  F x[3], y[3];
  for (int k = 0; k < n; k++) { memcpy(x[k].v, a + 8 * k, 32); memcpy(y[k].v, b + 8 * k, 32); if (neg_mask >> k & 1) x[k] = x[k].neg_lazy(); }
  F r = n == 2 ? F::dot2(x[0], y[0], x[1], y[1]) : F::dot3(x[0], y[0], x[1], y[1], x[2], y[2]);
  memcpy(out, r.v, 32);
}
extern "C" void host_pasta_fe_dot(int field, int n, int neg_mask, const u32* a, const u32* b, u32* out) {
  if (field == 2) fe_dot_t<PallasFp>(n, neg_mask, a, b, out); else fe_dot_t<VestaFp>(n, neg_mask, a, b, out);
}

// ---- Straus (signed 5-bit windows) over the BN254 G1 templates: out = sum_t k_t * P_t --------------------
#include "../../snark_verifier_axiom_b200/csrc/straus.cuh"
extern "C" int host_straus(int nt, const u32* pts_xy, const u32* ks, u32* out_xy) {
  if (nt > 16) return 1;
  static G1Jac tables[16 * STRAUS_TABLE];
  u32 k[16][8];
  for (int t = 0; t < nt; t++) {
    memcpy(k[t], ks + 8 * t, 32);
    straus_recode(k[t]);
    straus_build_table(tables + (size_t)t * STRAUS_TABLE, 1, load_aff(pts_xy + 16 * t));
  }
  G1Jac acc = straus_run<false>(&k[0][0], (u32)nt, tables, 1);
  static Fq prefix[16 * STRAUS_TABLE];
  straus_normalize(tables, prefix, 1, (u32)nt * STRAUS_TABLE);
  G1Jac acc2 = straus_run<true>(&k[0][0], (u32)nt, tables, 1);
  G1Affine a1 = acc.to_affine(), a2 = acc2.to_affine();
  if (!(a1.x == a2.x) || !(a1.y == a2.y)) return 2;  // Jacobian-table and normalised-table runs must agree
  store_aff(out_xy, a1);
  return 0;
}

// ---- GLV decomposition (csrc/glv.cuh): out = { |k1| (4 limbs), neg1, |k2| (4 limbs), neg2 }, returns 1 when both < 2^128
#include "../../snark_verifier_axiom_b200/csrc/glv.cuh"
extern "C" int host_glv_decompose(const u32* k, u32* out) {
  u32 k1[4], k2[4], n1, n2;
  bool ok = glv_decompose(k, k1, n1, k2, n2);
  memcpy(out, k1, 16); out[4] = n1; memcpy(out + 5, k2, 16); out[9] = n2;
  return ok ? 1 : 0;
}
extern "C" void host_glv_beta(u32* out) { Fq b = glv_beta_mont().from_mont(); memcpy(out, b.v, 32); }

// one-term GLV Straus (straus.cuh straus_run_glv1): out = k * P
extern "C" int host_straus_glv1(const u32* p_xy, const u32* k, u32* out_xy) {
  static G1Jac table[STRAUS_TABLE];
  u32 k1[4], k2[4], n1, n2;
  if (!glv_decompose(k, k1, n1, k2, n2)) return 1;
  straus_build_table(table, 1, load_aff(p_xy));
  store_aff(out_xy, straus_run_glv1(k1, n1, k2, n2, table, 1, glv_beta_mont()).to_affine());
  return 0;
}
