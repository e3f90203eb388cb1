// Host protocol compiler ("TapeLoader"): runs the reference's GENERIC verifier algorithm once,
// symbolically, over a `PlonkProtocol`, and records the straight-line tape the device executes for
// every proof (tape.cuh).  This is the third `Loader` back-end idea of the reference
// (snark-verifier/src/loader.rs:252-260; EvmLoader records Yul the same way,
// loader/evm/loader.rs:117-135) done for a B200: constants are folded on the host
// (cf. `Value::Constant` folding, loader/evm/loader.rs:384-410), everything that depends on the proof
// becomes a tape register.
//
// Restated reference logic (file:line in each function):
//   PlonkProof::read / evaluations / commitments / queries   verifier/plonk/proof.rs:52-318
//   PlonkProtocol::langranges, CommonPolynomialEvaluation     verifier/plonk/protocol.rs:70-98, 201-279
//   Expression::evaluate / degree / used_*                    verifier/plonk/protocol.rs:322-417
//   PlonkSuccinctVerifier::verify                             verifier/plonk.rs:58-92
//   Bdfg21 (SHPLONK)                                          pcs/kzg/multiopen/bdfg21.rs:47-367
//   Gwc19                                                     pcs/kzg/multiopen/gwc19.rs:43-158
//   Msm                                                       util/msm.rs:20-205
//   Domain::rotate_scalar, Fraction                           util/arithmetic.rs:131-234
#pragma once
#include <algorithm>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/svk.h"
#include "tape.cuh"

namespace svk_host {

struct CompileError : std::runtime_error {
  int kind;  // SVK_INVALID_PROTOCOL, or -1 for malformed/unsupported input
  CompileError(int k, const std::string& m) : std::runtime_error(m), kind(k) {}
};

// ------------------------------------------------------------------ host Fr helpers
inline Fr fr_from_u64(u64 x) {
  Fr r = Fr::zero();
  r.v[0] = (u32)x;
  r.v[1] = (u32)(x >> 32);
  return r.to_mont();
}
inline Fr fr_pow_u64(Fr a, u64 e) {
  Fr r = Fr::one();
  while (e) {
    if (e & 1) r = r * a;
    a = a.sqr();
    e >>= 1;
  }
  return r;
}
struct FrLess {
  bool operator()(const Fr& a, const Fr& b) const { return memcmp(a.v, b.v, 32) < 0; }
};

// ------------------------------------------------------------------ serialized protocol (see snark_verifier_axiom_b200/protocol.py)
struct Reader {
  const uint8_t* p;
  size_t n, pos = 0;
  Reader(const uint8_t* p_, size_t n_) : p(p_), n(n_) {}
  void need(size_t k) {
    if (pos + k > n) throw CompileError(-1, "protocol blob truncated");
  }
  u32 u32_() {
    need(4);
    u32 v;
    memcpy(&v, p + pos, 4);
    pos += 4;
    return v;
  }
  int32_t i32_() { return (int32_t)u32_(); }
  bool at_end() const { return pos >= n; }
  uint8_t u8_() {
    need(1);
    return p[pos++];
  }
  Fr fr_() {  // canonical LE -> Montgomery
    need(32);
    Fr x;
    fe_load_le(x.v, p + pos);
    pos += 32;
    if (!Fr::is_canonical(x.v)) throw CompileError(-1, "non-canonical Fr in protocol");
    return x.to_mont();
  }
  void bytes(uint8_t* out, size_t k) {
    need(k);
    memcpy(out, p + pos, k);
    pos += k;
  }
};

struct Query {
  u32 poly;
  int32_t rot;
  bool operator<(const Query& o) const { return poly != o.poly ? poly < o.poly : rot < o.rot; }
  bool operator==(const Query& o) const { return poly == o.poly && rot == o.rot; }
};

struct Expr {
  enum Tag { CONST = 0, IDENTITY = 1, LAGRANGE = 2, POLY = 3, CHALLENGE = 4, NEG = 5, SUM = 6, PRODUCT = 7, SCALED = 8, DISTRIBUTE = 9 } tag;
  Fr c;                 // CONST value / SCALED factor
  int32_t i = 0;        // LAGRANGE index / CHALLENGE index
  Query q{0, 0};        // POLY
  std::vector<std::unique_ptr<Expr>> kids;  // DISTRIBUTE: kids[0..n-1] exprs, kids[n] base
};

inline std::unique_ptr<Expr> parse_expr(Reader& r, int depth = 0) {
  if (depth > 4096) throw CompileError(-1, "expression too deep");
  auto e = std::make_unique<Expr>();
  uint8_t t = r.u8_();
  if (t > 9) throw CompileError(-1, "bad expression tag");
  e->tag = (Expr::Tag)t;
  switch (e->tag) {
    case Expr::CONST: e->c = r.fr_(); break;
    case Expr::IDENTITY: break;
    case Expr::LAGRANGE: e->i = r.i32_(); break;
    case Expr::POLY: e->q.poly = r.u32_(); e->q.rot = r.i32_(); break;
    case Expr::CHALLENGE: e->i = (int32_t)r.u32_(); break;
    case Expr::NEG: e->kids.push_back(parse_expr(r, depth + 1)); break;
    case Expr::SUM:
    case Expr::PRODUCT:
      e->kids.push_back(parse_expr(r, depth + 1));
      e->kids.push_back(parse_expr(r, depth + 1));
      break;
    case Expr::SCALED:
      e->kids.push_back(parse_expr(r, depth + 1));
      e->c = r.fr_();
      break;
    case Expr::DISTRIBUTE: {
      u32 n = r.u32_();
      if (n == 0 || n > 65536) throw CompileError(-1, "bad DistributePowers arity");
      for (u32 k = 0; k < n + 1; k++) e->kids.push_back(parse_expr(r, depth + 1));
      break;
    }
  }
  return e;
}

// protocol.rs:372-386
inline size_t expr_degree(const Expr& e) {
  switch (e.tag) {
    case Expr::CONST: case Expr::CHALLENGE: return 0;
    case Expr::IDENTITY: case Expr::LAGRANGE: case Expr::POLY: return 1;
    case Expr::NEG: case Expr::SCALED: return expr_degree(*e.kids[0]);
    case Expr::SUM: return std::max(expr_degree(*e.kids[0]), expr_degree(*e.kids[1]));
    case Expr::PRODUCT: return expr_degree(*e.kids[0]) + expr_degree(*e.kids[1]);
    case Expr::DISTRIBUTE: {
      size_t d = 0;
      for (auto& k : e.kids) d = std::max(d, expr_degree(*k));
      return d;
    }
  }
  return 0;
}
// protocol.rs:388-417
inline void expr_collect(const Expr& e, std::set<int32_t>& lag, std::set<Query>& qs) {
  if (e.tag == Expr::LAGRANGE) lag.insert(e.i);
  if (e.tag == Expr::POLY) qs.insert(e.q);
  for (auto& k : e.kids) expr_collect(*k, lag, qs);
}

struct ProtocolDesc {
  u32 k = 0;
  Fr gen, gen_inv, n_inv;
  u64 n = 0;
  std::vector<svk_g1> preprocessed;
  std::vector<u32> num_instance, num_witness, num_challenge;
  std::vector<Query> evaluations, queries;
  u32 chunk_degree = 1;
  std::unique_ptr<Expr> numerator;
  bool has_initial_state = false;
  Fr initial_state;
  uint8_t linearization = 0;  // 0 None, 1 WithoutConstant, 2 MinusVanishingTimesQuotient
  std::vector<std::vector<std::pair<u32, u32>>> accumulator_indices;
  u32 acc_limbs = 3, acc_bits = 88;  // `LimbsEncoding<LIMBS, BITS>` (pcs/kzg/accumulator.rs:34); SDK: snark-verifier-sdk/src/lib.rs:33-40

  Fr rotate_one(int32_t rot) const {  // Domain::rotate_scalar(1, rot), arithmetic.rs:154-161
    if (rot == 0) return Fr::one();
    if (rot > 0) return fr_pow_u64(gen, (u64)rot);
    return fr_pow_u64(gen_inv, (u64)(-(int64_t)rot));
  }
  size_t num_chunk() const {  // protocol.rs:288-293
    size_t d = expr_degree(*numerator);
    d = d ? d - 1 : 0;
    return (d + chunk_degree - 1) / chunk_degree;
  }
};

inline void validate_accumulator_indices(const ProtocolDesc& p) {
  if (p.accumulator_indices.empty()) return;
  if (p.acc_limbs == 0 || p.acc_bits == 0 || (p.acc_limbs - 1) * p.acc_bits + 256 > 1024)
    throw CompileError(-1, "LimbsEncoding parameters out of range");
  for (auto& v : p.accumulator_indices) {
    // `assert_eq!(limbs.len(), 4 * LIMBS)` (accumulator.rs:61) and `instances[i][j]` (proof.rs:139-146) panic in the reference
    if (v.size() != 4 * (size_t)p.acc_limbs) throw CompileError(-1, "accumulator_indices entry must name 4 * LIMBS instances");
    for (auto& ij : v)
      if (ij.first >= p.num_instance.size() || ij.second >= p.num_instance[ij.first])
        throw CompileError(-1, "accumulator index out of the instance ranges");
  }
}

inline ProtocolDesc parse_protocol(const uint8_t* blob, size_t len) {
  Reader r(blob, len);
  if (r.u32_() != 0x504b5653u) throw CompileError(-1, "bad magic (want 'SVKP')");
  if (r.u32_() != 1) throw CompileError(-1, "unsupported protocol blob version");
  ProtocolDesc p;
  p.k = r.u32_();
  if (p.k > 28) throw CompileError(-1, "domain k > 28");
  p.n = 1ull << p.k;
  p.gen = r.fr_();
  p.gen_inv = p.gen.inv();
  p.n_inv = fr_from_u64(p.n).inv();
  auto count = [&](u32 max) {
    u32 c = r.u32_();
    if (c > max) throw CompileError(-1, "count out of range");
    return c;
  };
  u32 np = count(1 << 16);
  p.preprocessed.resize(np);
  for (auto& g : p.preprocessed) r.bytes((uint8_t*)&g, 64);
  for (u32 i = count(1 << 16); i--;) p.num_instance.push_back(r.u32_());
  for (u32 i = count(1 << 16); i--;) p.num_witness.push_back(r.u32_());
  for (u32 i = count(1 << 16); i--;) p.num_challenge.push_back(r.u32_());
  for (u32 i = count(1 << 20); i--;) { Query q; q.poly = r.u32_(); q.rot = r.i32_(); p.evaluations.push_back(q); }
  for (u32 i = count(1 << 20); i--;) { Query q; q.poly = r.u32_(); q.rot = r.i32_(); p.queries.push_back(q); }
  p.chunk_degree = r.u32_();
  if (p.chunk_degree == 0) throw CompileError(-1, "chunk_degree == 0");
  p.numerator = parse_expr(r);
  p.has_initial_state = r.u8_() != 0;
  if (p.has_initial_state) p.initial_state = r.fr_();
  if (r.u8_() != 0) throw CompileError(-1, "instance_committing_key is not supported on the KZG path (SURVEY App. A)");
  p.linearization = r.u8_();
  if (p.linearization > 2) throw CompileError(-1, "bad linearization tag");
  for (u32 i = count(1 << 16); i--;) {
    std::vector<std::pair<u32, u32>> v;
    for (u32 j = count(1 << 16); j--;) { u32 a = r.u32_(), b = r.u32_(); v.push_back({a, b}); }
    p.accumulator_indices.push_back(v);
  }
  if (!r.at_end()) {  // optional trailer: the accumulator encoding's type parameters
    p.acc_limbs = r.u8_();
    p.acc_bits = r.u8_();
  }
  validate_accumulator_indices(p);
  return p;
}

// ---- the reference's own wire format: `bincode::serialize(&PlonkProtocol<G1Affine>)` ----------------------------------------
// bincode 1.3.3 with default options (snark-verifier-sdk/Cargo.toml:18; `bincode::deserialize_from`, sdk/src/halo2.rs:262-269):
// fixed-width little-endian integers, usize as u64, enum variant index as u32, Option as one tag byte, Vec as u64 length +
// elements, structs / tuples / Box as the plain concatenation of their fields in declaration order.  Field order from the
// serde derives: PlonkProtocol (verifier/plonk/protocol.rs:20-63), Domain (util/arithmetic.rs:130-142), Query (:296-300),
// Rotation (arithmetic.rs:99-100), QuotientPolynomial (:281-285), Expression (:308-319), CommonPolynomial (:180-185),
// LinearizationStrategy (:503-513), InstanceCommittingKey (:515-519).
// Field elements: halo2curves' `derive_serde` representation is NOT in the tree (Cargo.lock:1803-1826).  Two encodings are
// accepted: SVK_FE_MONTGOMERY = the derived `Fr([u64; 4])` / `G1Affine { x, y }` (raw Montgomery limbs, R = 2^256 -- the 0.3.x
// derive as far as it can be recalled without the source) and SVK_FE_CANONICAL = 32-byte little-endian `to_repr()` (later
// halo2curves).  SVK_FE_AUTO picks the one under which `domain.n_inv * n == 1`.  FORMAT UNPINNED: no reference-written
// file exists in the tree to check against; tests round-trip through a writer that follows the same rules.
// SVK_FE_AUTO / SVK_FE_MONTGOMERY / SVK_FE_CANONICAL: include/svk.h

struct BincodeReader {
  Reader r;
  int fe;
  BincodeReader(const uint8_t* p, size_t n, int fe_) : r(p, n), fe(fe_) {}
  u64 u64_() {
    u64 lo = r.u32_(), hi = r.u32_();
    return lo | (hi << 32);
  }
  u32 len_(u64 max) {
    u64 v = u64_();
    if (v > max) throw CompileError(-1, "bincode: length out of range");
    return (u32)v;
  }
  u32 usize_() { return len_(0xffffffffull); }
  bool option_() {
    uint8_t t = r.u8_();
    if (t > 1) throw CompileError(-1, "bincode: bad Option tag");
    return t == 1;
  }
  template <class F>
  F fe_() {  // -> Montgomery form
    F x;
    uint8_t b[32];
    r.bytes(b, 32);
    fe_load_le(x.v, b);
    if (!F::is_canonical(x.v)) throw CompileError(-1, "bincode: field element out of range");
    return fe == SVK_FE_CANONICAL ? x.to_mont() : x;
  }
  svk_g1 g1_() {  // -> canonical bytes (identity = zeros in both encodings)
    Fq x = fe_<Fq>().from_mont(), y = fe_<Fq>().from_mont();
    svk_g1 g;
    fe_store_le(g.x.b, x.v);
    fe_store_le(g.y.b, y.v);
    return g;
  }
  Query query_() {
    Query q;
    q.poly = usize_();
    q.rot = r.i32_();
    return q;
  }
};

inline std::unique_ptr<Expr> parse_expr_bincode(BincodeReader& b, int depth = 0) {
  if (depth > 20000) throw CompileError(-1, "expression too deep");
  auto e = std::make_unique<Expr>();
  u32 v = b.r.u32_();
  switch (v) {
    case 0: e->tag = Expr::CONST; e->c = b.fe_<Fr>(); break;
    case 1: {
      u32 cp = b.r.u32_();
      if (cp == 0) e->tag = Expr::IDENTITY;
      else if (cp == 1) { e->tag = Expr::LAGRANGE; e->i = b.r.i32_(); }
      else throw CompileError(-1, "bincode: bad CommonPolynomial variant");
      break;
    }
    case 2: e->tag = Expr::POLY; e->q = b.query_(); break;
    case 3: e->tag = Expr::CHALLENGE; e->i = (int32_t)b.usize_(); break;
    case 4: e->tag = Expr::NEG; e->kids.push_back(parse_expr_bincode(b, depth + 1)); break;
    case 5:
    case 6:
      e->tag = v == 5 ? Expr::SUM : Expr::PRODUCT;
      e->kids.push_back(parse_expr_bincode(b, depth + 1));
      e->kids.push_back(parse_expr_bincode(b, depth + 1));
      break;
    case 7:
      e->tag = Expr::SCALED;
      e->kids.push_back(parse_expr_bincode(b, depth + 1));
      e->c = b.fe_<Fr>();
      break;
    case 8: {
      e->tag = Expr::DISTRIBUTE;
      u32 n = b.len_(65536);
      if (n == 0) throw CompileError(-1, "bad DistributePowers arity");
      for (u32 k = 0; k < n + 1; k++) e->kids.push_back(parse_expr_bincode(b, depth + 1));
      break;
    }
    default: throw CompileError(-1, "bincode: bad Expression variant");
  }
  return e;
}

inline ProtocolDesc parse_protocol_bincode_as(const uint8_t* bytes, size_t len, int fe, size_t* consumed) {
  BincodeReader b(bytes, len, fe);
  ProtocolDesc p;
  p.k = b.usize_();
  if (p.k > 28) throw CompileError(-1, "domain k > 28");
  p.n = 1ull << p.k;
  if (b.u64_() != p.n) throw CompileError(-1, "bincode: domain.n != 2^k");
  p.n_inv = b.fe_<Fr>();
  p.gen = b.fe_<Fr>();
  p.gen_inv = b.fe_<Fr>();
  if (!(p.n_inv * fr_from_u64(p.n) == Fr::one()) || !(p.gen * p.gen_inv == Fr::one()))
    throw CompileError(-3, "bincode: inconsistent Domain (wrong field-element encoding?)");
  for (u32 i = b.len_(1 << 16); i--;) p.preprocessed.push_back(b.g1_());
  for (u32 i = b.len_(1 << 16); i--;) p.num_instance.push_back(b.usize_());
  for (u32 i = b.len_(1 << 16); i--;) p.num_witness.push_back(b.usize_());
  for (u32 i = b.len_(1 << 16); i--;) p.num_challenge.push_back(b.usize_());
  for (u32 i = b.len_(1 << 20); i--;) p.evaluations.push_back(b.query_());
  for (u32 i = b.len_(1 << 20); i--;) p.queries.push_back(b.query_());
  p.chunk_degree = b.usize_();
  if (p.chunk_degree == 0) throw CompileError(-1, "chunk_degree == 0");
  p.numerator = parse_expr_bincode(b);
  p.has_initial_state = b.option_();
  if (p.has_initial_state) p.initial_state = b.fe_<Fr>();
  if (b.option_()) throw CompileError(-1, "instance_committing_key is not supported on the KZG path (SURVEY App. A)");
  p.linearization = 0;
  if (b.option_()) {
    u32 v = b.r.u32_();
    if (v > 1) throw CompileError(-1, "bincode: bad LinearizationStrategy variant");
    p.linearization = (uint8_t)(v + 1);
  }
  for (u32 i = b.len_(1 << 16); i--;) {
    std::vector<std::pair<u32, u32>> v;
    for (u32 j = b.len_(1 << 16); j--;) { u32 a = b.usize_(), c = b.usize_(); v.push_back({a, c}); }
    p.accumulator_indices.push_back(v);
  }
  validate_accumulator_indices(p);
  if (consumed) *consumed = b.r.pos;
  return p;
}

inline ProtocolDesc parse_protocol_bincode(const uint8_t* bytes, size_t len, int fe_encoding, size_t* consumed, int* fe_used = nullptr) {
  if (fe_encoding < 0 || fe_encoding > 2) throw CompileError(-1, "unknown field-element encoding");
  if (fe_encoding != SVK_FE_AUTO) {
    if (fe_used) *fe_used = fe_encoding;
    return parse_protocol_bincode_as(bytes, len, fe_encoding, consumed);
  }
  try {
    if (fe_used) *fe_used = SVK_FE_MONTGOMERY;
    return parse_protocol_bincode_as(bytes, len, SVK_FE_MONTGOMERY, consumed);
  } catch (CompileError& e) {
    if (e.kind != -3 && std::string(e.what()).find("out of range") == std::string::npos) throw;
  }
  if (fe_used) *fe_used = SVK_FE_CANONICAL;
  return parse_protocol_bincode_as(bytes, len, SVK_FE_CANONICAL, consumed);
}

// ------------------------------------------------------------------ symbolic scalars + tape builder
struct Sym {
  bool is_const = true;
  Fr c = Fr::zero();
  int id = -1;
};

struct SsaOp {
  uint16_t op;
  int dst = -1, a = -1, b = -1;  // value ids (or const index in b for *C ops / a for T_CONST)
  u32 imm = 0;
  std::vector<std::pair<int, int>> binv;  // (src value, dst value)
};

struct PointRead {
  u32 byte_offset;  // in the proof
  int val_x, val_y; // value ids holding (x mod r, y mod r) in Fr Montgomery form
};

#define SVK_BASE_G (-1)

struct MsmTerm {
  int base;      // SVK_BASE_G, [0, n_pre) preprocessed, n_pre + j = j-th point read from the proof
  Sym scalar;
  int slot = -1; // out_scalar slot (after finalize), -1 when the scalar is the constant 1
};

class TapeBuilder {
 public:
  std::vector<SsaOp> ops;
  std::vector<Fr> consts;
  int n_values = 0;
  int transcript_kind = 0;  // 0 Poseidon (halo2.rs), 1 Keccak EvmTranscript (evm.rs)
  size_t kbuf_len = 0;      // EvmTranscript: current length of `buf` in bytes (known statically)
  std::vector<Sym> tbuf;  // transcript buffer (Poseidon::buf)
  std::vector<PointRead> points;
  u32 cursor = 0;         // proof byte cursor
  u32 n_instances = 0;
  int n_perm = 0;

  int new_value() { return n_values++; }
  int const_index(const Fr& c) {
    for (size_t i = 0; i < consts.size(); i++)
      if (consts[i] == c) return (int)i;
    consts.push_back(c);
    if (consts.size() > 65535) throw CompileError(-1, "too many constants");
    return (int)consts.size() - 1;
  }
  Sym cst(const Fr& c) { Sym s; s.c = c; return s; }
  Sym val(int id) { Sym s; s.is_const = false; s.id = id; return s; }
  int materialize(const Sym& s) {
    if (!s.is_const) return s.id;
    SsaOp o; o.op = T_CONST; o.dst = new_value(); o.a = const_index(s.c);
    ops.push_back(o);
    return o.dst;
  }
  Sym emit2(uint16_t op, int a, int b) {
    SsaOp o; o.op = op; o.dst = new_value(); o.a = a; o.b = b;
    ops.push_back(o);
    return val(o.dst);
  }
  Sym add(const Sym& x, const Sym& y) {
    if (x.is_const && y.is_const) return cst(x.c + y.c);
    if (x.is_const) return x.c.is_zero() ? y : emit2(T_ADDC, y.id, const_index(x.c));
    if (y.is_const) return y.c.is_zero() ? x : emit2(T_ADDC, x.id, const_index(y.c));
    return emit2(T_ADD, x.id, y.id);
  }
  Sym sub(const Sym& x, const Sym& y) {
    if (x.is_const && y.is_const) return cst(x.c - y.c);
    if (y.is_const) return y.c.is_zero() ? x : emit2(T_SUBC, x.id, const_index(y.c));
    if (x.is_const) return emit2(T_CSUB, y.id, const_index(x.c));
    return emit2(T_SUB, x.id, y.id);
  }
  Sym mul(const Sym& x, const Sym& y) {
    if (x.is_const && y.is_const) return cst(x.c * y.c);
    if (x.is_const) {
      if (x.c.is_zero()) return cst(Fr::zero());
      if (x.c == Fr::one()) return y;
      return emit2(T_MULC, y.id, const_index(x.c));
    }
    if (y.is_const) return mul(y, x);
    return emit2(T_MUL, x.id, y.id);
  }
  Sym neg(const Sym& x) {
    if (x.is_const) return cst(x.c.neg());
    SsaOp o; o.op = T_NEG; o.dst = new_value(); o.a = x.id;
    ops.push_back(o);
    return val(o.dst);
  }
  // L::batch_invert (loader.rs:241-248): each element replaced by its inverse, 0 stays 0
  void batch_invert(std::vector<Sym*> xs) {
    SsaOp o; o.op = T_BINV;
    for (Sym* x : xs) {
      if (x->is_const) { x->c = x->c.inv(); continue; }
      int d = new_value();
      o.binv.push_back({x->id, d});
      *x = val(d);
    }
    if (o.binv.empty()) return;
    if (o.binv.size() > SVK_BINV_MAX) {  // split: value-identical
      for (size_t i = 0; i < o.binv.size(); i += SVK_BINV_MAX) {
        SsaOp p; p.op = T_BINV;
        p.binv.assign(o.binv.begin() + i, o.binv.begin() + std::min(o.binv.size(), i + SVK_BINV_MAX));
        ops.push_back(p);
      }
    } else ops.push_back(o);
  }
  // LoadedScalar::pow_const (loader.rs:49-68)
  Sym pow_const(Sym base, u64 exp) {
    if (exp == 0) throw CompileError(SVK_INVALID_PROTOCOL, "pow_const(0)");
    while ((exp & 1) == 0) { base = mul(base, base); exp >>= 1; }
    Sym acc = base;
    while (exp > 1) {
      exp >>= 1;
      base = mul(base, base);
      if (exp & 1) acc = mul(acc, base);
    }
    return acc;
  }
  // LoadedScalar::powers(n) (loader.rs:71-78)
  std::vector<Sym> powers(const Sym& x, size_t n) {
    if (n == 0) throw CompileError(SVK_INVALID_PROTOCOL, "powers(0)");
    std::vector<Sym> out{cst(Fr::one())};
    Sym cur = x;
    for (size_t i = 1; i < n; i++) {
      out.push_back(cur);
      if (i + 1 < n) cur = mul(cur, x);
    }
    return out;
  }

  // ---- transcript (transcript/halo2.rs:198-261 over poseidon.rs:449-467)
  void common_scalar(const Sym& s) {
    if (transcript_kind == 1) {  // evm.rs:198-202
      SsaOp o; o.op = T_KABSORB_REG; o.a = materialize(s);
      ops.push_back(o);
      kbuf_len += 32;
      return;
    }
    tbuf.push_back(s);
  }
  void kabsorb_proof(u32 byte_offset) {
    SsaOp o; o.op = T_KABSORB_PROOF; o.imm = byte_offset / 32;
    ops.push_back(o);
    kbuf_len += 32;
  }
  Sym squeeze_challenge() {
    if (transcript_kind == 1) {  // evm.rs:172-182
      SsaOp q; q.op = T_KSQUEEZE; q.dst = new_value(); q.imm = kbuf_len == 32 ? 1 : 0;
      ops.push_back(q);
      kbuf_len = 32;
      n_perm++;
      return val(q.dst);
    }
    std::vector<Sym> buf;
    buf.swap(tbuf);
    bool exact = buf.size() % SVK_POSEIDON_RATE == 0;
    for (size_t i = 0; i < buf.size(); i += SVK_POSEIDON_RATE) {
      SsaOp o; o.op = T_PERM;
      size_t k = std::min((size_t)SVK_POSEIDON_RATE, buf.size() - i);
      o.imm = (u32)k;
      o.a = materialize(buf[i]);
      if (k > 1) o.b = materialize(buf[i + 1]);
      ops.push_back(o);
      n_perm++;
    }
    if (exact) { SsaOp o; o.op = T_PERM; o.imm = 0; ops.push_back(o); n_perm++; }
    SsaOp q; q.op = T_SQUEEZE; q.dst = new_value();
    ops.push_back(q);
    return val(q.dst);
  }
  Sym read_scalar() {
    if (transcript_kind == 1) {  // evm.rs:210-221: 32 B big-endian; the bytes read are the bytes absorbed
      SsaOp o; o.op = T_READ_SCALAR_BE; o.dst = new_value(); o.imm = cursor / 32;
      ops.push_back(o);
      kabsorb_proof(cursor);
      cursor += 32;
      return val(o.dst);
    }
    SsaOp o; o.op = T_READ_SCALAR; o.dst = new_value(); o.imm = cursor / 32;
    ops.push_back(o);
    cursor += 32;
    Sym s = val(o.dst);
    common_scalar(s);
    return s;
  }
  int read_ec_point() {  // returns the proof-point ordinal
    if (transcript_kind == 1) {  // evm.rs:223-242: x || y, 64 B big-endian, validated by k_load_points_be
      PointRead pr;
      pr.byte_offset = cursor;
      pr.val_x = pr.val_y = -1;
      points.push_back(pr);
      kabsorb_proof(cursor);
      kabsorb_proof(cursor + 32);
      cursor += 64;
      return (int)points.size() - 1;
    }
    PointRead pr;
    pr.byte_offset = cursor;
    pr.val_x = new_value();
    pr.val_y = new_value();
    cursor += 32;
    points.push_back(pr);
    common_scalar(val(pr.val_x));
    common_scalar(val(pr.val_y));
    return (int)points.size() - 1;
  }
  Sym instance(u32 flat_index) {
    SsaOp o; o.op = T_INSTANCE; o.dst = new_value(); o.imm = flat_index;
    ops.push_back(o);
    return val(o.dst);
  }
};

// ------------------------------------------------------------------ util/msm.rs:20-205 (symbolic)
struct SMsm {
  bool has_const = false;
  Sym constant;
  std::vector<Sym> scalars;
  std::vector<int> bases;
  static SMsm of_constant(const Sym& c) { SMsm m; m.has_const = true; m.constant = c; return m; }
  static SMsm of_base(TapeBuilder& tb, int base) { SMsm m; m.scalars.push_back(tb.cst(Fr::one())); m.bases.push_back(base); return m; }
  size_t size() const { return bases.size(); }
  void scale(TapeBuilder& tb, const Sym& f) {
    if (has_const) constant = tb.mul(constant, f);
    for (auto& s : scalars) s = tb.mul(s, f);
  }
  // msm.rs:88-95 dedups equal bases BY VALUE; at compile time only identity of the source is known
  // (same id => same value).  Distinct ids holding equal points simply stay two terms: same sum.
  void push(TapeBuilder& tb, const Sym& s, int base) {
    for (size_t i = 0; i < bases.size(); i++)
      if (bases[i] == base) { scalars[i] = tb.add(scalars[i], s); return; }
    scalars.push_back(s);
    bases.push_back(base);
  }
  void extend(TapeBuilder& tb, const SMsm& o) {
    if (has_const && o.has_const) constant = tb.add(constant, o.constant);
    else if (!has_const && o.has_const) { has_const = true; constant = o.constant; }
    for (size_t i = 0; i < o.bases.size(); i++) push(tb, o.scalars[i], o.bases[i]);
  }
  SMsm negated(TapeBuilder& tb) const {
    SMsm m = *this;
    if (m.has_const) m.constant = tb.neg(m.constant);
    for (auto& s : m.scalars) s = tb.neg(s);
    return m;
  }
  SMsm plus(TapeBuilder& tb, const SMsm& o) const { SMsm m = *this; m.extend(tb, o); return m; }
  SMsm minus(TapeBuilder& tb, const SMsm& o) const { SMsm m = *this; m.extend(tb, o.negated(tb)); return m; }
  SMsm times(TapeBuilder& tb, const Sym& f) const { SMsm m = *this; m.scale(tb, f); return m; }
};
inline SMsm smsm_sum(TapeBuilder& tb, const std::vector<SMsm>& v) {
  if (v.empty()) return SMsm();
  SMsm acc = v[0];
  for (size_t i = 1; i < v.size(); i++) acc.extend(tb, v[i]);
  return acc;
}

// util/arithmetic.rs:166-234
struct SFraction {
  bool has_numer = false;
  Sym numer, denom, eval;
  bool has_eval = false, inv = false;
  static SFraction make(const Sym& n, const Sym& d) { SFraction f; f.has_numer = true; f.numer = n; f.denom = d; return f; }
  static SFraction one_over(const Sym& d) { SFraction f; f.denom = d; return f; }
  Sym* denom_mut() {
    if (inv) return nullptr;
    inv = true;
    return &denom;
  }
  void evaluate(TapeBuilder& tb) {
    if (!inv) throw CompileError(-1, "Fraction::evaluate before denom_mut");
    if (!has_eval) {
      eval = has_numer ? tb.mul(numer, denom) : denom;
      has_numer = false;
      has_eval = true;
    }
  }
  const Sym& evaluated() const {
    if (!has_eval) throw CompileError(-1, "Fraction::evaluated before evaluate");
    return eval;
  }
};

// ------------------------------------------------------------------ compiled output
struct CompiledProtocol {
  int mos = 0;
  int transcript_kind = 0;
  bool verify_valid = true;     // false => Error::InvalidProtocol for every proof that reads fine
  std::string invalid_reason;
  std::vector<TapeOp> ops;      // physical registers
  u32 read_ops_end = 0;         // ops [0, read_ops_end) = PlonkProof::read; the rest = verify
  std::vector<uint16_t> aux;
  std::vector<Fr> consts;
  u32 n_regs = 0;
  std::vector<PointRead> points;   // val_x / val_y now hold PHYSICAL registers
  u32 proof_len = 0;               // bytes consumed by read_proof (trailing bytes are ignored)
  std::vector<u32> num_instance;
  u32 n_instances = 0;
  u32 n_challenges = 0;            // challenges + z + pcs challenges, in squeeze order
  u32 n_scalar_slots = 0;
  std::vector<MsmTerm> lhs, rhs;   // final `lhs.evaluate(Some(g))`, `rhs.evaluate(Some(g))`
  std::vector<svk_g1> preprocessed;
  int n_perm = 0;
  size_t n_fr_mul = 0;             // tape statistics
  // old accumulators (`protocol.accumulator_indices`, proof.rs:139-146): flat instance indices, 4 * acc_limbs per accumulator
  std::vector<u32> old_acc_idx;
  u32 n_old = 0, acc_limbs = 3, acc_bits = 88;
};

struct PcsQuery {
  u32 poly;
  Fr shift;
  Sym eval;
};

class Compiler {
 public:
  const ProtocolDesc& P;
  int mos;
  TapeBuilder tb;
  std::vector<Sym> challenges_out;  // every squeezed challenge in order (for the ABI's out_challenges)

  Compiler(const ProtocolDesc& p, int mos_, int transcript_kind = 0) : P(p), mos(mos_) { tb.transcript_kind = transcript_kind; }

  Sym squeeze() {
    Sym s = tb.squeeze_challenge();
    challenges_out.push_back(s);
    return s;
  }

  // ---------------- proof.rs:52-153
  std::vector<std::vector<Sym>> instances;
  std::vector<int> witnesses, quotients;  // proof-point ordinals
  std::vector<Sym> challenges, evaluations;
  Sym z;
  // pcs proof
  Sym mu, gamma, z_prime, v, u;
  int w = -1, w_prime = -1;
  std::vector<int> ws;

  int base_of_point(int ordinal) const { return (int)P.preprocessed.size() + ordinal; }

  void read_proof() {
    if (P.has_initial_state) tb.common_scalar(tb.cst(P.initial_state));
    u32 flat = 0;
    for (u32 n : P.num_instance) {
      std::vector<Sym> col;
      for (u32 i = 0; i < n; i++) col.push_back(tb.instance(flat++));
      instances.push_back(col);
    }
    tb.n_instances = flat;
    for (auto& col : instances)
      for (auto& x : col) tb.common_scalar(x);
    size_t phases = std::min(P.num_witness.size(), P.num_challenge.size());  // `zip`
    for (size_t ph = 0; ph < phases; ph++) {
      for (u32 i = 0; i < P.num_witness[ph]; i++) witnesses.push_back(tb.read_ec_point());
      for (u32 i = 0; i < P.num_challenge[ph]; i++) challenges.push_back(squeeze());
    }
    for (size_t i = 0; i < P.num_chunk(); i++) quotients.push_back(tb.read_ec_point());
    z = squeeze();
    for (size_t i = 0; i < P.evaluations.size(); i++) evaluations.push_back(tb.read_scalar());
    if (mos == SVK_MOS_BDFG21) {  // bdfg21.rs:101-114
      mu = squeeze();
      gamma = squeeze();
      w = tb.read_ec_point();
      z_prime = squeeze();
      w_prime = tb.read_ec_point();
    } else {  // gwc19.rs:100-108
      v = squeeze();
      size_t nsets = gwc_sets(empty_queries()).size();
      for (size_t i = 0; i < nsets; i++) ws.push_back(tb.read_ec_point());
      u = squeeze();
    }
  }

  std::vector<PcsQuery> empty_queries() {  // proof.rs:156-165
    std::vector<PcsQuery> out;
    for (auto& q : P.queries) out.push_back({q.poly, P.rotate_one(q.rot), Sym()});
    return out;
  }

  // ---------------- protocol.rs:70-98
  std::vector<int32_t> langranges() {
    std::set<int32_t> lag;
    std::set<Query> qs;
    expr_collect(*P.numerator, lag, qs);
    std::vector<int32_t> out(lag.begin(), lag.end());
    size_t offset = P.preprocessed.size();
    int32_t mn = 0, mx = 0;
    for (auto& q : qs) {
      if (q.poly < offset || q.poly >= offset + P.num_instance.size()) continue;
      if (q.rot < mn) mn = q.rot;
      else if (q.rot > mx) mx = q.rot;
    }
    int32_t max_len = 0;
    for (u32 n : P.num_instance) max_len = std::max(max_len, (int32_t)n);
    for (int32_t i = -mx; i < max_len + (mn < 0 ? -mn : mn); i++) out.push_back(i);
    return out;
  }

  // ---------------- CommonPolynomialEvaluation (protocol.rs:201-279)
  Sym zn, zn_minus_one;
  SFraction zn_minus_one_inv;
  std::map<int32_t, SFraction> lagrange;

  void common_poly_eval() {
    zn = tb.pow_const(z, P.n);
    std::vector<int32_t> ls = langranges();
    std::sort(ls.begin(), ls.end());
    ls.erase(std::unique(ls.begin(), ls.end()), ls.end());
    zn_minus_one = tb.sub(zn, tb.cst(Fr::one()));
    zn_minus_one_inv = SFraction::one_over(zn_minus_one);
    Sym numer = tb.mul(zn_minus_one, tb.cst(P.n_inv));
    for (int32_t i : ls) {
      Sym omega = tb.cst(P.rotate_one(i));
      lagrange[i] = SFraction::make(tb.mul(numer, omega), tb.sub(z, omega));
    }
    std::vector<Sym*> denoms;  // plonk.rs:68
    for (auto& kv : lagrange) denoms.push_back(kv.second.denom_mut());
    denoms.push_back(zn_minus_one_inv.denom_mut());
    tb.batch_invert(denoms);
    for (auto& kv : lagrange) kv.second.evaluate(tb);
    zn_minus_one_inv.evaluate(tb);
  }
  const Sym& lagrange_eval(int32_t i) {
    auto it = lagrange.find(i);
    if (it == lagrange.end()) throw CompileError(SVK_INVALID_PROTOCOL, "missing Lagrange evaluation");  // `.unwrap()` panic
    return it->second.evaluated();
  }

  // ---------------- proof.rs:283-318
  std::map<Query, Sym> eval_map;
  void build_evaluations() {
    std::set<int32_t> lag;
    std::set<Query> qs;
    expr_collect(*P.numerator, lag, qs);
    size_t offset = P.preprocessed.size();
    for (auto& q : qs) {
      if (q.poly < offset || q.poly >= offset + P.num_instance.size()) continue;
      auto& inst = instances[q.poly - offset];
      Sym acc = tb.cst(Fr::zero());
      bool first = true;
      for (size_t k2 = 0; k2 < inst.size(); k2++) {  // loader.sum_products
        Sym t = tb.mul(inst[k2], lagrange_eval(-q.rot + (int32_t)k2));
        acc = first ? t : tb.add(acc, t);
        first = false;
      }
      eval_map[q] = acc;
    }
    for (size_t i = 0; i < P.evaluations.size(); i++) eval_map[P.evaluations[i]] = evaluations[i];
  }

  // ---------------- proof.rs:179-281
  std::vector<SMsm> commitments;
  SMsm eval_expr(const Expr& e) {  // Expression::evaluate over the Msm algebra (proof.rs:203-236)
    switch (e.tag) {
      case Expr::CONST: return SMsm::of_constant(tb.cst(e.c));
      case Expr::IDENTITY: return SMsm::of_constant(z);
      case Expr::LAGRANGE: return SMsm::of_constant(lagrange_eval(e.i));
      case Expr::POLY: {
        auto it = eval_map.find(e.q);
        if (it != eval_map.end()) return SMsm::of_constant(it->second);
        if (e.q.rot == 0 && e.q.poly < commitments.size()) return commitments[e.q.poly];
        throw CompileError(SVK_INVALID_PROTOCOL, "Missing query");
      }
      case Expr::CHALLENGE:
        if (e.i >= 0 && (size_t)e.i < challenges.size()) return SMsm::of_constant(challenges[e.i]);
        throw CompileError(SVK_INVALID_PROTOCOL, "Missing challenge");
      case Expr::NEG: return eval_expr(*e.kids[0]).negated(tb);
      case Expr::SUM: {
        SMsm a = eval_expr(*e.kids[0]);
        SMsm b = eval_expr(*e.kids[1]);
        return a.plus(tb, b);
      }
      case Expr::PRODUCT: {
        SMsm a = eval_expr(*e.kids[0]);
        SMsm b = eval_expr(*e.kids[1]);
        return product(a, b);
      }
      case Expr::SCALED: return eval_expr(*e.kids[0]).times(tb, tb.cst(e.c));
      case Expr::DISTRIBUTE: {
        size_t n = e.kids.size() - 1;
        if (n == 1) return eval_expr(*e.kids[0]);
        SMsm acc = eval_expr(*e.kids[0]);
        SMsm scalar = eval_expr(*e.kids[n]);
        for (size_t i = 1; i < n; i++) acc = product(acc, scalar).plus(tb, eval_expr(*e.kids[i]));
        return acc;
      }
    }
    throw CompileError(-1, "bad expression");
  }
  SMsm product(const SMsm& a, const SMsm& b) {  // proof.rs:227-234
    if (a.size() == 0) {
      if (!a.has_const) throw CompileError(SVK_INVALID_PROTOCOL, "try_into_constant on empty Msm");  // `.unwrap()` panic
      return b.times(tb, a.constant);
    }
    if (b.size() == 0) {
      if (!b.has_const) throw CompileError(SVK_INVALID_PROTOCOL, "try_into_constant on empty Msm");
      return a.times(tb, b.constant);
    }
    throw CompileError(SVK_INVALID_PROTOCOL, "Invalid linearization");
  }

  void build_commitments() {
    for (size_t i = 0; i < P.preprocessed.size(); i++) commitments.push_back(SMsm::of_base(tb, (int)i));
    for (size_t i = 0; i < P.num_instance.size(); i++) commitments.push_back(SMsm());
    for (int wi : witnesses) commitments.push_back(SMsm::of_base(tb, base_of_point(wi)));
    SMsm numerator = eval_expr(*P.numerator);
    Query quotient_query{(u32)(P.preprocessed.size() + P.num_instance.size() + witnesses.size()), 0};
    std::vector<Sym> coeffs = tb.powers(tb.pow_const(zn, P.chunk_degree), quotients.empty() ? 1 : quotients.size());
    std::vector<SMsm> chunks;
    for (size_t i = 0; i < quotients.size(); i++) chunks.push_back(SMsm::of_base(tb, base_of_point(quotients[i])).times(tb, coeffs[i]));
    SMsm quotient = smsm_sum(tb, chunks);
    if (P.linearization == 1) {  // WithoutConstant
      Query lin{quotient_query.poly + 1, 0};
      SMsm msm = numerator;
      bool hc = msm.has_const;
      Sym c = hc ? msm.constant : tb.cst(Fr::zero());
      msm.has_const = false;
      commitments.push_back(quotient);
      commitments.push_back(msm);
      auto it = eval_map.find(lin);
      if (it == eval_map.end()) throw CompileError(SVK_INVALID_PROTOCOL, "missing linearization evaluation");
      eval_map[quotient_query] = tb.mul(tb.add(c, it->second), zn_minus_one_inv.evaluated());
    } else if (P.linearization == 2) {  // MinusVanishingTimesQuotient
      SMsm msm = numerator.minus(tb, quotient.times(tb, zn_minus_one));
      Sym c = msm.has_const ? msm.constant : tb.cst(Fr::zero());
      msm.has_const = false;
      commitments.push_back(msm);
      eval_map[quotient_query] = c;
    } else {
      commitments.push_back(quotient);
      if (numerator.size() != 0) throw CompileError(SVK_INVALID_PROTOCOL, "Invalid linearization");
      if (!numerator.has_const) throw CompileError(SVK_INVALID_PROTOCOL, "try_into_constant on empty Msm");
      eval_map[quotient_query] = tb.mul(numerator.constant, zn_minus_one_inv.evaluated());
    }
  }

  std::vector<PcsQuery> build_queries() {  // proof.rs:167-177
    std::vector<PcsQuery> qs = empty_queries();
    for (size_t i = 0; i < qs.size(); i++) {
      auto it = eval_map.find(P.queries[i]);
      if (it == eval_map.end()) throw CompileError(SVK_INVALID_PROTOCOL, "query without evaluation");  // `.unwrap()` panic
      qs[i].eval = it->second;
      eval_map.erase(it);
    }
    return qs;
  }

  // ---------------- SHPLONK (bdfg21.rs)
  struct BSet {
    std::vector<Fr> shifts;
    std::vector<u32> polys;
    std::vector<std::vector<Sym>> evals;
  };
  static bool same_shift_set(const std::vector<Fr>& a, const std::vector<Fr>& b) {
    std::set<Fr, FrLess> sa(a.begin(), a.end()), sb(b.begin(), b.end());
    if (sa.size() != sb.size()) return false;
    auto ia = sa.begin();
    auto ib = sb.begin();
    for (; ia != sa.end(); ++ia, ++ib)
      if (!(*ia == *ib)) return false;
    return true;
  }
  static std::vector<BSet> bdfg_sets(const std::vector<PcsQuery>& queries) {  // bdfg21.rs:117-167
    struct PS { u32 poly; std::vector<Fr> shifts; std::vector<Sym> evals; };
    std::vector<PS> ps;
    for (auto& q : queries) {
      auto it = std::find_if(ps.begin(), ps.end(), [&](const PS& x) { return x.poly == q.poly; });
      if (it != ps.end()) {
        if (std::find(it->shifts.begin(), it->shifts.end(), q.shift) == it->shifts.end()) {
          it->shifts.push_back(q.shift);
          it->evals.push_back(q.eval);
        }
      } else ps.push_back({q.poly, {q.shift}, {q.eval}});
    }
    std::vector<BSet> sets;
    for (auto& p : ps) {
      auto it = std::find_if(sets.begin(), sets.end(), [&](const BSet& s) { return same_shift_set(s.shifts, p.shifts); });
      if (it != sets.end()) {
        if (std::find(it->polys.begin(), it->polys.end(), p.poly) == it->polys.end()) {
          it->polys.push_back(p.poly);
          std::vector<Sym> ev;
          for (auto& lhs : it->shifts) {
            size_t idx = std::find(p.shifts.begin(), p.shifts.end(), lhs) - p.shifts.begin();
            ev.push_back(p.evals[idx]);
          }
          it->evals.push_back(ev);
        }
      } else sets.push_back({p.shifts, {p.poly}, {p.evals}});
    }
    return sets;
  }
  struct BCoeff {
    Sym z_s;
    std::vector<SFraction> eval_coeffs;
    bool has_cc = false;
    SFraction commitment_coeff;
    bool has_rc = false;
    SFraction r_eval_coeff;
  };
  void bdfg_verify(const std::vector<PcsQuery>& queries, SMsm& lhs, SMsm& rhs) {  // bdfg21.rs:47-79
    std::vector<BSet> sets = bdfg_sets(queries);
    if (sets.empty()) throw CompileError(SVK_INVALID_PROTOCOL, "no queries");
    // query_set_coeffs (bdfg21.rs:169-219)
    std::set<Fr, FrLess> superset;
    size_t size = 2;
    for (auto& s : sets) {
      superset.insert(s.shifts.begin(), s.shifts.end());
      size = std::max(size, s.shifts.size());
    }
    std::vector<Sym> powers_of_z = tb.powers(z, size);
    std::map<Fr, Sym, FrLess> zpm;
    for (auto& sh : superset) zpm[sh] = tb.sub(z_prime, tb.mul(z, tb.cst(sh)));
    std::vector<BCoeff> coeffs;
    bool have_zs1 = false;
    Sym z_s_1;
    for (auto& s : sets) {  // QuerySetCoeff::new (bdfg21.rs:276-329)
      BCoeff c;
      size_t kk = s.shifts.size();
      const Sym& zz = powers_of_z[1];
      const Sym& z_pow = powers_of_z[kk - 1];
      for (size_t j = 0; j < kk; j++) {
        Fr nep = Fr::one();
        bool any = false;
        for (size_t i = 0; i < kk; i++)
          if (i != j) { Fr d = s.shifts[j] - s.shifts[i]; nep = any ? nep * d : d; any = true; }
        // sum_products_with_coeff([(nep, z^(k-1), z'), (-(nep*shift), z^(k-1), z)])
        Sym t1 = tb.mul(tb.mul(tb.cst(nep), z_pow), z_prime);
        Sym t2 = tb.mul(tb.mul(tb.cst((nep * s.shifts[j]).neg()), z_pow), zz);
        c.eval_coeffs.push_back(SFraction::one_over(tb.add(t1, t2)));
      }
      Sym zs = tb.cst(Fr::one());
      for (auto& sh : s.shifts) zs = tb.mul(zs, zpm[sh]);
      c.z_s = zs;
      if (have_zs1) { c.has_cc = true; c.commitment_coeff = SFraction::make(z_s_1, zs); }
      if (!have_zs1) { have_zs1 = true; z_s_1 = zs; }
      coeffs.push_back(c);
    }
    {  // first batch_invert (bdfg21.rs:214, denoms() :331-339)
      std::vector<Sym*> d;
      for (auto& c : coeffs) {
        for (auto& f : c.eval_coeffs) d.push_back(f.denom_mut());
        if (c.has_cc) d.push_back(c.commitment_coeff.denom_mut());
      }
      tb.batch_invert(d);
    }
    {  // second batch_invert (bdfg21.rs:215, denoms() :341-359)
      std::vector<Sym*> d;
      for (auto& c : coeffs) {
        for (auto& f : c.eval_coeffs) f.evaluate(tb);
        if (c.has_cc) c.commitment_coeff.evaluate(tb);
        Sym sum = c.eval_coeffs[0].evaluated();
        for (size_t j = 1; j < c.eval_coeffs.size(); j++) sum = tb.add(sum, c.eval_coeffs[j].evaluated());
        c.r_eval_coeff = c.has_cc ? SFraction::make(c.commitment_coeff.evaluated(), sum) : SFraction::one_over(sum);
        c.has_rc = true;
        d.push_back(c.r_eval_coeff.denom_mut());
      }
      tb.batch_invert(d);
    }
    for (auto& c : coeffs) c.r_eval_coeff.evaluate(tb);

    size_t max_polys = 0;
    for (auto& s : sets) max_polys = std::max(max_polys, s.polys.size());
    std::vector<Sym> powers_of_mu = tb.powers(mu, max_polys);
    std::vector<Sym> powers_of_gamma = tb.powers(gamma, sets.size());
    std::vector<SMsm> msms;
    for (size_t si = 0; si < sets.size(); si++) {  // QuerySet::msm (bdfg21.rs:229-259)
      auto& s = sets[si];
      auto& c = coeffs[si];
      std::vector<SMsm> terms;
      for (size_t l = 0; l < s.polys.size(); l++) {
        if (s.polys[l] >= commitments.size()) throw CompileError(SVK_INVALID_PROTOCOL, "query poly out of range");
        SMsm commitment = c.has_cc ? commitments[s.polys[l]].times(tb, c.commitment_coeff.evaluated()) : commitments[s.polys[l]];
        Sym acc;
        for (size_t j = 0; j < c.eval_coeffs.size(); j++) {
          Sym t = tb.mul(c.eval_coeffs[j].evaluated(), s.evals[l][j]);
          acc = j == 0 ? t : tb.add(acc, t);
        }
        Sym r_eval = tb.mul(acc, c.r_eval_coeff.evaluated());
        terms.push_back(commitment.minus(tb, SMsm::of_constant(r_eval)).times(tb, powers_of_mu[l]));
      }
      msms.push_back(smsm_sum(tb, terms).times(tb, powers_of_gamma[si]));
    }
    SMsm f = smsm_sum(tb, msms).minus(tb, SMsm::of_base(tb, base_of_point(w)).times(tb, coeffs[0].z_s));
    rhs = SMsm::of_base(tb, base_of_point(w_prime));
    lhs = f.plus(tb, rhs.times(tb, z_prime));
  }

  // ---------------- GWC (gwc19.rs)
  struct GSet {
    Fr shift;
    std::vector<u32> polys;
    std::vector<Sym> evals;
  };
  static std::vector<GSet> gwc_sets(const std::vector<PcsQuery>& queries) {  // gwc19.rs:140-158
    std::vector<GSet> sets;
    for (auto& q : queries) {
      auto it = std::find_if(sets.begin(), sets.end(), [&](const GSet& s) { return s.shift == q.shift; });
      if (it != sets.end()) { it->polys.push_back(q.poly); it->evals.push_back(q.eval); }
      else sets.push_back({q.shift, {q.poly}, {q.eval}});
    }
    return sets;
  }
  void gwc_verify(const std::vector<PcsQuery>& queries, SMsm& lhs, SMsm& rhs) {  // gwc19.rs:43-80
    std::vector<GSet> sets = gwc_sets(queries);
    if (sets.empty()) throw CompileError(SVK_INVALID_PROTOCOL, "no queries");
    std::vector<Sym> powers_of_u = tb.powers(u, sets.size());
    size_t max_polys = 0;
    for (auto& s : sets) max_polys = std::max(max_polys, s.polys.size());
    std::vector<Sym> powers_of_v = tb.powers(v, max_polys);
    std::vector<SMsm> per_set;
    for (size_t si = 0; si < sets.size(); si++) {
      auto& s = sets[si];
      std::vector<SMsm> terms;
      for (size_t l = 0; l < s.polys.size(); l++) {
        if (s.polys[l] >= commitments.size()) throw CompileError(SVK_INVALID_PROTOCOL, "query poly out of range");
        terms.push_back(commitments[s.polys[l]].minus(tb, SMsm::of_constant(s.evals[l])).times(tb, powers_of_v[l]));
      }
      per_set.push_back(smsm_sum(tb, terms).times(tb, powers_of_u[si]));
    }
    SMsm f = smsm_sum(tb, per_set);
    std::vector<SMsm> rhs_terms, lhs_terms;
    for (size_t si = 0; si < sets.size() && si < ws.size(); si++) {
      SMsm uw = SMsm::of_base(tb, base_of_point(ws[si])).times(tb, powers_of_u[si]);
      rhs_terms.push_back(uw);
      Sym z_omega = tb.mul(tb.cst(sets[si].shift), z);
      lhs_terms.push_back(uw.times(tb, z_omega));
    }
    lhs = f.plus(tb, smsm_sum(tb, lhs_terms));
    rhs = smsm_sum(tb, rhs_terms);
  }

  // ---------------- drive + register allocation
  CompiledProtocol run() {
    CompiledProtocol out;
    out.mos = mos;
    read_proof();
    size_t read_ssa_end = tb.ops.size();
    SMsm lhs, rhs;
    try {
      common_poly_eval();
      build_evaluations();
      build_commitments();
      std::vector<PcsQuery> qs = build_queries();
      if (mos == SVK_MOS_BDFG21) bdfg_verify(qs, lhs, rhs);
      else gwc_verify(qs, lhs, rhs);
    } catch (CompileError& e) {
      if (e.kind != SVK_INVALID_PROTOCOL) throw;
      out.verify_valid = false;
      out.invalid_reason = e.what();
      tb.ops.resize(read_ssa_end);
      lhs = SMsm();
      rhs = SMsm();
    }
    // outputs: challenges, then MSM scalars.  `evaluate(Some(g))` prepends (constant, g) (msm.rs:70-77)
    std::vector<SsaOp> tail;
    for (size_t i = 0; i < challenges_out.size(); i++) {
      SsaOp o; o.op = T_OUT_CHALLENGE; o.imm = (u32)i; o.a = tb.materialize(challenges_out[i]);
      tb.ops.push_back(o);
    }
    out.n_challenges = (u32)challenges_out.size();
    u32 slot = 0;
    auto flatten = [&](const SMsm& m, std::vector<MsmTerm>& terms) {
      if (m.has_const) terms.push_back({SVK_BASE_G, m.constant, -1});
      for (size_t i = 0; i < m.bases.size(); i++) terms.push_back({m.bases[i], m.scalars[i], -1});
      for (auto& t : terms) {
        if (t.scalar.is_const && t.scalar.c == Fr::one()) continue;
        t.slot = (int)slot++;
        SsaOp o; o.op = T_OUT_SCALAR; o.imm = (u32)t.slot; o.a = tb.materialize(t.scalar);
        tb.ops.push_back(o);
      }
    };
    if (out.verify_valid) {
      flatten(lhs, out.lhs);
      flatten(rhs, out.rhs);
    }
    out.n_scalar_slots = slot;
    allocate(out, read_ssa_end);
    out.proof_len = tb.cursor;
    out.num_instance = P.num_instance;
    out.n_instances = tb.n_instances;
    out.preprocessed = P.preprocessed;
    out.n_perm = tb.n_perm;
    return out;
  }

  // Linear-scan allocation of physical registers over the SSA tape.  A destination never aliases an
  // operand of the same op (T_BINV reads its sources after writing destinations).
  void allocate(CompiledProtocol& out, size_t read_ssa_end) {
    auto& ops = tb.ops;
    std::vector<int> last_use(tb.n_values, -1);
    auto use = [&](int v, int at) { if (v >= 0) last_use[v] = std::max(last_use[v], at); };
    for (size_t i = 0; i < ops.size(); i++) {
      auto& o = ops[i];
      switch (o.op) {
        case T_ADD: case T_SUB: case T_MUL: use(o.a, (int)i); use(o.b, (int)i); break;
        case T_NEG: case T_ADDC: case T_SUBC: case T_CSUB: case T_MULC: case T_OUT_SCALAR: case T_OUT_CHALLENGE: case T_KABSORB_REG:
          use(o.a, (int)i);
          break;
        case T_PERM: if (o.imm >= 1) use(o.a, (int)i); if (o.imm >= 2) use(o.b, (int)i); break;
        case T_BINV: for (auto& pr : o.binv) use(pr.first, (int)i); break;
        default: break;
      }
    }
    std::vector<int> phys(tb.n_values, -1);
    std::vector<int> free_list;
    int n_phys = 0;
    auto alloc = [&]() {
      if (!free_list.empty()) { int r = free_list.back(); free_list.pop_back(); return r; }
      return n_phys++;
    };
    // values written by the decompress kernel before the tape starts
    for (auto& pr : tb.points)
      if (pr.val_x >= 0) { phys[pr.val_x] = alloc(); phys[pr.val_y] = alloc(); }
    auto release_dead = [&](int v, int at) {
      if (v >= 0 && last_use[v] == at && phys[v] >= 0) { free_list.push_back(phys[v]); last_use[v] = -2; }
    };
    auto P16 = [&](int v) {
      if (v < 0 || phys[v] < 0) throw CompileError(-1, "internal: use of unallocated value");
      if (phys[v] > 65535) throw CompileError(-1, "too many registers");
      return (uint16_t)phys[v];
    };
    for (size_t i = 0; i < ops.size(); i++) {
      if (i == read_ssa_end) out.read_ops_end = (u32)out.ops.size();
      auto& o = ops[i];
      TapeOp t{o.op, 0, 0, 0};
      int at = (int)i;
      switch (o.op) {
        case T_CONST: phys[o.dst] = alloc(); t.dst = P16(o.dst); t.a = (uint16_t)o.a; break;
        case T_ADD: case T_SUB: case T_MUL:
          phys[o.dst] = alloc(); t.dst = P16(o.dst); t.a = P16(o.a); t.b = P16(o.b);
          release_dead(o.a, at); release_dead(o.b, at);
          if (o.op == T_MUL) out.n_fr_mul++;
          break;
        case T_NEG:
          phys[o.dst] = alloc(); t.dst = P16(o.dst); t.a = P16(o.a); release_dead(o.a, at); break;
        case T_ADDC: case T_SUBC: case T_CSUB: case T_MULC:
          phys[o.dst] = alloc(); t.dst = P16(o.dst); t.a = P16(o.a); t.b = (uint16_t)o.b; release_dead(o.a, at);
          if (o.op == T_MULC) out.n_fr_mul++;
          break;
        case T_BINV: {
          t.a = (uint16_t)out.aux.size();
          t.b = (uint16_t)o.binv.size();
          if (out.aux.size() + 2 * o.binv.size() > 65535) throw CompileError(-1, "aux table overflow");
          for (auto& pr : o.binv) phys[pr.second] = alloc();
          for (auto& pr : o.binv) { out.aux.push_back(P16(pr.first)); out.aux.push_back(P16(pr.second)); }
          for (auto& pr : o.binv) release_dead(pr.first, at);
          out.n_fr_mul += 3 * o.binv.size() + 380;
          break;
        }
        case T_READ_SCALAR: case T_INSTANCE: case T_READ_SCALAR_BE:
          phys[o.dst] = alloc(); t.dst = P16(o.dst); t.a = (uint16_t)(o.imm & 0xffff); t.b = (uint16_t)(o.imm >> 16); break;
        case T_KABSORB_REG: t.a = P16(o.a); release_dead(o.a, at); break;
        case T_KABSORB_PROOF: t.a = (uint16_t)(o.imm & 0xffff); t.b = (uint16_t)(o.imm >> 16); break;
        case T_KSQUEEZE: phys[o.dst] = alloc(); t.dst = P16(o.dst); t.a = (uint16_t)o.imm; break;
        case T_PERM:
          t.dst = (uint16_t)o.imm;
          if (o.imm >= 1) t.a = P16(o.a);
          if (o.imm >= 2) t.b = P16(o.b);
          if (o.imm >= 1) release_dead(o.a, at);
          if (o.imm >= 2) release_dead(o.b, at);
          out.n_fr_mul += 600;
          break;
        case T_SQUEEZE: phys[o.dst] = alloc(); t.dst = P16(o.dst); break;
        case T_OUT_SCALAR: case T_OUT_CHALLENGE:
          t.dst = (uint16_t)o.imm; t.a = P16(o.a); release_dead(o.a, at); break;
        default: throw CompileError(-1, "internal: unknown SSA op");
      }
      // a value that is never used can be recycled immediately
      if (o.dst >= 0 && last_use[o.dst] == -1 && phys[o.dst] >= 0) free_list.push_back(phys[o.dst]);
      if (o.op == T_BINV)
        for (auto& pr : o.binv)
          if (last_use[pr.second] == -1) free_list.push_back(phys[pr.second]);
      out.ops.push_back(t);
    }
    if (read_ssa_end >= ops.size()) out.read_ops_end = (u32)out.ops.size();
    out.n_regs = (u32)n_phys;
    out.consts = tb.consts;
    out.points = tb.points;
    for (auto& pr : out.points)
      if (pr.val_x >= 0) { pr.val_x = phys[pr.val_x]; pr.val_y = phys[pr.val_y]; }
  }
};

inline CompiledProtocol compile_protocol_desc(const ProtocolDesc& p, int mos, int transcript_kind);
inline CompiledProtocol compile_protocol(const uint8_t* blob, size_t len, int mos, int transcript_kind = 0) {
  ProtocolDesc p = parse_protocol(blob, len);
  return compile_protocol_desc(p, mos, transcript_kind);
}
inline CompiledProtocol compile_protocol_bincode(const uint8_t* bytes, size_t len, int fe_encoding, int mos, int transcript_kind,
                                                 size_t* consumed, int* fe_used = nullptr) {
  ProtocolDesc p = parse_protocol_bincode(bytes, len, fe_encoding, consumed, fe_used);
  return compile_protocol_desc(p, mos, transcript_kind);
}
inline CompiledProtocol compile_protocol_desc(const ProtocolDesc& p, int mos, int transcript_kind) {
  if (mos != SVK_MOS_BDFG21 && mos != SVK_MOS_GWC19) throw CompileError(-1, "unknown multi-open scheme");
  if (transcript_kind != 0 && transcript_kind != 1) throw CompileError(-1, "unknown transcript kind");
  Compiler c(p, mos, transcript_kind);
  CompiledProtocol out = c.run();
  out.transcript_kind = transcript_kind;
  std::vector<u32> col_off(p.num_instance.size() + 1, 0);
  for (size_t i = 0; i < p.num_instance.size(); i++) col_off[i + 1] = col_off[i] + p.num_instance[i];
  for (auto& v : p.accumulator_indices)
    for (auto& ij : v) out.old_acc_idx.push_back(col_off[ij.first] + ij.second);
  out.n_old = (u32)p.accumulator_indices.size();
  out.acc_limbs = p.acc_limbs;
  out.acc_bits = p.acc_bits;
  return out;
}

}  // namespace svk_host
