/* oracle/c/cref.c -- C restatement of the reference's NativeLoader CPU path.  TEST INFRASTRUCTURE:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * The reference (Rust) cannot be built here (no cargo/rustc; arithmetic in un-vendored halo2curves
 * 0.3.1, Cargo.lock:1803-1826), so this is `kind: "port"`: 4x64-bit Montgomery limbs (unsigned
 * __int128), and -- deliberately -- the REFERENCE's algorithms, not the GPU's:
 *   scalar mul   : MSB-first double-and-add over all 256 bits, add always computed then selected
 *                  (halo2curves `impl Mul<&Fr> for &G1Affine`), one per (scalar, base) pair and summed
 *                  = NativeLoader::multi_scalar_multiplication, snark-verifier/src/loader/native.rs:61-71
 *   inversion    : Fermat x^(r-2), ONE PER ELEMENT (loader.rs:241-248 is not batched natively)
 *   transcript   : Poseidon sponge, optimised permutation (util/hash/poseidon.rs:455-501), strictly serial
 *   fold         : KzgAs::verify, accumulation.rs:45-61 (naive MSM again)
 *   decide       : 2-pair Miller loop + final exponentiation, pcs/kzg/decider.rs:60-68
 * The per-proof operation sequence itself is not re-derived here: the Python oracle (oracle/loader.py
 * `Tracer`) records exactly the primitive operations NativeLoader performs for one proof of the protocol,
 * and `cref_replay` executes that trace for every proof.  Bit-exact against the Python oracle
 * (tests/test_oracle_c.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef struct { u64 v[4]; } fe;

typedef struct { u64 m[4]; u64 inv; fe one; fe r2; } field;

static const field FQ = {{0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}, 0x87d20782e4866389ULL,
  {{0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL}},
  {{0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL}}};
static const field FR = {{0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}, 0xc2e1f593efffffffULL,
  {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}},
  {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}}};

static inline int fe_is_zero(const fe* a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b) { return ((a->v[0] ^ b->v[0]) | (a->v[1] ^ b->v[1]) | (a->v[2] ^ b->v[2]) | (a->v[3] ^ b->v[3])) == 0; }
static inline int geq(const u64* a, const u64* m) {
  for (int i = 3; i >= 0; i--) { if (a[i] > m[i]) return 1; if (a[i] < m[i]) return 0; }
  return 1;
}
static inline void sub_mod_raw(u64* r, const u64* a, const u64* m) {
  u128 b = 0;
  for (int i = 0; i < 4; i++) { u128 t = (u128)a[i] - m[i] - (u64)b; r[i] = (u64)t; b = (t >> 64) & 1; }
}
static inline void f_add(const field* F, fe* r, const fe* a, const fe* b) {
  u128 c = 0; u64 t[4];
  for (int i = 0; i < 4; i++) { c += (u128)a->v[i] + b->v[i]; t[i] = (u64)c; c >>= 64; }
  if (geq(t, F->m)) sub_mod_raw(r->v, t, F->m); else memcpy(r->v, t, 32);
}
static inline void f_sub(const field* F, fe* r, const fe* a, const fe* b) {
  u128 br = 0; u64 t[4];
  for (int i = 0; i < 4; i++) { u128 x = (u128)a->v[i] - b->v[i] - (u64)br; t[i] = (u64)x; br = (x >> 64) & 1; }
  if (br) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)t[i] + F->m[i]; t[i] = (u64)c; c >>= 64; } }
  memcpy(r->v, t, 32);
}
static inline void f_neg(const field* F, fe* r, const fe* a) { fe z = {{0, 0, 0, 0}}; f_sub(F, r, &z, a); }
static inline void f_mul(const field* F, fe* r, const fe* a, const fe* b) {  /* CIOS */
  u64 t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a->v[j] * b->v[i] + t[j]; t[j] = (u64)c; c >>= 64; }
    c += t[4]; t[4] = (u64)c; t[5] = (u64)(c >> 64);
    u64 m = t[0] * F->inv;
    c = ((u128)m * F->m[0] + t[0]) >> 64;
    for (int j = 1; j < 4; j++) { c += (u128)m * F->m[j] + t[j]; t[j - 1] = (u64)c; c >>= 64; }
    c += t[4]; t[3] = (u64)c; t[4] = t[5] + (u64)(c >> 64);
  }
  if (t[4] || geq(t, F->m)) sub_mod_raw(r->v, t, F->m); else memcpy(r->v, t, 32);
}
static inline void f_sqr(const field* F, fe* r, const fe* a) { f_mul(F, r, a, a); }
static void f_pow(const field* F, fe* r, const fe* a, const u64* e) {
  fe acc = F->one;
  for (int w = 3; w >= 0; w--) for (int b = 63; b >= 0; b--) { f_sqr(F, &acc, &acc); if ((e[w] >> b) & 1) f_mul(F, &acc, &acc, a); }
  *r = acc;
}
static void f_inv(const field* F, fe* r, const fe* a) { u64 e[4] = {F->m[0] - 2, F->m[1], F->m[2], F->m[3]}; f_pow(F, r, a, e); }
static void f_to_mont(const field* F, fe* r, const fe* a) { f_mul(F, r, a, &F->r2); }
static void f_from_mont(const field* F, fe* r, const fe* a) { fe o = {{1, 0, 0, 0}}; f_mul(F, r, a, &o); }
static int f_canonical(const field* F, const fe* a) { return !geq(a->v, F->m); }

/* ---------------------------------------------------------------- G1 */
typedef struct { fe x, y; int inf; } g1a;
typedef struct { fe X, Y, Z; } g1j;
#define Q (&FQ)
static void j_identity(g1j* r) { r->X = FQ.one; r->Y = FQ.one; memset(&r->Z, 0, 32); }
static void j_dbl(g1j* r, const g1j* p) {
  if (fe_is_zero(&p->Z)) { *r = *p; return; }
  fe A, B, C, D, E, F, t, X3, Y3, Z3;
  f_sqr(Q, &A, &p->X); f_sqr(Q, &B, &p->Y); f_sqr(Q, &C, &B);
  f_add(Q, &t, &p->X, &B); f_sqr(Q, &t, &t); f_sub(Q, &t, &t, &A); f_sub(Q, &t, &t, &C); f_add(Q, &D, &t, &t);
  f_add(Q, &E, &A, &A); f_add(Q, &E, &E, &A); f_sqr(Q, &F, &E);
  f_sub(Q, &X3, &F, &D); f_sub(Q, &X3, &X3, &D);
  f_sub(Q, &t, &D, &X3); f_mul(Q, &Y3, &E, &t); f_add(Q, &C, &C, &C); f_add(Q, &C, &C, &C); f_add(Q, &C, &C, &C); f_sub(Q, &Y3, &Y3, &C);
  f_mul(Q, &Z3, &p->Y, &p->Z); f_add(Q, &Z3, &Z3, &Z3);
  r->X = X3; r->Y = Y3; r->Z = Z3;
}
static void j_add_affine(g1j* r, const g1j* p, const g1a* q) {
  if (q->inf) { *r = *p; return; }
  if (fe_is_zero(&p->Z)) { r->X = q->x; r->Y = q->y; r->Z = FQ.one; return; }
  fe Z1Z1, U2, S2, H, HH, I, J, rr, V, t, X3, Y3, Z3;
  f_sqr(Q, &Z1Z1, &p->Z); f_mul(Q, &U2, &q->x, &Z1Z1); f_mul(Q, &S2, &q->y, &p->Z); f_mul(Q, &S2, &S2, &Z1Z1);
  if (fe_eq(&U2, &p->X)) { if (fe_eq(&S2, &p->Y)) { j_dbl(r, p); return; } j_identity(r); return; }
  f_sub(Q, &H, &U2, &p->X); f_sqr(Q, &HH, &H); f_add(Q, &I, &HH, &HH); f_add(Q, &I, &I, &I); f_mul(Q, &J, &H, &I);
  f_sub(Q, &rr, &S2, &p->Y); f_add(Q, &rr, &rr, &rr); f_mul(Q, &V, &p->X, &I);
  f_sqr(Q, &X3, &rr); f_sub(Q, &X3, &X3, &J); f_sub(Q, &X3, &X3, &V); f_sub(Q, &X3, &X3, &V);
  f_sub(Q, &t, &V, &X3); f_mul(Q, &Y3, &rr, &t); f_mul(Q, &t, &p->Y, &J); f_add(Q, &t, &t, &t); f_sub(Q, &Y3, &Y3, &t);
  f_add(Q, &Z3, &p->Z, &H); f_sqr(Q, &Z3, &Z3); f_sub(Q, &Z3, &Z3, &Z1Z1); f_sub(Q, &Z3, &Z3, &HH);
  r->X = X3; r->Y = Y3; r->Z = Z3;
}
static void j_add(g1j* r, const g1j* p, const g1j* q) {
  if (fe_is_zero(&q->Z)) { *r = *p; return; }
  if (fe_is_zero(&p->Z)) { *r = *q; return; }
  fe Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, rr, V, t, X3, Y3, Z3;
  f_sqr(Q, &Z1Z1, &p->Z); f_sqr(Q, &Z2Z2, &q->Z); f_mul(Q, &U1, &p->X, &Z2Z2); f_mul(Q, &U2, &q->X, &Z1Z1);
  f_mul(Q, &S1, &p->Y, &q->Z); f_mul(Q, &S1, &S1, &Z2Z2); f_mul(Q, &S2, &q->Y, &p->Z); f_mul(Q, &S2, &S2, &Z1Z1);
  if (fe_eq(&U1, &U2)) { if (fe_eq(&S1, &S2)) { j_dbl(r, p); return; } j_identity(r); return; }
  f_sub(Q, &H, &U2, &U1); f_add(Q, &I, &H, &H); f_sqr(Q, &I, &I); f_mul(Q, &J, &H, &I);
  f_sub(Q, &rr, &S2, &S1); f_add(Q, &rr, &rr, &rr); f_mul(Q, &V, &U1, &I);
  f_sqr(Q, &X3, &rr); f_sub(Q, &X3, &X3, &J); f_sub(Q, &X3, &X3, &V); f_sub(Q, &X3, &X3, &V);
  f_sub(Q, &t, &V, &X3); f_mul(Q, &Y3, &rr, &t); f_mul(Q, &t, &S1, &J); f_add(Q, &t, &t, &t); f_sub(Q, &Y3, &Y3, &t);
  f_add(Q, &Z3, &p->Z, &q->Z); f_sqr(Q, &Z3, &Z3); f_sub(Q, &Z3, &Z3, &Z1Z1); f_sub(Q, &Z3, &Z3, &Z2Z2); f_mul(Q, &Z3, &Z3, &H);
  r->X = X3; r->Y = Y3; r->Z = Z3;
}
static void j_to_affine(g1a* r, const g1j* p) {
  if (fe_is_zero(&p->Z)) { memset(r, 0, sizeof *r); r->inf = 1; return; }
  fe zi, zi2; f_inv(Q, &zi, &p->Z); f_sqr(Q, &zi2, &zi);
  f_mul(Q, &r->x, &p->X, &zi2); f_mul(Q, &r->y, &p->Y, &zi2); f_mul(Q, &r->y, &r->y, &zi); r->inf = 0;
}
/* Optional "optimised CPU" mode (bench.py cpu_baseline.optimised_value): the same group elements
 * through a 4-bit fixed-window multiplication (15-entry table, 252 doublings + 64 additions instead of 256 doublings + 256
 * additions) -- what a tuned CPU verifier would do without changing the reference's structure.  Default 0 = literal port. */
static int CREF_OPT = 0;
void cref_set_optimised(int on) { CREF_OPT = on; }
static void g1_mul_win4(g1j* r, const g1a* p, const fe* k_canon) {
  g1j tbl[16], acc;
  j_identity(&tbl[0]);
  tbl[1].X = p->x; tbl[1].Y = p->y; tbl[1].Z = FQ.one;
  if (p->inf) j_identity(&tbl[1]);
  for (int i = 2; i < 16; i++) j_add_affine(&tbl[i], &tbl[i - 1], p);
  j_identity(&acc);
  for (int w = 63; w >= 0; w--) {
    if (w != 63) { j_dbl(&acc, &acc); j_dbl(&acc, &acc); j_dbl(&acc, &acc); j_dbl(&acc, &acc); }
    unsigned d = (unsigned)(k_canon->v[w >> 4] >> ((w & 15) * 4)) & 0xf;
    if (d) j_add(&acc, &acc, &tbl[d]);
  }
  *r = acc;
}
/* halo2curves `&G1Affine * &Fr`: all 256 bits of to_repr(), always add then conditional_select */
static void g1_mul_ref(g1j* r, const g1a* p, const fe* k_canon) {
  if (CREF_OPT) { g1_mul_win4(r, p, k_canon); return; }
  g1j acc, t; j_identity(&acc);
  for (int w = 3; w >= 0; w--) for (int b = 63; b >= 0; b--) {
    j_dbl(&acc, &acc);
    j_add_affine(&t, &acc, p);
    if ((k_canon->v[w] >> b) & 1) acc = t;
  }
  *r = acc;
}

/* ---------------------------------------------------------------- Poseidon (constants uploaded by Python, Montgomery) */
typedef struct { fe start[5][3], partial[57], end[3][3], mds[3][3], pre[3][3], row[57][3], col[57][2], cap; } pconsts;
static pconsts PK;
void cref_set_poseidon(const u64* canon /* 381 elements x 4 limbs in struct order */) {
  fe* dst = (fe*)&PK;
  for (int i = 0; i < (int)(sizeof(pconsts) / sizeof(fe)); i++) { fe c; memcpy(c.v, canon + 4 * i, 32); f_to_mont(&FR, &dst[i], &c); }
}
#define Rf (&FR)
static void pow5c(fe* x, const fe* c) { fe x2, x4; f_sqr(Rf, &x2, x); f_sqr(Rf, &x4, &x2); f_mul(Rf, x, &x4, x); if (c) f_add(Rf, x, x, c); }
static void mds3(fe* s, fe m[3][3]) {
  fe r[3], t;
  for (int i = 0; i < 3; i++) { f_mul(Rf, &r[i], &m[i][0], &s[0]); f_mul(Rf, &t, &m[i][1], &s[1]); f_add(Rf, &r[i], &r[i], &t); f_mul(Rf, &t, &m[i][2], &s[2]); f_add(Rf, &r[i], &r[i], &t); }
  s[0] = r[0]; s[1] = r[1]; s[2] = r[2];
}
static void permute(fe* s, int n_in, const fe* in) {
  f_add(Rf, &s[0], &s[0], &PK.start[0][0]);
  for (int i = 0; i < n_in; i++) { f_add(Rf, &s[i + 1], &s[i + 1], &in[i]); f_add(Rf, &s[i + 1], &s[i + 1], &PK.start[0][i + 1]); }
  for (int i = n_in + 1, k = 0; i < 3; i++, k++) { f_add(Rf, &s[i], &s[i], &PK.start[0][i]); if (k == 0) f_add(Rf, &s[i], &s[i], &FR.one); }
  for (int r = 1; r < 4; r++) { for (int i = 0; i < 3; i++) pow5c(&s[i], &PK.start[r][i]); mds3(s, PK.mds); }
  for (int i = 0; i < 3; i++) pow5c(&s[i], &PK.start[4][i]);
  mds3(s, PK.pre);
  for (int r = 0; r < 57; r++) {
    pow5c(&s[0], &PK.partial[r]);
    fe n0, t, n1, n2;
    f_mul(Rf, &n0, &PK.row[r][0], &s[0]); f_mul(Rf, &t, &PK.row[r][1], &s[1]); f_add(Rf, &n0, &n0, &t); f_mul(Rf, &t, &PK.row[r][2], &s[2]); f_add(Rf, &n0, &n0, &t);
    f_mul(Rf, &n1, &PK.col[r][0], &s[0]); f_add(Rf, &n1, &n1, &s[1]);
    f_mul(Rf, &n2, &PK.col[r][1], &s[0]); f_add(Rf, &n2, &n2, &s[2]);
    s[0] = n0; s[1] = n1; s[2] = n2;
  }
  for (int r = 0; r < 3; r++) { for (int i = 0; i < 3; i++) pow5c(&s[i], &PK.end[r][i]); mds3(s, PK.mds); }
  for (int i = 0; i < 3; i++) pow5c(&s[i], NULL);
  mds3(s, PK.mds);
}
typedef struct { fe s[3]; fe buf[64]; int n; } sponge;
static void sp_init(sponge* sp) { sp->s[0] = PK.cap; memset(&sp->s[1], 0, 64); sp->n = 0; }
static void sp_absorb(sponge* sp, const fe* x) { if (sp->n < 64) sp->buf[sp->n++] = *x; }
static void sp_squeeze(sponge* sp, fe* out) {
  int n = sp->n; sp->n = 0;
  for (int i = 0; i < n; i += 2) permute(sp->s, (n - i >= 2) ? 2 : 1, &sp->buf[i]);
  if (n % 2 == 0) permute(sp->s, 0, NULL);
  *out = sp->s[1];
}
static void fq_to_fr(fe* out, const fe* fq_mont) {  /* fe_to_fe: value mod r (p < 2r) */
  fe c; f_from_mont(Q, &c, fq_mont);
  if (geq(c.v, FR.m)) sub_mod_raw(c.v, c.v, FR.m);
  f_to_mont(Rf, out, &c);
}
static void sp_absorb_point(sponge* sp, const g1a* p) { fe x, y; fq_to_fr(&x, &p->x); fq_to_fr(&y, &p->y); sp_absorb(sp, &x); sp_absorb(sp, &y); }

/* halo2curves G1Affine::from_bytes.  0 ok, 1 invalid, 2 identity */
static int g1_decompress(g1a* out, const uint8_t* b) {
  fe x; memcpy(x.v, b, 32);
  int ysign = (int)(x.v[3] >> 63); x.v[3] &= 0x7fffffffffffffffULL;
  memset(out, 0, sizeof *out); out->inf = 1;
  if (!f_canonical(Q, &x)) return 1;
  if (fe_is_zero(&x) && !ysign) return 2;
  fe xm, rhs, y, t, three; f_to_mont(Q, &xm, &x);
  f_sqr(Q, &rhs, &xm); f_mul(Q, &rhs, &rhs, &xm); f_add(Q, &three, &FQ.one, &FQ.one); f_add(Q, &three, &three, &FQ.one); f_add(Q, &rhs, &rhs, &three);
  u64 e[4]; { u128 c = 1; for (int i = 0; i < 4; i++) { c += FQ.m[i]; e[i] = (u64)c; c >>= 64; } for (int i = 0; i < 3; i++) e[i] = (e[i] >> 2) | (e[i + 1] << 62); e[3] >>= 2; }
  f_pow(Q, &y, &rhs, e); f_sqr(Q, &t, &y);
  if (!fe_eq(&t, &rhs)) return 1;
  fe yc; f_from_mont(Q, &yc, &y);
  if ((int)(yc.v[0] & 1) != ysign) f_neg(Q, &y, &y);
  out->x = xm; out->y = y; out->inf = 0;
  return 0;
}

/* ---------------------------------------------------------------- trace replay */
enum { C_ADD, C_SUB, C_MUL, C_NEG, C_INV, C_MSM, C_SQUEEZE, C_COMMON_SCALAR, C_COMMON_POINT, C_READ_SCALAR, C_READ_POINT, C_INPUT, C_CLEAR };
typedef struct {
  int n_ops; const int32_t* ops;  /* 4 ints per op */
  const int32_t* msm_pairs;       /* (sreg, preg) */
  int n_s, n_p;
  int n_const_s; const int32_t* const_s_reg; const u64* const_s_val;  /* canonical */
  int n_const_p; const int32_t* const_p_reg; const u64* const_p_xy;   /* canonical x,y; (0,0) = identity */
  int out_lhs, out_rhs;
} trace_t;

static void load_const_state(const trace_t* T, fe* S, g1a* Pt) {
  for (int i = 0; i < T->n_const_s; i++) { fe c; memcpy(c.v, T->const_s_val + 4 * i, 32); f_to_mont(Rf, &S[T->const_s_reg[i]], &c); }
  for (int i = 0; i < T->n_const_p; i++) {
    g1a* p = &Pt[T->const_p_reg[i]]; fe x, y; memcpy(x.v, T->const_p_xy + 8 * i, 32); memcpy(y.v, T->const_p_xy + 8 * i + 4, 32);
    if (fe_is_zero(&x) && fe_is_zero(&y)) { memset(p, 0, sizeof *p); p->inf = 1; } else { f_to_mont(Q, &p->x, &x); f_to_mont(Q, &p->y, &y); p->inf = 0; }
  }
}
/* status: 0 ok, 4 transcript error (sub-code in bits 8..) */
static int replay_one(const trace_t* T, fe* S, g1a* Pt, const uint8_t* proof, int proof_len, const u64* inputs, uint8_t* out_acc) {
  sponge sp; sp_init(&sp);
  int pos = 0;
  for (int i = 0; i < T->n_ops; i++) {
    const int32_t* o = T->ops + 4 * i;
    switch (o[0]) {
      case C_ADD: f_add(Rf, &S[o[1]], &S[o[2]], &S[o[3]]); break;
      case C_SUB: f_sub(Rf, &S[o[1]], &S[o[2]], &S[o[3]]); break;
      case C_MUL: f_mul(Rf, &S[o[1]], &S[o[2]], &S[o[3]]); break;
      case C_NEG: f_neg(Rf, &S[o[1]], &S[o[2]]); break;
      case C_INV: f_inv(Rf, &S[o[1]], &S[o[2]]); break;
      case C_MSM: {
        g1j acc, t; j_identity(&acc);
        for (int k = 0; k < o[3]; k++) {
          const int32_t* pr = T->msm_pairs + 2 * (o[2] + k);
          fe kc; f_from_mont(Rf, &kc, &S[pr[0]]);
          g1_mul_ref(&t, &Pt[pr[1]], &kc);
          if (k == 0) acc = t; else j_add(&acc, &acc, &t);
        }
        j_to_affine(&Pt[o[1]], &acc);
        break;
      }
      case C_SQUEEZE: sp_squeeze(&sp, &S[o[1]]); break;
      case C_COMMON_SCALAR: sp_absorb(&sp, &S[o[2]]); break;
      case C_COMMON_POINT: if (Pt[o[2]].inf) return 4 | (4 << 8); sp_absorb_point(&sp, &Pt[o[2]]); break;
      case C_READ_SCALAR: {
        if (pos + 32 > proof_len) return 4 | (1 << 8);
        fe c; memcpy(c.v, proof + pos, 32); pos += 32;
        if (!f_canonical(Rf, &c)) return 4 | (2 << 8);
        f_to_mont(Rf, &S[o[1]], &c); sp_absorb(&sp, &S[o[1]]);
        break;
      }
      case C_READ_POINT: {
        if (pos + 32 > proof_len) return 4 | (1 << 8);
        int rc = g1_decompress(&Pt[o[1]], proof + pos); pos += 32;
        if (rc == 1) return 4 | (3 << 8);
        if (rc == 2) return 4 | (4 << 8);
        sp_absorb_point(&sp, &Pt[o[1]]);
        break;
      }
      case C_INPUT: { fe c; memcpy(c.v, inputs + 4 * o[2], 32); if (!f_canonical(Rf, &c)) return 1; f_to_mont(Rf, &S[o[1]], &c); break; }
      case C_CLEAR: sp_init(&sp); break;
    }
  }
  const g1a* outs[2] = {&Pt[T->out_lhs], &Pt[T->out_rhs]};
  for (int h = 0; h < 2; h++) {
    fe x, y; memset(out_acc + 64 * h, 0, 64);
    if (!outs[h]->inf) { f_from_mont(Q, &x, &outs[h]->x); f_from_mont(Q, &y, &outs[h]->y); memcpy(out_acc + 64 * h, x.v, 32); memcpy(out_acc + 64 * h + 32, y.v, 32); }
  }
  return 0;
}

/* PlonkSuccinctVerifier::{read_proof, verify} for n proofs; `threads` pthreads, proofs strided over them
 * (the reference's path itself is single-threaded; callers parallelise over proofs). */
#include <pthread.h>
typedef struct {
  const trace_t* T; int n, tid, nthreads; const uint8_t* proofs; int stride; const int32_t* lens; const u64* inputs; int n_inputs;
  uint8_t* out_accs; int32_t* out_status;
} job_t;
static void* replay_worker(void* arg) {
  job_t* j = (job_t*)arg;
  const trace_t* T = j->T;
  fe* S = (fe*)calloc(T->n_s + 1, sizeof(fe));
  g1a* Pt = (g1a*)calloc(T->n_p + 1, sizeof(g1a));
  load_const_state(T, S, Pt);
  for (int i = j->tid; i < j->n; i += j->nthreads) {
    j->out_status[i] = replay_one(T, S, Pt, j->proofs + (size_t)i * j->stride, j->lens ? j->lens[i] : j->stride,
                                  j->inputs + (size_t)4 * j->n_inputs * i, j->out_accs + (size_t)128 * i);
    if (j->out_status[i]) memset(j->out_accs + (size_t)128 * i, 0, 128);
  }
  free(S); free(Pt);
  return NULL;
}
int cref_replay(const trace_t* T, int n, const uint8_t* proofs, int stride, const int32_t* lens, const u64* inputs, int n_inputs,
                uint8_t* out_accs, int32_t* out_status, int threads) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  pthread_t th[256]; job_t jobs[256];
  for (int t = 0; t < threads; t++) {
    job_t j = {T, n, t, threads, proofs, stride, lens, inputs, n_inputs, out_accs, out_status};
    jobs[t] = j;
    if (t > 0) pthread_create(&th[t], NULL, replay_worker, &jobs[t]);
  }
  replay_worker(&jobs[0]);
  for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
  return 0;
}

static void load_acc_point(g1a* p, const uint8_t* b) {
  fe x, y; memcpy(x.v, b, 32); memcpy(y.v, b + 32, 32);
  if (fe_is_zero(&x) && fe_is_zero(&y)) { memset(p, 0, sizeof *p); p->inf = 1; return; }
  f_to_mont(Q, &p->x, &x); f_to_mont(Q, &p->y, &y); p->inf = 0;
}
/* `util::msm::multi_scalar_multiplication` (snark-verifier/src/util/msm.rs:238-317): window = ceil(ln n) + 2 bits over the 256-bit
 * `to_repr()`, 2^w - 1 buckets per window (MSB window first: w doublings of the running result, bucket fill with mixed additions,
 * running-sum reduction), and with the `parallel` feature one serial MSM per chunk of n / threads points, results added.
 * scalars: n x 32 B LE canonical, points: n x 64 B affine canonical ((0, 0) = identity), out: 64 B affine canonical. */
#include <math.h>
static void msm_serial(int n, const uint8_t* scalars, const g1a* bases, g1j* result) {
  int w = (int)ceil(log((double)n)) + 2;
  if (w < 1) w = 1;
  int nbuckets = (1 << w) - 1, num_window = (256 + w - 1) / w;
  g1j* buckets = (g1j*)malloc(sizeof(g1j) * (size_t)nbuckets);
  for (int idx = num_window - 1; idx >= 0; idx--) {
    for (int k = 0; k < w; k++) j_dbl(result, result);
    for (int b = 0; b < nbuckets; b++) j_identity(&buckets[b]);
    for (int i = 0; i < n; i++) {
      int skip_bits = idx * w, skip_bytes = skip_bits / 8;
      uint64_t v = 0;
      for (int q = 0; q < 8 && skip_bytes + q < 32; q++) v |= (uint64_t)scalars[32 * (size_t)i + skip_bytes + q] << (8 * q);
      uint64_t d = (v >> (skip_bits - skip_bytes * 8)) & (uint64_t)nbuckets;
      if (d) j_add_affine(&buckets[d - 1], &buckets[d - 1], &bases[i]);
    }
    g1j running; j_identity(&running);
    for (int b = nbuckets - 1; b >= 0; b--) { j_add(&running, &running, &buckets[b]); j_add(result, result, &running); }
  }
  free(buckets);
}
typedef struct { int n; const uint8_t* scalars; const g1a* bases; g1j result; } msm_job_t;
static void* msm_worker(void* arg) { msm_job_t* j = (msm_job_t*)arg; j_identity(&j->result); if (j->n) msm_serial(j->n, j->scalars, j->bases, &j->result); return NULL; }
int cref_msm(int n, const uint8_t* scalars, const uint8_t* points, uint8_t* out, int threads) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  g1a* bases = (g1a*)malloc(sizeof(g1a) * (size_t)(n ? n : 1));
  for (int i = 0; i < n; i++) load_acc_point(&bases[i], points + 64 * (size_t)i);
  if (n < threads) threads = 1;  /* msm.rs:296-300 */
  int chunk = (n + threads - 1) / threads;
  pthread_t th[256]; msm_job_t jobs[256];
  for (int t = 0; t < threads; t++) {
    int lo = t * chunk, hi = lo + chunk > n ? n : lo + chunk;
    if (lo > n) lo = n;
    msm_job_t j = {hi > lo ? hi - lo : 0, scalars + 32 * (size_t)lo, bases + lo, {{{0}}}};
    jobs[t] = j;
    if (t > 0) pthread_create(&th[t], NULL, msm_worker, &jobs[t]);
  }
  msm_worker(&jobs[0]);
  g1j acc = jobs[0].result;
  for (int t = 1; t < threads; t++) { pthread_join(th[t], NULL); j_add(&acc, &acc, &jobs[t].result); }
  g1a a; j_to_affine(&a, &acc);
  memset(out, 0, 64);
  if (!a.inf) { fe x, y; f_from_mont(Q, &x, &a.x); f_from_mont(Q, &y, &a.y); memcpy(out, x.v, 32); memcpy(out + 32, y.v, 32); }
  free(bases);
  return 0;
}

/* KzgAs::create_proof / verify, zk = false, one group (accumulation.rs:45-61,113-136). returns status */
int cref_fold_group(int n, const uint8_t* accs, uint8_t* out_acc, u64* out_r_canon) {
  sponge sp; sp_init(&sp);
  g1a* L = (g1a*)malloc(sizeof(g1a) * n), *Rr = (g1a*)malloc(sizeof(g1a) * n);
  for (int i = 0; i < n; i++) {
    load_acc_point(&L[i], accs + 128 * (size_t)i); load_acc_point(&Rr[i], accs + 128 * (size_t)i + 64);
    if (L[i].inf || Rr[i].inf) { free(L); free(Rr); return 4 | (4 << 8); }
    /* the sponge buffer is unbounded in the reference; absorb pairwise to keep ours small */
    sp_absorb_point(&sp, &L[i]); permute(sp.s, 2, sp.buf); sp.n = 0;
    sp_absorb_point(&sp, &Rr[i]); permute(sp.s, 2, sp.buf); sp.n = 0;
  }
  permute(sp.s, 0, NULL);
  fe r = sp.s[1], pw = FR.one, rc;
  f_from_mont(Rf, &rc, &r); memcpy(out_r_canon, rc.v, 32);
  g1j al, ar, t; j_identity(&al); j_identity(&ar);
  for (int i = 0; i < n; i++) {
    fe kc; f_from_mont(Rf, &kc, &pw);
    g1_mul_ref(&t, &L[i], &kc); if (i == 0) al = t; else j_add(&al, &al, &t);
    g1_mul_ref(&t, &Rr[i], &kc); if (i == 0) ar = t; else j_add(&ar, &ar, &t);
    f_mul(Rf, &pw, &pw, &r);
  }
  g1a a; fe x, y;
  j_to_affine(&a, &al); memset(out_acc, 0, 128);
  if (!a.inf) { f_from_mont(Q, &x, &a.x); f_from_mont(Q, &y, &a.y); memcpy(out_acc, x.v, 32); memcpy(out_acc + 32, y.v, 32); }
  j_to_affine(&a, &ar);
  if (!a.inf) { f_from_mont(Q, &x, &a.x); f_from_mont(Q, &y, &a.y); memcpy(out_acc + 64, x.v, 32); memcpy(out_acc + 96, y.v, 32); }
  free(L); free(Rr);
  return 0;
}

/* ---------------------------------------------------------------- tower + pairing (decider.rs:60-68) */
typedef struct { fe c0, c1; } f2;
typedef struct { f2 c0, c1, c2; } f6;
typedef struct { f6 c0, c1; } f12;
static void f2_add(f2* r, const f2* a, const f2* b) { f_add(Q, &r->c0, &a->c0, &b->c0); f_add(Q, &r->c1, &a->c1, &b->c1); }
static void f2_sub(f2* r, const f2* a, const f2* b) { f_sub(Q, &r->c0, &a->c0, &b->c0); f_sub(Q, &r->c1, &a->c1, &b->c1); }
static void f2_neg(f2* r, const f2* a) { f_neg(Q, &r->c0, &a->c0); f_neg(Q, &r->c1, &a->c1); }
static void f2_mul(f2* r, const f2* a, const f2* b) {
  fe t0, t1, t2, s0, s1; f_mul(Q, &t0, &a->c0, &b->c0); f_mul(Q, &t1, &a->c1, &b->c1);
  f_add(Q, &s0, &a->c0, &a->c1); f_add(Q, &s1, &b->c0, &b->c1); f_mul(Q, &t2, &s0, &s1);
  f_sub(Q, &r->c0, &t0, &t1); f_sub(Q, &t2, &t2, &t0); f_sub(Q, &r->c1, &t2, &t1);
}
static void f2_sqr(f2* r, const f2* a) { f2_mul(r, a, a); }
static void f2_muls(f2* r, const f2* a, const fe* s) { f_mul(Q, &r->c0, &a->c0, s); f_mul(Q, &r->c1, &a->c1, s); }
static void f2_conj(f2* r, const f2* a) { r->c0 = a->c0; f_neg(Q, &r->c1, &a->c1); }
static void f2_mulxi(f2* r, const f2* a) {
  fe a8, b8, t0, t1; f_add(Q, &a8, &a->c0, &a->c0); f_add(Q, &a8, &a8, &a8); f_add(Q, &a8, &a8, &a8);
  f_add(Q, &b8, &a->c1, &a->c1); f_add(Q, &b8, &b8, &b8); f_add(Q, &b8, &b8, &b8);
  f_add(Q, &t0, &a8, &a->c0); f_sub(Q, &t0, &t0, &a->c1); f_add(Q, &t1, &b8, &a->c1); f_add(Q, &t1, &t1, &a->c0);
  r->c0 = t0; r->c1 = t1;
}
static void f2_inv(f2* r, const f2* a) {
  fe n, t; f_sqr(Q, &n, &a->c0); f_sqr(Q, &t, &a->c1); f_add(Q, &n, &n, &t); f_inv(Q, &n, &n);
  f_mul(Q, &r->c0, &a->c0, &n); f_mul(Q, &t, &a->c1, &n); f_neg(Q, &r->c1, &t);
}
static void f6_add(f6* r, const f6* a, const f6* b) { f2_add(&r->c0, &a->c0, &b->c0); f2_add(&r->c1, &a->c1, &b->c1); f2_add(&r->c2, &a->c2, &b->c2); }
static void f6_sub(f6* r, const f6* a, const f6* b) { f2_sub(&r->c0, &a->c0, &b->c0); f2_sub(&r->c1, &a->c1, &b->c1); f2_sub(&r->c2, &a->c2, &b->c2); }
static void f6_neg(f6* r, const f6* a) { f2_neg(&r->c0, &a->c0); f2_neg(&r->c1, &a->c1); f2_neg(&r->c2, &a->c2); }
static void f6_mul(f6* r, const f6* a, const f6* b) {
  f2 t0, t1, t2, s, u, x, r0, r1, r2;
  f2_mul(&t0, &a->c0, &b->c0); f2_mul(&t1, &a->c1, &b->c1); f2_mul(&t2, &a->c2, &b->c2);
  f2_add(&s, &a->c1, &a->c2); f2_add(&u, &b->c1, &b->c2); f2_mul(&x, &s, &u); f2_sub(&x, &x, &t1); f2_sub(&x, &x, &t2); f2_mulxi(&x, &x); f2_add(&r0, &x, &t0);
  f2_add(&s, &a->c0, &a->c1); f2_add(&u, &b->c0, &b->c1); f2_mul(&x, &s, &u); f2_sub(&x, &x, &t0); f2_sub(&x, &x, &t1); f2_mulxi(&s, &t2); f2_add(&r1, &x, &s);
  f2_add(&s, &a->c0, &a->c2); f2_add(&u, &b->c0, &b->c2); f2_mul(&x, &s, &u); f2_sub(&x, &x, &t0); f2_sub(&x, &x, &t2); f2_add(&r2, &x, &t1);
  r->c0 = r0; r->c1 = r1; r->c2 = r2;
}
static void f6_mulv(f6* r, const f6* a) { f2 t; f2_mulxi(&t, &a->c2); f2 a0 = a->c0, a1 = a->c1; r->c0 = t; r->c1 = a0; r->c2 = a1; }
static void f6_inv(f6* r, const f6* a) {
  f2 t0, t1, t2, x, d;
  f2_sqr(&t0, &a->c0); f2_mul(&x, &a->c1, &a->c2); f2_mulxi(&x, &x); f2_sub(&t0, &t0, &x);
  f2_sqr(&t1, &a->c2); f2_mulxi(&t1, &t1); f2_mul(&x, &a->c0, &a->c1); f2_sub(&t1, &t1, &x);
  f2_sqr(&t2, &a->c1); f2_mul(&x, &a->c0, &a->c2); f2_sub(&t2, &t2, &x);
  f2 y; f2_mul(&x, &a->c2, &t1); f2_mul(&y, &a->c1, &t2); f2_add(&x, &x, &y); f2_mulxi(&x, &x); f2_mul(&d, &a->c0, &t0); f2_add(&d, &d, &x);
  f2_inv(&d, &d); f2_mul(&r->c0, &t0, &d); f2_mul(&r->c1, &t1, &d); f2_mul(&r->c2, &t2, &d);
}
static void f12_mul(f12* r, const f12* a, const f12* b) {
  f6 t0, t1, s, u, x; f6_mul(&t0, &a->c0, &b->c0); f6_mul(&t1, &a->c1, &b->c1);
  f6_add(&s, &a->c0, &a->c1); f6_add(&u, &b->c0, &b->c1); f6_mul(&x, &s, &u); f6_sub(&x, &x, &t0); f6_sub(&x, &x, &t1);
  f6_mulv(&s, &t1); f6_add(&r->c0, &t0, &s); r->c1 = x;
}
static void f12_one(f12* r) { memset(r, 0, sizeof *r); r->c0.c0.c0 = FQ.one; }
static int f12_is_one(const f12* a) { f12 o; f12_one(&o); return memcmp(a, &o, sizeof o) == 0; }
static void f12_conj(f12* r, const f12* a) { r->c0 = a->c0; f6_neg(&r->c1, &a->c1); }
static void f12_inv(f12* r, const f12* a) {
  f6 t0, t1, d; f6_mul(&t0, &a->c0, &a->c0); f6_mul(&t1, &a->c1, &a->c1); f6_mulv(&t1, &t1); f6_sub(&d, &t0, &t1); f6_inv(&d, &d);
  f6_mul(&r->c0, &a->c0, &d); f6_mul(&t0, &a->c1, &d); f6_neg(&r->c1, &t0);
}
typedef struct { f2 x, y; } g2a;
static f2 G12, G13;  /* xi^((p-1)/3), xi^((p-1)/2) */
static f2 GAM[5]; static fe GAM2[5]; static int tower_init = 0;
static void f2_pow(f2* r, const f2* a, const u64* e, int n) {
  f2 acc; memset(&acc, 0, sizeof acc); acc.c0 = FQ.one;
  for (int w = n - 1; w >= 0; w--) for (int b = 63; b >= 0; b--) { f2_sqr(&acc, &acc); if ((e[w] >> b) & 1) f2_mul(&acc, &acc, a); }
  *r = acc;
}
static void init_tower(void) {
  if (tower_init) return;
  u64 e[4] = {FQ.m[0] - 1, FQ.m[1], FQ.m[2], FQ.m[3]}; u128 rem = 0;
  for (int i = 3; i >= 0; i--) { u128 cur = (rem << 64) | e[i]; e[i] = (u64)(cur / 6); rem = cur % 6; }
  f2 xi; fe nine = FQ.one; for (int i = 0; i < 3; i++) f_add(Q, &nine, &nine, &nine); f_add(Q, &nine, &nine, &FQ.one); xi.c0 = nine; xi.c1 = FQ.one;
  f2 g; f2_pow(&g, &xi, e, 4); f2 acc = g;
  for (int i = 0; i < 5; i++) { GAM[i] = acc; f2 cj, nn; f2_conj(&cj, &acc); f2_mul(&nn, &acc, &cj); GAM2[i] = nn.c0; f2_mul(&acc, &acc, &g); }
  G12 = GAM[1]; G13 = GAM[2]; tower_init = 1;
}
static void f12_frob(f12* r, const f12* a, int k) {  /* k = 1 or 2 */
  const f2* in[6] = {&a->c0.c0, &a->c1.c0, &a->c0.c1, &a->c1.c1, &a->c0.c2, &a->c1.c2};
  f2* out[6] = {&r->c0.c0, &r->c1.c0, &r->c0.c1, &r->c1.c1, &r->c0.c2, &r->c1.c2};
  f12 tmp = *a; (void)tmp;
  f2 v[6];
  for (int i = 0; i < 6; i++) {
    if (k == 1) { f2_conj(&v[i], in[i]); if (i) f2_mul(&v[i], &v[i], &GAM[i - 1]); }
    else { v[i] = *in[i]; if (i) f2_muls(&v[i], &v[i], &GAM2[i - 1]); }
  }
  for (int i = 0; i < 6; i++) *out[i] = v[i];
}
/* line through T and Q (T==Q: tangent) at P; advances T.  l = yP - lam xP w + (lam xT - yT) w^3 */
static void line_step(f12* l, g2a* T, const g2a* Qp, int dbl, const g1a* P) {
  f2 lam, t, x3, y3;
  if (dbl) { f2 x2, d; f2_sqr(&x2, &T->x); f2_add(&t, &x2, &x2); f2_add(&t, &t, &x2); f2_add(&d, &T->y, &T->y); f2_inv(&d, &d); f2_mul(&lam, &t, &d); }
  else { f2 n, d; f2_sub(&n, &Qp->y, &T->y); f2_sub(&d, &Qp->x, &T->x); f2_inv(&d, &d); f2_mul(&lam, &n, &d); }
  memset(l, 0, sizeof *l);
  l->c0.c0.c0 = P->y;
  f2_muls(&t, &lam, &P->x); f2_neg(&l->c1.c0, &t);
  f2_mul(&t, &lam, &T->x); f2_sub(&l->c1.c1, &t, &T->y);
  f2_sqr(&x3, &lam); f2_sub(&x3, &x3, &T->x); f2_sub(&x3, &x3, dbl ? &T->x : &Qp->x);
  f2_sub(&t, &T->x, &x3); f2_mul(&y3, &lam, &t); f2_sub(&y3, &y3, &T->y);
  T->x = x3; T->y = y3;
}
static void miller(f12* f, const g1a* P, const g2a* Qp) {
  static const u64 ATE[2] = {0x9d797039be763ba8ULL, 0x1ULL};
  f12_one(f);
  if (P->inf) return;
  g2a T = *Qp; f12 l;
  for (int i = 63; i >= 0; i--) {
    f12_mul(f, f, f); line_step(&l, &T, &T, 1, P); f12_mul(f, f, &l);
    if ((ATE[0] >> i) & 1) { line_step(&l, &T, Qp, 0, P); f12_mul(f, f, &l); }
  }
  g2a Q1, Q2; f2 t;
  f2_conj(&t, &Qp->x); f2_mul(&Q1.x, &t, &G12); f2_conj(&t, &Qp->y); f2_mul(&Q1.y, &t, &G13);
  f2_conj(&t, &Q1.x); f2_mul(&Q2.x, &t, &G12); f2_conj(&t, &Q1.y); f2_mul(&Q2.y, &t, &G13); f2_neg(&Q2.y, &Q2.y);
  line_step(&l, &T, &Q1, 0, P); f12_mul(f, f, &l);
  line_step(&l, &T, &Q2, 0, P); f12_mul(f, f, &l);
}
static void f12_powx(f12* r, const f12* a) {
  const u64 X = 0x44e992b44a6909f1ULL; f12 acc = *a;
  for (int i = 61; i >= 0; i--) { f12_mul(&acc, &acc, &acc); if ((X >> i) & 1) f12_mul(&acc, &acc, a); }
  *r = acc;
}
static void f12_pown(f12* r, const f12* a, int n) { f12 acc; f12_one(&acc); for (int b = 7; b >= 0; b--) { f12_mul(&acc, &acc, &acc); if ((n >> b) & 1) f12_mul(&acc, &acc, a); } *r = acc; }
static void final_exp(f12* r, const f12* f0) {
  f12 t, f, inv, fx, fx2, fx3, a, b, c, e0, e1, e2, u;
  f12_conj(&t, f0); f12_inv(&inv, f0); f12_mul(&t, &t, &inv); f12_frob(&f, &t, 2); f12_mul(&f, &f, &t);
  f12_powx(&fx, &f); f12_powx(&fx2, &fx); f12_powx(&fx3, &fx2);
  f12_pown(&a, &fx2, 6); f12_mul(&e2, &a, &f);
  f12_pown(&a, &fx3, 36); f12_pown(&b, &fx2, 18); f12_pown(&c, &fx, 12); f12_mul(&u, &a, &b); f12_mul(&u, &u, &c); f12_conj(&u, &u); f12_mul(&e1, &u, &f);
  f12_pown(&b, &fx2, 30); f12_pown(&c, &fx, 18); f12_mul(&u, &a, &b); f12_mul(&u, &u, &c); f12_mul(&t, &f, &f); f12_mul(&u, &u, &t); f12_conj(&e0, &u);
  f12 p3, p2, p1; f12_frob(&p3, &f, 2); f12_frob(&p3, &p3, 1); f12_frob(&p2, &e2, 2); f12_frob(&p1, &e1, 1);
  f12_mul(r, &p3, &p2); f12_mul(r, r, &p1); f12_mul(r, r, &e0);
}
/* dk: g2 (4 fe canonical), s_g2 (4 fe canonical); acc: 128 bytes.  returns 1 accept / 0 reject */
int cref_decide(const uint8_t* acc, const u64* g2_canon, const u64* sg2_canon) {
  init_tower();
  g1a L, Rr; load_acc_point(&L, acc); load_acc_point(&Rr, acc + 64);
  g2a G2, SG2; fe c;
  fe* dst[8] = {&G2.x.c0, &G2.x.c1, &G2.y.c0, &G2.y.c1, &SG2.x.c0, &SG2.x.c1, &SG2.y.c0, &SG2.y.c1};
  for (int i = 0; i < 4; i++) { memcpy(c.v, g2_canon + 4 * i, 32); f_to_mont(Q, dst[i], &c); memcpy(c.v, sg2_canon + 4 * i, 32); f_to_mont(Q, dst[4 + i], &c); }
  f2_neg(&SG2.y, &SG2.y);  /* -s_g2 */
  f12 f1, f2v, f, e; miller(&f1, &L, &G2); miller(&f2v, &Rr, &SG2); f12_mul(&f, &f1, &f2v); final_exp(&e, &f);
  return f12_is_one(&e);
}
