// Poseidon permutation, latency form: the three state words of ONE sponge live in three different warps of a block.
//
// Same function as poseidon.cuh (`Poseidon::permutation`, snark-verifier/src/util/hash/poseidon.rs:469-501); only the
// schedule differs.  A transcript is a serial chain of permutations (27 per StandardPlonk proof, 2m + 1 per fold group),
// and a permutation run by one thread is ~475 dependent Montgomery products.  With few sponges in a launch (a 4096-proof
// batch = 128 warps on 592 sub-partitions) that chain is the whole kernel time.  Here lane l of warp w holds word s_w of
// sponge l ("warp-specialised": warp 0 = the transcript owner, warps 1, 2 = helpers), so the warps never diverge and the
// register-file layouts ([reg][item]) of the callers stay coalesced:
//   full round     every warp: y_w = s_w^5 + c_w, exchange through shared memory, s_w = <M[w], y>          (4 products deep)
//   partial round  warp 0 holds t with s_0 = sigma t (the S-box commutes with scaling, poseidon.cuh sc_*): v = t^5, post v, t = v + Q
//                                                                                                             (3 products + 1 addition deep)
//                  warp 2: Q_2 = (row_2 / sigma') s_2 -> warp 1;   warp 1: Q = (row_1 / sigma') s_1 + Q_2 + row_0 c / sigma' -> warp 0
//                                                                                                             (posted BEFORE v is needed)
//                  warp w: s_w = (col_w sigma^5) v + (s_w + col_w c)     (c = the round constant: y = x + c never materialises)
// Warp 0 never waits for a helper in the partial rounds (P of round r only needs x of round r - 1); the helpers wait for
// x.  Synchronisation = named barriers (bar.arrive / bar.sync, producer/consumer form), mailboxes double-buffered.
// Dependent chain per permutation: 57 x 3 + 8 x 4 products + the exchanges, instead of ~475 product-equivalents.
#pragma once
#include "poseidon.cuh"

#if defined(__CUDACC__)
#define PCOOP_THREADS 96
enum { PC_BAR_CMD = 1, PC_BAR_FULL = 2, PC_BAR_Y = 3 /* 3, 4 */, PC_BAR_P = 5 /* 5, 6 */, PC_BAR_Q = 7 /* 7, 8 */ };
#define PC_CMD_RESET 4
#define PC_CMD_EXIT 8

struct __align__(16) PoseidonCoopShared {
  uint4 mail[2][3][32][2];  // [buffer][word][sponge] x 32 B
  uint4 in[2][32][2];       // absorbed inputs of the next permutation
  uint4 out1[32][2];        // word 1 after the permutation (the squeeze output, poseidon.rs:465)
  int n_in[32];             // inputs absorbed by the next permutation, per sponge (0, 1 or 2)
  int cmd;                  // PC_CMD_RESET | PC_CMD_EXIT
};

__device__ __forceinline__ void pc_bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(PCOOP_THREADS) : "memory"); }
__device__ __forceinline__ void pc_bar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(PCOOP_THREADS) : "memory"); }
// two-warp producer/consumer pairs (64 participants)
__device__ __forceinline__ void pc_bar_sync2(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void pc_bar_arrive2(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
// products stay out-of-line calls (field.cuh): with them inlined the three role loops no longer fit the instruction caches of
// their lone warps (measured: k_tape_coop 4.0 -> 5.5 ms)
__device__ __forceinline__ Fr pc_pow5(const Fr& x) { return fr_pow5(x); }
__device__ __forceinline__ void pc_st(uint4* p, const Fr& x) {
  p[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  p[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}
__device__ __forceinline__ Fr pc_ld(const uint4* p) {
  uint4 lo = p[0], hi = p[1];
  Fr x;
  x.v[0] = lo.x; x.v[1] = lo.y; x.v[2] = lo.z; x.v[3] = lo.w;
  x.v[4] = hi.x; x.v[5] = hi.y; x.v[6] = hi.z; x.v[7] = hi.w;
  return x;
}

// One permutation, executed by all three warps with their own `role` (= state word index) and word `s`.
static __device__ __noinline__ void pc_permute(Fr& s, int role, int lane, PoseidonCoopShared* sh, const PoseidonConsts& k, int n_in) {
  // absorb_with_pre_constants incl. the "+1" padding on the first unused slot (poseidon.rs:362-384)
  if (role == 0) {
    s = s + k.start[0][0];
  } else {
    const Fr& c = k.start[0][role];
    if (n_in >= role) s = s + pc_ld(sh->in[role - 1][lane]) + c;
    else if (n_in == role - 1) s = s + c + Fr::one();
    else s = s + c;
  }
  int buf = 0;
  auto full = [&](const Fr(*M)[3], const Fr* cr) {
    Fr y = pc_pow5(s);
    if (cr) y = y + cr[role];
    pc_st(sh->mail[buf][role][lane], y);
    pc_bar_sync(PC_BAR_FULL);
    Fr y0 = pc_ld(sh->mail[buf][0][lane]), y1 = pc_ld(sh->mail[buf][1][lane]), y2 = pc_ld(sh->mail[buf][2][lane]);
    s = Fr::dot3(M[role][0], y0, M[role][1], y1, M[role][2], y2);
    buf ^= 1;
  };
  for (int r = 1; r < SVK_POSEIDON_RF / 2; r++) full(k.mds, k.start[r]);
  full(k.pre_sparse_mds, k.start[SVK_POSEIDON_RF / 2]);
  // partial rounds with the sparse MDS factorisation (poseidon.rs:398-410)
  // scaled form (poseidon.cuh: sc_*): warp 0 holds t with s0 = sigma_r t
  if (role == 0) {
    for (int r = 0; r < SVK_POSEIDON_RP; r++) {
      Fr v = pc_pow5(s);
      pc_st(sh->mail[buf][0][lane], v);
      pc_bar_arrive(PC_BAR_Y + buf);
      pc_bar_sync2(PC_BAR_P + buf);
      s = v + pc_ld(sh->mail[buf][1][lane]);
      buf ^= 1;
    }
    s = k.sc_end * s;
  } else if (role == 1) {
    for (int r = 0; r < SVK_POSEIDON_RP; r++) {
      Fr P = k.sc_r[r][0] * s + k.sc_k[r];
      pc_bar_sync2(PC_BAR_Q + buf);
      P = P + pc_ld(sh->mail[buf][2][lane]);
      pc_st(sh->mail[buf][1][lane], P);
      pc_bar_arrive2(PC_BAR_P + buf);
      Fr u = s + k.coop_cc[r][0];
      pc_bar_sync(PC_BAR_Y + buf);
      s = k.sc_a[r][0] * pc_ld(sh->mail[buf][0][lane]) + u;
      buf ^= 1;
    }
  } else {
    for (int r = 0; r < SVK_POSEIDON_RP; r++) {
      Fr P = k.sc_r[r][1] * s;
      pc_st(sh->mail[buf][2][lane], P);
      pc_bar_arrive2(PC_BAR_Q + buf);
      Fr u = s + k.coop_cc[r][1];
      pc_bar_sync(PC_BAR_Y + buf);
      s = k.sc_a[r][1] * pc_ld(sh->mail[buf][0][lane]) + u;
      buf ^= 1;
    }
  }
  for (int r = 0; r < SVK_POSEIDON_RF / 2 - 1; r++) full(k.mds, k.end[r]);
  full(k.mds, nullptr);
}

// Transcript owner's side (warp 0).  `s1` shadows the helper's word 1 after the last permutation.
struct PoseidonCoopMain {
  Fr s0, s1;
  PoseidonCoopShared* sh;
  int lane;
  int reset_pending;
};

__device__ __forceinline__ void pcm_init(PoseidonCoopMain& m, const PoseidonConsts& k) {
  m.s0 = k.capacity;
  m.s1 = Fr::zero();
  m.reset_pending = PC_CMD_RESET;
}

__device__ __forceinline__ void pcm_permute(PoseidonCoopMain& m, const PoseidonConsts& k, int n_in, const Fr& in0, const Fr& in1) {
  if (n_in >= 1) pc_st(m.sh->in[0][m.lane], in0);
  if (n_in >= 2) pc_st(m.sh->in[1][m.lane], in1);
  m.sh->n_in[m.lane] = n_in;
  if (m.lane == 0) m.sh->cmd = m.reset_pending;
  m.reset_pending = 0;
  pc_bar_sync(PC_BAR_CMD);
  pc_permute(m.s0, 0, m.lane, m.sh, k, n_in);
  pc_bar_sync(PC_BAR_FULL);  // helper 1 has posted word 1
  m.s1 = pc_ld(m.sh->out1[m.lane]);
}

__device__ __forceinline__ void pcm_exit(PoseidonCoopMain& m) {
  if (m.lane == 0) m.sh->cmd = PC_CMD_EXIT;
  pc_bar_sync(PC_BAR_CMD);
}

// Helper warps (role 1, 2): serve permutations until the owner exits.
__device__ __forceinline__ void pc_helper(PoseidonCoopShared* sh, const PoseidonConsts& k, int role, int lane) {
  Fr s = Fr::zero();
  for (;;) {
    pc_bar_sync(PC_BAR_CMD);
    int cmd = sh->cmd;
    if (cmd & PC_CMD_EXIT) return;
    if (cmd & PC_CMD_RESET) s = Fr::zero();
    pc_permute(s, role, lane, sh, k, sh->n_in[lane]);
    if (role == 1) pc_st(sh->out1[lane], s);
    pc_bar_sync(PC_BAR_FULL);
  }
}
#endif
