// K3 + group MSM: `KzgAs::{read_proof, verify}` / `create_proof` with zk = false
// (snark-verifier/src/pcs/kzg/accumulation.rs:29-62, 113-136, 139-196; call pattern
// snark-verifier-sdk/src/halo2/aggregation.rs:235-245):
//     absorb lhs_i, rhs_i of every accumulator (as [x mod r, y mod r]) -> r = squeeze
//     lhs = sum r^i lhs_i ,  rhs = sum r^i rhs_i
// `group_size` = 0 (or >= n) is the reference's flat fold: one serial sponge over all 4n coordinates.
// Otherwise consecutive groups of `group_size` are folded independently (each with a fresh transcript,
// as aggregation.rs:216 constructs one per aggregation) and the group results are folded again until
// one accumulator remains -- the same tree the oracle's `api.fold` performs.
//
// Per tree level, by its size (thresholds in svk_ctx.h; both forms give the same bytes, tests/test_gpu_schedules.py):
//   wide    k_fold_sponge (one group per thread) or k_fold_sponge_coop (32 groups per block of three warps): sponge -> r, scalars r^j;
//           k_group_var: the group MSMs as interleaved (Straus) multiplications; k_group_sum: base_0 + partial sums, to_affine
//   narrow  k_fold_sponge_dbl: the sponge blocks, and beside them blocks that double every base 255 times; k_fold_add: r^j A_j as
//           additions of those doublings over the NAF of r^j, 8 or 32 lanes per point; k_group_sum
#include "straus.cuh"
#include "poseidon_coop.cuh"
#include "svk_ctx.h"

__device__ __forceinline__ void load_canon32(u32* v, const uint8_t* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 lo = q[0], hi = q[1];
  v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
}

// status word: 0 ok, else SVK_TRANSCRIPT | sub << 8 (identity point: common_ec_point fails,
// transcript/halo2.rs:214-224; non-canonical coordinate: not a G1Affine at all)
// Segments: `n_seg` independent batches of `n` accumulators each, laid out back to back; groups never cross a
// segment boundary, so one launch folds the same tree level of every batch.
__global__ void __launch_bounds__(32) k_fold_sponge(size_t n_seg, size_t n, size_t m, const uint8_t* accs, const PoseidonConsts* pk, u32* scalars,
                                                    u32* out_r, size_t out_r_stride_words, int32_t* status, size_t status_stride_words) {
  size_t G = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t n_groups = (n + m - 1) / m;
  if (G >= n_groups * n_seg) return;
  size_t seg = G / n_groups, g = G % n_groups;
  size_t begin = seg * n + g * m, end = (g * m + m < n ? g * m + m : n) + seg * n;
  PoseidonState st;
  poseidon_init(st, *pk);
  int32_t bad = 0;
  for (size_t i = begin; i < end; i++) {
    for (int h = 0; h < 2; h++) {  // lhs, rhs
      u32 x[8], y[8];
      load_canon32(x, accs + i * 128 + h * 64);
      load_canon32(y, accs + i * 128 + h * 64 + 32);
      u32 any = 0;
      for (int k = 0; k < 8; k++) any |= x[k] | y[k];
      if (any == 0) bad = SVK_TRANSCRIPT | (SVK_T_POINT_IDENTITY << 8);
      if (!Fq::is_canonical(x) || !Fq::is_canonical(y)) bad = SVK_TRANSCRIPT | (SVK_T_POINT_INVALID << 8);
      Fr fx, fy;
      fq_canon_to_fr_canon(fx.v, x);
      fq_canon_to_fr_canon(fy.v, y);
      poseidon_permute(st, *pk, 2, fx.to_mont(), fy.to_mont());
    }
  }
  poseidon_permute(st, *pk, 0, Fr::zero(), Fr::zero());  // buffer length is a multiple of RATE (poseidon.rs:462-464)
  Fr r = st.s[1];
  Fr rc = r.from_mont();
  if (out_r)
    for (int k = 0; k < 8; k++) out_r[seg * out_r_stride_words + k] = rc.v[k];
  // powers r^j, j = 0.. (loader.rs:71-78); r^0 = 1 is not stored (the MSM adds the base directly)
  Fr p = r;
  for (size_t i = begin + 1; i < end; i++) {
    Fr pc = p.from_mont();
    uint4* o = reinterpret_cast<uint4*>(scalars + i * 8);
    o[0] = make_uint4(pc.v[0], pc.v[1], pc.v[2], pc.v[3]);
    o[1] = make_uint4(pc.v[4], pc.v[5], pc.v[6], pc.v[7]);
    p = p * r;
  }
  if (bad) atomicMax(status + seg * status_stride_words, bad);
}

// Latency form (poseidon_coop.cuh): 32 groups per block of three warps.  The levels of the fold tree of a 4096-proof batch
// (groups of 4) have 1024, 256, 64, 16, 4 and 1 groups, each a lone serial chain of 2m + 1 permutations.  The permutation count is made uniform
// across the warp (named barriers need whole warps): a lane whose group is shorter than the longest of its warp (the last
// group of a segment, lanes past the end) captures its r after its own 2 len + 1 permutations and keeps permuting a dead state.
__device__ __forceinline__ void fold_sponge_coop_block(PoseidonCoopShared& sh, size_t n_seg, size_t n, size_t m, const uint8_t* accs,
                                                       const PoseidonConsts* pk, u32* scalars, u32* out_r, size_t out_r_stride_words,
                                                       int32_t* status, size_t status_stride_words) {
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp) {
    pc_helper(&sh, *pk, warp, lane);
    return;
  }
  size_t n_groups = (n + m - 1) / m;
  size_t G = (size_t)blockIdx.x * 32 + lane;
  if (G >= n_groups * n_seg) G = n_groups * n_seg - 1;  // duplicates store identical values
  size_t seg = G / n_groups, g = G % n_groups;
  size_t begin = seg * n + g * m, end = (g * m + m < n ? g * m + m : n) + seg * n;
  u32 len = (u32)(end - begin), max_len = len;
  for (int d = 16; d >= 1; d >>= 1) max_len = max(max_len, __shfl_xor_sync(0xffffffffu, max_len, d));
  PoseidonCoopMain st;
  st.sh = &sh;
  st.lane = lane;
  pcm_init(st, *pk);
  int32_t bad = 0;
  Fr r = Fr::zero();
  for (u32 q = 0; q < 2 * max_len + 1; q++) {
    Fr fx = Fr::zero(), fy = Fr::zero();
    int n_in = 0;
    if (q < 2 * len) {
      size_t i = begin + (q >> 1);
      int h = (int)(q & 1);
      u32 x[8], y[8];
      load_canon32(x, accs + i * 128 + h * 64);
      load_canon32(y, accs + i * 128 + h * 64 + 32);
      u32 any = 0;
      for (int k = 0; k < 8; k++) any |= x[k] | y[k];
      if (any == 0) bad = SVK_TRANSCRIPT | (SVK_T_POINT_IDENTITY << 8);
      if (!Fq::is_canonical(x) || !Fq::is_canonical(y)) bad = SVK_TRANSCRIPT | (SVK_T_POINT_INVALID << 8);
      fq_canon_to_fr_canon(fx.v, x);
      fq_canon_to_fr_canon(fy.v, y);
      fx = fx.to_mont();
      fy = fy.to_mont();
      n_in = 2;
    }
    pcm_permute(st, *pk, n_in, fx, fy);
    if (q == 2 * len) r = st.s1;
  }
  pcm_exit(st);
  Fr rc = r.from_mont();
  if (out_r)
    for (int k = 0; k < 8; k++) out_r[seg * out_r_stride_words + k] = rc.v[k];
  Fr p = r;
  for (size_t i = begin + 1; i < end; i++) {
    Fr pc = p.from_mont();
    uint4* o = reinterpret_cast<uint4*>(scalars + i * 8);
    o[0] = make_uint4(pc.v[0], pc.v[1], pc.v[2], pc.v[3]);
    o[1] = make_uint4(pc.v[4], pc.v[5], pc.v[6], pc.v[7]);
    p = p * r;
  }
  if (bad) atomicMax(status + seg * status_stride_words, bad);
}

__global__ void __launch_bounds__(PCOOP_THREADS) k_fold_sponge_coop(size_t n_seg, size_t n, size_t m, const uint8_t* accs, const PoseidonConsts* pk,
                                                                    u32* scalars, u32* out_r, size_t out_r_stride_words, int32_t* status,
                                                                    size_t status_stride_words) {
  __shared__ PoseidonCoopShared sh;
  fold_sponge_coop_block(sh, n_seg, n, m, accs, pk, scalars, out_r, out_r_stride_words, status, status_stride_words);
}

// Group MSM, split like the per-proof MSM (verify.cu):
//   k_group_var  `vpl` threads per (group, side); a thread owns terms j = 1 + lane, 1 + lane + vpl, ... of its group
//                and runs them as ONE interleaved (Straus) multiplication (straus.cuh): 255 shared doublings + per term a
//                16-entry Jacobian table and <= 52 signed-window additions.  vpl = 1 where groups are plentiful (minimal work:
//                1785 + 7 x ~800 M for a group of 8 instead of 7 full multiplications), vpl = m - 1 on the upper, narrow levels
//                (minimal latency).  The first version ran one block of 32 threads per (group, side) with 7 active
//                lanes and was the largest consumer of issue slots of the whole step (profiles/r1_notes.md).
//   k_group_sum  one (group, side) per thread: base_0 (scalar r^0 = 1) + partial sums, to_affine.
#define FOLD_TERMS_MAX 16
__device__ __forceinline__ bool load_acc_point(G1Affine& b, const uint8_t* p) {
  load_canon32(b.x.v, p);
  load_canon32(b.y.v, p + 32);
  bool canon = Fq::is_canonical(b.x.v) && Fq::is_canonical(b.y.v);
  if (!b.is_identity()) { b.x = b.x.to_mont(); b.y = b.y.to_mont(); }
  return canon && g1_on_curve(b);
}

template <bool AFFINE>
__global__ void __launch_bounds__(64) k_group_var(size_t n_seg, size_t n, size_t m, u32 vpl, const uint8_t* accs, const u32* scalars,
                                                  G1Jac* tables, Fq* prefix, G1Jac* partials, int32_t* status, size_t status_stride_words) {
  size_t n_groups = (n + m - 1) / m;
  size_t n_threads = n_seg * n_groups * 2 * vpl;
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_threads) return;
  u32 lane = (u32)(gid % vpl);
  size_t gs = gid / vpl;  // (G, side)
  int h = (int)(gs & 1);
  size_t G = gs >> 1;
  size_t seg = G / n_groups, g = G % n_groups;
  size_t begin = seg * n + g * m, end = (g * m + m < n ? g * m + m : n) + seg * n;
  u32 k[FOLD_TERMS_MAX][8];
  u32 nt = 0;
  if (!AFFINE && vpl + 1 >= m) {
    // one member per thread: GLV halves the doubling chain (straus.cuh straus_run_glv1)
    size_t i = begin + 1 + lane;
    G1Jac acc = G1Jac::identity();
    if (i < end) {
      u32 kk[8], k1[4], k2[4], n1, n2;
      load_canon32(kk, reinterpret_cast<const uint8_t*>(scalars + i * 8));
      G1Affine base;
      if (!load_acc_point(base, accs + i * 128 + h * 64)) {
        atomicMax(status + seg * status_stride_words, SVK_TRANSCRIPT | (SVK_T_POINT_INVALID << 8));
        base = G1Affine::identity();
      }
      if (glv_decompose(kk, k1, n1, k2, n2)) {
        straus_build_table(tables + gid, n_threads, base);
        acc = straus_run_glv1(k1, n1, k2, n2, tables + gid, n_threads, glv_beta_mont());
      } else {  // cannot happen for a canonical scalar; the plain path gives the same point
        straus_recode(kk);
        straus_build_table(tables + gid, n_threads, base);
        acc = straus_run<false>(kk, 1, tables + gid, n_threads);
      }
    }
    partials[gid] = acc;
    return;
  }
  for (size_t i = begin + 1 + lane; i < end && nt < FOLD_TERMS_MAX; i += vpl, nt++) {
    load_canon32(k[nt], reinterpret_cast<const uint8_t*>(scalars + i * 8));
    G1Affine base;
    if (!load_acc_point(base, accs + i * 128 + h * 64)) {
      atomicMax(status + seg * status_stride_words, SVK_TRANSCRIPT | (SVK_T_POINT_INVALID << 8));
      base = G1Affine::identity();
    }
    straus_recode(k[nt]);
    straus_build_table(tables + ((size_t)nt * STRAUS_TABLE) * n_threads + gid, n_threads, base);
  }
  if (AFFINE) straus_normalize(tables + gid, prefix + gid, n_threads, nt * STRAUS_TABLE);
  G1Jac acc = straus_run<AFFINE>(&k[0][0], nt, tables + gid, n_threads);
  partials[gid] = acc;
}

__global__ void __launch_bounds__(128) k_group_sum(size_t n_seg, size_t n, size_t m, u32 vpl, const uint8_t* accs, const G1Jac* partials,
                                                   uint8_t* out_accs, size_t out_stride, int32_t* status, size_t status_stride_words) {
  size_t n_groups = (n + m - 1) / m;
  size_t gs = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gs >= n_seg * n_groups * 2) return;
  int h = (int)(gs & 1);
  size_t G = gs >> 1;
  size_t seg = G / n_groups, g = G % n_groups;
  size_t begin = seg * n + g * m;
  G1Affine b0;
  if (!load_acc_point(b0, accs + begin * 128 + h * 64)) {
    atomicMax(status + seg * status_stride_words, SVK_TRANSCRIPT | (SVK_T_POINT_INVALID << 8));
    b0 = G1Affine::identity();
  }
  G1Jac acc = G1Jac::from_affine(b0);
  for (u32 l = 0; l < vpl; l++) acc = acc.add(partials[gs * vpl + l]);
  G1Affine a = acc.to_affine();
  Fq x = a.x.from_mont(), y = a.y.from_mont();
  uint4* o = reinterpret_cast<uint4*>(out_accs + G * out_stride + h * 64);
  o[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  o[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
  o[2] = make_uint4(y.v[0], y.v[1], y.v[2], y.v[3]);
  o[3] = make_uint4(y.v[4], y.v[5], y.v[6], y.v[7]);
}

// ---- Narrow levels: the doublings leave the critical path ---------------------------------------------------------------
// A level of the tree is  sponge -> r -> sum_j r^j A_j.  The 254 doublings of a scalar multiplication do not depend on the
// scalar when they are applied to the BASE (right-to-left): D_b = 2^b A_j.  So while the sponge blocks hash (a lone chain of
// 2m + 1 permutations, ~2 ms), the other blocks of the SAME launch compute D_0..D_254 of every accumulator; once r is
// known, r^j A_j = sum_b naf_b(r^j) D_b is ~85 independent additions, spread over 8 lanes and tree-reduced: ~0.1 ms
// instead of a 255-doubling chain (1.4 ms) after the sponge.  Used while the table (255 x 96 B per point) stays small.
#define FOLD_DBL_ENTRIES 255
__global__ void __launch_bounds__(PCOOP_THREADS) k_fold_sponge_dbl(unsigned sponge_blocks, size_t n_seg, size_t n, size_t m, const uint8_t* accs,
                                                                   const PoseidonConsts* pk, u32* scalars, u32* out_r, size_t out_r_stride_words,
                                                                   int32_t* status, size_t status_stride_words, G1Jac* dbl) {
  __shared__ PoseidonCoopShared sh;
  if (blockIdx.x < sponge_blocks) {
    fold_sponge_coop_block(sh, n_seg, n, m, accs, pk, scalars, out_r, out_r_stride_words, status, status_stride_words);
    return;
  }
  size_t n_threads = n_seg * n * 2;
  size_t tid = (size_t)(blockIdx.x - sponge_blocks) * PCOOP_THREADS + threadIdx.x;
  if (tid >= n_threads) return;
  size_t i = tid >> 1;
  int h = (int)(tid & 1);
  size_t seg = i / n, li = i % n;
  if (li % m == 0) return;  // r^0 = 1: k_group_sum adds this base as it is
  G1Affine base;
  if (!load_acc_point(base, accs + i * 128 + h * 64)) {
    atomicMax(status + seg * status_stride_words, SVK_TRANSCRIPT | (SVK_T_POINT_INVALID << 8));
    base = G1Affine::identity();
  }
  G1Jac d = G1Jac::from_affine(base);
  for (int b = 0; b < FOLD_DBL_ENTRIES; b++) {
    dbl[(size_t)b * n_threads + tid] = d;
    d = d.dbl();
  }
}

__device__ __forceinline__ G1Jac shfl_xor_jac(const G1Jac& p, int mask) {
  G1Jac r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.X.v[i] = __shfl_xor_sync(0xffffffffu, p.X.v[i], mask);
    r.Y.v[i] = __shfl_xor_sync(0xffffffffu, p.Y.v[i], mask);
    r.Z.v[i] = __shfl_xor_sync(0xffffffffu, p.Z.v[i], mask);
  }
  return r;
}

// partials[(G * 2 + h) * vpl + (j - 1)] = r^j A_j for member j >= 1 of group G; LANES lanes per (accumulator, side).
// A lane first packs the positions of its non-zero NAF digits, then adds: the trip count of a warp is the longest list of its
// lanes (~ a third of the positions), not the number of positions (a digit test inside the loop made every position cost an
// addition for the whole warp: 0.41 ms per level instead of 0.1).
template <int LANES>
__global__ void __launch_bounds__(64) k_fold_add(size_t n_seg, size_t n, size_t m, u32 vpl, const u32* scalars, const G1Jac* dbl, G1Jac* partials) {
  size_t n_threads = n_seg * n * 2;
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t tid = gid / LANES;
  u32 lane = (u32)(gid % LANES);
  bool active = tid < n_threads;
  if (!active) tid = n_threads - 1;  // whole warps for the shuffles
  size_t i = tid >> 1;
  int h = (int)(tid & 1);
  size_t seg = i / n, li = i % n, n_groups = (n + m - 1) / m;
  size_t g = li / m, j = li % m;
  G1Jac acc = G1Jac::identity();
  constexpr int per = 256 / LANES;
  // non-zero digits of this lane: bit q of `nz`, sign in `ng`
  u32 nz = 0, ng = 0;
  if (j != 0) {
    u32 k[9], k3[10];
    load_canon32(k, reinterpret_cast<const uint8_t*>(scalars + i * 8));
    k[8] = 0;
    // 3k; NAF digit at weight 2^pos = bit(3k, pos + 1) - bit(k, pos + 1)
    u32 c = 0;
    for (int w = 0; w < 9; w++) {
      u64 t = (u64)k[w] * 3 + c;
      k3[w] = (u32)t;
      c = (u32)(t >> 32);
    }
    k3[9] = 0;
    for (int q = 0; q < per; q++) {
      int pos = (int)lane * per + q;
      if (pos >= FOLD_DBL_ENTRIES) break;
      int bit = pos + 1;
      u32 a = (k3[bit >> 5] >> (bit & 31)) & 1, b = (k[bit >> 5] >> (bit & 31)) & 1;
      nz |= (a ^ b) << q;
      ng |= b << q;
    }
  }
  while (__any_sync(0xffffffffu, nz != 0)) {
    if (nz) {
      int q = __ffs(nz) - 1;
      nz &= nz - 1;
      G1Jac d = dbl[((size_t)lane * per + q) * n_threads + tid];
      if ((ng >> q) & 1) d.Y = d.Y.neg();
      acc = acc.add(d);
    }
  }
#pragma unroll
  for (int d = LANES / 2; d >= 1; d >>= 1) acc = acc.add(shfl_xor_jac(acc, d));  // butterfly: every lane ends with the sum of its group
  if (active && lane == 0 && j != 0) partials[((seg * n_groups + g) * 2 + h) * vpl + (j - 1)] = acc;
}

// Folds `n_seg` independent batches of `n` accumulators each (d_accs: [seg][n] x 128 B) down to one accumulator per
// batch, written to d_out + seg * out_stride: { svk_acc (128) ; r of the last fold call (32) ; int32 status }.
// The fold owns scratch slots 8 (ping-pong accumulators) and 9 (scalars).
int svk_fold_launch_seg(svk_ctx* ctx, size_t n_seg, size_t n, const uint8_t* d_accs, size_t group_size, uint8_t* d_out, size_t out_stride) {
  if (n == 0 || n_seg == 0) return svk_fail(ctx, "fold of zero accumulators (`assert!(!instances.is_empty())`, accumulation.rs:121)");
  if (out_stride < 164 || out_stride % 4) return svk_fail(ctx, "fold: bad output stride");
  cudaStream_t s = ctx->stream;
  size_t m = (group_size <= 1 || group_size >= n) ? n : group_size;
  size_t first_groups = (n + m - 1) / m;
  uint8_t* d_tmp;
  u32* d_scal;
  size_t lvl0 = n_seg * first_groups, lvl1 = n_seg * ((first_groups + m - 1) / m);
  if (svk_scratch(ctx, 8, (lvl0 + lvl1 + 2) * 128, (void**)&d_tmp)) return -1;
  if (svk_scratch(ctx, 9, n_seg * n * 32 + 32, (void**)&d_scal)) return -1;
  SVK_CUDA(ctx, cudaMemset2DAsync(d_out + 160, out_stride, 0, 4, n_seg, s));
  int32_t* d_status = (int32_t*)(d_out + 160);
  u32* d_r = (u32*)(d_out + 128);
  const uint8_t* cur = d_accs;
  size_t cnt = n;
  uint8_t* bufs[2] = {d_tmp, d_tmp + (lvl0 + 1) * 128};
  int which = 0;
  for (;;) {
    size_t groups = (cnt + m - 1) / m;
    bool last = groups == 1;
    uint8_t* dst = last ? d_out : bufs[which];
    size_t total_groups = groups * n_seg;
    size_t gm = cnt < m ? cnt : m;
    size_t pt_threads = n_seg * cnt * 2;
    G1Jac *d_tables, *d_partials;
    u32 vpl = 1;
    if (gm > 1 && gm - 1 <= 64 && pt_threads <= ctx->fold_dbl_threads_max && total_groups <= ctx->tape_coop_max) {
      // narrow level: doublings of the bases run beside the sponge, the scalar multiplications become additions
      vpl = (u32)(gm - 1);
      size_t n_part = total_groups * 2 * vpl;
      if (svk_scratch(ctx, 20, (size_t)FOLD_DBL_ENTRIES * pt_threads * sizeof(G1Jac), (void**)&d_tables)) return -1;
      if (svk_scratch(ctx, 18, (n_part + 1) * sizeof(G1Jac), (void**)&d_partials)) return -1;
      SVK_CUDA(ctx, cudaMemsetAsync(d_partials, 0, n_part * sizeof(G1Jac), s));  // absent members of a short last group: identity (Z = 0)
      unsigned sponge_blocks = (unsigned)((total_groups + 31) / 32), dbl_blocks = (unsigned)((pt_threads + PCOOP_THREADS - 1) / PCOOP_THREADS);
      SVK_LAUNCH(ctx, "k_fold_sponge_dbl",
                 k_fold_sponge_dbl<<<sponge_blocks + dbl_blocks, PCOOP_THREADS, 0, s>>>(sponge_blocks, n_seg, cnt, m, cur, ctx->d_poseidon, d_scal,
                                                                                       last ? d_r : nullptr, out_stride / 4, d_status, out_stride / 4,
                                                                                       d_tables));
      // 8 lanes per point while the points alone fill the machine, 32 on the narrow levels
      if (pt_threads >= 4096)
        SVK_LAUNCH(ctx, "k_fold_add", k_fold_add<8><<<(unsigned)((pt_threads * 8 + 63) / 64), 64, 0, s>>>(n_seg, cnt, m, vpl, d_scal, d_tables, d_partials));
      else
        SVK_LAUNCH(ctx, "k_fold_add", k_fold_add<32><<<(unsigned)((pt_threads * 32 + 63) / 64), 64, 0, s>>>(n_seg, cnt, m, vpl, d_scal, d_tables, d_partials));
    } else {
    if (total_groups <= ctx->tape_coop_max)
      SVK_LAUNCH(ctx, "k_fold_sponge_coop",
                 k_fold_sponge_coop<<<(unsigned)((total_groups + 31) / 32), PCOOP_THREADS, 0, s>>>(n_seg, cnt, m, cur, ctx->d_poseidon, d_scal,
                                                                                                   last ? d_r : nullptr, out_stride / 4, d_status,
                                                                                                   out_stride / 4));
    else
      SVK_LAUNCH(ctx, "k_fold_sponge",
                 k_fold_sponge<<<(unsigned)((total_groups + 31) / 32), 32, 0, s>>>(n_seg, cnt, m, cur, ctx->d_poseidon, d_scal, last ? d_r : nullptr,
                                                                                  out_stride / 4, d_status, out_stride / 4));
    // lanes per (group, side): 1 while there are enough groups to fill the machine, else one lane per term
    if (total_groups * 2 < ctx->fold_lanes_groups_max && gm > 1) vpl = (u32)(gm - 1);
    size_t terms_per_thread = gm > 1 ? (gm - 1 + vpl - 1) / vpl : 0;
    if (terms_per_thread > FOLD_TERMS_MAX) vpl = (u32)((gm - 1 + FOLD_TERMS_MAX - 1) / FOLD_TERMS_MAX), terms_per_thread = (gm - 1 + vpl - 1) / vpl;
    size_t var_threads = total_groups * 2 * vpl;
    if (svk_scratch(ctx, 17, (terms_per_thread * 16 * var_threads + 1) * sizeof(G1Jac), (void**)&d_tables)) return -1;
    if (svk_scratch(ctx, 18, (var_threads + 1) * sizeof(G1Jac), (void**)&d_partials)) return -1;
    // threads that own >= 3 terms normalise their tables to affine (one inversion per thread) and use mixed additions
    if (terms_per_thread >= 3) {
      Fq* d_prefix;
      if (svk_scratch(ctx, 22, (terms_per_thread * 16 * var_threads + 1) * sizeof(Fq), (void**)&d_prefix)) return -1;
      SVK_LAUNCH(ctx, "k_group_var",
                 k_group_var<true><<<(unsigned)((var_threads + 63) / 64), 64, 0, s>>>(n_seg, cnt, m, vpl, cur, d_scal, d_tables, d_prefix, d_partials,
                                                                                      d_status, out_stride / 4));
    } else {
      SVK_LAUNCH(ctx, "k_group_var",
                 k_group_var<false><<<(unsigned)((var_threads + 63) / 64), 64, 0, s>>>(n_seg, cnt, m, vpl, cur, d_scal, d_tables, nullptr, d_partials,
                                                                                       d_status, out_stride / 4));
    }
    }
    SVK_LAUNCH(ctx, "k_group_sum",
               k_group_sum<<<(unsigned)((total_groups * 2 + 127) / 128), 128, 0, s>>>(n_seg, cnt, m, vpl, cur, d_partials, dst,
                                                                                      last ? out_stride : 128, d_status, out_stride / 4));
    if (last) break;
    cur = dst;
    cnt = groups;
    which ^= 1;
  }
  SVK_CUDA(ctx, cudaGetLastError());
  return 0;
}

// `KzgAsProof::read` with zk = true (pcs/kzg/accumulation.rs:124-133): the two blind points are read from the accumulation proof
// as compressed G1 (32 B each) and absorbed like the instances; `verify` appends them as the LAST pair (:45-49).  So the zk fold
// of n accumulators IS the flat fold of n + 1 pairs, the last one decoded here.  An undecodable point is written as a
// non-canonical coordinate, an identity as (0, 0): the fold reports them as POINT_INVALID / POINT_IDENTITY (transcript/halo2.rs:214-260).
__global__ void k_blind_points(const uint8_t* blind, uint8_t* slot) {
  int t = threadIdx.x;
  if (t >= 2) return;
  G1Affine pt;
  u32 xc[8], yc[8];
  int rc = g1_decompress(blind + 32 * t, pt, xc, yc);
  uint4* o = reinterpret_cast<uint4*>(slot + 64 * t);
  if (rc == 0) {
    o[0] = make_uint4(xc[0], xc[1], xc[2], xc[3]);
    o[1] = make_uint4(xc[4], xc[5], xc[6], xc[7]);
    o[2] = make_uint4(yc[0], yc[1], yc[2], yc[3]);
    o[3] = make_uint4(yc[4], yc[5], yc[6], yc[7]);
  } else {
    u32 v = rc == 2 ? 0u : 0xffffffffu;
    for (int i = 0; i < 4; i++) o[i] = make_uint4(v, v, v, v);
  }
}

int svk_blind_points_launch(svk_ctx* ctx, const uint8_t* d_blind, uint8_t* d_slot) {
  SVK_LAUNCH(ctx, "k_blind_points", k_blind_points<<<1, 32, 0, ctx->stream>>>(d_blind, d_slot));
  SVK_CUDA(ctx, cudaGetLastError());
  return 0;
}

// single batch, separate output pointers (svk_kzg_as_fold_dev): staged through a 256-byte record
int svk_fold_launch(svk_ctx* ctx, size_t n, const uint8_t* d_accs, size_t group_size, uint8_t* d_out_acc, u32* d_out_r, int32_t* d_status) {
  uint8_t* rec;
  if (svk_scratch(ctx, 16, 256, (void**)&rec)) return -1;
  if (svk_fold_launch_seg(ctx, 1, n, d_accs, group_size, rec, 256)) return -1;
  cudaStream_t s = ctx->stream;
  SVK_CUDA(ctx, cudaMemcpyAsync(d_out_acc, rec, 128, cudaMemcpyDeviceToDevice, s));
  if (d_out_r) SVK_CUDA(ctx, cudaMemcpyAsync(d_out_r, rec + 128, 32, cudaMemcpyDeviceToDevice, s));
  SVK_CUDA(ctx, cudaMemcpyAsync(d_status, rec + 160, 4, cudaMemcpyDeviceToDevice, s));
  return 0;
}

// Per batch: ok = decide_ok && every proof status == 0 && fold_status == 0   (PlonkVerifier::verify over the batch).
// One block per batch; records: { acc 128 ; r 32 ; fold_status 4 ; decide_ok 1 ; ok 1 } with stride rec_stride.
__global__ void k_batch_verdict(size_t batch, const int32_t* status, uint8_t* records, size_t rec_stride) {
  __shared__ int any_bad;
  if (threadIdx.x == 0) any_bad = 0;
  __syncthreads();
  size_t seg = blockIdx.x;
  int bad = 0;
  for (size_t i = threadIdx.x; i < batch; i += blockDim.x) bad |= status[seg * batch + i] != 0;
  if (bad) atomicOr(&any_bad, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint8_t* rec = records + seg * rec_stride;
    int32_t fold_status = *reinterpret_cast<const int32_t*>(rec + 160);
    rec[165] = (!any_bad && fold_status == 0 && rec[164]) ? 1 : 0;
  }
}

int svk_batch_verdict_launch(svk_ctx* ctx, size_t n_seg, size_t batch, const int32_t* d_status, uint8_t* d_records, size_t rec_stride) {
  SVK_LAUNCH(ctx, "k_batch_verdict", k_batch_verdict<<<(unsigned)n_seg, 256, 0, ctx->stream>>>(batch, d_status, d_records, rec_stride));
  SVK_CUDA(ctx, cudaGetLastError());
  return 0;
}
