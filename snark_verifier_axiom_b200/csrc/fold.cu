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
//   k_fold_sponge : one group per thread: Poseidon sponge -> r, then the scalars r^j (canonical)
//   k_group_msm   : one block per (group, lhs|rhs): windowed scalar multiplications strided over the
//                   block's threads, shared-memory tree reduction of the partial sums, to_affine
#include "g1.cuh"
#include "poseidon.cuh"
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

__device__ __noinline__ G1Jac fold_mul_window4(const G1Affine& p, const u32* k) {
  G1Jac tbl[16];
  tbl[0] = G1Jac::identity();
  tbl[1] = G1Jac::from_affine(p);
  tbl[2] = tbl[1].dbl();
  for (int i = 3; i < 16; i++) tbl[i] = tbl[i - 1].add_affine(p);
  G1Jac acc = G1Jac::identity();
  for (int w = 63; w >= 0; w--) {
    if (w != 63) acc = acc.dbl().dbl().dbl().dbl();
    u32 d = (k[w >> 3] >> ((w & 7) * 4)) & 0xf;
    acc = acc.add(tbl[d]);
  }
  return acc;
}

// grid = (n_groups, 2); blockDim = L (power of two); dynamic smem = L * sizeof(G1Jac)
__global__ void k_group_msm(size_t n, size_t m, const uint8_t* accs, const u32* scalars, uint8_t* out_accs, size_t out_stride,
                            int32_t* status, size_t status_stride_words) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  G1Jac* sm = reinterpret_cast<G1Jac*>(smem_raw);
  size_t G = blockIdx.x;
  int h = blockIdx.y;
  size_t n_groups = (n + m - 1) / m;
  size_t seg = G / n_groups, g = G % n_groups;
  size_t begin = seg * n + g * m, end = (g * m + m < n ? g * m + m : n) + seg * n;
  G1Jac acc = G1Jac::identity();
  for (size_t i = begin + threadIdx.x; i < end; i += blockDim.x) {
    G1Affine b;
    load_canon32(b.x.v, accs + i * 128 + h * 64);
    load_canon32(b.y.v, accs + i * 128 + h * 64 + 32);
    bool canon = Fq::is_canonical(b.x.v) && Fq::is_canonical(b.y.v);
    if (!b.is_identity()) { b.x = b.x.to_mont(); b.y = b.y.to_mont(); }
    if (!canon || !g1_on_curve(b)) {
      atomicMax(status + seg * status_stride_words, SVK_TRANSCRIPT | (SVK_T_POINT_INVALID << 8));
      continue;
    }
    if (i == begin) {
      acc = acc.add_affine(b);
    } else {
      u32 k[8];
      load_canon32(k, reinterpret_cast<const uint8_t*>(scalars + i * 8));
      acc = acc.add(fold_mul_window4(b, k));
    }
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (unsigned s = blockDim.x / 2; s >= 1; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] = sm[threadIdx.x].add(sm[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    G1Affine a = sm[0].to_affine();
    Fq x = a.x.from_mont(), y = a.y.from_mont();
    uint4* o = reinterpret_cast<uint4*>(out_accs + G * out_stride + h * 64);
    o[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    o[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
    o[2] = make_uint4(y.v[0], y.v[1], y.v[2], y.v[3]);
    o[3] = make_uint4(y.v[4], y.v[5], y.v[6], y.v[7]);
  }
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
    SVK_LAUNCH(ctx, "k_fold_sponge",
               k_fold_sponge<<<(unsigned)((total_groups + 31) / 32), 32, 0, s>>>(n_seg, cnt, m, cur, ctx->d_poseidon, d_scal, last ? d_r : nullptr,
                                                                                out_stride / 4, d_status, out_stride / 4));
    unsigned L = 32;
    while (L < m && L < 256) L <<= 1;
    dim3 grid((unsigned)total_groups, 2);
    SVK_LAUNCH(ctx, "k_group_msm",
               k_group_msm<<<grid, L, L * sizeof(G1Jac), s>>>(cnt, m, cur, d_scal, dst, last ? out_stride : 128, d_status, out_stride / 4));
    if (last) break;
    cur = dst;
    cnt = groups;
    which ^= 1;
  }
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
