// K4: Pippenger bucket MSM over BN254 G1 (and the Pasta curves of the IPA decider) + batched scalar multiplication.
//
// Replaces `util::msm::multi_scalar_multiplication` (snark-verifier/src/util/msm.rs:238-317; window
// ceil(ln n)+2, 2^w-1 buckets, running-sum reduction, rayon chunking) for BASELINE config 3, and is the
// large-n engine for `KzgAs::verify`'s fold MSM (pcs/kzg/accumulation.rs:51-59).  Results are canonical
// affine points, so any correct schedule is bit-exact (SURVEY finding 1).
//
// Pipeline (all on the context stream, no host round trips):
//   k_msm_prepare   one point per thread (coalesced 16-byte loads): Montgomery copy of the point; BN254: the GLV split
//                   k = k1 + k2 lambda (glv.cuh) and the second point phi(P) = (beta x, y) -- 2n "virtual" points with
//                   128-bit scalars: half the windows, half the Horner doublings; signed c-bit digits of every window
//                   (the top window stays unsigned), per-(window, bucket) histogram
//   k_msm_scan      one block per window: exclusive scan of the histogram -> bucket offsets
//   k_msm_scatter   counting sort of (point index | sign) by bucket id, per window
//   k_msm_order_*   counting sort of the BUCKETS by population (descending; warp-aggregated atomics: one atomic per distinct
//                   bin of a warp): the threads of a warp then own buckets of the same size and stay converged -- with
//                   buckets in index order a warp ran as long as its fullest bucket (~2x the mean at 8 points per bucket)
//   k_msm_buckets   one bucket per thread in that order: XYZZ mixed additions (8M + 2S) over its sorted run; buckets far
//                   above the mean (skewed scalars) are listed instead and
//   k_msm_heavy     summed by a whole warp each: lanes stride over the run, warp-shuffle butterfly of the 32 partial sums
//   k_msm_reduce    blocks of 256 threads per window: sum_b b*B_b over a bucket range by chunked running sums + shared-memory tree
//   k_msm_combine   one warp per window adds the block partials (shuffle butterfly), then Horner over the windows + to_affine
#include "glv.cuh"
#include "pasta.cuh"
#include "svk_ctx.h"

struct MsmPlan {
  u32 c;            // window bits
  u32 windows;      // number of windows
  u32 buckets;      // regular windows: signed digits, bucket ids 1..2^(c-1) (0 = digit zero)
  u32 top_buckets;  // top window: GLV: unsigned, ids 1..2^tb ; legacy: = buckets
  u32 nb;           // ids allocated per window = max(buckets, top_buckets) + 1
  u32 glv;          // 1: BN254, scalars split in two 128-bit halves over 2n virtual points
  u32 heavy;        // buckets with at least this many points go to k_msm_heavy
};

template <class C>
struct CurveGlv { static constexpr bool value = false; };
template <>
struct CurveGlv<CurveBn254> { static constexpr bool value = true; };

// Window choice.  GLV (BN254): 128-bit half-scalars over 2n points; the top window holds 128 - c (W - 1) bits and is left
// UNSIGNED (ids up to 2^tb), so no carry leaves it.  Legacy (Pasta, 255-bit scalars): the TOP window only holds
// 255 - c (W - 1) significant bits; c = 16 (W = 16) and c = 15 (W = 17) are the well-filled choices, small inputs use c = 8.
static MsmPlan msm_plan(size_t n, bool glv) {
  MsmPlan p;
  p.glv = glv ? 1 : 0;
  if (glv) {
    // measured on B200 (profiles/r2_notes.md): a FULL top window (tb = c: c = 16) beats every thinner plan from 2^16 points up --
    // with tb < c the top window has few, very full buckets
    u32 c = n >= (1u << 16) ? 16 : (n >= (1u << 13) ? 13 : (n >= (1u << 10) ? 10 : 7));
    if (const char* e = getenv("SVK_MSM_C")) c = (u32)atoi(e);  // tuning sweeps (profiles/r2_notes.md)
    p.c = c;
    p.windows = (128 + c - 1) / c;
    p.buckets = 1u << (c - 1);
    u32 tb = 128 - c * (p.windows - 1);
    p.top_buckets = 1u << tb;
  } else {
    u32 c = n >= (1u << 18) ? 16 : (n >= (1u << 13) ? 15 : 8);
    p.c = c;
    p.windows = (255 + c - 1) / c;  // 254/255-bit scalars + one carry bit
    p.buckets = 1u << (c - 1);
    p.top_buckets = p.buckets;
  }
  p.nb = (p.buckets > p.top_buckets ? p.buckets : p.top_buckets) + 1;
  size_t nv = glv ? 2 * n : n;
  // a bucket far above the mean (skewed scalars; the thinly populated top window of a small input: 2^16 points put 128 in
  // each of ~1000 top buckets against a mean of 32) would be one long serial chain: it goes to a warp instead
  size_t avg = nv / p.buckets + 1;
  p.heavy = (u32)(avg * 3 > 64 ? avg * 3 : 64);
  return p;
}

__device__ __forceinline__ void load32(u32* v, const uint8_t* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 lo = q[0], hi = q[1];
  v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
}

// digits of one (virtual) point: keys[w * nv + idx] = bucket id (0 = skip) | sign << 31.  `k`: limbs, zero-padded two limbs past
// the last window bit.
__device__ __forceinline__ void msm_digits(const MsmPlan& plan, const u32* k, u32 base_neg, bool skip, size_t nv, size_t idx, u32* keys, u32* hist) {
  u32 carry = 0;
  for (u32 w = 0; w < plan.windows; w++) {
    u32 bit = w * plan.c;
    u32 word = bit >> 5, sh = bit & 31;
    u64 two = (u64)k[word] | ((u64)k[word + 1] << 32);
    u32 d = (u32)((two >> sh) & ((1u << plan.c) - 1)) + carry;
    u32 neg = 0;
    carry = 0;
    bool top_unsigned = plan.glv && w == plan.windows - 1;
    if (!top_unsigned && d > plan.buckets) { d = (1u << plan.c) - d; neg = 1; carry = 1; }
    if (skip) d = 0;
    keys[(size_t)w * nv + idx] = d | ((neg ^ base_neg) << 31);
    if (d) atomicAdd(&hist[(size_t)w * plan.nb + d], 1u);
  }
}

template <class C>
__global__ void __launch_bounds__(256) k_msm_prepare(size_t n, MsmPlan plan, const uint8_t* scalars, const uint8_t* points,
                                                     AffT<typename C::Base>* pts_m, u32* keys, u32* hist, int* bad) {
  typedef typename C::Base Fq;
  typedef typename C::Scalar Fr;
  typedef AffT<Fq> G1Affine;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine p;
  load32(p.x.v, points + i * 64);
  load32(p.y.v, points + i * 64 + 32);
  bool canon = Fq::is_canonical(p.x.v) && Fq::is_canonical(p.y.v);
  bool ident = p.is_identity();
  if (!ident) { p.x = p.x.to_mont(); p.y = p.y.to_mont(); }
  if (!canon || !curve_on_curve<C>(p)) { *bad = 1; p = G1Affine::identity(); ident = true; }
  pts_m[i] = p;
  u32 k[11];
  load32(k, scalars + i * 32);
  k[8] = k[9] = k[10] = 0;
  if (!Fr::is_canonical(k)) { *bad = 2; ident = true; }
  if constexpr (CurveGlv<C>::value) {
    size_t nv = 2 * n;
    G1Affine q = p;
    if (!ident) q.x = p.x * glv_beta_mont();  // phi(P)
    pts_m[n + i] = q;
    u32 k1[7], k2[7], n1, n2;
    for (int j = 4; j < 7; j++) k1[j] = k2[j] = 0;
    if (!glv_decompose(k, k1, n1, k2, n2)) { *bad = 3; ident = true; }  // cannot happen for k < r (tools/glv_constants.py)
    msm_digits(plan, k1, n1, ident, nv, i, keys, hist);
    msm_digits(plan, k2, n2, ident, nv, n + i, keys, hist);
  } else {
    msm_digits(plan, k, 0, ident, n, i, keys, hist);
  }
}

// exclusive scan of hist[w][0..nb) -> offs[w][..]; one block (1024 threads) per window
__global__ void __launch_bounds__(1024) k_msm_scan(MsmPlan plan, const u32* hist, u32* offs, u32* cursor) {
  __shared__ u32 part[1024];
  u32 w = blockIdx.x, nb = plan.nb;
  const u32* h = hist + (size_t)w * nb;
  u32* o = offs + (size_t)w * nb;
  u32* cu = cursor + (size_t)w * nb;
  u32 per = (nb + blockDim.x - 1) / blockDim.x;
  u32 lo = min(threadIdx.x * per, nb), hi = min(lo + per, nb);
  u32 s = 0;
  for (u32 b = lo; b < hi; b++) s += h[b];
  part[threadIdx.x] = s;
  __syncthreads();
  for (u32 d = 1; d < blockDim.x; d <<= 1) {  // Hillis-Steele inclusive scan
    u32 v = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  u32 base = threadIdx.x ? part[threadIdx.x - 1] : 0;
  for (u32 b = lo; b < hi; b++) {
    o[b] = base;
    cu[b] = base;
    base += h[b];
  }
}

__global__ void __launch_bounds__(256) k_msm_scatter(size_t nv, MsmPlan plan, const u32* keys, u32* cursor, u32* sorted) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  u32 w = blockIdx.y;
  if (i >= nv) return;
  u32 key = keys[(size_t)w * nv + i];
  u32 d = key & 0x7fffffffu;
  if (!d) return;
  u32 pos = atomicAdd(&cursor[(size_t)w * plan.nb + d], 1u);
  sorted[(size_t)w * nv + pos] = (u32)i | (key & 0x80000000u);
}

// ---- buckets ordered by population (descending).  bin = min(count, 63); order[] lists bucket slots (w * nb + id).
#define ORDER_BINS 64
__global__ void __launch_bounds__(256) k_msm_order_hist(size_t total, const u32* hist, u32* bin_cnt) {
  __shared__ u32 sm[ORDER_BINS];
  if (threadIdx.x < ORDER_BINS) sm[threadIdx.x] = 0;
  __syncthreads();
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < total) atomicAdd(&sm[min(hist[g], (u32)ORDER_BINS - 1)], 1u);
  __syncthreads();
  if (threadIdx.x < ORDER_BINS && sm[threadIdx.x]) atomicAdd(&bin_cnt[threadIdx.x], sm[threadIdx.x]);
}
__global__ void k_msm_order_scan(const u32* bin_cnt, u32* bin_cursor) {
  if (threadIdx.x) return;
  u32 acc = 0;
  for (int b = ORDER_BINS - 1; b >= 0; b--) { bin_cursor[b] = acc; acc += bin_cnt[b]; }
}
__global__ void __launch_bounds__(256) k_msm_order_scatter(size_t total, const u32* hist, u32* bin_cursor, u32* order) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  u32 bin = g < total ? min(hist[g], (u32)ORDER_BINS - 1) : ORDER_BINS;  // lanes past the end share a dummy bin
  // warp-aggregated: one atomic per distinct bin of the warp
  u32 mask = __match_any_sync(0xffffffffu, bin);
  u32 lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
  u32 rank = __popc(mask & ((1u << lane) - 1));
  u32 base = 0;
  if (lane == leader && bin < ORDER_BINS) base = atomicAdd(&bin_cursor[bin], __popc(mask));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (bin < ORDER_BINS) order[base + rank] = (u32)g;
}

// one bucket per thread, in population order; bucket slot g = w * nb + id (XYZZ).  A bucket far above the mean is cut into chunks of
// HEAVY_CHUNK points, each summed by one warp (k_msm_heavy) and added up by k_msm_heavy_sum: skewed scalars (all equal, all
// small) or a thin top window put most points of a window into a handful of buckets.
#define HEAVY_CHUNK 1024
struct HeavyItem {
  u32 g, chunk, n_chunks, first;  // bucket slot, chunk index, chunks of this bucket, list index of its chunk 0
};
template <class Fq>
__global__ void __launch_bounds__(128) k_msm_buckets(size_t nv, MsmPlan plan, size_t total, const u32* order, const u32* hist, const u32* offs,
                                                     const u32* sorted, const AffT<Fq>* pts_m, XyzzT<Fq>* buckets, HeavyItem* heavy_list, u32* n_heavy,
                                                     u32 heavy_cap) {
  typedef AffT<Fq> G1Affine;
  typedef XyzzT<Fq> G1Xyzz;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  u32 g = order[t];
  u32 cnt = hist[g], off = offs[g];
  G1Xyzz acc = G1Xyzz::identity();
  if (cnt >= plan.heavy) {
    u32 nc = (cnt + HEAVY_CHUNK - 1) / HEAVY_CHUNK;
    u32 slot = atomicAdd(n_heavy, nc);
    if (slot + nc <= heavy_cap) {  // warps will sum this bucket
      for (u32 j = 0; j < nc; j++) heavy_list[slot + j] = HeavyItem{g, j, nc, slot};
      return;
    }
  }
  u32 w = g / plan.nb;
  const u32* run = sorted + (size_t)w * nv + off;
  for (u32 j = 0; j < cnt; j++) {
    u32 e = run[j];
    G1Affine p = pts_m[e & 0x7fffffffu];
    if (e >> 31) p.y = p.y.neg();
    acc = acc.add_affine(p);
  }
  buckets[g] = acc;
}

template <class Fq>
__device__ __forceinline__ XyzzT<Fq> shfl_xor_xyzz(const XyzzT<Fq>& p, int m) {
  XyzzT<Fq> r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.X.v[i] = __shfl_xor_sync(0xffffffffu, p.X.v[i], m);
    r.Y.v[i] = __shfl_xor_sync(0xffffffffu, p.Y.v[i], m);
    r.ZZ.v[i] = __shfl_xor_sync(0xffffffffu, p.ZZ.v[i], m);
    r.ZZZ.v[i] = __shfl_xor_sync(0xffffffffu, p.ZZZ.v[i], m);
  }
  return r;
}

// one warp per listed chunk: lanes stride over the chunk, then a shuffle butterfly of the 32 partial sums -> heavy_part[item]
template <class Fq>
__global__ void __launch_bounds__(32) k_msm_heavy(size_t nv, MsmPlan plan, const HeavyItem* heavy_list, const u32* n_heavy, u32 heavy_cap,
                                                  const u32* hist, const u32* offs, const u32* sorted, const AffT<Fq>* pts_m, XyzzT<Fq>* heavy_part) {
  typedef AffT<Fq> G1Affine;
  typedef XyzzT<Fq> G1Xyzz;
  u32 count = min(*n_heavy, heavy_cap);
  for (u32 h = blockIdx.x; h < count; h += gridDim.x) {
    HeavyItem it = heavy_list[h];
    if (it.n_chunks == 0) continue;  // slots past an overflowing bucket stay zero
    u32 cnt = hist[it.g], off = offs[it.g], w = it.g / plan.nb;
    u32 lo = it.chunk * HEAVY_CHUNK, hi = min(lo + HEAVY_CHUNK, cnt);
    const u32* run = sorted + (size_t)w * nv + off;
    G1Xyzz acc = G1Xyzz::identity();
    for (u32 j = lo + threadIdx.x; j < hi; j += 32) {
      u32 e = run[j];
      G1Affine p = pts_m[e & 0x7fffffffu];
      if (e >> 31) p.y = p.y.neg();
      acc = acc.add_affine(p);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) acc = acc.add(shfl_xor_xyzz(acc, d));
    if (threadIdx.x == 0) heavy_part[h] = acc;
  }
}

// one warp per heavy bucket (its chunk-0 item): lanes stride over the chunk partials, butterfly, bucket written
template <class Fq>
__global__ void __launch_bounds__(32) k_msm_heavy_sum(const HeavyItem* heavy_list, const u32* n_heavy, u32 heavy_cap, const XyzzT<Fq>* heavy_part,
                                                      XyzzT<Fq>* buckets) {
  typedef XyzzT<Fq> G1Xyzz;
  u32 count = min(*n_heavy, heavy_cap);
  for (u32 h = blockIdx.x; h < count; h += gridDim.x) {
    HeavyItem it = heavy_list[h];
    if (it.n_chunks == 0 || it.chunk != 0) continue;
    G1Xyzz acc = G1Xyzz::identity();
    for (u32 j = threadIdx.x; j < it.n_chunks; j += 32) acc = acc.add(heavy_part[it.first + j]);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) acc = acc.add(shfl_xor_xyzz(acc, d));
    if (threadIdx.x == 0) buckets[it.g] = acc;
  }
}

// S_w = sum_{b=1..B_w} b * bucket[w][b], in two stages: `rb` blocks of REDUCE_T threads per window each reduce a contiguous range of
// buckets to one partial (chunked running sums per thread, + lo * (plain sum) for the chunk's offset, shared-memory tree), then
// k_msm_combine adds the partials of a window.
#define REDUCE_T 256
// Blocks per window: `rb` for a regular window, `rbt` for the top one (twice the buckets when it is unsigned); all blocks of the
// launch are resident at once (one per SM: 185 registers), so the kernel time is one block's chain.
template <class Fq>
__global__ void __launch_bounds__(REDUCE_T) k_msm_reduce(MsmPlan plan, u32 rb, u32 rbt, const XyzzT<Fq>* buckets, XyzzT<Fq>* partials) {
  typedef XyzzT<Fq> G1Xyzz;
  __shared__ G1Xyzz sm[REDUCE_T];
  u32 w = blockIdx.y, B = w == plan.windows - 1 ? plan.top_buckets : plan.buckets;
  u32 nblk = w == plan.windows - 1 ? rbt : rb;
  if (blockIdx.x >= nblk) return;
  const G1Xyzz* bk = buckets + (size_t)w * plan.nb + 1;  // bk[b - 1] = bucket id b
  u32 chunk = (B + nblk - 1) / nblk;
  u32 c_lo = min(blockIdx.x * chunk, B), c_hi = min(c_lo + chunk, B);
  u32 per = (chunk + REDUCE_T - 1) / REDUCE_T;
  u32 lo = min(c_lo + threadIdx.x * per, c_hi), hi = min(lo + per, c_hi);  // bucket ids lo+1 .. hi
  G1Xyzz run = G1Xyzz::identity(), acc = G1Xyzz::identity();
  for (u32 b = hi; b > lo; b--) {
    run = run.add(bk[b - 1]);
    acc = acc.add(run);  // acc = sum (id - lo) * B_id
  }
  // + lo * run  (run = plain sum of the chunk): double-and-add on the small integer lo < 2^c
  if (lo && !run.is_identity()) {
    G1Xyzz m = run;
    for (int bit = 30 - __clz(lo); bit >= 0; bit--) {
      m = m.dbl();
      if ((lo >> bit) & 1) m = m.add(run);
    }
    acc = acc.add(m);
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (u32 s = REDUCE_T / 2; s >= 1; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] = sm[threadIdx.x].add(sm[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[(size_t)w * rbt + blockIdx.x] = sm[0];
}

template <class Fq>
__device__ __forceinline__ Fq bcast_fe(const Fq& a, int src) {
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xffffffffu, a.v[i], src);
  return r;
}

// Jacobian doubling (dbl-2009-l, as JacT::dbl) by a whole warp that holds the SAME point on every lane: the independent products of
// each level run on lanes 0, 1, 2 and are broadcast -- a lone warp pays per instruction, not per lane, so three products cost one.
// Depth 3 products (X^2 | Y^2 | YZ ; B^2 | (X+B)^2 | E^2 ; E (D - X3)) instead of 7.  `l3` = lane % 3.
template <class Fq>
__device__ __forceinline__ JacT<Fq> dbl_lanes(const JacT<Fq>& p, u32 l3) {
  Fq a = l3 == 0 ? p.X : p.Y, b = l3 == 2 ? p.Z : a;
  Fq m = a * b;
  Fq A = bcast_fe(m, 0), B = bcast_fe(m, 1), YZ = bcast_fe(m, 2);
  Fq E = A.dbl() + A;
  Fq s = l3 == 0 ? B : (l3 == 1 ? p.X + B : E);
  Fq q = s.sqr();
  Fq C = bcast_fe(q, 0), t = bcast_fe(q, 1), EE = bcast_fe(q, 2);
  Fq D = (t - A - C).dbl();
  JacT<Fq> r;
  r.X = EE - D.dbl();
  r.Y = E * (D - r.X) - C.dbl().dbl().dbl();
  r.Z = YZ.dbl();  // an identity stays one (Z = 0)
  return r;
}

// warp w adds the rb partials of window w (lanes stride, shuffle butterfly); warp 0 then runs Horner over the windows
// (c lane-parallel doublings each, `dbl_lanes`) + to_affine
#define COMBINE_MAX_WINDOWS 32
#define COMBINE_T 256
template <class Fq>
__global__ void __launch_bounds__(COMBINE_T) k_msm_combine(MsmPlan plan, u32 rb, u32 rbt, const XyzzT<Fq>* partials, uint8_t* out) {
  typedef AffT<Fq> G1Affine;
  typedef XyzzT<Fq> G1Xyzz;
  __shared__ G1Xyzz wsum[COMBINE_MAX_WINDOWS];
  u32 lane = threadIdx.x & 31;
  for (u32 w = threadIdx.x >> 5; w < plan.windows; w += COMBINE_T / 32) {
    G1Xyzz sacc = G1Xyzz::identity();
    u32 nblk = w == plan.windows - 1 ? rbt : rb;
    for (u32 i = lane; i < nblk; i += 32) sacc = sacc.add(partials[(size_t)w * rbt + i]);
    if (rbt > 1) {
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) sacc = sacc.add(shfl_xor_xyzz(sacc, d));
    }
    if (lane == 0) wsum[w] = sacc;
  }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  // Horner in Jacobian coordinates (its doubling is 2M + 5S against 6M + 3S + a dot product for XYZZ); every lane of warp 0 carries
  // the same accumulator, the additions run redundantly
  const u32 l3 = lane % 3;
  JacT<Fq> acc = wsum[plan.windows - 1].to_jac();
  for (int ww = (int)plan.windows - 2; ww >= 0; ww--) {
    for (u32 k = 0; k < plan.c; k++) acc = dbl_lanes(acc, l3);
    acc = acc.add(wsum[ww].to_jac());
  }
  if (lane != 0) return;
  G1Affine a = acc.to_affine();
  Fq x = a.x.from_mont(), y = a.y.from_mont();
  uint4* o = reinterpret_cast<uint4*>(out);
  o[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  o[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
  o[2] = make_uint4(y.v[0], y.v[1], y.v[2], y.v[3]);
  o[3] = make_uint4(y.v[4], y.v[5], y.v[6], y.v[7]);
}

// d_out: 64 B affine canonical; d_status: int (0 ok, 1 bad point, 2 bad scalar)
template <class C>
static int msm_launch_t(svk_ctx* ctx, size_t n, const uint8_t* d_scalars, const uint8_t* d_points, uint8_t* d_out, int* d_status) {
  typedef AffT<typename C::Base> G1Affine;
  typedef XyzzT<typename C::Base> G1Xyzz;
  typedef typename C::Base F;
  if (n == 0) {
    SVK_CUDA(ctx, cudaMemsetAsync(d_out, 0, 64, ctx->stream));
    SVK_CUDA(ctx, cudaMemsetAsync(d_status, 0, 4, ctx->stream));
    return 0;
  }
  if (n >= (1ull << 30)) return svk_fail(ctx, "msm: n too large");
  const bool glv = CurveGlv<C>::value;
  MsmPlan plan = msm_plan(n, glv);
  if (plan.windows > COMBINE_MAX_WINDOWS) return svk_fail(ctx, "msm: window plan");
  cudaStream_t s = ctx->stream;
  size_t nv = glv ? 2 * n : n, nb = plan.nb, total = (size_t)plan.windows * nb;
  const u32 heavy_cap = 8192;  // chunks of HEAVY_CHUNK points
  G1Affine* pts_m;
  u32 *keys, *sorted, *hist;
  G1Xyzz* buckets;
  // <= 4 buckets per thread of the reduction while all its blocks are resident at once (one per SM); the top window gets blocks
  // in proportion to its buckets
  u32 ratio = plan.top_buckets > plan.buckets ? plan.top_buckets / plan.buckets : 1;
  u32 rb = (plan.buckets + REDUCE_T * 4 - 1) / (REDUCE_T * 4);
  u32 rb_cap = (u32)ctx->sm_count / (plan.windows - 1 + ratio);
  if (rb > rb_cap) rb = rb_cap ? rb_cap : 1;
  u32 rbt = rb * ratio;
  if (svk_scratch(ctx, 10, nv * sizeof(G1Affine), (void**)&pts_m)) return -1;
  if (svk_scratch(ctx, 11, (size_t)plan.windows * nv * 4, (void**)&keys)) return -1;
  if (svk_scratch(ctx, 12, (size_t)plan.windows * nv * 4, (void**)&sorted)) return -1;
  // hist | offs | cursor | order (total each) | bin_cnt, bin_cursor (64 + 65) | n_heavy | heavy_list
  size_t words = total * 4 + 2 * ORDER_BINS + 8 + heavy_cap * 4;
  if (svk_scratch(ctx, 13, words * 4, (void**)&hist)) return -1;
  if (svk_scratch(ctx, 14, (total + (size_t)plan.windows * rbt + heavy_cap) * sizeof(G1Xyzz), (void**)&buckets)) return -1;
  u32* offs = hist + total;
  u32* cursor = offs + total;
  u32* order = cursor + total;
  u32* bin_cnt = order + total;
  u32* bin_cursor = bin_cnt + ORDER_BINS;
  u32* n_heavy = bin_cursor + ORDER_BINS + 4;
  HeavyItem* heavy_list = reinterpret_cast<HeavyItem*>(n_heavy + 4);
  G1Xyzz* wsums = buckets + total;
  G1Xyzz* heavy_part = wsums + (size_t)plan.windows * rbt;
  SVK_CUDA(ctx, cudaMemsetAsync(hist, 0, total * 4, s));
  SVK_CUDA(ctx, cudaMemsetAsync(bin_cnt, 0, (2 * ORDER_BINS + 8 + (size_t)heavy_cap * 4) * 4, s));
  SVK_CUDA(ctx, cudaMemsetAsync(d_status, 0, 4, s));
  unsigned gb = (unsigned)((n + 255) / 256), gv = (unsigned)((nv + 255) / 256), gt = (unsigned)((total + 255) / 256);
  SVK_LAUNCH(ctx, "k_msm_prepare", k_msm_prepare<C><<<gb, 256, 0, s>>>(n, plan, d_scalars, d_points, pts_m, keys, hist, d_status));
  SVK_LAUNCH(ctx, "k_msm_scan", k_msm_scan<<<plan.windows, 1024, 0, s>>>(plan, hist, offs, cursor));
  SVK_LAUNCH(ctx, "k_msm_scatter", k_msm_scatter<<<dim3(gv, plan.windows), 256, 0, s>>>(nv, plan, keys, cursor, sorted));
  SVK_LAUNCH(ctx, "k_msm_order", k_msm_order_hist<<<gt, 256, 0, s>>>(total, hist, bin_cnt));
  SVK_LAUNCH(ctx, "k_msm_order", k_msm_order_scan<<<1, 32, 0, s>>>(bin_cnt, bin_cursor));
  SVK_LAUNCH(ctx, "k_msm_order", k_msm_order_scatter<<<gt, 256, 0, s>>>(total, hist, bin_cursor, order));
  SVK_LAUNCH(ctx, "k_msm_buckets", k_msm_buckets<F><<<(unsigned)((total + 127) / 128), 128, 0, s>>>(nv, plan, total, order, hist, offs, sorted, pts_m, buckets,
                                                                                                   heavy_list, n_heavy, heavy_cap));
  SVK_LAUNCH(ctx, "k_msm_heavy", k_msm_heavy<F><<<ctx->sm_count * 8, 32, 0, s>>>(nv, plan, heavy_list, n_heavy, heavy_cap, hist, offs, sorted, pts_m, heavy_part));
  SVK_LAUNCH(ctx, "k_msm_heavy", k_msm_heavy_sum<F><<<ctx->sm_count * 2, 32, 0, s>>>(heavy_list, n_heavy, heavy_cap, heavy_part, buckets));
  SVK_LAUNCH(ctx, "k_msm_reduce", k_msm_reduce<F><<<dim3(rbt, plan.windows), REDUCE_T, 0, s>>>(plan, rb, rbt, buckets, wsums));
  SVK_LAUNCH(ctx, "k_msm_combine", k_msm_combine<F><<<1, COMBINE_T, 0, s>>>(plan, rb, rbt, wsums, d_out));
  SVK_CUDA(ctx, cudaGetLastError());
  return 0;
}

int svk_msm_launch(svk_ctx* ctx, size_t n, const uint8_t* d_scalars, const uint8_t* d_points, uint8_t* d_out, int* d_status) {
  return msm_launch_t<CurveBn254>(ctx, n, d_scalars, d_points, d_out, d_status);
}

// curve: SVK_CURVE_BN254_G1 (0), SVK_CURVE_PALLAS (1), SVK_CURVE_VESTA (2)
int svk_msm_curve_launch(svk_ctx* ctx, int curve, size_t n, const uint8_t* d_scalars, const uint8_t* d_points, uint8_t* d_out, int* d_status) {
  switch (curve) {
    case 0: return msm_launch_t<CurveBn254>(ctx, n, d_scalars, d_points, d_out, d_status);
    case 1: return msm_launch_t<CurvePallas>(ctx, n, d_scalars, d_points, d_out, d_status);
    case 2: return msm_launch_t<CurveVesta>(ctx, n, d_scalars, d_points, d_out, d_status);
    default: return svk_fail(ctx, "msm: unknown curve id");
  }
}

// ---- batched scalar multiplication: out[i] = scalars[i] * points[i % n_points]  (`base * scalar`,
// loader/native.rs:67); used by the synthetic-workload generator and as a building block.
__device__ __noinline__ G1Jac mulb_window4(const G1Affine& p, const u32* k) {
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

__global__ void __launch_bounds__(128) k_g1_mul_batch(size_t n, const uint8_t* scalars, const uint8_t* points, size_t n_points, uint8_t* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine p;
  size_t pi = i % n_points;
  load32(p.x.v, points + pi * 64);
  load32(p.y.v, points + pi * 64 + 32);
  if (!p.is_identity()) { p.x = p.x.to_mont(); p.y = p.y.to_mont(); }
  u32 k[8];
  load32(k, scalars + i * 32);
  G1Affine a = mulb_window4(p, k).to_affine();
  Fq x = a.x.from_mont(), y = a.y.from_mont();
  uint4* o = reinterpret_cast<uint4*>(out + i * 64);
  o[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  o[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
  o[2] = make_uint4(y.v[0], y.v[1], y.v[2], y.v[3]);
  o[3] = make_uint4(y.v[4], y.v[5], y.v[6], y.v[7]);
}

int svk_g1_mul_batch_launch(svk_ctx* ctx, size_t n, const uint8_t* d_scalars, const uint8_t* d_points, size_t n_points, uint8_t* d_out) {
  if (n == 0) return 0;
  if (n_points == 0) return svk_fail(ctx, "g1_mul_batch: no points");
  SVK_LAUNCH(ctx, "k_g1_mul_batch", k_g1_mul_batch<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(n, d_scalars, d_points, n_points, d_out));
  SVK_CUDA(ctx, cudaGetLastError());
  return 0;
}
