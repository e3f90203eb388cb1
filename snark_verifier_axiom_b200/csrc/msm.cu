// K4: Pippenger bucket MSM over BN254 G1 + batched scalar multiplication.
//
// Replaces `util::msm::multi_scalar_multiplication` (snark-verifier/src/util/msm.rs:238-317; window
// ceil(ln n)+2, 2^w-1 buckets, running-sum reduction, rayon chunking) for BASELINE config 3, and is the
// large-n engine for `KzgAs::verify`'s fold MSM (pcs/kzg/accumulation.rs:51-59).  Results are canonical
// affine points, so any correct schedule is bit-exact (SURVEY finding 1).
//
// Pipeline (all on the context stream, no host round trips):
//   k_msm_prepare   one point per thread: Montgomery copy of the point, signed c-bit digits of the
//                   scalar for every window, per-(window,bucket) histogram
//   k_msm_scan      one block per window: exclusive scan of the histogram -> bucket offsets
//   k_msm_scatter   counting sort of (point index | sign) by bucket id, per window
//   k_msm_buckets   one bucket per thread: XYZZ mixed additions (8M+2S) over its sorted run
//   k_msm_reduce    up to 8 blocks per window: sum_b b*B_b over a bucket range by chunked running sums + shared-memory tree
//   k_msm_combine   one warp: per-window sum of the block partials, then Horner over the windows (c doublings each) + to_affine
#include "pasta.cuh"
#include "svk_ctx.h"

struct MsmPlan {
  u32 c;        // window bits
  u32 windows;  // number of windows
  u32 buckets;  // 2^(c-1) (bucket ids 1..buckets; 0 = digit zero)
};

// Window choice.  254-bit scalars: the TOP window only holds 254 - c*(W-1) significant bits, and a
// top window with few significant bits has few populated buckets, each with n / 2^bits points -- a serial
// tail for the one-bucket-per-thread accumulation (measured: c = 13 at n = 2^16 spent 6.4 ms there).
// c = 16 (W = 16, 14 top bits) and c = 15 (W = 17, 14 top bits) are the well-filled choices; small inputs
// use c = 8 (W = 32, 6 top bits) where the bucket reduction would otherwise dominate.
static MsmPlan msm_plan(size_t n) {
  u32 c = n >= (1u << 18) ? 16 : (n >= (1u << 13) ? 15 : 8);
  MsmPlan p;
  p.c = c;
  p.windows = (255 + c - 1) / c;  // 254-bit scalars + one carry bit
  p.buckets = 1u << (c - 1);
  return p;
}

__device__ __forceinline__ void load32(u32* v, const uint8_t* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 lo = q[0], hi = q[1];
  v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
}

// keys[w * n + i] = bucket id (0 = skip) | sign << 31
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
  u32 k[9];
  load32(k, scalars + i * 32);
  k[8] = 0;
  if (!Fr::is_canonical(k)) { *bad = 2; ident = true; }
  u32 carry = 0;
  for (u32 w = 0; w < plan.windows; w++) {
    u32 bit = w * plan.c;
    u32 word = bit >> 5, sh = bit & 31;
    u64 two = (u64)k[word] | ((u64)(word + 1 <= 8 ? k[word + 1] : 0) << 32);
    u32 d = (u32)((two >> sh) & ((1u << plan.c) - 1)) + carry;
    u32 neg = 0;
    carry = 0;
    if (d > plan.buckets) { d = (1u << plan.c) - d; neg = 1; carry = 1; }
    if (ident) d = 0;
    keys[(size_t)w * n + i] = d | (neg << 31);
    if (d) atomicAdd(&hist[(size_t)w * (plan.buckets + 1) + d], 1u);
  }
}

// exclusive scan of hist[w][0..buckets] -> offs[w][..]; one block (1024 threads) per window
__global__ void __launch_bounds__(1024) k_msm_scan(MsmPlan plan, const u32* hist, u32* offs, u32* cursor) {
  __shared__ u32 part[1024];
  u32 w = blockIdx.x, nb = plan.buckets + 1;
  const u32* h = hist + (size_t)w * nb;
  u32* o = offs + (size_t)w * nb;
  u32* cu = cursor + (size_t)w * nb;
  u32 per = (nb + blockDim.x - 1) / blockDim.x;
  u32 lo = threadIdx.x * per, hi = min(lo + per, nb);
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

__global__ void __launch_bounds__(256) k_msm_scatter(size_t n, MsmPlan plan, const u32* keys, u32* cursor, u32* sorted) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  u32 w = blockIdx.y;
  if (i >= n) return;
  u32 key = keys[(size_t)w * n + i];
  u32 d = key & 0x7fffffffu;
  if (!d) return;
  u32 pos = atomicAdd(&cursor[(size_t)w * (plan.buckets + 1) + d], 1u);
  sorted[(size_t)w * n + pos] = (u32)i | (key & 0x80000000u);
}

// one bucket per thread; buckets[w][b-1] (XYZZ)
template <class Fq>
__global__ void __launch_bounds__(128) k_msm_buckets(size_t n, MsmPlan plan, const u32* hist, const u32* offs, const u32* sorted,
                                                     const AffT<Fq>* pts_m, XyzzT<Fq>* buckets) {
  typedef AffT<Fq> G1Affine;
  typedef XyzzT<Fq> G1Xyzz;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)plan.windows * plan.buckets;
  if (t >= total) return;
  u32 w = (u32)(t / plan.buckets), b = (u32)(t % plan.buckets) + 1;
  size_t hidx = (size_t)w * (plan.buckets + 1) + b;
  u32 cnt = hist[hidx], off = offs[hidx];
  const u32* run = sorted + (size_t)w * n + off;
  G1Xyzz acc = G1Xyzz::identity();
  for (u32 j = 0; j < cnt; j++) {
    u32 e = run[j];
    G1Affine p = pts_m[e & 0x7fffffffu];
    if (e >> 31) p.y = p.y.neg();
    acc = acc.add_affine(p);
  }
  buckets[t] = acc;
}

// S_w = sum_{b=1..B} b * bucket[w][b-1], in two stages: `rb` blocks of REDUCE_T threads per window each reduce a contiguous range of
// buckets to one partial (chunked running sums per thread, + lo * (plain sum) for the chunk's offset, shared-memory tree), then
// k_msm_combine adds the partials of a window.  (The first version used ONE block per window: 128 buckets per thread at c = 16 made
// this serial tail 2.8 ms -- a third of a 2^20-point MSM.)
#define REDUCE_T 256
template <class Fq>
__global__ void __launch_bounds__(REDUCE_T) k_msm_reduce(MsmPlan plan, u32 rb, const XyzzT<Fq>* buckets, XyzzT<Fq>* partials) {
  typedef XyzzT<Fq> G1Xyzz;
  __shared__ G1Xyzz sm[REDUCE_T];
  u32 w = blockIdx.y, B = plan.buckets;
  const G1Xyzz* bk = buckets + (size_t)w * B;
  u32 chunk = (B + rb - 1) / rb;
  u32 c_lo = blockIdx.x * chunk, c_hi = min(c_lo + chunk, B);
  u32 per = (chunk + REDUCE_T - 1) / REDUCE_T;
  u32 lo = min(c_lo + threadIdx.x * per, c_hi), hi = min(lo + per, c_hi);  // bucket ids lo+1 .. hi
  G1Xyzz run = G1Xyzz::identity(), acc = G1Xyzz::identity();
  for (u32 b = hi; b > lo; b--) {
    run = run.add(bk[b - 1]);
    acc = acc.add(run);  // acc = sum (id - lo) * B_id
  }
  // + lo * run  (run = plain sum of the chunk): double-and-add on the small integer lo < 2^(c-1)
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
  if (threadIdx.x == 0) partials[(size_t)w * rb + blockIdx.x] = sm[0];
}

// one warp: lane w adds the rb partials of window w; lane 0 then runs Horner over the windows (c doublings each) + to_affine
template <class Fq>
__global__ void __launch_bounds__(32) k_msm_combine(MsmPlan plan, u32 rb, const XyzzT<Fq>* partials, uint8_t* out) {
  typedef AffT<Fq> G1Affine;
  typedef XyzzT<Fq> G1Xyzz;
  __shared__ G1Xyzz wsum[32];
  u32 lane = threadIdx.x;
  if (lane < plan.windows) {
    G1Xyzz sacc = partials[(size_t)lane * rb];
    for (u32 i = 1; i < rb; i++) sacc = sacc.add(partials[(size_t)lane * rb + i]);
    wsum[lane] = sacc;
  }
  __syncthreads();
  if (lane != 0) return;
  G1Xyzz acc = G1Xyzz::identity();
  for (int w = (int)plan.windows - 1; w >= 0; w--) {
    for (u32 k = 0; k < plan.c; k++) acc = acc.dbl();
    acc = acc.add(wsum[w]);
  }
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
  if (n == 0) {
    SVK_CUDA(ctx, cudaMemsetAsync(d_out, 0, 64, ctx->stream));
    SVK_CUDA(ctx, cudaMemsetAsync(d_status, 0, 4, ctx->stream));
    return 0;
  }
  if (n >= (1ull << 31)) return svk_fail(ctx, "msm: n too large");
  MsmPlan plan = msm_plan(n);
  cudaStream_t s = ctx->stream;
  size_t nb = plan.buckets + 1;
  G1Affine* pts_m;
  u32 *keys, *sorted, *hist;
  G1Xyzz* buckets;
  if (svk_scratch(ctx, 10, n * sizeof(G1Affine), (void**)&pts_m)) return -1;
  if (svk_scratch(ctx, 11, (size_t)plan.windows * n * 4, (void**)&keys)) return -1;
  if (svk_scratch(ctx, 12, (size_t)plan.windows * n * 4, (void**)&sorted)) return -1;
  if (svk_scratch(ctx, 13, (size_t)plan.windows * nb * 4 * 3, (void**)&hist)) return -1;
  u32 rb = plan.buckets >= 4096 ? 8 : (plan.buckets >= 1024 ? 4 : 1);  // reduce blocks per window: <= 16 buckets per thread
  if (svk_scratch(ctx, 14, ((size_t)plan.windows * plan.buckets + (size_t)plan.windows * rb) * sizeof(G1Xyzz), (void**)&buckets)) return -1;
  u32* offs = hist + (size_t)plan.windows * nb;
  u32* cursor = offs + (size_t)plan.windows * nb;
  G1Xyzz* wsums = buckets + (size_t)plan.windows * plan.buckets;
  SVK_CUDA(ctx, cudaMemsetAsync(hist, 0, (size_t)plan.windows * nb * 4, s));
  SVK_CUDA(ctx, cudaMemsetAsync(d_status, 0, 4, s));
  unsigned gb = (unsigned)((n + 255) / 256);
  SVK_LAUNCH(ctx, "k_msm_prepare", k_msm_prepare<C><<<gb, 256, 0, s>>>(n, plan, d_scalars, d_points, pts_m, keys, hist, d_status));
  SVK_LAUNCH(ctx, "k_msm_scan", k_msm_scan<<<plan.windows, 1024, 0, s>>>(plan, hist, offs, cursor));
  SVK_LAUNCH(ctx, "k_msm_scatter", k_msm_scatter<<<dim3(gb, plan.windows), 256, 0, s>>>(n, plan, keys, cursor, sorted));
  size_t total = (size_t)plan.windows * plan.buckets;
  SVK_LAUNCH(ctx, "k_msm_buckets", k_msm_buckets<typename C::Base><<<(unsigned)((total + 127) / 128), 128, 0, s>>>(n, plan, hist, offs, sorted, pts_m, buckets));
  SVK_LAUNCH(ctx, "k_msm_reduce", k_msm_reduce<typename C::Base><<<dim3(rb, plan.windows), REDUCE_T, 0, s>>>(plan, rb, buckets, wsums));
  SVK_LAUNCH(ctx, "k_msm_combine", k_msm_combine<typename C::Base><<<1, 32, 0, s>>>(plan, rb, wsums, d_out));
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
