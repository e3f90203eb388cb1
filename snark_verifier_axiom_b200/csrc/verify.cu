// K1/K2 kernels: batched `PlonkSuccinctVerifier::{read_proof, verify}`
// (snark-verifier/src/verifier/plonk.rs:32-93) for N proofs of ONE protocol.
//
//   k_decompress   : every G1 point of every proof: halo2curves `G1Affine::from_bytes` (Fq sqrt) +
//                    the [x mod r, y mod r] transcript encoding (transcript/halo2.rs:214-226, 247-260)
//   k_tape         : one proof per thread runs the compiled verifier tape (tape.cuh); k_tape_coop: the same tape with the
//                    Poseidon sponge spread over three warps per 32 proofs (poseidon_coop.cuh) for small launches
//   k_msm_var / k_msm_sum / k_to_affine : the final `lhs.evaluate(Some(g))` / `rhs.evaluate(Some(g))`
//                    (util/msm.rs:70-77 -> NativeLoader::multi_scalar_multiplication, loader/native.rs:61-71); two schedules
//                    (svk_protocol.h MsmSched): all variable-base terms of a proof on one thread (Straus, shared doublings),
//                    or one thread per term (GLV halves, 125 doublings) + 8 summing lanes when the launch is small
//   k_status       : per-proof `Result` -> status word (include/svk.h)
#include "compiler.h"
#include "poseidon_coop.cuh"
#include "straus.cuh"
#include "svk_ctx.h"
#include "svk_protocol.h"

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_decompress(size_t n_items, u32 n_points, const PointSched* sched, const uint8_t* proofs,
                                                    size_t proof_stride, const u32* proof_lens, u32* regs, G1Affine* pts, u32* err) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_items * n_points) return;
  size_t item = idx % n_items;
  u32 pi = (u32)(idx / n_items);
  PointSched s = sched[pi];
  u32 len = proof_lens ? min(proof_lens[item], (u32)proof_stride) : (u32)proof_stride;
  G1Affine pt = G1Affine::identity();
  Fr fx = Fr::zero(), fy = Fr::zero();
  if (s.byte_offset + 32 > len) {
    atomicMin(&err[item], (s.byte_offset << 8) | SVK_T_EOF);
  } else {
    u32 xc[8], yc[8];
    int rc = g1_decompress(proofs + item * proof_stride + s.byte_offset, pt, xc, yc);
    if (rc == 1) atomicMin(&err[item], (s.byte_offset << 8) | SVK_T_POINT_INVALID);
    else if (rc == 2) atomicMin(&err[item], (s.byte_offset << 8) | SVK_T_POINT_IDENTITY);
    else {
      fq_canon_to_fr_canon(fx.v, xc);
      fq_canon_to_fr_canon(fy.v, yc);
      fx = fx.to_mont();
      fy = fy.to_mont();
    }
  }
  RegFile rf{regs, n_items, item};
  rf.store(s.reg_x, fx);
  rf.store(s.reg_y, fy);
  pts[(size_t)pi * n_items + item] = pt;
}

// Uncompressed points of the Keccak `EvmTranscript` (transcript/evm.rs:223-242): x || y, 32 B big-endian each,
// both < p, on the curve; (0, 0) decodes to the identity, which `common_ec_point` rejects (evm.rs:184-196).
__global__ void __launch_bounds__(128) k_load_points_be(size_t n_items, u32 n_points, const PointSched* sched, const uint8_t* proofs,
                                                        size_t proof_stride, const u32* proof_lens, G1Affine* pts, u32* err) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_items * n_points) return;
  size_t item = idx % n_items;
  u32 pi = (u32)(idx / n_items);
  PointSched s = sched[pi];
  u32 len = proof_lens ? min(proof_lens[item], (u32)proof_stride) : (u32)proof_stride;
  G1Affine pt = G1Affine::identity();
  if (s.byte_offset + 64 > len) {
    atomicMin(&err[item], (s.byte_offset << 8) | SVK_T_EOF);
  } else {
    const uint8_t* p = proofs + item * proof_stride + s.byte_offset;
    Fq x, y;
    for (int i = 0; i < 8; i++) {
      const uint8_t* q = p + 4 * (7 - i);
      x.v[i] = ((u32)q[0] << 24) | ((u32)q[1] << 16) | ((u32)q[2] << 8) | (u32)q[3];
      q += 32;
      y.v[i] = ((u32)q[0] << 24) | ((u32)q[1] << 16) | ((u32)q[2] << 8) | (u32)q[3];
    }
    bool canon = Fq::is_canonical(x.v) && Fq::is_canonical(y.v);
    if (canon && x.is_zero() && y.is_zero()) {
      atomicMin(&err[item], (s.byte_offset << 8) | SVK_T_POINT_IDENTITY);
    } else {
      G1Affine c{x.to_mont(), y.to_mont()};
      if (!canon || !g1_on_curve(c)) atomicMin(&err[item], (s.byte_offset << 8) | SVK_T_POINT_INVALID);
      else pt = c;
    }
  }
  pts[(size_t)pi * n_items + item] = pt;
}

// ------------------------------------------------------------------------------------------------
template <bool KECCAK>
__global__ void __launch_bounds__(32) k_tape(size_t n_items, const TapeOp* ops, u32 n_ops, const uint16_t* aux, const Fr* consts,
                                             const PoseidonConsts* pk, u32* regs, const uint8_t* proofs, size_t proof_stride,
                                             const u32* proof_lens, const uint8_t* instances, u32 n_instances, u32* out_scalars,
                                             u32* out_challenges, u32 n_challenges, u32* err) {
  size_t item = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= n_items) return;
  RegFile rf{regs, n_items, item};
  TapeIo io;
  io.proof = proofs + item * proof_stride;
  io.proof_len = proof_lens ? min(proof_lens[item], (u32)proof_stride) : (u32)proof_stride;
  io.instances = instances + item * (size_t)n_instances * 32;
  io.n_instances = n_instances;
  io.out_scalars = out_scalars;
  io.out_challenges = out_challenges;
  io.n_challenge_slots = n_challenges;
  TranscriptState<KECCAK> st;
  if constexpr (KECCAK) keccak_reset(st.ks);
  else poseidon_init(st.ps, *pk);
  u32 e = SVK_NO_ERR;
  tape_exec<KECCAK>(ops, 0, n_ops, aux, consts, *pk, rf, io, st, e);
  if (e != SVK_NO_ERR) atomicMin(&err[item], e);
}

// Latency form of k_tape for the Poseidon transcript (poseidon_coop.cuh): 32 proofs per block of three warps; warp 0 runs the
// tape, warps 1 and 2 hold the other two sponge words during the permutations.  Same tape, same values; lanes past the end
// of the batch repeat the last proof (identical stores) so that every warp stays whole for the named barriers.
__device__ __forceinline__ void ts_permute(PoseidonCoopMain& st, const PoseidonConsts& pk, int n_in, const Fr& in0, const Fr& in1) { pcm_permute(st, pk, n_in, in0, in1); }
__device__ __forceinline__ Fr ts_squeeze(const PoseidonCoopMain& st) { return st.s1; }
__device__ __forceinline__ void ts_reset(PoseidonCoopMain& st, const PoseidonConsts& pk) { pcm_init(st, pk); }

__global__ void __launch_bounds__(PCOOP_THREADS) k_tape_coop(size_t n_items, const TapeOp* ops, u32 n_ops, const uint16_t* aux, const Fr* consts,
                                                             const PoseidonConsts* pk, u32* regs, const uint8_t* proofs, size_t proof_stride,
                                                             const u32* proof_lens, const uint8_t* instances, u32 n_instances, u32* out_scalars,
                                                             u32* out_challenges, u32 n_challenges, u32* err) {
  __shared__ PoseidonCoopShared sh;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp) {
    pc_helper(&sh, *pk, warp, lane);
    return;
  }
  size_t item = (size_t)blockIdx.x * 32 + lane;
  if (item >= n_items) item = n_items - 1;
  RegFile rf{regs, n_items, item};
  TapeIo io;
  io.proof = proofs + item * proof_stride;
  io.proof_len = proof_lens ? min(proof_lens[item], (u32)proof_stride) : (u32)proof_stride;
  io.instances = instances + item * (size_t)n_instances * 32;
  io.n_instances = n_instances;
  io.out_scalars = out_scalars;
  io.out_challenges = out_challenges;
  io.n_challenge_slots = n_challenges;
  PoseidonCoopMain st;
  st.sh = &sh;
  st.lane = lane;
  pcm_init(st, *pk);
  u32 e = SVK_NO_ERR;
  tape_exec<false>(ops, 0, n_ops, aux, consts, *pk, rf, io, st, e);
  pcm_exit(st);
  if (e != SVK_NO_ERR) atomicMin(&err[item], e);
}

// ------------------------------------------------------------------------------------------------
template <int MSM_LANES>
__device__ __forceinline__ G1Jac shfl_down_jac(const G1Jac& p, int delta) {
  G1Jac r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.X.v[i] = __shfl_down_sync(0xffffffffu, p.X.v[i], delta, MSM_LANES);
    r.Y.v[i] = __shfl_down_sync(0xffffffffu, p.Y.v[i], delta, MSM_LANES);
    r.Z.v[i] = __shfl_down_sync(0xffffffffu, p.Z.v[i], delta, MSM_LANES);
  }
  return r;
}

// s * P with 4-bit fixed windows; uniform control flow across lanes (table lookup instead of a branch).
__device__ __noinline__ G1Jac g1_mul_window4(const G1Affine& p, const u32* k) {
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

// Fixed-base window tables: tables[((b * W + w) << bits) + d] = d * 2^(bits w) * fixed_bases[b]  (d = 0: identity), W = 256 / bits.
// One (base, window, chunk of 256 digits) per thread; run once per compiled protocol and key (shared process-wide).
__global__ void __launch_bounds__(64) k_fixed_tables(u32 n_fixed, u32 bits, const G1Affine* fixed_bases, G1Affine* tables) {
  u32 W = 256 / bits, chunks = (1u << bits) / 256;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n_fixed * W * chunks) return;
  u32 ch = (u32)(t % chunks);
  u32 bw = (u32)(t / chunks);
  u32 b = bw / W, w = bw % W;
  G1Jac pw = G1Jac::from_affine(fixed_bases[b]);
  for (u32 k = 0; k < bits * w; k++) pw = pw.dbl();
  G1Affine base = pw.to_affine();
  u32 k0[8] = {ch * 256, 0, 0, 0, 0, 0, 0, 0};
  G1Jac acc = g1_scalar_mul(base, k0);
  G1Affine* out = tables + ((size_t)bw << bits) + (size_t)ch * 256;
  for (u32 d = 0; d < 256; d++) {
    out[d] = acc.to_affine();  // identity -> (0, 0)
    acc = acc.add_affine(base);
  }
}

// ---- Per-proof MSM, split by uniformity of work (ncu on the first version -- one fused kernel, 16 lanes per
// proof -- showed 10.5 of 32 lanes active per instruction: variable-base and fixed-base lanes serialised, and
// the final inversion ran on 2 lanes; profiles/r1_notes.md):
//   k_msm_var    `vpl` threads per proof; a thread owns up to SVK_VAR_TERMS_MAX variable-base terms of ONE proof and runs
//                them as one interleaved (Straus) multiplication (straus.cuh): per term a 16-entry table normalised to affine
//                together with all other tables of the thread, <= 52 signed 5-bit-window mixed additions; 255 shared
//                doublings and one inversion per thread.  Each thread has its own host-scheduled item list (lanes never mix the
//                lhs and rhs sides); fewer lanes minimise total work, more lanes shorten the latency.
//                Jacobian partial -> partials[(lane * n_items) + proof]
//   k_msm_sum    MSM_LANES threads per (proof, side) -- 1 in the throughput schedule, 8 + a shuffle tree in the latency schedule
//                (155 serial additions = 0.9 ms on a lone thread) --: the fixed-base table windows (mixed additions only,
//                entries prefetched), then the partials and the scalar == 1 bases -> sums[side][proof]
//   k_to_affine  one proof per thread: one inversion for both sides, canonical accumulator bytes
#define SVK_VAR_TERMS_MAX 16
#ifndef SVK_MSMVAR_MINBLOCKS
#define SVK_MSMVAR_MINBLOCKS 1
#endif
template <bool AFFINE>
__global__ void __launch_bounds__(64, SVK_MSMVAR_MINBLOCKS) k_msm_var(size_t n_items, const MsmWork* var_items, const u32* var_lane_off, u32 vpl, const G1Affine* pts,
                                                const u32* scalars, G1Jac* tables, Fq* prefix, G1Jac* partials) {
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_items * vpl) return;
  size_t it = gid % n_items;
  u32 lane = (u32)(gid / n_items);
  // this thread's table area: [term slot][16] entries, interleaved over threads for coalescing
  size_t n_threads = n_items * vpl;
  u32 k[SVK_VAR_TERMS_MAX][8];
  u32 nt = 0;
  if (!AFFINE && var_lane_off[lane + 1] - var_lane_off[lane] == 1) {
    // latency schedule, one term per thread: GLV halves the doubling chain (straus.cuh straus_run_glv1)
    MsmWork wk = var_items[var_lane_off[lane]];
    u32 kk[8], k1[4], k2[4], n1, n2;
    const uint4* sp = reinterpret_cast<const uint4*>(scalars + ((size_t)wk.slot * n_items + it) * 8);
    uint4 lo = sp[0], hi = sp[1];
    kk[0] = lo.x; kk[1] = lo.y; kk[2] = lo.z; kk[3] = lo.w; kk[4] = hi.x; kk[5] = hi.y; kk[6] = hi.z; kk[7] = hi.w;
    if (glv_decompose(kk, k1, n1, k2, n2)) {
      G1Affine base = pts[(size_t)wk.base * n_items + it];
      straus_build_table(tables + gid, n_threads, base);
      partials[gid] = straus_run_glv1(k1, n1, k2, n2, tables + gid, n_threads, glv_beta_mont());
      return;
    }
  }
  for (u32 vi = var_lane_off[lane]; vi < var_lane_off[lane + 1] && nt < SVK_VAR_TERMS_MAX; vi++, nt++) {
    MsmWork wk = var_items[vi];
    const uint4* sp = reinterpret_cast<const uint4*>(scalars + ((size_t)wk.slot * n_items + it) * 8);
    uint4 lo = sp[0], hi = sp[1];
    k[nt][0] = lo.x; k[nt][1] = lo.y; k[nt][2] = lo.z; k[nt][3] = lo.w; k[nt][4] = hi.x; k[nt][5] = hi.y; k[nt][6] = hi.z; k[nt][7] = hi.w;
    straus_recode(k[nt]);
    G1Affine base = pts[(size_t)wk.base * n_items + it];
    straus_build_table(tables + ((size_t)nt * STRAUS_TABLE) * n_threads + gid, n_threads, base);
  }
  if (AFFINE) straus_normalize(tables + gid, prefix + gid, n_threads, nt * STRAUS_TABLE);
  G1Jac acc = straus_run<AFFINE>(&k[0][0], nt, tables + gid, n_threads);
  partials[gid] = acc;
}

// work: per side, per lane a list of items of kind 1 (fixed-base window slice), 2 (add base), 3 (add partial #base)
template <int MSM_LANES>
__global__ void __launch_bounds__(128) k_msm_sum(size_t n_items, const MsmWork* work_lhs, const u32* off_lhs, const MsmWork* work_rhs,
                                                 const u32* off_rhs, const FixedSlot* fixed_lhs, u32 per_lhs, const FixedSlot* fixed_rhs,
                                                 u32 per_rhs, u32 fixed_bits, const G1Affine* fixed_bases, const G1Affine* tables,
                                                 const G1Affine* pts, const u32* scalars, const G1Jac* partials, G1Jac* sums) {
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t item = gid / MSM_LANES;
  u32 lane = (u32)(gid % MSM_LANES);
  bool active = item < n_items;
  size_t it = active ? item : n_items - 1;  // keep whole warps in the shuffles
  const MsmWork* work = blockIdx.y ? work_rhs : work_lhs;
  const u32* off = blockIdx.y ? off_rhs : off_lhs;
  G1Jac acc = G1Jac::identity();
  // fixed-base table additions: the same trip count on every lane
  const FixedSlot* fsched = (blockIdx.y ? fixed_rhs : fixed_lhs);
  u32 per = blockIdx.y ? per_rhs : per_lhs;
  // table entries are fetched one step ahead of their use (the address depends on the scalar only): the 16-bit tables are
  // HBM-resident (67 MB per base) and an L2 / DRAM round trip is shorter than the mixed addition that covers it
  const u32 FW = 256 / fixed_bits, dmask = (1u << fixed_bits) - 1;
  auto fetch = [&](u32 j, G1Affine& e) -> bool {
    FixedSlot fs = fsched[lane * per + j];
    if (fs.base < 0) return false;
    u32 bit = (u32)fs.w * fixed_bits;
    u32 d = (scalars[((size_t)fs.slot * n_items + it) * 8 + (bit >> 5)] >> (bit & 31)) & dmask;
    if (d == 0) return false;
    e = tables[(((size_t)fs.base * FW + fs.w) << fixed_bits) + d];
    return true;
  };
  G1Affine nxt = G1Affine::identity();
  bool have = per ? fetch(0, nxt) : false;
  for (u32 j = 0; j < per; j++) {
    G1Affine cur = nxt;
    bool use = have;
    if (j + 1 < per) have = fetch(j + 1, nxt);
    if (use) acc = acc.add_affine(cur);
  }
  for (u32 wi = off[lane]; wi < off[lane + 1]; wi++) {
    MsmWork wk = work[wi];
    if (wk.kind == 1) {
      u32 k[8];
      const uint4* sp = reinterpret_cast<const uint4*>(scalars + ((size_t)wk.slot * n_items + it) * 8);
      uint4 lo = sp[0], hi = sp[1];
      k[0] = lo.x; k[1] = lo.y; k[2] = lo.z; k[3] = lo.w; k[4] = hi.x; k[5] = hi.y; k[6] = hi.z; k[7] = hi.w;
      const G1Affine* tb = tables + (((size_t)wk.base * FW) << fixed_bits);
      for (int w = wk.w0; w < wk.w1; w++) {
        u32 bit = (u32)w * fixed_bits;
        u32 d = (k[bit >> 5] >> (bit & 31)) & dmask;
        acc = acc.add_affine(tb[((size_t)w << fixed_bits) + d]);
      }
    } else if (wk.kind == 2) {
      acc = acc.add_affine(wk.fixed ? fixed_bases[wk.base] : pts[(size_t)wk.base * n_items + it]);
    } else if (wk.kind == 3) {
      acc = acc.add(partials[(size_t)wk.base * n_items + it]);
    } else {  // kind 0: a full variable-base multiplication (the few rhs terms of GWC)
      u32 k[8];
      const uint4* sp = reinterpret_cast<const uint4*>(scalars + ((size_t)wk.slot * n_items + it) * 8);
      uint4 lo = sp[0], hi = sp[1];
      k[0] = lo.x; k[1] = lo.y; k[2] = lo.z; k[3] = lo.w; k[4] = hi.x; k[5] = hi.y; k[6] = hi.z; k[7] = hi.w;
      acc = acc.add(g1_mul_window4(pts[(size_t)wk.base * n_items + it], k));
    }
  }
  if (off[MSM_LANES] > 1 || per) {  // a side with a single item (SHPLONK rhs = W') has nothing to reduce
#pragma unroll
    for (int d = MSM_LANES / 2; d >= 1; d >>= 1) {
      G1Jac o = shfl_down_jac<MSM_LANES>(acc, d);
      if (lane < (u32)d) acc = acc.add(o);
    }
  }
  if (active && lane == 0) sums[(size_t)blockIdx.y * n_items + item] = acc;
}

// One proof per thread: both sides share ONE inversion (Montgomery's trick on Z_lhs * Z_rhs; an identity side takes Z := 1).
__global__ void __launch_bounds__(128) k_to_affine(size_t n_items, const G1Jac* sums, const u32* err, uint8_t* out_acc, size_t acc_stride) {
  size_t item = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= n_items) return;
  bool bad = err[item] != SVK_NO_ERR;
  G1Jac l = sums[item], r = sums[n_items + item];
  Fq zl = l.is_identity() ? Fq::one() : l.Z, zr = r.is_identity() ? Fq::one() : r.Z;
  Fq inv = (zl * zr).inv();
  G1Affine a[2] = {l.to_affine_with_zinv(inv * zr), r.to_affine_with_zinv(inv * zl)};
#pragma unroll
  for (int side = 0; side < 2; side++) {
    G1Affine p = bad ? G1Affine::identity() : a[side];
    Fq x = p.x.from_mont(), y = p.y.from_mont();
    uint4* o = reinterpret_cast<uint4*>(out_acc + item * acc_stride + (side ? 64 : 0));
    o[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    o[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
    o[2] = make_uint4(y.v[0], y.v[1], y.v[2], y.v[3]);
    o[3] = make_uint4(y.v[4], y.v[5], y.v[6], y.v[7]);
  }
}

// ------------------------------------------------------------------------------------------------
// Old accumulators carried in the instances: `LimbsEncoding<LIMBS, BITS>::from_repr` (pcs/kzg/accumulator.rs:57-77) over
// `protocol.accumulator_indices` (verifier/plonk/proof.rs:139-146).  Each coordinate is sum_k limb_k << (BITS k) as an
// integer (`fe_from_limbs`, util/arithmetic.rs:262-274); the reference PANICS when that integer does not fit 32 bytes or is
// >= p (`fe_from_big`, :237-243) or when (x, y) is not on the curve (`from_xy(..).unwrap()`, accumulator.rs:72-73; (0, 0) is
// the identity and passes).  Here those proofs get status SVK_ACCUMULATOR_PANIC.  One proof per thread; its old accumulators
// land behind its new one: out_acc[item][1 + a].  A proof that already failed gets zeros, like its new accumulator.
#define SVK_ERR_ACC_PANIC 0xfffffefeu
__global__ void __launch_bounds__(64) k_old_accumulators(size_t n_items, u32 n_old, u32 limbs, u32 bits, const u32* idx, const uint8_t* instances,
                                                        u32 n_instances, u32* err, uint8_t* out_acc, size_t acc_stride) {
  size_t item = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= n_items) return;
  bool bad = err[item] != SVK_NO_ERR;
  const uint8_t* inst = instances + item * (size_t)n_instances * 32;
  for (u32 a = 0; a < n_old && !bad; a++) {
    G1Affine pt[2];
    for (u32 c = 0; c < 4 && !bad; c++) {
      u32 acc[33];
      for (int i = 0; i < 33; i++) acc[i] = 0;
      for (u32 k = 0; k < limbs; k++) {
        u32 w[8];
        fe_load_le(w, inst + 32 * (size_t)idx[(a * 4 + c) * limbs + k]);
        if (!Fr::is_canonical(w)) bad = true;  // not an `Fr`: the tape reports it as InvalidInstances first (lower error word)
        u32 sh = bits * k, wo = sh >> 5, bo = sh & 31;
        u64 carry = 0;
        for (int i = 0; i < 9; i++) {
          u32 lo = i < 8 ? w[i] : 0, prev = i > 0 ? w[i - 1] : 0;
          u32 piece = bo ? ((lo << bo) | (prev >> (32 - bo))) : lo;
          u64 t = (u64)acc[wo + i] + piece + carry;
          acc[wo + i] = (u32)t;
          carry = t >> 32;
        }
        for (u32 i = wo + 9; carry && i < 33; i++) {
          u64 t = (u64)acc[i] + carry;
          acc[i] = (u32)t;
          carry = t >> 32;
        }
      }
      for (int i = 8; i < 33; i++) bad = bad || acc[i] != 0;
      Fq v;
      for (int i = 0; i < 8; i++) v.v[i] = acc[i];
      bad = bad || !Fq::is_canonical(v.v);
      if (c & 1) pt[c >> 1].y = v; else pt[c >> 1].x = v;
    }
    for (int sde = 0; sde < 2 && !bad; sde++) {
      G1Affine m = pt[sde];
      if (!(m.x.is_zero() && m.y.is_zero())) {
        m.x = m.x.to_mont();
        m.y = m.y.to_mont();
        bad = !g1_on_curve(m);
      }
    }
    if (!bad) {
      uint4* o = reinterpret_cast<uint4*>(out_acc + item * acc_stride + 128 * (size_t)(1 + a));
      for (int sde = 0; sde < 2; sde++) {
        const Fq &x = pt[sde].x, &y = pt[sde].y;
        o[4 * sde + 0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
        o[4 * sde + 1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
        o[4 * sde + 2] = make_uint4(y.v[0], y.v[1], y.v[2], y.v[3]);
        o[4 * sde + 3] = make_uint4(y.v[4], y.v[5], y.v[6], y.v[7]);
      }
    }
  }
  if (bad) {
    if (err[item] == SVK_NO_ERR) err[item] = SVK_ERR_ACC_PANIC;
    uint4* o = reinterpret_cast<uint4*>(out_acc + item * acc_stride + 128);
    for (u32 i = 0; i < n_old * 8; i++) o[i] = make_uint4(0, 0, 0, 0);
  }
}

int svk_fixed_tables_launch(svk_ctx* ctx, ProtocolDevice* pd) {
  u32 n_fixed = pd->n_pre + 1;
  size_t total = (size_t)n_fixed * (256 / pd->fixed_bits) * ((1u << pd->fixed_bits) / 256);
  SVK_LAUNCH(ctx, "k_fixed_tables",
             k_fixed_tables<<<(unsigned)((total + 63) / 64), 64, 0, ctx->stream>>>(n_fixed, pd->fixed_bits, pd->d_fixed, pd->d_fixed_tables));
  SVK_CUDA(ctx, cudaGetLastError());
  SVK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// `Poseidon::update(inputs); squeeze()` on a fresh sponge (util/hash/poseidon.rs:448-467), one sponge per thread: the buffer is
// absorbed RATE elements at a time; a length that is a multiple of RATE gets one more permutation of the empty chunk.
// mode 1: the same sponges through the warp-cooperative schedule (poseidon_coop.cuh), 32 per block of three warps.
__global__ void __launch_bounds__(PCOOP_THREADS) k_poseidon_squeeze(size_t n, u32 n_in, const uint8_t* inputs, const PoseidonConsts* pk, uint8_t* out,
                                                                    int coop, int* bad) {
  __shared__ PoseidonCoopShared sh;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp) {
    if (coop) pc_helper(&sh, *pk, warp, lane);
    return;
  }
  size_t i = (size_t)blockIdx.x * 32 + lane;
  if (i >= n) i = n - 1;
  PoseidonState st;
  PoseidonCoopMain cm;
  cm.sh = &sh;
  cm.lane = lane;
  if (coop) pcm_init(cm, *pk); else poseidon_init(st, *pk);
  u32 n_perm = n_in / 2 + 1;  // chunks of RATE = 2, plus the empty chunk when the length is exact
  for (u32 q = 0; q < n_perm; q++) {
    Fr a = Fr::zero(), b = Fr::zero();
    int k = (int)min(2u, n_in - min(n_in, 2 * q));
    if (k >= 1) { fe_load_le(a.v, inputs + (i * n_in + 2 * q) * 32); if (!Fr::is_canonical(a.v)) *bad = 1; a = a.to_mont(); }
    if (k >= 2) { fe_load_le(b.v, inputs + (i * n_in + 2 * q + 1) * 32); if (!Fr::is_canonical(b.v)) *bad = 1; b = b.to_mont(); }
    if (coop) pcm_permute(cm, *pk, k, a, b); else poseidon_permute(st, *pk, k, a, b);
  }
  if (coop) pcm_exit(cm);
  Fr r = (coop ? cm.s1 : st.s[1]).from_mont();
  fe_store_le(out + i * 32, r.v);
}

int svk_poseidon_squeeze_launch(svk_ctx* ctx, size_t n, u32 n_in, const uint8_t* d_inputs, uint8_t* d_out, int coop, int* d_bad) {
  SVK_LAUNCH(ctx, "k_poseidon_squeeze",
             k_poseidon_squeeze<<<(unsigned)((n + 31) / 32), PCOOP_THREADS, 0, ctx->stream>>>(n, n_in, d_inputs, ctx->d_poseidon, d_out, coop, d_bad));
  SVK_CUDA(ctx, cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// mode: 0 normal, 1 InvalidInstances for all, 2 InvalidProtocol unless the read failed
__global__ void k_status(size_t n_items, const u32* err, int mode, int32_t* status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_items) return;
  u32 e = err[i];
  int32_t s = SVK_OK;
  if (mode == 1) s = SVK_INVALID_INSTANCES;
  else if (e == SVK_ERR_ACC_PANIC) s = SVK_ACCUMULATOR_PANIC;
  else if (e != SVK_NO_ERR) s = ((e & 0xff) == 0) ? SVK_INVALID_INSTANCES : (SVK_TRANSCRIPT | (int32_t)((e & 0xff) << 8));
  else if (mode == 2) s = SVK_INVALID_PROTOCOL;
  status[i] = s;
}

__global__ void k_fill_u32(size_t n, u32* p, u32 v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ------------------------------------------------------------------------------------------------
int svk_succinct_verify_launch(svk_ctx* ctx, ProtocolDevice* pd, size_t n, const uint8_t* d_instances, u32 n_instances_given,
                               const uint8_t* d_proofs, size_t proof_stride, const u32* d_proof_lens, uint8_t* d_out_acc,
                               u32* d_out_challenges, int32_t* d_out_status) {
  if (n == 0) return 0;
  cudaStream_t s = ctx->stream;
  const size_t acc_stride = 128 * (size_t)pd->accs_per_proof();
  u32 *d_err, *d_regs, *d_scalars, *d_chal_scratch;
  G1Affine* d_pts;
  size_t n_pts = pd->points.size();
  if (svk_scratch(ctx, 2, n * 4, (void**)&d_err)) return -1;
  if (svk_scratch(ctx, 3, (size_t)pd->n_regs * n * 32 + 32, (void**)&d_regs)) return -1;
  if (svk_scratch(ctx, 4, n_pts * n * sizeof(G1Affine) + 64, (void**)&d_pts)) return -1;
  if (svk_scratch(ctx, 5, (size_t)pd->n_scalar_slots * n * 32 + 32, (void**)&d_scalars)) return -1;
  if (!d_out_challenges) {
    if (svk_scratch(ctx, 1, (size_t)pd->n_challenges * n * 32 + 32, (void**)&d_chal_scratch)) return -1;
    d_out_challenges = d_chal_scratch;
  }
  unsigned b = 256;
  SVK_LAUNCH(ctx, "k_fill_u32", k_fill_u32<<<(unsigned)((n + b - 1) / b), b, 0, s>>>(n, d_err, SVK_NO_ERR));
  int mode = 0;
  if (n_instances_given != pd->n_instances) mode = 1;  // proof.rs:66-69
  else if (!pd->verify_valid) mode = 2;
  if (mode != 1) {
    if (n_pts) {
      size_t total = n * n_pts;
      if (pd->transcript_kind == 1)
        SVK_LAUNCH(ctx, "k_load_points_be",
                   k_load_points_be<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(n, (u32)n_pts, pd->d_sched, d_proofs, proof_stride,
                                                                                   d_proof_lens, d_pts, d_err));
      else
        SVK_LAUNCH(ctx, "k_decompress",
                   k_decompress<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(n, (u32)n_pts, pd->d_sched, d_proofs, proof_stride, d_proof_lens,
                                                                               d_regs, d_pts, d_err));
    }
    u32 n_ops = pd->verify_valid ? pd->n_ops : pd->read_ops_end;
    if (pd->transcript_kind == 1)
      SVK_LAUNCH(ctx, "k_tape_keccak",
                 k_tape<true><<<(unsigned)((n + 31) / 32), 32, 0, s>>>(n, pd->d_ops, n_ops, pd->d_aux, pd->d_consts, ctx->d_poseidon, d_regs,
                                                                      d_proofs, proof_stride, d_proof_lens, d_instances, pd->n_instances,
                                                                      d_scalars, d_out_challenges, pd->n_challenges, d_err));
    else if (n <= ctx->tape_coop_max)
      SVK_LAUNCH(ctx, "k_tape_coop",
                 k_tape_coop<<<(unsigned)((n + 31) / 32), PCOOP_THREADS, 0, s>>>(n, pd->d_ops, n_ops, pd->d_aux, pd->d_consts, ctx->d_poseidon, d_regs,
                                                                                d_proofs, proof_stride, d_proof_lens, d_instances, pd->n_instances,
                                                                                d_scalars, d_out_challenges, pd->n_challenges, d_err));
    else
      SVK_LAUNCH(ctx, "k_tape",
                 k_tape<false><<<(unsigned)((n + 31) / 32), 32, 0, s>>>(n, pd->d_ops, n_ops, pd->d_aux, pd->d_consts, ctx->d_poseidon, d_regs,
                                                                       d_proofs, proof_stride, d_proof_lens, d_instances, pd->n_instances,
                                                                       d_scalars, d_out_challenges, pd->n_challenges, d_err));
  }
  // `PlonkProof::read` ends with the old accumulators (proof.rs:139-146): their panic precedes anything `verify` reports
  if (pd->n_old && mode != 1)
    SVK_LAUNCH(ctx, "k_old_accumulators",
               k_old_accumulators<<<(unsigned)((n + 63) / 64), 64, 0, s>>>(n, pd->n_old, pd->acc_limbs, pd->acc_bits, pd->d_old_idx, d_instances,
                                                                          pd->n_instances, d_err, d_out_acc, acc_stride));
  if (mode == 0) {
    G1Jac *d_partials, *d_sums, *d_tables;
    // latency schedule (one k_msm_var thread per term) while its threads still fit the machine a few times over
    const MsmSched& sc = pd->sched[(n * pd->sched[1].var_lanes_total <= ctx->msm_latency_threads_max) ? 1 : 0];
    u32 vpl = sc.var_lanes_total;
    u32 terms_per_thread = sc.var_terms_per_thread;
    if (svk_scratch(ctx, 7, (size_t)std::max<u32>(vpl, 1) * n * sizeof(G1Jac), (void**)&d_partials)) return -1;
    if (svk_scratch(ctx, 15, 2 * n * sizeof(G1Jac), (void**)&d_sums)) return -1;
    if (pd->n_var) {
      size_t total = n * vpl;
      Fq* d_prefix;
      if (svk_scratch(ctx, 6, (size_t)terms_per_thread * 16 * total * sizeof(G1Jac), (void**)&d_tables)) return -1;
      // threads that own >= 3 terms normalise their tables to affine (one inversion per thread) and use mixed additions
      if (terms_per_thread >= 3) {
        if (svk_scratch(ctx, 21, (size_t)terms_per_thread * 16 * total * sizeof(Fq), (void**)&d_prefix)) return -1;
        SVK_LAUNCH(ctx, "k_msm_var",
                   k_msm_var<true><<<(unsigned)((total + 63) / 64), 64, 0, s>>>(n, sc.d_var_items, sc.d_var_lane_off, vpl, d_pts, d_scalars, d_tables, d_prefix, d_partials));
      } else {
        SVK_LAUNCH(ctx, "k_msm_var",
                   k_msm_var<false><<<(unsigned)((total + 63) / 64), 64, 0, s>>>(n, sc.d_var_items, sc.d_var_lane_off, vpl, d_pts, d_scalars, d_tables, nullptr, d_partials));
      }
    }
    dim3 grid((unsigned)((n * sc.msm_lanes + 127) / 128), 2);
#define SVK_MSM_SUM_ARGS                                                                                                                \
  n, sc.d_work_lhs, sc.d_lane_off_lhs, sc.d_work_rhs, sc.d_lane_off_rhs, sc.d_fixed_lhs, sc.fixed_per_lhs, sc.d_fixed_rhs, sc.fixed_per_rhs, \
      pd->fixed_bits, pd->d_fixed, pd->d_fixed_tables, d_pts, d_scalars, d_partials, d_sums
    if (sc.msm_lanes == 1)
      SVK_LAUNCH(ctx, "k_msm_sum", k_msm_sum<1><<<grid, 128, 0, s>>>(SVK_MSM_SUM_ARGS));
    else if (sc.msm_lanes == 2)
      SVK_LAUNCH(ctx, "k_msm_sum", k_msm_sum<2><<<grid, 128, 0, s>>>(SVK_MSM_SUM_ARGS));
    else if (sc.msm_lanes == 4)
      SVK_LAUNCH(ctx, "k_msm_sum", k_msm_sum<4><<<grid, 128, 0, s>>>(SVK_MSM_SUM_ARGS));
    else
      SVK_LAUNCH(ctx, "k_msm_sum",
                 k_msm_sum<SVK_MSM_LANES_LATENCY><<<grid, 128, 0, s>>>(n, sc.d_work_lhs, sc.d_lane_off_lhs, sc.d_work_rhs, sc.d_lane_off_rhs,
                                                                       sc.d_fixed_lhs, sc.fixed_per_lhs, sc.d_fixed_rhs, sc.fixed_per_rhs,
                                                                       pd->fixed_bits, pd->d_fixed, pd->d_fixed_tables, d_pts, d_scalars,
                                                                       d_partials, d_sums));
    SVK_LAUNCH(ctx, "k_to_affine", k_to_affine<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(n, d_sums, d_err, d_out_acc, acc_stride));
  } else {
    SVK_CUDA(ctx, cudaMemsetAsync(d_out_acc, 0, n * acc_stride, s));
  }
  SVK_LAUNCH(ctx, "k_status", k_status<<<(unsigned)((n + b - 1) / b), b, 0, s>>>(n, d_err, mode, d_out_status));
  SVK_CUDA(ctx, cudaGetLastError());
  return 0;
}
