/* libsvk -- C ABI of the B200-native batch verifier for the NativeLoader KZG/PLONK path of
 * snark-verifier.  Plain pointers and sizes only.  Every entry point cites the reference
 * interface it replaces (paths relative to the reference root).
 *
 * Data layout at the ABI: field elements are 32-byte little-endian CANONICAL values
 * (= halo2curves `to_repr()`), never Montgomery.  Batches are arrays of structs.
 * Return value: 0 on success, < 0 for CUDA / argument faults (see svk_last_error).  Verification
 * outcomes are reported per item in status / ok arrays, never through the return code, and the
 * process is never aborted.
 *
 * `*_dev` variants take DEVICE pointers (inputs already resident in HBM) and enqueue on the
 * context's stream without synchronising; the plain variants take HOST pointers, copy in, run,
 * copy out and synchronise.
 */
#ifndef SVK_H
#define SVK_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct svk_ctx svk_ctx;

typedef struct { uint8_t b[32]; } svk_fe;                 /* Fr or Fq, LE canonical */
typedef struct { svk_fe x, y; } svk_g1;                   /* G1Affine; identity = all zero */
typedef struct { svk_fe x_c0, x_c1, y_c0, y_c1; } svk_g2; /* G2Affine over Fq2 = c0 + c1 u */
typedef struct { svk_g1 lhs, rhs; } svk_acc;              /* KzgAccumulator  (pcs/kzg/accumulator.rs:6-26) */
typedef struct { svk_g1 g1; svk_g2 g2; svk_g2 s_g2; } svk_deciding_key; /* KzgDecidingKey (pcs/kzg/decider.rs:6-36) */

/* per-item status = `Error` of snark-verifier/src/lib.rs:21-30 */
enum {
  SVK_OK = 0,
  SVK_INVALID_INSTANCES = 1, /* Error::InvalidInstances   verifier/plonk/proof.rs:66-69 */
  SVK_INVALID_PROTOCOL = 2,  /* Error::InvalidProtocol    verifier/plonk/proof.rs:216,223,232,273 */
  SVK_ASSERTION_FAILURE = 3, /* Error::AssertionFailure   pcs/kzg/decider.rs:67 */
  SVK_TRANSCRIPT = 4,        /* Error::Transcript         system/halo2/transcript/halo2.rs:214-260 */
  SVK_ACCUMULATOR_PANIC = 5  /* the reference PANICS here: an old accumulator's limbs do not decode to curve points
                                (`fe_from_big` / `from_xy(..).unwrap()`, util/arithmetic.rs:237-243, pcs/kzg/accumulator.rs:72-73);
                                a Rust shim re-raises it as a panic */
};
/* sub-codes stored in bits 8.. of a status word when the low byte is SVK_TRANSCRIPT */
enum { SVK_T_EOF = 1, SVK_T_SCALAR_RANGE = 2, SVK_T_POINT_INVALID = 3, SVK_T_POINT_IDENTITY = 4 };

/* Fiat-Shamir transcript of the proofs.  POSEIDON: `PoseidonTranscript` (system/halo2/transcript/halo2.rs:163-304; 32 B LE scalars,
 * 32 B compressed points).  EVM: Keccak-256 `EvmTranscript` (system/halo2/transcript/evm.rs:152-243; 32 B BIG-endian scalars,
 * 64 B uncompressed big-endian points) -- the format of `gen_evm_proof_*` / examples/evm-verifier.rs proofs. */
enum { SVK_TRANSCRIPT_POSEIDON = 0, SVK_TRANSCRIPT_EVM = 1 };

enum { SVK_MOS_BDFG21 = 0 /* SHPLONK, pcs/kzg/multiopen/bdfg21.rs */, SVK_MOS_GWC19 = 1 /* pcs/kzg/multiopen/gwc19.rs */ };

/* ---- context ------------------------------------------------------------------------------- */
/* One context = one device + one stream.  Calls on one context are serialised by an internal (recursive) mutex; different
 * contexts may be used concurrently from different threads.
 * Fails (returns < 0, *out = NULL) when no sm_100 GPU is present: there is no CPU fallback. */
int svk_create(int device, svk_ctx** out);
void svk_destroy(svk_ctx* ctx);
const char* svk_last_error(svk_ctx* ctx);
/* Use an external CUDA stream (cudaStream_t as void*), e.g. torch's current stream. */
int svk_set_stream(svk_ctx* ctx, void* cuda_stream);
int svk_sync(svk_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py `gpu_launches`). */
uint64_t svk_launch_count(svk_ctx* ctx);

/* Per-kernel device timing (CUDA events on the launching stream around every launch).  The report is
 * a JSON object {"kernel": {"count": launches, "ms": total device ms}}; reading it resets it. */
int svk_profile_enable(svk_ctx* ctx, int on);
int svk_profile_report(svk_ctx* ctx, char* buf, size_t buf_len);
/* The same launches as a timeline: JSON [["kernel", start_ms, end_ms], ...], times relative to svk_profile_enable(ctx, 1) on the
 * device clock, so that the timelines of several contexts (streams) enabled back to back can be laid side by side. */
int svk_profile_timeline(svk_ctx* ctx, char* buf, size_t buf_len);

/* Test hook: out[i] = `Poseidon::new().update(inputs[i]).squeeze()` (util/hash/poseidon.rs:448-467; T = 3, RATE = 2, R_F = 8,
 * R_P = 57, the SDK transcript's hash: snark-verifier-sdk/src/halo2.rs:52-56), n sponges of n_inputs elements each.
 * schedule 0 = one thread per sponge (poseidon.cuh), 1 = the warp-cooperative permutation (poseidon_coop.cuh). */
int svk_poseidon_squeeze(svk_ctx* ctx, size_t n, const svk_fe* inputs, uint32_t n_inputs, int schedule, svk_fe* out);

/* ---- KzgDecidingKey -------------------------------------------------------------------------
 * `KzgDecidingKey::new(g1, g2, s_g2)` (pcs/kzg/decider.rs:15-24) + halo2curves `G2Prepared::from`
 * (decider.rs:64): validates the G2 points, precomputes the line tables of g2 and -s_g2 and
 * uploads them.  Returns a key id >= 0. */
int svk_dk_load(svk_ctx* ctx, const svk_deciding_key* dk);

/* ---- AccumulationDecider::decide / decide_all (pcs/kzg/decider.rs:60-81) ----------------------
 * out_ok[i] = 1 iff e(lhs_i, g2) * e(rhs_i, -s_g2) == 1.  No fail-fast: every accumulator gets
 * its own answer (`decide_all` == all ones). */
int svk_kzg_decide_batch(svk_ctx* ctx, int dk, size_t n, const svk_acc* accs, uint8_t* out_ok);
int svk_kzg_decide_batch_dev(svk_ctx* ctx, int dk, size_t n, const void* d_accs, void* d_out_ok);

/* ---- PlonkProtocol ingestion ---------------------------------------------------------------------
 * Compiles a serialized `PlonkProtocol<G1Affine>` (verifier/plonk/protocol.rs:21-63; byte layout in
 * snark_verifier_axiom_b200/protocol.py) for `PlonkSuccinctVerifier<KzgAs<Bn256, MOS>>` into a device
 * "verifier tape" -- what `protocol.loaded(&loader)` + the generic verifier do per call in the
 * reference (protocol.rs:106-130, verifier/plonk.rs:58-92) is done once here.  `dk` supplies
 * `svk.g` (pcs/kzg.rs:21-37).  Returns a protocol id >= 0.  A protocol whose expressions the
 * reference would reject with Error::InvalidProtocol still compiles; its proofs get that status. */
int svk_protocol_compile(svk_ctx* ctx, const uint8_t* blob, size_t len, int mos, int dk); /* Poseidon transcript */
int svk_protocol_compile_ex(svk_ctx* ctx, const uint8_t* blob, size_t len, int mos, int transcript_kind, int dk);
/* The same from the reference's own wire format: `bincode::serialize(&protocol)` of a `PlonkProtocol<G1Affine>` (bincode 1.3.3 default
 * options: what `Snark` files hold, snark-verifier-sdk/src/lib.rs:44-50, sdk/src/halo2.rs:262-269) -- the call a Rust shim
 * makes.  fe_encoding: how halo2curves' serde writes Fr / Fq -- SVK_FE_MONTGOMERY (raw `[u64; 4]` Montgomery
 * limbs, the `derive_serde` of halo2curves 0.3.x), SVK_FE_CANONICAL (32-byte LE `to_repr`), SVK_FE_AUTO (whichever makes
 * `domain.n_inv * n == 1`).  *consumed (may be NULL) = bytes of `bytes` the protocol occupied, so that the `instances` and
 * `proof` fields of a `Snark` that follow can be located.  The accumulator encoding is `LimbsEncoding<3, 88>` (sdk/src/lib.rs:33-40). */
enum { SVK_FE_AUTO = 0, SVK_FE_MONTGOMERY = 1, SVK_FE_CANONICAL = 2 };
int svk_protocol_compile_bincode(svk_ctx* ctx, const uint8_t* bytes, size_t len, int fe_encoding, int mos, int transcript_kind, int dk,
                                 size_t* consumed, int* fe_used /* may be NULL: the encoding AUTO settled on */);
/* out[20] = { proof_len, n_instances, n_challenges, n_regs, n_ops, n_poseidon_perms, verify_valid,
 *             n_fr_mul, n_lhs_terms, n_rhs_terms, n_points, n_scalar_slots, msm_modmul_per_proof (all three MSM
 *             kernels), msm_var_modmul_per_proof (k_msm_var only), n_var_terms, var_lanes,
 *             n_old_accumulators, acc_limbs, acc_bits, transcript_kind } */
int svk_protocol_info(svk_ctx* ctx, int proto, uint32_t* out);

/* ---- PlonkSuccinctVerifier::{read_proof, verify} (verifier/plonk.rs:32-93) over a batch ----------
 * For each of the n proofs (proof i = proofs + i*proof_stride, length proof_lens[i] or proof_stride
 * when proof_lens == NULL; trailing bytes are ignored like the reference's reader):
 *   out_acc[i][0]     = the KzgAccumulator {lhs, rhs} (zeros when status != 0); a protocol with `accumulator_indices`
 *                       (aggregation snarks; n_old = svk_protocol_info[16]) yields 1 + n_old records per proof:
 *                       out_acc[i][1 + a] = old accumulator a, `LimbsEncoding<LIMBS, BITS>::from_repr` of the instances it
 *                       names (pcs/kzg/accumulator.rs:57-77) -- the `Vec<KzgAccumulator>` of verifier/plonk.rs:86-91
 *   out_challenges[i] = every squeezed challenge in order (n_challenges each; may be NULL)
 *   out_status[i]     = SVK_OK / SVK_INVALID_INSTANCES / SVK_INVALID_PROTOCOL /
 *                       SVK_TRANSCRIPT | subcode << 8 / SVK_ACCUMULATOR_PANIC
 * instances: n * n_instances field elements (all instance columns of a proof, concatenated);
 * n_instances != sum(protocol.num_instance) => SVK_INVALID_INSTANCES for every proof (proof.rs:66-69).  The batch calls
 * carry the FLAT count; the per-column comparison of proof.rs:66-69 is svk_plonk_instance_shape_ok below, which a host
 * binding applies to every snark before flattening its columns (snark_verifier_axiom_b200/verifier.py: PlonkVerifier.pack).
 * proof_lens[i] > proof_stride is clamped to proof_stride (a proof never reads its neighbour's bytes). */
/* 1 when the column lengths equal protocol.num_instance (proof.rs:66-69), 0 when not, < 0 on a bad protocol id. */
int svk_plonk_instance_shape_ok(svk_ctx* ctx, int proto, uint32_t n_cols, const uint32_t* col_lens);
int svk_plonk_succinct_verify_batch(svk_ctx* ctx, int proto, size_t n, const svk_fe* instances, uint32_t n_instances,
                                    const uint8_t* proofs, size_t proof_stride, const uint32_t* proof_lens, svk_acc* out_acc,
                                    svk_fe* out_challenges, int32_t* out_status);
int svk_plonk_succinct_verify_batch_dev(svk_ctx* ctx, int proto, size_t n, const void* d_instances, uint32_t n_instances,
                                        const void* d_proofs, size_t proof_stride, const void* d_proof_lens, void* d_out_acc,
                                        void* d_out_challenges, void* d_out_status);

/* The unevaluated `Msm` (util/msm.rs:20-24) behind out_acc: the terms of `lhs` (side 0) / `rhs` (side 1) as
 * (fixed, base, slot) triples -- fixed = 1: base indexes protocol.preprocessed (base == len: the generator `svk.g`),
 * fixed = 0: base is the ordinal of a G1 point read from the proof; slot = index into the per-proof scalar vector,
 * -1 when the scalar is the constant 1 -- and the scalar vectors themselves ([n][n_scalar_slots], canonical).
 * Debug / tooling surface (the synthetic-workload generator solves for the last opening point with it). */
int svk_protocol_msm_terms(svk_ctx* ctx, int proto, int side, int32_t* out, size_t max_terms);
int svk_plonk_msm_scalars_batch(svk_ctx* ctx, int proto, size_t n, const svk_fe* instances, uint32_t n_instances, const uint8_t* proofs,
                                size_t proof_stride, const uint32_t* proof_lens, svk_fe* out_scalars, svk_fe* out_challenges,
                                int32_t* out_status);

/* ---- KzgAs accumulation (pcs/kzg/accumulation.rs:17-63, 97-137, 139-196) ------------------------
 * `KzgAs::create_proof` / `read_proof`+`verify` with `KzgAsProvingKey::default()` (zk = false, the
 * SDK's use at snark-verifier-sdk/src/halo2/aggregation.rs:235-245): a fresh Poseidon transcript
 * absorbs every accumulator, r = squeeze, out = (sum r^i lhs_i, sum r^i rhs_i).
 * group_size 0 (or >= n): exactly that flat fold.  group_size m >= 2: consecutive groups of m are
 * folded independently (fresh transcript each), then the group results, until one remains.
 * out_r: challenge of the last (root) fold call.  out_status: 0, or SVK_TRANSCRIPT|sub<<8 when an
 * accumulator point is the identity / not a canonical curve point (reference: Err / unrepresentable). */
int svk_kzg_as_fold(svk_ctx* ctx, size_t n, const svk_acc* accs, size_t group_size, svk_acc* out_acc, svk_fe* out_r,
                    int32_t* out_status);
int svk_kzg_as_fold_dev(svk_ctx* ctx, size_t n, const void* d_accs, size_t group_size, void* d_out_acc, void* d_out_r,
                        void* d_out_status);
/* `KzgAs::{read_proof, verify}` with a zero-knowledge accumulation proof (`KzgAsVerifyingKey::zk()`, pcs/kzg/accumulation.rs:45-49,
 * 124-133): `as_proof` holds the two compressed blind points the prover wrote; they are absorbed after the n instances and
 * folded in as the last pair with r^n.  Short stream / bad encoding / identity => SVK_TRANSCRIPT | sub << 8 in *out_status. */
int svk_kzg_as_fold_zk(svk_ctx* ctx, size_t n, const svk_acc* accs, const uint8_t* as_proof, size_t as_proof_len, svk_acc* out_acc,
                       svk_fe* out_r, int32_t* out_status);

/* ---- PlonkVerifier::verify (verifier/plonk.rs:98-135) over a batch: succinct-verify every proof,
 * fold the accumulators (above), decide the folded accumulator with ONE pairing.
 * out_ok = 1 iff every proof's status is 0 and the folded accumulator is accepted.
 * locate_failures != 0: when all proofs read fine but the folded pairing fails, every accumulator is
 * decided on its own and the offenders get SVK_ASSERTION_FAILURE (what per-proof
 * `PlonkVerifier::verify` would have returned).
 * With old accumulators every proof contributes 1 + n_old accumulators to the fold, new one first (the flattening of
 * snark-verifier-sdk/src/halo2/aggregation.rs:216-245), and `locate_failures` is `decide_all` over each proof's own list.
 * _dev: d_out_accs n*(1+n_old)*128 B, d_out_status n*4 B, d_out_folded 256 B =
 *       { svk_acc folded; svk_fe r; int32 fold_status; uint8 decide_ok; uint8 ok }. */
int svk_plonk_verify_batch(svk_ctx* ctx, int proto, size_t n, const svk_fe* instances, uint32_t n_instances, const uint8_t* proofs,
                           size_t proof_stride, const uint32_t* proof_lens, size_t group_size, int locate_failures,
                           int32_t* out_status, svk_acc* out_folded, uint8_t* out_ok);
int svk_plonk_verify_batch_dev(svk_ctx* ctx, int proto, size_t n, const void* d_instances, uint32_t n_instances, const void* d_proofs,
                               size_t proof_stride, const void* d_proof_lens, size_t group_size, void* d_out_accs, void* d_out_status,
                               void* d_out_folded);

/* ---- G1 multi-scalar multiplication -----------------------------------------------------------------
 * `util::msm::multi_scalar_multiplication(scalars, bases)` (util/msm.rs:238-317), value-identical:
 * out = sum_i scalars[i] * points[i] as a canonical affine point (identity = zeros).
 * out_status: 0 ok, 1 a point is not a canonical curve point, 2 a scalar is >= r (both unrepresentable as
 * halo2curves values; the offending terms are skipped). */
int svk_msm_g1(svk_ctx* ctx, size_t n, const svk_fe* scalars, const svk_g1* points, svk_g1* out, int32_t* out_status);
int svk_msm_g1_dev(svk_ctx* ctx, size_t n, const void* d_scalars, const void* d_points, void* d_out, void* d_status);
/* Batched `base * scalar` (loader/native.rs:67): out[i] = scalars[i] * points[i % n_points]. */
int svk_g1_mul_batch(svk_ctx* ctx, size_t n, const svk_fe* scalars, const svk_g1* points, size_t n_points, svk_g1* out);
int svk_g1_mul_batch_dev(svk_ctx* ctx, size_t n, const void* d_scalars, const void* d_points, size_t n_points, void* d_out);

/* ---- MSM over a chosen curve, and the IPA accumulation decider -------------------------------
 * The reference's `multi_scalar_multiplication` is generic over `CurveAffine` (util/msm.rs:238-317); its one in-tree caller on
 * a verification path is `IpaAs::decide` over the Pasta curves (pcs/ipa/decider.rs:47-56; the reference test uses
 * `pasta::pallas`, pcs/ipa.rs:407-446).  Points are 64-byte affine (x, y LE canonical in the curve's BASE field; identity = zeros),
 * scalars 32-byte LE canonical in its SCALAR field. */
enum { SVK_CURVE_BN254_G1 = 0, SVK_CURVE_PALLAS = 1, SVK_CURVE_VESTA = 2 };
int svk_msm_curve(svk_ctx* ctx, int curve, size_t n, const svk_fe* scalars, const svk_g1* points, svk_g1* out, int32_t* out_status);
int svk_msm_curve_dev(svk_ctx* ctx, int curve, size_t n, const void* d_scalars, const void* d_points, void* d_out, void* d_status);
/* `<IpaAs<C, MOS> as AccumulationDecider<C, NativeLoader>>::decide` per accumulator (decider.rs:47-56; `decide_all`, :58-67, is
 * "every status == 0"): g = `IpaDecidingKey::g` (2^k points, decider.rs:5-16); accumulator i = `IpaAccumulator { xi, u }`
 * (pcs/ipa/accumulator.rs:5-25) = xi[i*k .. (i+1)*k], u[i].  out_status[i] = SVK_OK iff
 * u == multi_scalar_multiplication(h_coeffs(xi, 1), g).to_affine(), else SVK_ASSERTION_FAILURE ("U == commit(G, h)").
 * *out_invalid (may be NULL) = 1 if some input was not a value the reference's types can hold (xi >= the scalar modulus, a point of
 * g off the curve or non-canonical); such accumulators fail.  n == 0 and k == 0 are argument faults (the reference asserts). */
int svk_ipa_decide_batch(svk_ctx* ctx, int curve, uint32_t k, const svk_g1* g, size_t n, const svk_fe* xi, const svk_g1* u,
                         int32_t* out_status, int32_t* out_invalid);
int svk_ipa_decide_batch_dev(svk_ctx* ctx, int curve, uint32_t k, const void* d_g, size_t n, const void* d_xi, const void* d_u,
                             void* d_out_status, void* d_invalid);

/* Several batches per call: n_batches batches of batch_size proofs laid out back to back.  Every batch is folded
 * and decided on its own (same results as n_batches separate svk_plonk_verify_batch calls); the kernels of one call
 * serve all batches, so the serial tail of a batch (fold levels, the pairing) is shared.  Records, 256 B per batch:
 * { svk_acc folded ; svk_fe r ; int32 fold_status ; uint8 decide_ok ; uint8 ok ; padding }. */
int svk_plonk_verify_multi(svk_ctx* ctx, int proto, size_t n_batches, size_t batch_size, const svk_fe* instances, uint32_t n_instances,
                           const uint8_t* proofs, size_t proof_stride, const uint32_t* proof_lens, size_t group_size, int locate_failures,
                           int32_t* out_status, uint8_t* out_records);
int svk_plonk_verify_multi_dev(svk_ctx* ctx, int proto, size_t n_batches, size_t batch_size, const void* d_instances, uint32_t n_instances,
                               const void* d_proofs, size_t proof_stride, const void* d_proof_lens, size_t group_size, void* d_out_accs,
                               void* d_out_status, void* d_out_records);
/* ---- proof-sharded jobs over the GPUs of one box (one context per rank + an ncclComm_t) ---------------------------
 * NCCL is bound at run time (dlopen of libnccl.so.2); single-GPU users never touch it.
 *   svk_nccl_unique_id   rank 0 creates the 128-byte ncclUniqueId and hands it to the other ranks (any side channel)
 *   svk_nccl_init        collective: creates this context's communicator (ncclCommInitRank)
 *   svk_nccl_attach      uses a communicator the host already owns (ncclComm_t as void*); not destroyed with the context
 *   svk_plonk_verify_sharded_dev
 *       this rank's shard (n_batches x batch_size proofs): succinct verify + per-batch fold locally (no pairing), ncclAllGather of
 *       the n_batches 256-byte records on the context stream, fold of batch b over the ranks (`KzgAs`, flat, rank order:
 *       pcs/kzg/accumulation.rs:29-62) and ONE pairing per batch (decider.rs:60-68).  d_gather: world x n_batches x 256 B
 *       ([rank][batch]), d_final_records: n_batches x 256 B { global accumulator ; r ; fold_status ; decide_ok ; ok } with
 *       ok = every rank's local ok && fold_status == 0 && decide_ok.  Everything is enqueued on the stream, nothing blocks. */
int svk_nccl_unique_id(uint8_t* out_id_128);
int svk_nccl_init(svk_ctx* ctx, int world, int rank, const uint8_t* id_128);
int svk_nccl_attach(svk_ctx* ctx, void* nccl_comm, int world, int rank);
int svk_plonk_verify_sharded_dev(svk_ctx* ctx, int proto, size_t n_batches, size_t batch_size, const void* d_instances, uint32_t n_instances,
                                 const void* d_proofs, size_t proof_stride, const void* d_proof_lens, size_t group_size, void* d_out_accs,
                                 void* d_out_status, void* d_out_records, void* d_gather, void* d_final_records);
/* The same without the pairing -- the per-rank half of a proof-sharded job: the per-batch folded accumulators of all
 * ranks are all-gathered, folded once more (svk_kzg_as_fold_multi_dev) and decided ONCE (svk_kzg_decide_records_dev).  Records
 * carry decide_ok = 1 ("not decided here"). */
int svk_plonk_fold_multi_dev(svk_ctx* ctx, int proto, size_t n_batches, size_t batch_size, const void* d_instances, uint32_t n_instances,
                               const void* d_proofs, size_t proof_stride, const void* d_proof_lens, size_t group_size, void* d_out_accs,
                               void* d_out_status, void* d_out_records);

/* Segmented fold / decide on 256-byte records (the layout above), device pointers: n_seg independent groups of n
 * accumulators each ([seg][n] x svk_acc) -> one record per segment; decide fills `decide_ok` of every record.
 * Used for the cross-GPU level of a sharded verification (distributed.py). */
int svk_kzg_as_fold_multi_dev(svk_ctx* ctx, size_t n_seg, size_t n, const void* d_accs, size_t group_size, void* d_out_records);
int svk_kzg_decide_records_dev(svk_ctx* ctx, int dk, size_t n_records, void* d_records);

/* ---- micro-benchmark of the integer-multiply roofline (DESIGN.md "IMAD peak") ------------------
 * Runs `iters` dependent Montgomery multiplications per thread on every SM; returns modmul/s. */
int svk_bench_modmul_peak(svk_ctx* ctx, int iters, double* out_modmul_per_s, double* out_ms);

#ifdef __cplusplus
}
#endif
#endif
