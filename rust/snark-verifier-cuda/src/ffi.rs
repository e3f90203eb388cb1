//! Raw C ABI of libsvk: one declaration per entry point of `include/svk.h` (first block: what the safe layer in lib.rs uses).
//! Field elements cross the boundary as 32-byte little-endian canonical values (`to_repr()`).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

#[repr(C)]
pub struct svk_ctx {
    _p: [u8; 0],
}
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct svk_fe {
    pub b: [u8; 32],
}
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct svk_g1 {
    pub x: svk_fe,
    pub y: svk_fe,
}
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct svk_g2 {
    pub x_c0: svk_fe,
    pub x_c1: svk_fe,
    pub y_c0: svk_fe,
    pub y_c1: svk_fe,
}
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct svk_acc {
    pub lhs: svk_g1,
    pub rhs: svk_g1,
}
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct svk_deciding_key {
    pub g1: svk_g1,
    pub g2: svk_g2,
    pub s_g2: svk_g2,
}

/// Status words (`include/svk.h`): the low byte is the reference's `Error` variant, bits 8.. a sub-code.
pub const SVK_OK: i32 = 0;
pub const SVK_INVALID_INSTANCES: i32 = 1;
pub const SVK_INVALID_PROTOCOL: i32 = 2;
pub const SVK_ASSERTION_FAILURE: i32 = 3;
pub const SVK_TRANSCRIPT: i32 = 4;
pub const SVK_ACCUMULATOR_PANIC: i32 = 5;
pub const SVK_T_EOF: i32 = 1;
pub const SVK_T_SCALAR_RANGE: i32 = 2;
pub const SVK_T_POINT_INVALID: i32 = 3;
pub const SVK_T_POINT_IDENTITY: i32 = 4;

extern "C" {
    pub fn svk_create(device: i32, out: *mut *mut svk_ctx) -> i32;
    pub fn svk_destroy(ctx: *mut svk_ctx);
    pub fn svk_last_error(ctx: *mut svk_ctx) -> *const c_char;
    pub fn svk_dk_load(ctx: *mut svk_ctx, dk: *const svk_deciding_key) -> i32;
    pub fn svk_protocol_compile_bincode(
        ctx: *mut svk_ctx, bytes: *const u8, len: usize, fe_encoding: i32, mos: i32, transcript_kind: i32, dk: i32,
        consumed: *mut usize, fe_used: *mut i32,
    ) -> i32;
    pub fn svk_protocol_info(ctx: *mut svk_ctx, proto: i32, out: *mut u32) -> i32;
    pub fn svk_plonk_instance_shape_ok(ctx: *mut svk_ctx, proto: i32, n_cols: u32, col_lens: *const u32) -> i32;
    pub fn svk_plonk_succinct_verify_batch(
        ctx: *mut svk_ctx, proto: i32, n: usize, instances: *const svk_fe, n_instances: u32, proofs: *const u8,
        proof_stride: usize, proof_lens: *const u32, out_acc: *mut svk_acc, out_challenges: *mut svk_fe, out_status: *mut i32,
    ) -> i32;
    pub fn svk_plonk_verify_batch(
        ctx: *mut svk_ctx, proto: i32, n: usize, instances: *const svk_fe, n_instances: u32, proofs: *const u8,
        proof_stride: usize, proof_lens: *const u32, group_size: usize, locate_failures: i32, out_status: *mut i32,
        out_folded: *mut svk_acc, out_ok: *mut u8,
    ) -> i32;
    pub fn svk_kzg_as_fold(
        ctx: *mut svk_ctx, n: usize, accs: *const svk_acc, group_size: usize, out_acc: *mut svk_acc, out_r: *mut svk_fe,
        out_status: *mut i32,
    ) -> i32;
    pub fn svk_kzg_decide_batch(ctx: *mut svk_ctx, dk: i32, n: usize, accs: *const svk_acc, out_ok: *mut u8) -> i32;
    pub fn svk_msm_g1(ctx: *mut svk_ctx, n: usize, scalars: *const svk_fe, points: *const svk_g1, out: *mut svk_g1, out_status: *mut i32) -> i32;
    // proof-sharded job over the GPUs of one box: one context per rank, NCCL inside the library (csrc/sharded.cu)
    pub fn svk_nccl_unique_id(out_id: *mut u8 /* [128] */) -> i32;
    pub fn svk_nccl_init(ctx: *mut svk_ctx, world: i32, rank: i32, id: *const u8 /* [128] */) -> i32;
    pub fn svk_plonk_verify_sharded_dev(
        ctx: *mut svk_ctx, proto: i32, n_batches: usize, batch_size: usize, d_instances: *const c_void, n_instances: u32,
        d_proofs: *const c_void, proof_stride: usize, d_proof_lens: *const c_void, group_size: usize, d_out_accs: *mut c_void,
        d_out_status: *mut c_void, d_out_records: *mut c_void, d_gather: *mut c_void, d_final_records: *mut c_void,
    ) -> i32;
}

// ---- the rest of `include/svk.h`: device-pointer variants (buffers already in HBM: `*_dev`, raw `void*` device pointers), the
// multi-batch calls, zk `KzgAs`, the curve-generic MSM and the IPA decider, stream binding and the profiling hooks.  Kept in step with
// the header by tests/test_abi_and_host.py::test_rust_ffi_declares_the_header.
extern "C" {
    pub fn svk_set_stream(ctx: *mut svk_ctx, cuda_stream: *mut c_void) -> i32;
    pub fn svk_sync(ctx: *mut svk_ctx) -> i32;
    pub fn svk_launch_count(ctx: *mut svk_ctx) -> u64;
    pub fn svk_profile_enable(ctx: *mut svk_ctx, on: i32) -> i32;
    pub fn svk_profile_report(ctx: *mut svk_ctx, buf: *mut c_char, buf_len: usize) -> i32;
    pub fn svk_profile_timeline(ctx: *mut svk_ctx, buf: *mut c_char, buf_len: usize) -> i32;
    pub fn svk_poseidon_squeeze(ctx: *mut svk_ctx, n: usize, inputs: *const svk_fe, n_inputs: u32, schedule: i32, out: *mut svk_fe) -> i32;
    pub fn svk_kzg_decide_batch_dev(ctx: *mut svk_ctx, dk: i32, n: usize, d_accs: *const c_void, d_out_ok: *mut c_void) -> i32;
    pub fn svk_protocol_compile(ctx: *mut svk_ctx, blob: *const u8, len: usize, mos: i32, dk: i32) -> i32;
    pub fn svk_protocol_compile_ex(ctx: *mut svk_ctx, blob: *const u8, len: usize, mos: i32, transcript_kind: i32, dk: i32) -> i32;
    pub fn svk_plonk_succinct_verify_batch_dev(
        ctx: *mut svk_ctx, proto: i32, n: usize, d_instances: *const c_void, n_instances: u32, d_proofs: *const c_void,
        proof_stride: usize, d_proof_lens: *const c_void, d_out_acc: *mut c_void, d_out_challenges: *mut c_void,
        d_out_status: *mut c_void,
    ) -> i32;
    pub fn svk_protocol_msm_terms(ctx: *mut svk_ctx, proto: i32, side: i32, out: *mut i32, max_terms: usize) -> i32;
    pub fn svk_plonk_msm_scalars_batch(
        ctx: *mut svk_ctx, proto: i32, n: usize, instances: *const svk_fe, n_instances: u32, proofs: *const u8, proof_stride: usize,
        proof_lens: *const u32, out_scalars: *mut svk_fe, out_challenges: *mut svk_fe, out_status: *mut i32,
    ) -> i32;
    pub fn svk_kzg_as_fold_dev(
        ctx: *mut svk_ctx, n: usize, d_accs: *const c_void, group_size: usize, d_out_acc: *mut c_void, d_out_r: *mut c_void,
        d_out_status: *mut c_void,
    ) -> i32;
    pub fn svk_kzg_as_fold_zk(
        ctx: *mut svk_ctx, n: usize, accs: *const svk_acc, as_proof: *const u8, as_proof_len: usize, out_acc: *mut svk_acc,
        out_r: *mut svk_fe, out_status: *mut i32,
    ) -> i32;
    pub fn svk_plonk_verify_batch_dev(
        ctx: *mut svk_ctx, proto: i32, n: usize, d_instances: *const c_void, n_instances: u32, d_proofs: *const c_void,
        proof_stride: usize, d_proof_lens: *const c_void, group_size: usize, d_out_accs: *mut c_void, d_out_status: *mut c_void,
        d_out_folded: *mut c_void,
    ) -> i32;
    pub fn svk_msm_g1_dev(
        ctx: *mut svk_ctx, n: usize, d_scalars: *const c_void, d_points: *const c_void, d_out: *mut c_void, d_status: *mut c_void,
    ) -> i32;
    pub fn svk_g1_mul_batch(
        ctx: *mut svk_ctx, n: usize, scalars: *const svk_fe, points: *const svk_g1, n_points: usize, out: *mut svk_g1,
    ) -> i32;
    pub fn svk_g1_mul_batch_dev(
        ctx: *mut svk_ctx, n: usize, d_scalars: *const c_void, d_points: *const c_void, n_points: usize, d_out: *mut c_void,
    ) -> i32;
    pub fn svk_msm_curve(
        ctx: *mut svk_ctx, curve: i32, n: usize, scalars: *const svk_fe, points: *const svk_g1, out: *mut svk_g1, out_status: *mut i32,
    ) -> i32;
    pub fn svk_msm_curve_dev(
        ctx: *mut svk_ctx, curve: i32, n: usize, d_scalars: *const c_void, d_points: *const c_void, d_out: *mut c_void,
        d_status: *mut c_void,
    ) -> i32;
    pub fn svk_ipa_decide_batch(
        ctx: *mut svk_ctx, curve: i32, k: u32, g: *const svk_g1, n: usize, xi: *const svk_fe, u: *const svk_g1, out_status: *mut i32,
        out_invalid: *mut i32,
    ) -> i32;
    pub fn svk_ipa_decide_batch_dev(
        ctx: *mut svk_ctx, curve: i32, k: u32, d_g: *const c_void, n: usize, d_xi: *const c_void, d_u: *const c_void,
        d_out_status: *mut c_void, d_invalid: *mut c_void,
    ) -> i32;
    pub fn svk_plonk_verify_multi(
        ctx: *mut svk_ctx, proto: i32, n_batches: usize, batch_size: usize, instances: *const svk_fe, n_instances: u32,
        proofs: *const u8, proof_stride: usize, proof_lens: *const u32, group_size: usize, locate_failures: i32, out_status: *mut i32,
        out_records: *mut u8,
    ) -> i32;
    pub fn svk_plonk_verify_multi_dev(
        ctx: *mut svk_ctx, proto: i32, n_batches: usize, batch_size: usize, d_instances: *const c_void, n_instances: u32,
        d_proofs: *const c_void, proof_stride: usize, d_proof_lens: *const c_void, group_size: usize, d_out_accs: *mut c_void,
        d_out_status: *mut c_void, d_out_records: *mut c_void,
    ) -> i32;
    pub fn svk_nccl_attach(ctx: *mut svk_ctx, nccl_comm: *mut c_void, world: i32, rank: i32) -> i32;
    pub fn svk_plonk_fold_multi_dev(
        ctx: *mut svk_ctx, proto: i32, n_batches: usize, batch_size: usize, d_instances: *const c_void, n_instances: u32,
        d_proofs: *const c_void, proof_stride: usize, d_proof_lens: *const c_void, group_size: usize, d_out_accs: *mut c_void,
        d_out_status: *mut c_void, d_out_records: *mut c_void,
    ) -> i32;
    pub fn svk_kzg_as_fold_multi_dev(
        ctx: *mut svk_ctx, n_seg: usize, n: usize, d_accs: *const c_void, group_size: usize, d_out_records: *mut c_void,
    ) -> i32;
    pub fn svk_kzg_decide_records_dev(ctx: *mut svk_ctx, dk: i32, n_records: usize, d_records: *mut c_void) -> i32;
    pub fn svk_bench_modmul_peak(ctx: *mut svk_ctx, iters: i32, out_modmul_per_s: *mut f64, out_ms: *mut f64) -> i32;
}
