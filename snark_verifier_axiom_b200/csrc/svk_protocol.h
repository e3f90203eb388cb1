// Device image of a compiled protocol (compiler.h -> CompiledProtocol).
#pragma once
#include <string>
#include <vector>

#include "g1.cuh"
#include "tape.cuh"

struct PointSched {
  u32 byte_offset;
  u32 reg_x, reg_y;
};

struct MsmTermDev {
  int32_t fixed;  // 1: fixed_bases[base] (g = index n_pre), 0: proof point ordinal `base`
  int32_t base;
  int32_t slot;   // out_scalar slot, -1 => scalar is the constant 1
};

// One unit of per-proof MSM work, executed by one lane of the proof's lane group (k_proof_msm).
struct MsmWork {
  int32_t kind;   // 0: variable base, Straus item (k_msm_var); 1: fixed base, table windows [w0, w1);
                  // 2: add the base (scalar == 1); 3: add partial number `base` produced by k_msm_var
  int32_t fixed;  // base lives in fixed_bases[] (1) or in the proof-point file (0)
  int32_t base;
  int32_t slot;
  int32_t w0, w1;
};
// One fixed-base table addition of k_msm_sum: every lane of a proof's lane group runs the same number of these
// (a uniform loop: per-lane item lists of different shapes diverged, ncu: 14 of 32 lanes active in that loop).
struct FixedSlot {
  int32_t base;  // index into fixed_bases / tables, -1 = padding (no-op)
  int32_t w;     // 8-bit window index
  int32_t slot;  // scalar slot
};
#define SVK_MSM_LANES_LATENCY 8
// Fixed-base tables use `fixed_bits`-wide windows (8 or 16, chosen per protocol): 256 / bits windows of 2^bits entries per base.
#define SVK_FIXED_BITS_SMALL 8    // 524 KB per base
#define SVK_FIXED_BITS_LARGE 16   // 67 MB per base, half the table additions; used while the tables stay under ~3 GB

// Lane schedule of the per-proof MSM (k_msm_var + k_msm_sum).  Two are built per protocol: [0] minimises total work (one
// Straus thread per proof and side: best when a launch has enough proofs to fill the machine), [1] minimises the dependent
// chain (one k_msm_var thread per variable-base term: 255 doublings + 52 additions deep instead of 255 + 52 x terms).
struct MsmSched {
  MsmWork *d_work_lhs = nullptr, *d_work_rhs = nullptr;  // k_msm_sum items of lane l = work[lane_off[l] .. lane_off[l+1])
  u32 *d_lane_off_lhs = nullptr, *d_lane_off_rhs = nullptr;
  MsmWork* d_var_items = nullptr;      // variable-base terms of both sides
  u32* d_var_lane_off = nullptr;       // items of var lane l = var_items[off[l] .. off[l+1]); partial index = lane
  u32 var_lanes = 1;                   // k_msm_var threads per proof serving the lhs side
  u32 var_lanes_total = 1;             // ... plus the lanes of the rhs side (GWC: rhs = sum u^i W_i)
  u32 var_terms_per_thread = 0;
  u32 msm_lanes = 1;                   // k_msm_sum threads per (proof, side): 1 (throughput) or 8 + shuffle tree (latency)
  FixedSlot *d_fixed_lhs = nullptr, *d_fixed_rhs = nullptr;  // [lane][per] fixed-base table additions
  u32 fixed_per_lhs = 0, fixed_per_rhs = 0;
  size_t msm_work_modmul = 0;          // algorithmic Fq mults per proof of this schedule (DESIGN.md work model)
};

struct ProtocolDevice {
  int mos = 0;
  int transcript_kind = 0;  // 0 Poseidon, 1 Keccak EvmTranscript
  bool verify_valid = true;
  std::string invalid_reason;
  TapeOp* d_ops = nullptr;
  u32 n_ops = 0, read_ops_end = 0;
  uint16_t* d_aux = nullptr;
  Fr* d_consts = nullptr;
  u32 n_regs = 0, n_instances = 0, n_challenges = 0, n_scalar_slots = 0, proof_len = 0;
  int n_perm = 0;
  size_t n_fr_mul = 0;
  std::vector<u32> num_instance;
  std::vector<PointSched> points;
  PointSched* d_sched = nullptr;
  MsmTermDev *d_lhs = nullptr, *d_rhs = nullptr;
  std::vector<MsmTermDev> h_lhs, h_rhs;  // host copies (svk_protocol_msm_terms)
  u32 n_lhs = 0, n_rhs = 0;
  MsmSched sched[2];            // [0] throughput, [1] latency (chosen per launch by the number of proofs)
  u32 fixed_bits = SVK_FIXED_BITS_SMALL;
  G1Affine* d_fixed_tables = nullptr;  // [n_pre + 1][256 / bits windows][2^bits digits]: d * 2^(bits w) * B, affine Montgomery
  u32 n_var = 0;                       // variable-base terms of both sides
  std::string table_key;               // key of the shared fixed-base table (svk_api.cu: g_tables)
  G1Affine* d_fixed = nullptr;  // preprocessed..., then g at index n_pre
  u32 n_pre = 0;
  int dk = -1;                  // deciding key whose g1 is baked in as `svk.g`
  // old accumulators carried in the instances (`LimbsEncoding::from_repr`, pcs/kzg/accumulator.rs:36-79)
  u32 n_old = 0, acc_limbs = 3, acc_bits = 88;
  u32* d_old_idx = nullptr;     // [n_old][4 * acc_limbs] flat instance indices
  u32 accs_per_proof() const { return 1 + n_old; }
};
