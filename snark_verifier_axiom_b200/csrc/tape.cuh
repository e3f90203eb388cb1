// K1 core: the per-proof "verifier tape" virtual machine.
//
// The host compiles a `PlonkProtocol` once (compiler.h -- the same idea as the reference's EvmLoader,
// which records a program instead of computing: snark-verifier/src/loader/evm/loader.rs:117-135)
// into a straight-line tape of Fr operations + transcript macro-ops.  Every proof of a batch runs the
// SAME tape (no divergence); error paths only set a status word and keep executing with
// dummy values.  The tape replaces, per proof:
//   PlonkProof::read                      snark-verifier/src/verifier/plonk/proof.rs:52-153
//   PoseidonTranscript (native)           snark-verifier/src/system/halo2/transcript/halo2.rs:198-261
//   Poseidon::squeeze                     snark-verifier/src/util/hash/poseidon.rs:455-467
//   PlonkSuccinctVerifier::verify scalars snark-verifier/src/verifier/plonk.rs:58-82
//   Bdfg21::verify / Gwc19::verify        snark-verifier/src/pcs/kzg/multiopen/{bdfg21.rs:47-79,gwc19.rs:43-80}
// up to (not including) the final group arithmetic, which is the per-proof MSM kernel (proof_msm.cu).
#pragma once
#include "keccak.cuh"
#include "poseidon.cuh"

enum TapeOpCode : uint16_t {
  T_CONST = 0,      // reg[dst] = consts[a]
  T_ADD,            // reg[dst] = reg[a] + reg[b]
  T_SUB,            // reg[dst] = reg[a] - reg[b]
  T_MUL,            // reg[dst] = reg[a] * reg[b]
  T_NEG,            // reg[dst] = -reg[a]
  T_ADDC,           // reg[dst] = reg[a] + consts[b]
  T_SUBC,           // reg[dst] = reg[a] - consts[b]
  T_CSUB,           // reg[dst] = consts[b] - reg[a]
  T_MULC,           // reg[dst] = reg[a] * consts[b]
  T_BINV,           // batch inversion of aux[a .. a+2b) = (src,dst) pairs; 0 -> 0 (loader.rs:241-248)
  T_READ_SCALAR,    // reg[dst] = proof scalar at byte offset 32*(a | b<<16); >= r => SVK_T_SCALAR_RANGE
  T_INSTANCE,       // reg[dst] = instances[a | b<<16]
  T_PERM,           // sponge permutation with dst = n_in inputs reg[a], reg[b]
  T_SQUEEZE,        // reg[dst] = state[1]
  T_RESET,          // sponge state = default
  T_OUT_SCALAR,     // msm_scalar[dst] = canonical(reg[a])
  T_OUT_CHALLENGE,  // challenge[dst] = canonical(reg[a])
  // ---- Keccak `EvmTranscript` (transcript/evm.rs:152-243); only emitted for transcript kind 1
  T_READ_SCALAR_BE, // reg[dst] = proof scalar, 32 B BIG-endian, at byte offset 32*(a | b<<16)
  T_KABSORB_REG,    // buf.extend(canonical(reg[a]) as 32 B big-endian)
  T_KABSORB_PROOF,  // buf.extend(proof[32*(a | b<<16) .. +32])   (a scalar or half of an uncompressed point, as read)
  T_KSQUEEZE,       // hash = keccak256(buf || [1 if a]); buf = hash; reg[dst] = hash (big-endian) mod r
  T_N_OPS
};

struct TapeOp {
  uint16_t op, dst, a, b;
};

// first-error word: (byte offset of the offending read << 8) | sub-code; 0xffffffff = no error
#define SVK_NO_ERR 0xffffffffu
HD void tape_note_error(u32& err, u32 byte_off, u32 sub) {
  u32 w = (byte_off << 8) | sub;
  if (w < err) err = w;
}

// Register file accessor: regs[reg][item] as 8 x u32, `n_items` items interleaved per register so that
// consecutive threads (items) touch consecutive 32-byte slots (coalesced 2 x 16 B per thread).
struct RegFile {
  u32* base;
  size_t n_items;
  size_t item;
  HD Fr load(u32 r) const {
    Fr x;
    const uint4* p = reinterpret_cast<const uint4*>(base + (r * n_items + item) * 8);
    uint4 lo = p[0], hi = p[1];
    x.v[0] = lo.x; x.v[1] = lo.y; x.v[2] = lo.z; x.v[3] = lo.w;
    x.v[4] = hi.x; x.v[5] = hi.y; x.v[6] = hi.z; x.v[7] = hi.w;
    return x;
  }
  HD void store(u32 r, const Fr& x) const {
    uint4* p = reinterpret_cast<uint4*>(base + (r * n_items + item) * 8);
    p[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    p[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
  }
};

struct TapeIo {
  const uint8_t* proof;      // this item's proof bytes
  u32 proof_len;
  const uint8_t* instances;  // this item's instances, 32 B each (LE canonical)
  u32 n_instances;
  u32* out_scalars;          // [slot][item] canonical, same interleaving as RegFile
  u32* out_challenges;       // [item][slot] canonical (ABI order)
  u32 n_challenge_slots;
};

#define SVK_BINV_MAX 48

// Transcript state of one proof.  KECCAK selects the member that exists at compile time, so the Poseidon kernel
// does not carry the 25-lane Keccak state.
template <bool KECCAK>
struct TranscriptState;
template <>
struct TranscriptState<false> {
  PoseidonState ps;
};
template <>
struct TranscriptState<true> {
  KeccakSponge ks;
};
// Poseidon sponge interface of the tape: permute / squeeze / reset.  Implemented by the one-thread sponge (here) and by the
// warp-cooperative one (poseidon_coop.cuh: PoseidonCoopMain, used when a launch has too few proofs to fill the machine).
HD void ts_permute(TranscriptState<false>& st, const PoseidonConsts& pk, int n_in, const Fr& in0, const Fr& in1) { poseidon_permute(st.ps, pk, n_in, in0, in1); }
HD Fr ts_squeeze(const TranscriptState<false>& st) { return st.ps.s[1]; }
HD void ts_reset(TranscriptState<false>& st, const PoseidonConsts& pk) { poseidon_init(st.ps, pk); }
HD void ts_permute(TranscriptState<true>&, const PoseidonConsts&, int, const Fr&, const Fr&) {}
HD Fr ts_squeeze(const TranscriptState<true>&) { return Fr::zero(); }
HD void ts_reset(TranscriptState<true>& st, const PoseidonConsts&) { keccak_reset(st.ks); }

// Executes ops [begin, end) for one item.  Sponge state persists in `st` across calls.
template <bool KECCAK, class Regs, class TS>
HD void tape_exec(const TapeOp* ops, u32 begin, u32 end, const uint16_t* aux, const Fr* consts, const PoseidonConsts& pk,
                  const Regs& regs, const TapeIo& io, TS& st, u32& err) {
  for (u32 pc = begin; pc < end; pc++) {
    TapeOp op = ops[pc];
    switch (op.op) {
      case T_CONST: regs.store(op.dst, consts[op.a]); break;
      case T_ADD: regs.store(op.dst, regs.load(op.a) + regs.load(op.b)); break;
      case T_SUB: regs.store(op.dst, regs.load(op.a) - regs.load(op.b)); break;
      case T_MUL: regs.store(op.dst, regs.load(op.a) * regs.load(op.b)); break;
      case T_NEG: regs.store(op.dst, regs.load(op.a).neg()); break;
      case T_ADDC: regs.store(op.dst, regs.load(op.a) + consts[op.b]); break;
      case T_SUBC: regs.store(op.dst, regs.load(op.a) - consts[op.b]); break;
      case T_CSUB: regs.store(op.dst, consts[op.b] - regs.load(op.a)); break;
      case T_MULC: regs.store(op.dst, regs.load(op.a) * consts[op.b]); break;
      case T_BINV: {
        // Montgomery batch inversion, zero entries skipped and mapped to zero.
        const uint16_t* pr = aux + op.a;
        u32 n = op.b;
        Fr acc = Fr::one();
        // forward: prefix products stored in the dst registers
        for (u32 i = 0; i < n; i++) {
          Fr x = regs.load(pr[2 * i]);
          regs.store(pr[2 * i + 1], acc);  // product of the non-zero entries before i
          if (!x.is_zero()) acc = acc * x;
        }
        Fr inv = acc.inv();
        for (u32 i = n; i-- > 0;) {
          Fr x = regs.load(pr[2 * i]);
          Fr pre = regs.load(pr[2 * i + 1]);
          if (x.is_zero()) {
            regs.store(pr[2 * i + 1], Fr::zero());
          } else {
            regs.store(pr[2 * i + 1], inv * pre);
            inv = inv * x;
          }
        }
        break;
      }
      case T_READ_SCALAR: {
        u32 off = 32u * ((u32)op.a | ((u32)op.b << 16));
        Fr x = Fr::zero();
        if (off + 32 > io.proof_len) {
          tape_note_error(err, off, SVK_T_EOF);
        } else {
          fe_load_le(x.v, io.proof + off);
          if (!Fr::is_canonical(x.v)) {
            tape_note_error(err, off, SVK_T_SCALAR_RANGE);
            x = Fr::zero();
          }
        }
        regs.store(op.dst, x.to_mont());
        break;
      }
      case T_INSTANCE: {
        u32 idx = (u32)op.a | ((u32)op.b << 16);
        Fr x = Fr::zero();
        if (idx < io.n_instances) {
          fe_load_le(x.v, io.instances + 32 * (size_t)idx);
          if (!Fr::is_canonical(x.v)) {
            tape_note_error(err, 0, 0);  // non-canonical instance: reported as SVK_INVALID_INSTANCES
            x = Fr::zero();
          }
        }
        regs.store(op.dst, x.to_mont());
        break;
      }
      case T_PERM:
        if constexpr (!KECCAK) {
          Fr in0 = Fr::zero(), in1 = Fr::zero();
          if (op.dst >= 1) in0 = regs.load(op.a);
          if (op.dst >= 2) in1 = regs.load(op.b);
          ts_permute(st, pk, op.dst, in0, in1);
        }
        break;
      case T_SQUEEZE:
        if constexpr (!KECCAK) regs.store(op.dst, ts_squeeze(st));
        break;
      case T_RESET:
        ts_reset(st, pk);
        break;
      case T_READ_SCALAR_BE:
        if constexpr (KECCAK) {
          u32 off = 32u * ((u32)op.a | ((u32)op.b << 16));
          Fr x = Fr::zero();
          if (off + 32 > io.proof_len) {
            tape_note_error(err, off, SVK_T_EOF);
          } else {
            const uint8_t* p = io.proof + off;
            for (int i = 0; i < 8; i++) {
              const uint8_t* q = p + 4 * (7 - i);
              x.v[i] = ((u32)q[0] << 24) | ((u32)q[1] << 16) | ((u32)q[2] << 8) | (u32)q[3];
            }
            if (!Fr::is_canonical(x.v)) {
              tape_note_error(err, off, SVK_T_SCALAR_RANGE);
              x = Fr::zero();
            }
          }
          regs.store(op.dst, x.to_mont());
        }
        break;
      case T_KABSORB_REG:
        if constexpr (KECCAK) {
          Fr x = regs.load(op.a).from_mont();
          keccak_absorb_limbs_be(st.ks, x.v);
        }
        break;
      case T_KABSORB_PROOF:
        if constexpr (KECCAK) {
          u32 off = 32u * ((u32)op.a | ((u32)op.b << 16));
          if (off + 32 <= io.proof_len) keccak_absorb(st.ks, io.proof + off, 32);
        }
        break;
      case T_KSQUEEZE:
        if constexpr (KECCAK) {
          if (op.a) keccak_absorb_byte(st.ks, 1);
          uint8_t h[32];
          keccak_finish(st.ks, h);
          keccak_absorb(st.ks, h, 32);  // buf = hash
          Fr x;
          for (int i = 0; i < 8; i++) {
            const uint8_t* q = h + 4 * (7 - i);
            x.v[i] = ((u32)q[0] << 24) | ((u32)q[1] << 16) | ((u32)q[2] << 8) | (u32)q[3];
          }
          // u256_to_fe: value mod r.  (R^2 * x + m r) / 2^256 < 2r for any 256-bit x, reduced once by the multiplier
          regs.store(op.dst, x.to_mont_wide());
        }
        break;
      case T_OUT_SCALAR: {
        Fr x = regs.load(op.a).from_mont();
        RegFile o{io.out_scalars, regs.n_items, regs.item};
        o.store(op.dst, x);
        break;
      }
      case T_OUT_CHALLENGE: {
        Fr x = regs.load(op.a).from_mont();
        u32* p = io.out_challenges + ((size_t)regs.item * io.n_challenge_slots + op.dst) * 8;
        for (int i = 0; i < 8; i++) p[i] = x.v[i];
        break;
      }
      default: break;
    }
  }
}
