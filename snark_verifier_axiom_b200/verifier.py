"""Host-side mirror of the reference's verification call surface, over libsvk (include/svk.h).

Mirrored names (reference file:line):
  KzgSuccinctVerifyingKey, KzgDecidingKey   snark-verifier/src/pcs/kzg.rs:21-37, pcs/kzg/decider.rs:6-36
  KzgAccumulator                            snark-verifier/src/pcs/kzg/accumulator.rs:6-26
  KzgAs.{create_proof, decide, decide_all}  snark-verifier/src/pcs/kzg/accumulation.rs:139-196, decider.rs:60-81
  PlonkSuccinctVerifier / PlonkVerifier     snark-verifier/src/verifier/plonk.rs:32-135
  Snark, SHPLONK, GWC                       snark-verifier-sdk/src/lib.rs:38-60
  Error                                     snark-verifier/src/lib.rs:21-30
The unit of work is a BATCH of snarks of one protocol (the B200 is a throughput device); a batch of
one reproduces the reference's single-proof calls.  numpy arrays are host buffers; the `*_dev`
methods take torch CUDA tensors (device memory + streams are PyTorch's job, nothing else is).
"""
import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from ._lib import SvkError, lib
from .protocol import FR_MODULUS, PlonkProtocol

SHPLONK = BDFG21 = 0
GWC = GWC19 = 1
POSEIDON_TRANSCRIPT = 0  # sdk `PoseidonTranscript` (snark-verifier-sdk/src/halo2.rs:58-67)
EVM_TRANSCRIPT = 1       # Keccak `EvmTranscript` (snark-verifier/src/system/halo2/transcript/evm.rs)

FE_AUTO, FE_MONTGOMERY, FE_CANONICAL = 0, 1, 2  # include/svk.h SVK_FE_*
_R_INV = pow(1 << 256, -1, FR_MODULUS)
STATUS_NAMES = {0: "Ok", 1: "InvalidInstances", 2: "InvalidProtocol", 3: "AssertionFailure", 4: "Transcript", 5: "AccumulatorPanic"}


class Error(Exception):
    """snark-verifier/src/lib.rs:21-30"""

    def __init__(self, status: int):
        self.status = int(status)
        self.kind = STATUS_NAMES.get(self.status & 0xFF, "Unknown")
        super().__init__(f"{self.kind} (status 0x{self.status:x})")


def _fe(v: int) -> bytes:
    return int(v).to_bytes(32, "little")


def _g1_bytes(pt) -> bytes:
    return bytes(64) if pt is None else _fe(pt[0]) + _fe(pt[1])


def _g1_from(b: bytes):
    x, y = int.from_bytes(b[:32], "little"), int.from_bytes(b[32:64], "little")
    return None if x == 0 and y == 0 else (x, y)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(ctypes.c_void_p)
    return ctypes.c_void_p(a.data_ptr())  # torch tensor


@dataclass
class KzgSuccinctVerifyingKey:
    g: Tuple[int, int]


@dataclass
class KzgDecidingKey:
    svk: KzgSuccinctVerifyingKey
    g2: Tuple[Tuple[int, int], Tuple[int, int]]
    s_g2: Tuple[Tuple[int, int], Tuple[int, int]]

    @classmethod
    def new(cls, g1, g2, s_g2):
        return cls(KzgSuccinctVerifyingKey(g1), g2, s_g2)

    def to_bytes(self) -> bytes:
        def g2b(q):
            return _fe(q[0][0]) + _fe(q[0][1]) + _fe(q[1][0]) + _fe(q[1][1])

        return _g1_bytes(self.svk.g) + g2b(self.g2) + g2b(self.s_g2)


@dataclass
class KzgAccumulator:
    lhs: Optional[Tuple[int, int]]
    rhs: Optional[Tuple[int, int]]

    def to_bytes(self) -> bytes:
        return _g1_bytes(self.lhs) + _g1_bytes(self.rhs)

    @classmethod
    def from_bytes(cls, b: bytes):
        return cls(_g1_from(b[:64]), _g1_from(b[64:128]))


@dataclass
class Snark:
    """snark-verifier-sdk/src/lib.rs:44-60 (the protocol is shared by the batch)"""

    instances: List[List[int]]
    proof: bytes


class Context:
    """One GPU + one stream (`svk_ctx`).  Raises SvkError when no B200 is present."""

    def __init__(self, device: int = 0):
        self._L = lib()
        self._c = ctypes.c_void_p()
        if self._L.svk_create(device, ctypes.byref(self._c)) != 0:
            raise SvkError(self._L.svk_last_error(None).decode())
        self.device = device

    def close(self):
        if self._c:
            self._L.svk_destroy(self._c)
            self._c = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise SvkError(self._L.svk_last_error(self._c).decode())
        return rc

    def set_stream(self, cuda_stream: int):
        self._check(self._L.svk_set_stream(self._c, ctypes.c_void_p(cuda_stream)))

    def sync(self):
        self._check(self._L.svk_sync(self._c))

    @property
    def launch_count(self) -> int:
        return int(self._L.svk_launch_count(self._c))

    def load_deciding_key(self, dk: KzgDecidingKey) -> int:
        return self._check(self._L.svk_dk_load(self._c, dk.to_bytes()))

    def compile_protocol(self, protocol: PlonkProtocol, mos: int, dk_id: int, transcript: int = 0) -> int:
        blob = protocol.to_bytes()
        return self._check(self._L.svk_protocol_compile_ex(self._c, blob, len(blob), mos, transcript, dk_id))

    def compile_protocol_bincode(self, data: bytes, mos: int, dk_id: int, transcript: int = 0, fe_encoding: int = 0):
        """`bincode::serialize(&PlonkProtocol<G1Affine>)` bytes (the head of a `Snark` file) -> (protocol id, bytes consumed,
        field-element encoding used).  include/svk.h `svk_protocol_compile_bincode`."""
        consumed, fe_used = ctypes.c_size_t(0), ctypes.c_int(0)
        pid = self._check(self._L.svk_protocol_compile_bincode(self._c, data, len(data), fe_encoding, mos, transcript, dk_id,
                                                                ctypes.byref(consumed), ctypes.byref(fe_used)))
        return pid, consumed.value, fe_used.value

    def protocol_info(self, pid: int) -> dict:
        out = (ctypes.c_uint32 * 20)()
        self._check(self._L.svk_protocol_info(self._c, pid, out))
        keys = ["proof_len", "n_instances", "n_challenges", "n_regs", "n_ops", "n_poseidon_perms", "verify_valid", "n_fr_mul",
                "n_lhs_terms", "n_rhs_terms", "n_points", "n_scalar_slots", "msm_modmul_per_proof", "msm_var_modmul_per_proof",
                "n_var_terms", "var_lanes", "n_old_accumulators", "acc_limbs", "acc_bits", "transcript_kind"]
        return dict(zip(keys, [int(x) for x in out]))

    def modmul_peak(self, iters: int = 4000):
        rate, ms = ctypes.c_double(), ctypes.c_double()
        self._check(self._L.svk_bench_modmul_peak(self._c, iters, ctypes.byref(rate), ctypes.byref(ms)))
        return rate.value, ms.value


class KzgAs:
    """`KzgAs<Bn256, MOS>`: accumulation + deciding (the multi-open part lives in the compiled protocol)."""

    def __init__(self, ctx: Context, dk: KzgDecidingKey):
        self.ctx = ctx
        self.dk = dk
        self.dk_id = ctx.load_deciding_key(dk)

    def create_proof(self, accumulators: Sequence[KzgAccumulator], group_size: int = 0):
        """accumulation.rs:139-196 with zk = false -> (KzgAccumulator, r).  Raises Error like the
        reference returns Err (identity point in an accumulator)."""
        n = len(accumulators)
        assert n > 0, "assert!(!instances.is_empty())"
        buf = np.frombuffer(b"".join(a.to_bytes() for a in accumulators), dtype=np.uint8).copy()
        out, r, st = np.zeros(128, np.uint8), np.zeros(32, np.uint8), np.zeros(1, np.int32)
        L, c = self.ctx._L, self.ctx._c
        self.ctx._check(L.svk_kzg_as_fold(c, n, _ptr(buf), group_size, _ptr(out), _ptr(r), _ptr(st)))
        if st[0] != 0:
            raise Error(st[0])
        return KzgAccumulator.from_bytes(out.tobytes()), int.from_bytes(r.tobytes(), "little")

    def verify_zk(self, accumulators: Sequence[KzgAccumulator], as_proof: bytes):
        """`KzgAs::read_proof` + `verify` with `KzgAsVerifyingKey(true)` (accumulation.rs:29-62, 113-136): `as_proof` = the two
        compressed blind points written by a zero-knowledge `create_proof` -> (KzgAccumulator, r)."""
        n = len(accumulators)
        assert n > 0, "assert!(!instances.is_empty())"
        buf = np.frombuffer(b"".join(a.to_bytes() for a in accumulators), dtype=np.uint8).copy()
        pf = np.frombuffer(bytes(as_proof) or b"\0", dtype=np.uint8).copy()
        out, r, st = np.zeros(128, np.uint8), np.zeros(32, np.uint8), np.zeros(1, np.int32)
        L, c = self.ctx._L, self.ctx._c
        self.ctx._check(L.svk_kzg_as_fold_zk(c, n, _ptr(buf), _ptr(pf), len(as_proof), _ptr(out), _ptr(r), _ptr(st)))
        if st[0] != 0:
            raise Error(st[0])
        return KzgAccumulator.from_bytes(out.tobytes()), int.from_bytes(r.tobytes(), "little")

    def decide_batch(self, accumulators: Sequence[KzgAccumulator]) -> List[bool]:
        n = len(accumulators)
        if n == 0:
            return []
        buf = np.frombuffer(b"".join(a.to_bytes() for a in accumulators), dtype=np.uint8).copy()
        out = np.zeros(n, np.uint8)
        self.ctx._check(self.ctx._L.svk_kzg_decide_batch(self.ctx._c, self.dk_id, n, _ptr(buf), _ptr(out)))
        return [bool(x) for x in out]

    def decide(self, accumulator: KzgAccumulator) -> None:
        """decider.rs:60-68"""
        if not self.decide_batch([accumulator])[0]:
            raise Error(3)

    def decide_all(self, accumulators: Sequence[KzgAccumulator]) -> None:
        """decider.rs:70-81"""
        assert len(accumulators) > 0, "assert!(!accumulators.is_empty())"
        if not all(self.decide_batch(accumulators)):
            raise Error(3)


@dataclass
class BatchResult:
    ok: bool
    status: np.ndarray  # int32 per proof
    folded: Optional[KzgAccumulator]

    def errors(self) -> List[Optional[Error]]:
        return [None if s == 0 else Error(s) for s in self.status]


class PlonkVerifier:
    """`PlonkVerifier<KzgAs<Bn256, MOS>>` / `PlonkSuccinctVerifier<..>` for ONE protocol, over batches."""

    def __init__(self, ctx: Context, dk: KzgDecidingKey, protocol: Optional[PlonkProtocol], mos: int = SHPLONK, kzg_as: Optional[KzgAs] = None,
                 transcript: int = POSEIDON_TRANSCRIPT, _pid: Optional[int] = None):
        self.ctx = ctx
        self.kzg_as = kzg_as or KzgAs(ctx, dk)
        self.mos = mos
        self.transcript = transcript
        self.pid = ctx.compile_protocol(protocol, mos, self.kzg_as.dk_id, transcript) if _pid is None else _pid
        self.info = ctx.protocol_info(self.pid)
        self.protocol = protocol

    @classmethod
    def from_snark_bincode(cls, ctx: Context, dk: KzgDecidingKey, data: bytes, mos: int = SHPLONK, kzg_as: Optional[KzgAs] = None,
                           transcript: int = POSEIDON_TRANSCRIPT, fe_encoding: int = FE_AUTO):
        """`read_snark` (snark-verifier-sdk/src/halo2.rs:262-269): a bincode `Snark { protocol, instances, proof }`
        (sdk/src/lib.rs:44-50) -> (verifier for its protocol, Snark).  The protocol is parsed and compiled by the library
        itself (`svk_protocol_compile_bincode`); only the two trailing vectors are read here.  As the reference warns, the
        caller must know which multi-open scheme the snark was made with."""
        kzg_as = kzg_as or KzgAs(ctx, dk)
        pid, pos, fe_used = ctx.compile_protocol_bincode(data, mos, kzg_as.dk_id, transcript, fe_encoding)
        pv = cls(ctx, dk, None, mos, kzg_as, transcript, _pid=pid)

        def u64():
            nonlocal pos
            if pos + 8 > len(data):
                raise ValueError("snark file truncated")
            v = int.from_bytes(data[pos : pos + 8], "little")
            pos += 8
            return v

        def fr():
            nonlocal pos
            if pos + 32 > len(data):
                raise ValueError("snark file truncated")
            v = int.from_bytes(data[pos : pos + 32], "little")
            pos += 32
            if v >= FR_MODULUS:
                raise ValueError("instance out of range")
            return v * _R_INV % FR_MODULUS if fe_used == FE_MONTGOMERY else v

        instances = []
        for _ in range(u64()):
            instances.append([fr() for _ in range(u64())])
        n = u64()
        if pos + n > len(data):
            raise ValueError("snark file truncated")
        return pv, Snark(instances, bytes(data[pos : pos + n]))

    # ---- packing ------------------------------------------------------------------------------
    def pack(self, snarks: Sequence[Snark]):
        """-> (instances u8[n, n_inst*32], n_inst, proofs u8[n, stride], lens u32[n]).
        `self.bad_shape` lists the snarks whose instance COLUMNS do not match `protocol.num_instance` (proof.rs:66-69: the
        comparison is per column; the ABI carries the flat count): they are packed with zero instances and reported as
        InvalidInstances by `_apply_shape`."""
        n = len(snarks)
        L, c = self.ctx._L, self.ctx._c
        self.bad_shape = []
        n_inst = self.info["n_instances"]
        shapes = {}
        for i, s in enumerate(snarks):
            key = tuple(len(col) for col in s.instances)
            if key not in shapes:
                lens_ = np.array(key, np.uint32)
                shapes[key] = L.svk_plonk_instance_shape_ok(c, self.pid, len(key), _ptr(lens_) if len(key) else None) == 1
            if not shapes[key]:
                self.bad_shape.append(i)
        if n and len(self.bad_shape) == n:
            # nothing in the batch has the protocol's shape: keep the caller's flat count so that libsvk reports the error itself
            n_inst = len([x for col in snarks[0].instances for x in col])
        stride = max([len(s.proof) for s in snarks] + [32])
        stride = (stride + 31) // 32 * 32
        proofs = np.zeros((n, stride), np.uint8)
        lens = np.zeros(n, np.uint32)
        inst = np.zeros((n, max(n_inst, 1) * 32), np.uint8)
        bad = set(self.bad_shape)
        for i, s in enumerate(snarks):
            flat = [x for col in s.instances for x in col]
            proofs[i, : len(s.proof)] = np.frombuffer(s.proof, np.uint8)
            lens[i] = len(s.proof)
            if flat and i not in bad:
                inst[i, : n_inst * 32] = np.frombuffer(b"".join(_fe(x) for x in flat), np.uint8)
            elif flat and len(bad) == n and len(flat) == n_inst:
                inst[i, : n_inst * 32] = np.frombuffer(b"".join(_fe(x) for x in flat), np.uint8)
        return inst, n_inst, proofs, lens

    def _apply_shape(self, st):
        """InvalidInstances precedes everything else `read_proof` reports (proof.rs:66-69 is its first check)."""
        for i in getattr(self, "bad_shape", []):
            st[i] = 1
        return st

    # ---- PlonkSuccinctVerifier::{read_proof, verify} ------------------------------------------------
    def succinct_verify(self, snarks: Sequence[Snark]):
        """-> (accumulators [KzgAccumulator|None], challenges [[int]], status int32[n]).
        For a protocol with `accumulator_indices` (an aggregation snark) entry i is the list [new, old_0, ...] of
        verifier/plonk.rs:86-91 instead of a single accumulator."""
        n = len(snarks)
        inst, n_inst, proofs, lens = self.pack(snarks)
        nch = self.info["n_challenges"]
        apk = 1 + self.info["n_old_accumulators"]
        acc = np.zeros((n, apk, 128), np.uint8)
        ch = np.zeros((n, max(nch, 1), 32), np.uint8)
        st = np.zeros(n, np.int32)
        L, c = self.ctx._L, self.ctx._c
        self.ctx._check(L.svk_plonk_succinct_verify_batch(c, self.pid, n, _ptr(inst), n_inst, _ptr(proofs), proofs.shape[1], _ptr(lens),
                                                          _ptr(acc), _ptr(ch), _ptr(st)))
        st = self._apply_shape(st)
        if apk == 1:
            accs = [KzgAccumulator.from_bytes(acc[i, 0].tobytes()) if st[i] == 0 else None for i in range(n)]
        else:
            accs = [[KzgAccumulator.from_bytes(acc[i, k].tobytes()) for k in range(apk)] if st[i] == 0 else None for i in range(n)]
        chals = [[int.from_bytes(ch[i, j].tobytes(), "little") for j in range(nch)] for i in range(n)]
        return accs, chals, st

    # ---- PlonkVerifier::verify -----------------------------------------------------------------------
    def verify(self, snarks: Sequence[Snark], group_size: int = 0, locate_failures: bool = True) -> BatchResult:
        n = len(snarks)
        inst, n_inst, proofs, lens = self.pack(snarks)
        st = np.zeros(n, np.int32)
        folded = np.zeros(128, np.uint8)
        ok = np.zeros(1, np.uint8)
        L, c = self.ctx._L, self.ctx._c
        self.ctx._check(L.svk_plonk_verify_batch(c, self.pid, n, _ptr(inst), n_inst, _ptr(proofs), proofs.shape[1], _ptr(lens), group_size,
                                                 1 if locate_failures else 0, _ptr(st), _ptr(folded), _ptr(ok)))
        st = self._apply_shape(st)
        if self.bad_shape:
            ok[0] = 0
        return BatchResult(bool(ok[0]), st, KzgAccumulator.from_bytes(folded.tobytes()) if ok[0] else None)

    def verify_batches(self, batches: Sequence[Sequence[Snark]], group_size: int = 0, locate_failures: bool = True) -> List[BatchResult]:
        """Several equally sized batches in ONE call (`svk_plonk_verify_multi`): each batch is folded and decided on
        its own; the kernels of the call serve all batches."""
        nb = len(batches)
        bs = len(batches[0])
        assert all(len(b) == bs for b in batches)
        flat = [s for b in batches for s in b]
        inst, n_inst, proofs, lens = self.pack(flat)
        st = np.zeros(nb * bs, np.int32)
        rec = np.zeros(nb * 256, np.uint8)
        L, c = self.ctx._L, self.ctx._c
        self.ctx._check(L.svk_plonk_verify_multi(c, self.pid, nb, bs, _ptr(inst), n_inst, _ptr(proofs), proofs.shape[1], _ptr(lens), group_size,
                                                 1 if locate_failures else 0, _ptr(st), _ptr(rec)))
        st = self._apply_shape(st)
        out = []
        for b in range(nb):
            r = rec[b * 256 : (b + 1) * 256]
            ok = bool(r[165]) and not any(b * bs <= i < (b + 1) * bs for i in self.bad_shape)
            out.append(BatchResult(ok, st[b * bs : (b + 1) * bs], KzgAccumulator.from_bytes(r[:128].tobytes()) if ok else None))
        return out

    def verify_one(self, snark: Snark) -> None:
        """The reference's single-proof `PlonkVerifier::verify(..)` -> Ok(()) or raises Error."""
        r = self.verify([snark])
        if r.status[0] != 0:
            raise Error(r.status[0])
        if not r.ok:
            raise Error(3)


# ---- EcPointLoader::multi_scalar_multiplication / util::msm (snark-verifier/src/util/msm.rs:238-317) ----
def multi_scalar_multiplication(ctx: Context, scalars: Sequence[int], bases: Sequence[Optional[Tuple[int, int]]]):
    """sum_i scalars[i] * bases[i] -> affine point (None = identity).  Pippenger bucket kernel (csrc/msm.cu)."""
    assert len(scalars) == len(bases)
    n = len(scalars)
    sc = np.frombuffer(b"".join(_fe(s) for s in scalars), dtype=np.uint8).copy() if n else np.zeros(0, np.uint8)
    pt = np.frombuffer(b"".join(_g1_bytes(p) for p in bases), dtype=np.uint8).copy() if n else np.zeros(0, np.uint8)
    out, st = np.zeros(64, np.uint8), np.zeros(1, np.int32)
    ctx._check(ctx._L.svk_msm_g1(ctx._c, n, _ptr(sc), _ptr(pt), _ptr(out), _ptr(st)))
    if st[0] != 0:
        raise ValueError(f"msm: invalid input (status {int(st[0])})")
    return _g1_from(out.tobytes())


def g1_mul_batch(ctx: Context, scalars: Sequence[int], bases: Sequence[Optional[Tuple[int, int]]]):
    """[scalars[i] * bases[i % len(bases)]] -- batched `base * scalar` (loader/native.rs:67)."""
    n = len(scalars)
    sc = np.frombuffer(b"".join(_fe(s) for s in scalars), dtype=np.uint8).copy()
    pt = np.frombuffer(b"".join(_g1_bytes(p) for p in bases), dtype=np.uint8).copy()
    out = np.zeros(n * 64, np.uint8)
    ctx._check(ctx._L.svk_g1_mul_batch(ctx._c, n, _ptr(sc), _ptr(pt), len(bases), _ptr(out)))
    b = out.tobytes()
    return [_g1_from(b[64 * i : 64 * i + 64]) for i in range(n)]


# ---- the same MSM over a chosen curve, and the IPA decider ----------------------------------
CURVE_BN254_G1, CURVE_PALLAS, CURVE_VESTA = 0, 1, 2  # include/svk.h SVK_CURVE_*


def multi_scalar_multiplication_on(ctx: Context, curve: int, scalars: Sequence[int], bases: Sequence[Optional[Tuple[int, int]]]):
    """`multi_scalar_multiplication` is generic over `CurveAffine` (util/msm.rs:238): the Pippenger kernels instantiated for `curve`."""
    assert len(scalars) == len(bases)
    n = len(scalars)
    sc = np.frombuffer(b"".join(_fe(s) for s in scalars), dtype=np.uint8).copy() if n else np.zeros(0, np.uint8)
    pt = np.frombuffer(b"".join(_g1_bytes(p) for p in bases), dtype=np.uint8).copy() if n else np.zeros(0, np.uint8)
    out, st = np.zeros(64, np.uint8), np.zeros(1, np.int32)
    ctx._check(ctx._L.svk_msm_curve(ctx._c, curve, n, _ptr(sc), _ptr(pt), _ptr(out), _ptr(st)))
    if st[0] != 0:
        raise ValueError(f"msm: invalid input (status {int(st[0])})")
    return _g1_from(out.tobytes())


@dataclass
class IpaAccumulator:
    """snark-verifier/src/pcs/ipa/accumulator.rs:5-25"""

    xi: Sequence[int]
    u: Optional[Tuple[int, int]]


@dataclass
class IpaDecidingKey:
    """snark-verifier/src/pcs/ipa/decider.rs:5-16 (`svk` is not used by `decide`; `g` has 2^k points)."""

    g: Sequence[Optional[Tuple[int, int]]]
    curve: int = CURVE_PALLAS


class IpaAs:
    """`<IpaAs<C, MOS> as AccumulationDecider<C, NativeLoader>>` (snark-verifier/src/pcs/ipa/decider.rs:33-67)."""

    @staticmethod
    def decide_batch(ctx: Context, dk: IpaDecidingKey, accumulators: Sequence[IpaAccumulator]) -> np.ndarray:
        """Per-accumulator status (0 ok, 3 AssertionFailure "U == commit(G, h)"), no fail-fast."""
        assert len(accumulators) > 0  # decider.rs:61
        k = len(accumulators[0].xi)
        assert k > 0 and len(dk.g) == 1 << k and all(len(a.xi) == k for a in accumulators)  # pcs/ipa.rs:380
        n = len(accumulators)
        g = np.frombuffer(b"".join(_g1_bytes(p) for p in dk.g), dtype=np.uint8).copy()
        xi = np.frombuffer(b"".join(_fe(x) for a in accumulators for x in a.xi), dtype=np.uint8).copy()
        u = np.frombuffer(b"".join(_g1_bytes(a.u) for a in accumulators), dtype=np.uint8).copy()
        st, inv = np.zeros(n, np.int32), np.zeros(1, np.int32)
        ctx._check(ctx._L.svk_ipa_decide_batch(ctx._c, dk.curve, k, _ptr(g), n, _ptr(xi), _ptr(u), _ptr(st), _ptr(inv)))
        return st

    @staticmethod
    def decide(ctx: Context, dk: IpaDecidingKey, accumulator: IpaAccumulator) -> None:
        st = IpaAs.decide_batch(ctx, dk, [accumulator])
        if st[0] != 0:
            raise Error(st[0])

    @staticmethod
    def decide_all(ctx: Context, dk: IpaDecidingKey, accumulators: Sequence[IpaAccumulator]) -> None:
        st = IpaAs.decide_batch(ctx, dk, accumulators)
        bad = np.nonzero(st)[0]
        if len(bad):
            raise Error(st[bad[0]])
