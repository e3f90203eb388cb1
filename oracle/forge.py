"""Trapdoor forging of VALID StandardPlonk proofs (fixtures without a prover).
TEST INFRASTRUCTURE ONLY.

The reference generates every proof at run time with the halo2 prover
(snark-verifier-sdk/src/halo2.rs:77-146, 178-260), which cannot run here.  With the SRS trapdoor
`s` known and every G1 point given a known discrete log, the whole verification collapses to
`lhs = s * rhs` (snark-verifier/src/pcs/kzg/decider.rs:64-67), and the last opening point(s) can
be solved for:
  * SHPLONK: W' is read last and no challenge depends on it (bdfg21.rs:101-106):
        w' = dlog(f) / (s - z')
  * GWC: every W_i precedes `u` (gwc19.rs:104-106), so each set must vanish on its own:
        w_i = dlog(sum_j v^j (C_ij - e_ij G)) / (s - z * shift_i)
Given the transcript prefix the accepting W' (W_i) is unique, so forged proofs are byte-identical
to honest ones with the same randomness.
"""
import random

from . import bn254
from .bn254 import R
from .halo2_system import compile_protocol, standard_plonk_protocol, standard_plonk_shape
from .kzg import KzgAsBdfg21, KzgAsGwc19, KzgDecidingKey, gwc19_query_sets
from .loader import EcPoint, NativeLoader
from .plonk import CommonPolynomialEvaluation, PlonkProof
from .transcript import PoseidonTranscript, make_transcript

# ---------------------------------------------------------------- fixed-base table for G (8-bit windows)
_TABLE = None


def _table():
    global _TABLE
    if _TABLE is None:
        tbl = []
        base = bn254.G1_GEN
        for _ in range(32):
            row = [None]
            acc = None
            for _d in range(255):
                acc = bn254.g1_add(acc, base)
                row.append(acc)
            tbl.append(row)
            base = bn254.g1_add(acc, base)  # 256 * base
        _TABLE = tbl
    return _TABLE


def g_mul(k):
    """k * G via the window table (value-identical to bn254.g1_mul(G1_GEN, k))."""
    k %= R
    tbl = _table()
    acc = (1, 1, 0)
    for w in range(32):
        d = (k >> (8 * w)) & 0xFF
        if d:
            acc = bn254._jac_add_affine(acc, tbl[w][d])
    return bn254._jac_to_affine(acc)


class DlogLoader(NativeLoader):
    """NativeLoader whose MSM is evaluated in the exponent (all bases carry `.dlog`)."""

    def multi_scalar_multiplication(self, pairs):
        d = 0
        for s, b in pairs:
            assert b.dlog is not None
            d = (d + s.v * b.dlog) % R
        return EcPoint(None, self, None, dlog=d)

    def ec_point_load_const(self, pt):
        e = super().ec_point_load_const(pt)
        e.dlog = self.known.get(pt)
        return e

    def __init__(self, known):
        super().__init__()
        self.known = known  # affine tuple -> dlog


class Setup:
    """SRS trapdoor + a StandardPlonk verifying key with known dlogs (seeded, reproducible)."""

    def __init__(self, seed=0, k=8, num_instance=1, accumulator_indices=(), shape=None, num_proof=1):
        """`num_instance` rows in the single instance column; `accumulator_indices`: lists of 12 (column, row) pairs naming
        the limbs of old accumulators (an aggregation-circuit-shaped protocol, sdk/src/halo2/aggregation.rs:423-425).
        `shape`: another circuit than StandardPlonk (a `ConstraintSystemShape`: lookups, phases, ...); `num_proof` > 1:
        one protocol verifying several proofs of that circuit at once (system/halo2.rs:57-63)."""
        rng = random.Random(seed)
        self.k = k
        self.s = rng.randrange(1, R)
        self.g1 = bn254.G1_GEN
        self.g2 = bn254.G2_GEN
        self.s_g2 = bn254.g2_mul(bn254.G2_GEN, self.s)
        n_pre = 8 if shape is None else shape.num_fixed + len(shape.permutation_columns)
        self.vk_dlogs = [rng.randrange(1, R) for _ in range(n_pre)]
        self.preprocessed = [g_mul(d) for d in self.vk_dlogs]
        self.transcript_initial_state = rng.randrange(R)  # stands in for the vk digest (system/halo2.rs:137)
        if shape is None and num_instance == 1 and not accumulator_indices and num_proof == 1:
            self.protocol = standard_plonk_protocol(k, self.preprocessed, self.transcript_initial_state)
        else:
            num_instance = [num_instance] if isinstance(num_instance, int) else list(num_instance)
            self.protocol = compile_protocol(k, shape or standard_plonk_shape(), self.preprocessed, self.transcript_initial_state,
                                             num_instance, num_proof=num_proof)
            self.protocol.accumulator_indices = [list(x) for x in accumulator_indices]
        self.dk = KzgDecidingKey.new(self.g1, self.g2, self.s_g2)
        self.known = {self.g1: 1}
        for d, p in zip(self.vk_dlogs, self.preprocessed):
            self.known[p] = d


def _msm_dlog(msm):
    d = 0 if msm.constant is None else msm.constant.v
    for s, b in zip(msm.scalars, msm.bases):
        d = (d + s.v * b.dlog) % R
    return d


def old_accumulator_limbs(setup, d, limbs=3, bits=88, valid=True):
    """The 4 * LIMBS instance values of an old accumulator (lhs, rhs) = (s d G, d G) -- or, with valid=False,
    ((s d + 1) G, d G): well-formed points that fail the pairing check: `fe_to_limbs` of x, y of both points
    (util/arithmetic.rs:278-290; what the aggregation circuit exposes)."""
    lhs, rhs = g_mul((setup.s * d + (0 if valid else 1)) % R), g_mul(d)
    out = []
    for v in (lhs[0], lhs[1], rhs[0], rhs[1]):
        out += [(v >> (bits * i)) & ((1 << bits) - 1) for i in range(limbs)]
    return out


def forge_proof(setup, scheme, seed, transcript="poseidon", old_valid=True):
    """Returns (instances [[int]], proof bytes) accepted by PlonkVerifier<KzgAs<Bn256, scheme>>.
    scheme in {"bdfg21", "gwc19"}; transcript "poseidon" (compressed LE points, LE scalars) or "evm"
    (Keccak EvmTranscript: uncompressed BE points, BE scalars)."""
    rng = random.Random(("proof", seed).__repr__())
    protocol = setup.protocol
    AS = KzgAsBdfg21 if scheme == "bdfg21" else KzgAsGwc19
    n_w = sum(protocol.num_witness)
    n_q = protocol.quotient.num_chunk()
    n_open = 2 if scheme == "bdfg21" else len(gwc19_query_sets(PlonkProof.empty_queries(protocol)))
    known = dict(setup.known)

    def rand_point():
        d = rng.randrange(1, R)
        p = g_mul(d)
        known[p] = d
        return p

    instances = [[rng.randrange(R) for _ in range(n)] for n in protocol.num_instance]
    for idx in protocol.accumulator_indices:
        for (i, j), limb in zip(idx, old_accumulator_limbs(setup, rng.randrange(1, R), valid=old_valid)):
            instances[i][j] = limb
    wit = [rand_point() for _ in range(n_w)]
    quo = [rand_point() for _ in range(n_q)]
    evals = [rng.randrange(R) for _ in protocol.evaluations]
    if scheme == "bdfg21":
        opens = [rand_point(), bn254.G1_GEN]  # W random, W' placeholder
    else:
        opens = [bn254.G1_GEN] * n_open  # placeholders

    def enc_pt(p):
        if transcript == "poseidon":
            return bn254.g1_to_bytes(p)
        return int(p[0]).to_bytes(32, "big") + int(p[1]).to_bytes(32, "big")

    def enc_fe(e):
        return bn254.fe_to_bytes(e) if transcript == "poseidon" else int(e).to_bytes(32, "big")

    def assemble(open_pts):
        out = bytearray()
        for p in wit + quo:
            out += enc_pt(p)
        for e in evals:
            out += enc_fe(e)
        for p in open_pts:
            out += enc_pt(p)
        return bytes(out)

    loader = DlogLoader(known)
    inst_loaded = [[loader.load_const(x) for x in col] for col in instances]
    tr = make_transcript(transcript, loader, assemble(opens))
    proof = PlonkProof.read(setup.dk.svk, protocol, inst_loaded, tr, AS)
    for pt in proof.witnesses + proof.quotients:
        pt.dlog = known[pt.pt]
    cpe = CommonPolynomialEvaluation(protocol.domain, protocol.langranges(), proof.z)
    NativeLoader.batch_invert(cpe.denoms())
    cpe.evaluate()
    evaluations = proof.evaluations_map(protocol, inst_loaded, cpe)
    commitments = proof.commitments(protocol, cpe, evaluations)
    queries = proof.queries(protocol, evaluations)
    s = setup.s
    if scheme == "bdfg21":
        proof.pcs.w.dlog = known[proof.pcs.w.pt]
        proof.pcs.w_prime.dlog = 1
        acc = AS.verify(setup.dk.svk, commitments, proof.z, queries, proof.pcs)
        zp = proof.pcs.z_prime.v
        phi = (acc.lhs.dlog - zp) % R
        w_prime = phi * pow((s - zp) % R, R - 2, R) % R
        opens = [opens[0], g_mul(w_prime)]
    else:
        sets = gwc19_query_sets(queries)
        powers_of_v = proof.pcs.v.powers(max(len(st.polys) for st in sets))
        new_opens = []
        for st in sets:
            phi = _msm_dlog(st.msm(commitments, powers_of_v))
            denom = (s - proof.z.v * st.shift) % R
            new_opens.append(g_mul(phi * pow(denom, R - 2, R) % R))
        opens = new_opens
    return instances, assemble(opens)


def forge_batch(setup, scheme, n, seed0=1, transcript="poseidon"):
    out = [forge_proof(setup, scheme, seed0 + i, transcript) for i in range(n)]
    return [o[0] for o in out], [o[1] for o in out]
