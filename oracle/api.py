"""Convenience entry points mirroring how the reference's callers drive the path.
TEST INFRASTRUCTURE ONLY.

  * `verify`            = snark-verifier/examples/recursion.rs:846-855
  * `succinct_verify`   = the loop body of snark-verifier-sdk/src/halo2/aggregation.rs:216-232
  * `fold`              = aggregation.rs:235-245 (`AS::create_proof` with zk = false) and its
                          tree variant (groups of `group_size`, each group a fresh transcript)
  * `decide`            = snark-verifier/src/pcs/kzg/decider.rs:60-68
"""
from . import bn254
from .kzg import KzgAccumulator, KzgAsBdfg21, KzgAsGwc19
from .loader import NativeLoader
from .plonk import PlonkSuccinctVerifier, PlonkVerifier
from .transcript import PoseidonTranscript, ReferencePanic, VerifyError, make_transcript

SCHEMES = {"bdfg21": KzgAsBdfg21, "gwc19": KzgAsGwc19}

# FFI status codes (include/svk.h) <- `Error` (snark-verifier/src/lib.rs:21-30)
STATUS = {"OK": 0, "InvalidInstances": 1, "InvalidProtocol": 2, "AssertionFailure": 3, "Transcript": 4, "Panic": 5}


def succinct_verify(svk, protocol, instances, proof_bytes, scheme, loader=None, want_proof=False, transcript="poseidon"):
    """-> (accumulators, PlonkProof).  Raises VerifyError like the reference returns Err."""
    loader = loader or NativeLoader()
    AS = SCHEMES[scheme]
    flat = 0
    inst = []
    for col in instances:
        inst.append([loader.load_input(x, flat + j) for j, x in enumerate(col)])
        flat += len(col)
    tr = make_transcript(transcript, loader, proof_bytes)
    proof = PlonkSuccinctVerifier.read_proof(svk, protocol, inst, tr, AS)
    accs = PlonkSuccinctVerifier.verify(svk, protocol, inst, proof, AS)
    return (accs, proof) if want_proof else accs


def verify(dk, protocol, instances, proof_bytes, scheme, transcript="poseidon"):
    """`PlonkVerifier::verify(...)`: returns None or raises VerifyError."""
    loader = NativeLoader()
    AS = SCHEMES[scheme]
    inst = [[loader.load_const(x) for x in col] for col in instances]
    tr = make_transcript(transcript, loader, proof_bytes)
    proof = PlonkVerifier.read_proof(dk, protocol, inst, tr, AS)
    return PlonkVerifier.verify(dk, protocol, inst, proof, AS)


def status_of(fn, *a, **kw):
    try:
        fn(*a, **kw)
        return STATUS["OK"]
    except VerifyError as e:
        return STATUS[e.kind]
    except ReferencePanic:
        return STATUS["Panic"]


def fold(accumulators, group_size=0, loader=None):
    """Fold accumulators (list of (lhs_affine, rhs_affine)) into one with `KzgAs::create_proof`
    (zk = false).  group_size == 0 or >= len: the flat fold of aggregation.rs:235-245.  Otherwise a
    tree: consecutive groups of `group_size` are folded independently (each with a fresh
    transcript, as aggregation.rs:216 constructs one per aggregation), then the group results are
    folded the same way until one accumulator remains.  Returns ((lhs, rhs), [r per fold call])."""
    loader = loader or NativeLoader()
    accs = [KzgAccumulator(loader.ec_point_load_const(l), loader.ec_point_load_const(r)) for l, r in accumulators]
    rs = []
    m = group_size if group_size and group_size > 1 else len(accs)
    while True:
        nxt = []
        for i in range(0, len(accs), m):
            grp = accs[i : i + m]
            tr = PoseidonTranscript(loader)
            proof = KzgAsBdfg21.as_read_proof(False, grp, tr)
            rs.append(proof[1].v)
            nxt.append(KzgAsBdfg21.as_verify(False, grp, proof))
        accs = nxt
        if len(accs) == 1:
            break
    return (accs[0].lhs.pt, accs[0].rhs.pt), rs


def decide(dk, acc):
    """-> bool (`KzgAs::decide(..).is_ok()`), acc = (lhs_affine, rhs_affine)"""
    return bn254.pairing_check([(acc[0], dk.g2), (acc[1], bn254.g2_neg(dk.s_g2))])
