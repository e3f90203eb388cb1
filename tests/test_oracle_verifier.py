"""The oracle's verifier restatement: protocol shape of StandardPlonk k=8, trapdoor-forged
proofs accept under both multi-open schemes, mutations reject with the reference's error kinds, folds
(flat = snark-verifier-sdk/src/halo2/aggregation.rs:235-245, and trees) stay valid, and the committed
golden fixture is what the oracle computes."""
import random

import pytest

from oracle import api, bn254, forge
from oracle.kzg import bdfg21_query_sets, gwc19_query_sets
from oracle.plonk import PlonkProof
from oracle.transcript import VerifyError


@pytest.fixture(scope="module")
def S():
    return forge.Setup(0)


def test_protocol_shape(S):
    P = S.protocol
    assert (P.domain.k, P.domain.n) == (8, 256)
    assert P.domain.gen == 0x1058A83D529BE585820B96FF0A13F2DBD8675A9E5DD2336A6692CC1E5A526C81
    assert P.num_instance == [1] and P.num_witness == [3, 0, 3] and P.num_challenge == [1, 2, 1]
    assert P.quotient.num_chunk() == 3 and len(P.preprocessed) == 8
    assert P.evaluations == [(9, 0), (10, 0), (11, 0), (0, 0), (1, 0), (2, 0), (3, 0), (4, 0), (14, 0), (5, 0), (6, 0), (7, 0),
                             (12, 0), (12, 1), (12, -6), (13, 0), (13, 1)]
    assert P.queries == [(9, 0), (10, 0), (11, 0), (12, 0), (12, 1), (13, 0), (13, 1), (12, -6)] + [(i, 0) for i in range(8)] + [(15, 0), (14, 0)]
    assert sorted(set(P.langranges())) == [-6, -5, -4, -3, -2, -1, 0]
    sets = bdfg21_query_sets(PlonkProof.empty_queries(P))
    assert [s.polys for s in sets] == [[9, 10, 11, 0, 1, 2, 3, 4, 5, 6, 7, 15, 14], [12], [13]]
    assert [len(s.shifts) for s in sets] == [1, 3, 2]
    g = gwc19_query_sets(PlonkProof.empty_queries(P))
    assert [len(s.polys) for s in g] == [15, 2, 1]


@pytest.mark.parametrize("scheme,plen,nperm", [("bdfg21", 896, 27), ("gwc19", 928, 28)])
def test_forged_proofs_accept_and_mutations_reject(S, scheme, plen, nperm):
    inst, pf = forge.forge_proof(S, scheme, 11)
    assert len(pf) == plen
    accs = api.succinct_verify(S.dk.svk, S.protocol, inst, pf, scheme)
    acc = accs[0]
    assert bn254.g1_mul(acc.rhs.pt, S.s) == acc.lhs.pt  # lhs = s * rhs  (decider.rs:64-67 in the exponent)
    assert api.status_of(api.verify, S.dk, S.protocol, inst, pf, scheme) == 0
    assert api.status_of(api.verify, S.dk, S.protocol, inst, pf + b"trailing bytes are ignored", scheme) == 0
    rng = random.Random(4)
    for _ in range(3):
        bad = bytearray(pf)
        bad[rng.randrange(9 * 32, 26 * 32 - 1)] ^= 1 << rng.randrange(8)  # an evaluation byte, below the top byte
        assert api.status_of(api.verify, S.dk, S.protocol, inst, bytes(bad), scheme) in (3, 4)
    bad = bytearray(pf)
    bad[9 * 32 : 10 * 32] = b"\xff" * 32
    assert api.status_of(api.verify, S.dk, S.protocol, inst, bytes(bad), scheme) == 4  # scalar >= r
    bad = bytearray(pf)
    bad[0:32] = bytes(32)
    assert api.status_of(api.verify, S.dk, S.protocol, inst, bytes(bad), scheme) == 4  # identity point
    assert api.status_of(api.verify, S.dk, S.protocol, inst, pf[:-1], scheme) == 4  # short read
    assert api.status_of(api.verify, S.dk, S.protocol, [[1, 2]], pf, scheme) == 1  # InvalidInstances
    assert api.status_of(api.verify, S.dk, S.protocol, [[(inst[0][0] + 1) % bn254.R]], pf, scheme) == 3


def test_same_commitments_both_schemes(S):
    """SHPLONK and GWC forgeries over the same seed share witness/quotient/evals and both accept."""
    i1, p1 = forge.forge_proof(S, "bdfg21", 5)
    i2, p2 = forge.forge_proof(S, "gwc19", 5)
    assert i1 == i2 and p1[: 26 * 32] == p2[: 26 * 32]
    assert api.status_of(api.verify, S.dk, S.protocol, i1, p1, "bdfg21") == 0
    assert api.status_of(api.verify, S.dk, S.protocol, i2, p2, "gwc19") == 0


def test_fold_flat_and_tree(S):
    insts, proofs = forge.forge_batch(S, "bdfg21", 9, seed0=40)
    pairs = []
    for i, p in zip(insts, proofs):
        a = api.succinct_verify(S.dk.svk, S.protocol, i, p, "bdfg21")[0]
        pairs.append((a.lhs.pt, a.rhs.pt))
    for m in (0, 2, 4):
        (l, r), rs = api.fold(pairs, m)
        assert bn254.g1_mul(r, S.s) == l
    (l, r), rs = api.fold(pairs, 0)
    assert len(rs) == 1
    exp_l = exp_r = None
    for j, (a, b) in enumerate(pairs):
        k = pow(rs[0], j, bn254.R)
        exp_l = bn254.g1_add(exp_l, bn254.g1_mul(a, k))
        exp_r = bn254.g1_add(exp_r, bn254.g1_mul(b, k))
    assert (l, r) == (exp_l, exp_r)
    bad = list(pairs)
    bad[3] = (bn254.g1_add(bad[3][0], bn254.G1_GEN), bad[3][1])
    (l, r), _ = api.fold(bad, 4)
    assert not api.decide(S.dk, (l, r))
    with pytest.raises(VerifyError):
        api.fold([(None, pairs[0][1])], 0)


def test_golden_fixture_matches_oracle(S):
    import json
    import os

    path = os.path.join(os.path.dirname(__file__), "golden", "standard_plonk_k8.json")
    g = json.load(open(path))
    assert [int(p[0], 16) for p in g["preprocessed"]] == [p[0] for p in S.preprocessed]
    for scheme in ("bdfg21", "gwc19"):
        s = g["schemes"][scheme]
        for i in (0, 7):
            inst = [[int(x, 16) for x in col] for col in s["instances"][i]]
            pf = bytes.fromhex(s["proofs"][i])
            accs, proof = api.succinct_verify(S.dk.svk, S.protocol, inst, pf, scheme, want_proof=True)
            e = s["expect"][i]
            assert int(e["lhs"][0], 16) == accs[0].lhs.pt[0] and int(e["rhs"][1], 16) == accs[0].rhs.pt[1]
            assert int(e["challenges"][4], 16) == proof.z.v


def test_old_accumulators_limbs_encoding():
    """`accumulator_indices` + `LimbsEncoding<3, 88>` (pcs/kzg/accumulator.rs:57-77, verifier/plonk.rs:86-91, 131-134)."""
    from oracle.loader import fe_to_limbs
    from oracle.transcript import ReferencePanic

    S2 = forge.Setup(0, num_instance=14, accumulator_indices=[[(0, 1 + i) for i in range(12)]])
    inst, proof = forge.forge_proof(S2, "bdfg21", 41)
    accs = api.succinct_verify(S2.dk.svk, S2.protocol, inst, proof, "bdfg21")
    assert len(accs) == 2
    old = accs[1]
    want = []
    for v in (old.lhs.pt[0], old.lhs.pt[1], old.rhs.pt[0], old.rhs.pt[1]):
        want += fe_to_limbs(v, 3, 88)
    assert inst[0][1:13] == want  # from_repr inverts fe_to_limbs
    assert api.status_of(api.verify, S2.dk, S2.protocol, inst, proof, "bdfg21") == 0
    # a well-formed old accumulator that fails the pairing: only decide_all notices
    inst_b, proof_b = forge.forge_proof(S2, "bdfg21", 41, old_valid=False)
    assert len(api.succinct_verify(S2.dk.svk, S2.protocol, inst_b, proof_b, "bdfg21")) == 2
    assert api.status_of(api.verify, S2.dk, S2.protocol, inst_b, proof_b, "bdfg21") == 3
    # off-curve limbs: the reference panics
    bad = [list(inst[0])]
    bad[0][1] ^= 1
    with pytest.raises(ReferencePanic):
        api.succinct_verify(S2.dk.svk, S2.protocol, bad, proof, "bdfg21")
    assert api.status_of(api.verify, S2.dk, S2.protocol, bad, proof, "bdfg21") == 5
