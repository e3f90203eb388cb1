"""GPU parity for old accumulators: a protocol with `accumulator_indices` -- the shape of an aggregation snark
(snark-verifier-sdk/src/halo2/aggregation.rs:423-425) -- yields [new, old...] per proof (verifier/plonk.rs:86-91) through
`LimbsEncoding<3, 88>::from_repr` (pcs/kzg/accumulator.rs:57-77), and `PlonkVerifier::verify` is `decide_all` over them."""
import pytest

from oracle import api, forge
from oracle.transcript import ReferencePanic, VerifyError

from .util import to_product_protocol

pytestmark = pytest.mark.gpu

N_INST = 27
ACC_IDX = [[(0, 1 + i) for i in range(12)], [(0, 14 + i) for i in range(12)]]  # two old accumulators, a free instance before each


@pytest.fixture(scope="module")
def env():
    from snark_verifier_axiom_b200 import verifier as V

    S = forge.Setup(0, num_instance=N_INST, accumulator_indices=ACC_IDX)
    ctx = V.Context(0)
    dk = V.KzgDecidingKey.new(S.dk.svk.g, S.dk.g2, S.dk.s_g2)
    AS = V.KzgAs(ctx, dk)
    proto = to_product_protocol(S.protocol)
    pv = {m: V.PlonkVerifier(ctx, dk, proto, m, kzg_as=AS) for m in (V.SHPLONK, V.GWC)}
    yield V, S, pv
    ctx.close()


def oracle_status_and_accs(S, inst, proof, scheme):
    try:
        accs = api.succinct_verify(S.dk.svk, S.protocol, inst, proof, scheme)
        return 0, [(a.lhs.pt, a.rhs.pt) for a in accs]
    except VerifyError as e:
        return api.STATUS[e.kind], None
    except ReferencePanic:
        return api.STATUS["Panic"], None


@pytest.mark.parametrize("scheme,mos", [("bdfg21", 0), ("gwc19", 1)])
def test_old_accumulators_match_oracle(env, scheme, mos):
    V, S, pv = env
    assert pv[mos].info["n_old_accumulators"] == 2 and (pv[mos].info["acc_limbs"], pv[mos].info["acc_bits"]) == (3, 88)
    n = 11
    pairs = [forge.forge_proof(S, scheme, 700 + i, old_valid=(i != 3)) for i in range(n)]
    insts = [[list(col) for col in p[0]] for p in pairs]
    proofs = [p[1] for p in pairs]
    insts[5][0][2] ^= 1                        # low limb of lhs.x of old accumulator 0: off the curve -> the reference panics
    insts[6][0][3 + 12] += 1 << 200            # second limb of a coordinate of old accumulator 1: the integer needs > 32 bytes
    insts[7][0][1:13] = [0] * 12               # old accumulator 0 = (identity, identity): `from_xy` accepts (0, 0)
    insts[8][0][14 + 2] = (1 << 88) - 1        # top limb of lhs.x all ones: >= p -> fe_from_big panics
    snarks = [V.Snark(i, p) for i, p in zip(insts, proofs)]
    accs, _, st = pv[mos].succinct_verify(snarks)
    for i in range(n):
        want, oaccs = oracle_status_and_accs(S, insts[i], proofs[i], scheme)
        assert (int(st[i]) & 0xFF) == want, (i, st[i], want)
        if want == 0:
            assert [(a.lhs, a.rhs) for a in accs[i]] == oaccs, i
        else:
            assert accs[i] is None
    assert [int(s) for s in st[[5, 6, 8]]] == [5, 5, 5] and int(st[7]) == 0
    # PlonkVerifier::verify == decide_all over [new, old_0, old_1] of every proof
    good = [snarks[i] for i in (0, 1, 2, 4, 9, 10)]
    res = pv[mos].verify(good, group_size=4)
    assert res.ok and (res.status == 0).all()
    res = pv[mos].verify(good[:3] + [snarks[3]] + good[3:], group_size=0)   # proof 3: the proof is fine, its old accumulator is not
    assert not res.ok and [int(x) for x in res.status] == [0, 0, 0, 3, 0, 0, 0]
    assert api.status_of(api.verify, S.dk, S.protocol, insts[3], proofs[3], scheme) == 3
    assert api.status_of(api.verify, S.dk, S.protocol, insts[0], proofs[0], scheme) == 0
    # the fold of a batch consumes the accumulators of every proof in order: new, old_0, old_1
    flat = [a for i in (0, 1, 2, 4) for a in accs[i]]
    folded, _ = api.fold([(a.lhs, a.rhs) for a in flat], 0)
    res = pv[mos].verify([snarks[i] for i in (0, 1, 2, 4)], group_size=0)
    assert res.ok and (res.folded.lhs, res.folded.rhs) == folded
