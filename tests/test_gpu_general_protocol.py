"""GPU parity on a protocol beyond StandardPlonk: lookup argument, two advice phases with a user challenge,
rotations, two permutation grand products and num_proof = 2 (system/halo2.rs:199-243, 372-408, 593-668) -- the shapes real
halo2 circuits (zkEVM, aggregation) produce.  Proofs are trapdoor-forged, so every structural path of the generic verifier
(verifier/plonk/proof.rs:179-318, pcs/kzg/multiopen/*) runs; challenges, accumulators, statuses vs the oracle."""
import pytest

from oracle import api, forge
from oracle.transcript import VerifyError

from .util import lookup_two_phase_shape, to_product_protocol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from snark_verifier_axiom_b200 import verifier as V

    S = forge.Setup(3, shape=lookup_two_phase_shape(), num_instance=[2], num_proof=2)
    ctx = V.Context(0)
    dk = V.KzgDecidingKey.new(S.dk.svk.g, S.dk.g2, S.dk.s_g2)
    AS = V.KzgAs(ctx, dk)
    proto = to_product_protocol(S.protocol)
    yield V, S, ctx, dk, AS, proto
    ctx.close()


@pytest.mark.parametrize("scheme,mos,transcript", [("bdfg21", 0, "poseidon"), ("gwc19", 1, "poseidon"), ("bdfg21", 0, "evm")])
def test_general_protocol_matches_oracle(env, scheme, mos, transcript):
    V, S, ctx, dk, AS, proto = env
    pv = V.PlonkVerifier(ctx, dk, proto, mos, kzg_as=AS, transcript=V.EVM_TRANSCRIPT if transcript == "evm" else V.POSEIDON_TRANSCRIPT)
    n = 9
    pairs = [forge.forge_proof(S, scheme, 40 + i, transcript=transcript) for i in range(n)]
    insts = [p[0] for p in pairs]
    proofs = [bytearray(p[1]) for p in pairs]
    pt_len, fe_off = (32, 24 * 32) if transcript == "poseidon" else (64, 24 * 64)  # 19 witness + 5 quotient points, then 40 evaluations
    proofs[1][fe_off + 5 * 32 + (0 if transcript == "poseidon" else 31)] ^= 1   # an evaluation: reads fine, pairing fails
    proofs[3][fe_off + 39 * 32 : fe_off + 40 * 32] = b"\xff" * 32               # last evaluation >= r
    proofs[5] = proofs[5][: len(proofs[5]) - 7]                                  # truncated inside the last opening point
    proofs[6][20 * pt_len + 1] ^= 0x40                                           # a lookup/permutation commitment: other point or invalid
    snarks = [V.Snark(i, bytes(p)) for i, p in zip(insts, proofs)]
    accs, chals, st = pv.succinct_verify(snarks)
    n_ok = 0
    for i in range(n):
        try:
            o, proof = api.succinct_verify(S.dk.svk, S.protocol, insts[i], bytes(proofs[i]), scheme, want_proof=True, transcript=transcript)
            want = 0
        except VerifyError as e:
            want = api.STATUS[e.kind]
        assert (int(st[i]) & 0xFF) == want, (i, st[i], want)
        if want == 0:
            n_ok += 1
            assert (accs[i].lhs, accs[i].rhs) == (o[0].lhs.pt, o[0].rhs.pt), i
            ch = [c.v for c in proof.challenges] + [proof.z.v]
            ch += [proof.pcs.mu.v, proof.pcs.gamma.v, proof.pcs.z_prime.v] if scheme == "bdfg21" else [proof.pcs.v.v, proof.pcs.u.v]
            assert chals[i] == ch, i
    assert n_ok >= 6 and int(st[3]) == (4 | 2 << 8) and int(st[5]) == (4 | 1 << 8)
    good = [s for i, s in enumerate(snarks) if i in (0, 2, 4, 7, 8)]
    res = pv.verify(good, group_size=2)
    assert res.ok and (res.status == 0).all()
    res = pv.verify(good + [snarks[1]], group_size=0)
    assert not res.ok and [int(x) for x in res.status] == [0, 0, 0, 0, 0, 3]
    assert api.status_of(api.verify, S.dk, S.protocol, insts[1], bytes(proofs[1]), scheme, transcript=transcript) == 3


@pytest.mark.parametrize("fe", ["montgomery", "canonical"])
def test_snark_file_ingestion(env, fe):
    """`read_snark` + `PlonkVerifier::verify` on a bincode `Snark` (sdk/src/halo2.rs:262-269, sdk/src/lib.rs:44-50): the protocol
    travels in the reference's own wire format and is compiled by the library."""
    from oracle import bincode as obc

    V, S, ctx, dk, AS, proto = env
    inst, proof = forge.forge_proof(S, "bdfg21", 77)
    data = obc.serialize_snark(S.protocol, inst, proof, fe)
    pv, snark = V.PlonkVerifier.from_snark_bincode(ctx, dk, data, V.SHPLONK, kzg_as=AS)
    assert snark.instances == inst and snark.proof == proof
    assert pv.info["proof_len"] == len(proof) and pv.info["n_instances"] == 4
    pv.verify_one(snark)
    accs, _, st = pv.succinct_verify([snark])
    o = api.succinct_verify(S.dk.svk, S.protocol, inst, proof, "bdfg21")
    assert int(st[0]) == 0 and (accs[0].lhs, accs[0].rhs) == (o[0].lhs.pt, o[0].rhs.pt)
    bad = bytearray(proof)
    bad[-1] ^= 1
    with pytest.raises(V.Error):
        pv.verify_one(V.Snark(inst, bytes(bad)))
    # an aggregation-shaped snark: old accumulators named by the protocol inside the file
    S2 = forge.Setup(0, num_instance=14, accumulator_indices=[[(0, 1 + i) for i in range(12)]])
    dk2 = V.KzgDecidingKey.new(S2.dk.svk.g, S2.dk.g2, S2.dk.s_g2)
    inst2, proof2 = forge.forge_proof(S2, "gwc19", 78)
    pv2, snark2 = V.PlonkVerifier.from_snark_bincode(ctx, dk2, obc.serialize_snark(S2.protocol, inst2, proof2, fe), V.GWC)
    assert pv2.info["n_old_accumulators"] == 1
    pv2.verify_one(snark2)
    with pytest.raises(Exception):
        V.PlonkVerifier.from_snark_bincode(ctx, dk, data[: len(data) // 2], V.SHPLONK, kzg_as=AS)
