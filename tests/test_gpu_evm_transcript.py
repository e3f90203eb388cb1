"""GPU parity for the Keccak `EvmTranscript` path (snark-verifier/src/system/halo2/transcript/evm.rs:152-243):
proofs in the EVM wire format (32 B big-endian scalars, 64 B uncompressed big-endian points) -- challenges, accumulators,
statuses and the composed batch verify against the oracle, SHPLONK and GWC."""
import pytest

from oracle import api, forge
from oracle.transcript import VerifyError

from .util import to_product_protocol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from snark_verifier_axiom_b200 import verifier as V

    S = forge.Setup(0)
    ctx = V.Context(0)
    dk = V.KzgDecidingKey.new(S.dk.svk.g, S.dk.g2, S.dk.s_g2)
    AS = V.KzgAs(ctx, dk)
    proto = to_product_protocol(S.protocol)
    pv = {m: V.PlonkVerifier(ctx, dk, proto, m, kzg_as=AS, transcript=V.EVM_TRANSCRIPT) for m in (V.SHPLONK, V.GWC)}
    yield V, S, pv
    ctx.close()


@pytest.mark.parametrize("scheme,mos,plen", [("bdfg21", 0, 1248), ("gwc19", 1, 1312)])
def test_evm_transcript_matches_oracle(env, scheme, mos, plen):
    V, S, pv = env
    assert pv[mos].info["proof_len"] == plen
    n = 19
    insts, proofs = forge.forge_batch(S, scheme, n, seed0=300, transcript="evm")
    proofs = [bytearray(p) for p in proofs]
    proofs[2][9 * 64 + 31] ^= 1                   # an evaluation (low byte): still a scalar, pairing fails later
    proofs[4][9 * 64 : 9 * 64 + 32] = b"\xff" * 32  # scalar >= r
    proofs[6][63] ^= 1                            # y of the first point: off the curve
    proofs[8][0:64] = bytes(64)                   # (0, 0): decodes to the identity, rejected by common_ec_point
    snarks = [V.Snark(i, bytes(p)) for i, p in zip(insts, proofs)]
    accs, chals, st = pv[mos].succinct_verify(snarks)
    for i in range(n):
        try:
            o, proof = api.succinct_verify(S.dk.svk, S.protocol, insts[i], bytes(proofs[i]), scheme, want_proof=True, transcript="evm")
            want = 0
        except VerifyError as e:
            want = api.STATUS[e.kind]
        assert (int(st[i]) & 0xFF) == want, (i, st[i], want)
        if want == 0:
            assert (accs[i].lhs, accs[i].rhs) == (o[0].lhs.pt, o[0].rhs.pt)
            ch = [c.v for c in proof.challenges] + [proof.z.v]
            ch += [proof.pcs.mu.v, proof.pcs.gamma.v, proof.pcs.z_prime.v] if scheme == "bdfg21" else [proof.pcs.v.v, proof.pcs.u.v]
            assert chals[i] == ch
    assert [int(s) & 0xFF for s in st[[4, 6, 8]]] == [4, 4, 4] and [int(s) >> 8 for s in st[[4, 6, 8]]] == [2, 3, 4]
    good = [s for i, s in enumerate(snarks) if i not in (2, 4, 6, 8)]
    res = pv[mos].verify(good, group_size=4)
    assert res.ok and (res.status == 0).all()
    res = pv[mos].verify(good[:5] + [snarks[2]], group_size=0)
    assert not res.ok and [int(x) for x in res.status] == [0, 0, 0, 0, 0, 3]
    assert api.status_of(api.verify, S.dk, S.protocol, insts[2], bytes(proofs[2]), scheme, transcript="evm") == 3
    # a Poseidon-format proof is not an EVM-format proof
    i2, p2 = forge.forge_proof(S, scheme, 1)
    _, _, st2 = pv[mos].succinct_verify([V.Snark(i2, p2)])
    assert (int(st2[0]) & 0xFF) == 4
