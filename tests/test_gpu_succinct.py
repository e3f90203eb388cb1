"""GPU parity: svk_plonk_succinct_verify_batch vs the oracle's PlonkSuccinctVerifier
(snark-verifier/src/verifier/plonk.rs:32-93) on trapdoor-forged StandardPlonk k=8 proofs:
transcript challenges, accumulators and per-proof status, SHPLONK and GWC."""
import ctypes

import numpy as np
import pytest

from oracle import api, forge
from oracle.transcript import VerifyError

from .util import dk_bytes, g1_from, np_u8, ptr, to_product_protocol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from snark_verifier_axiom_b200._lib import lib

    L = lib()
    c = ctypes.c_void_p()
    assert L.svk_create(0, ctypes.byref(c)) == 0, L.svk_last_error(None)
    S = forge.Setup(0)
    kid = L.svk_dk_load(c, dk_bytes(S.dk))
    assert kid >= 0
    blob = to_product_protocol(S.protocol).to_bytes()
    yield L, c, S, kid, blob
    L.svk_destroy(c)


def oracle_result(S, inst, pf, scheme):
    try:
        accs, proof = api.succinct_verify(S.dk.svk, S.protocol, inst, pf, scheme, want_proof=True)
    except VerifyError as e:
        return api.STATUS[e.kind], None, None
    ch = [c.v for c in proof.challenges] + [proof.z.v]
    if scheme == "bdfg21":
        ch += [proof.pcs.mu.v, proof.pcs.gamma.v, proof.pcs.z_prime.v]
    else:
        ch += [proof.pcs.v.v, proof.pcs.u.v]
    return 0, (accs[0].lhs.pt, accs[0].rhs.pt), ch


@pytest.mark.parametrize("scheme,mos", [("bdfg21", 0), ("gwc19", 1)])
def test_succinct_verify_matches_oracle(env, scheme, mos):
    L, c, S, kid, blob = env
    pid = L.svk_protocol_compile(c, blob, len(blob), mos, kid)
    assert pid >= 0, L.svk_last_error(c)
    info = (ctypes.c_uint32 * 20)()
    assert L.svk_protocol_info(c, pid, info) == 0
    plen, n_inst, n_ch = info[0], info[1], info[2]
    assert plen == (896 if scheme == "bdfg21" else 928) and n_inst == 1
    n = 37  # ragged vs the 16-lane groups / 32-thread blocks
    insts, proofs = forge.forge_batch(S, scheme, n, seed0=100)
    proofs = [bytearray(p) for p in proofs]
    # corruptions: evaluation flip (still decodes), scalar >= r, invalid point, identity point
    proofs[3][9 * 32 + 5] ^= 1
    proofs[5][9 * 32 : 10 * 32] = b"\xff" * 32
    from oracle import bn254

    bad_x = next(x for x in range(1, 100) if not bn254.g1_from_bytes(x.to_bytes(32, "little"))[0])
    proofs[7][0:32] = bad_x.to_bytes(32, "little")  # x^3 + 3 is a non-residue -> invalid point encoding
    proofs[9][32:64] = bytes(32)
    stride = plen + 32  # trailing bytes are ignored
    buf = np.zeros((n, stride), dtype=np.uint8)
    for i, p in enumerate(proofs):
        buf[i, :plen] = np.frombuffer(bytes(p), dtype=np.uint8)
        buf[i, plen:] = 0xAB
    inst_buf = np_u8(b"".join(int(x).to_bytes(32, "little") for col in insts for x in col[0]))
    out_acc = np.zeros((n, 128), dtype=np.uint8)
    out_ch = np.zeros((n, n_ch, 32), dtype=np.uint8)
    out_st = np.full(n, -7, dtype=np.int32)
    rc = L.svk_plonk_succinct_verify_batch(c, pid, n, ptr(inst_buf), 1, ptr(buf), stride, None, ptr(out_acc), ptr(out_ch), ptr(out_st))
    assert rc == 0, L.svk_last_error(c)
    for i in range(n):
        st, acc, ch = oracle_result(S, insts[i], bytes(proofs[i]), scheme)
        assert (out_st[i] & 0xFF) == st, (i, out_st[i], st)
        if st == 0:
            got = (g1_from(out_acc[i, :64].tobytes()), g1_from(out_acc[i, 64:].tobytes()))
            assert got == acc, i
            got_ch = [int.from_bytes(out_ch[i, j].tobytes(), "little") for j in range(n_ch)]
            assert got_ch == ch, i
        else:
            assert not out_acc[i].any()
    # short proofs -> Transcript(EOF); wrong instance count -> InvalidInstances for all
    lens = np.full(n, plen, dtype=np.uint32)
    lens[2] = plen - 1
    lens[4] = 40
    rc = L.svk_plonk_succinct_verify_batch(c, pid, n, ptr(inst_buf), 1, ptr(buf), stride, ptr(lens), ptr(out_acc), ptr(out_ch), ptr(out_st))
    assert rc == 0
    assert (out_st[2] & 0xFF) == 4 and (out_st[2] >> 8) == 1 and (out_st[4] & 0xFF) == 4 and out_st[0] == 0
    inst2 = np.zeros(n * 2 * 32, dtype=np.uint8)
    rc = L.svk_plonk_succinct_verify_batch(c, pid, n, ptr(inst2), 2, ptr(buf), stride, None, ptr(out_acc), ptr(out_ch), ptr(out_st))
    assert rc == 0 and (out_st == 1).all()


@pytest.mark.gpu
def test_instance_columns_and_length_clamp(env):
    """proof.rs:66-69 compares the instance COLUMNS with protocol.num_instance ([[x], []] has the right flat count and the wrong
    shape); a proof length beyond the row stride is clamped instead of reading the neighbour's bytes."""
    from snark_verifier_axiom_b200 import verifier as V
    from snark_verifier_axiom_b200.standard_plonk import standard_plonk_protocol

    L, c, S, _kid, _blob = env
    ctx = V.Context(0)
    dk = V.KzgDecidingKey.new(S.dk.svk.g, S.dk.g2, S.dk.s_g2)
    pv = V.PlonkVerifier(ctx, dk, standard_plonk_protocol(8, S.preprocessed, S.transcript_initial_state), V.SHPLONK)
    insts, proofs = forge.forge_batch(S, "bdfg21", 3, seed0=4242)
    good = [V.Snark(i, p) for i, p in zip(insts, proofs)]
    bad = V.Snark([insts[1][0], []], proofs[1])  # two columns [1, 0] against num_instance [1]
    accs, _, st = pv.succinct_verify([good[0], bad, good[2]])
    assert list(st) == [0, 1, 0] and accs[1] is None
    assert api.status_of(api.verify, S.dk, S.protocol, [insts[1][0], []], proofs[1], "bdfg21") == 1
    res = pv.verify([good[0], bad, good[2]], group_size=0)
    assert not res.ok and list(res.status) == [0, 1, 0]
    assert pv.verify(good, group_size=0).ok
    # lengths beyond the stride: identical to the stride itself
    inst, n_inst, buf, lens = pv.pack(good)
    out_acc, out_st = np.zeros((2, 3, 128), np.uint8), np.zeros((2, 3), np.int32)
    n_ch = pv.info["n_challenges"]
    ch = np.zeros((3, n_ch, 32), np.uint8)
    for k, ln in enumerate((lens, lens + 1000)):
        rc = L.svk_plonk_succinct_verify_batch(ctx._c, pv.pid, 3, ptr(inst), n_inst, ptr(buf), buf.shape[1], ptr(ln.astype(np.uint32)), ptr(out_acc[k]),
                                               ptr(ch), ptr(out_st[k]))
        assert rc == 0
    assert (out_st == 0).all() and (out_acc[0] == out_acc[1]).all()
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("schedule", [0, 1])
def test_poseidon_squeeze_hook(env, schedule):
    """svk_poseidon_squeeze (SURVEY 8b test hook): `Poseidon::update(..); squeeze()` (util/hash/poseidon.rs:448-467) for sponges of
    0..7 inputs, through the one-thread permutation and through the warp-cooperative one (poseidon_coop.cuh)."""
    import random

    from oracle import poseidon
    from oracle.bn254 import R

    L, c, S, _kid, _blob = env
    rng = random.Random(11 + schedule)
    for n_in in range(0, 8):
        n = 70  # more than two blocks of 32, with a ragged tail
        vals = [[rng.randrange(R) for _ in range(n_in)] for _ in range(n)]
        if n_in:
            vals[0] = [0] * n_in
            vals[1] = [R - 1] * n_in
        buf = np_u8(b"".join(int(x).to_bytes(32, "little") for row in vals for x in row) or b"\0")
        out = np.zeros((n, 32), np.uint8)
        assert L.svk_poseidon_squeeze(c, n, ptr(buf), n_in, schedule, ptr(out)) == 0, L.svk_last_error(c)
        for i in range(n):
            h = poseidon.Poseidon()
            h.update(vals[i])
            assert int.from_bytes(out[i].tobytes(), "little") == h.squeeze(), (n_in, i)
    bad = np_u8((R).to_bytes(32, "little"))
    out = np.zeros((1, 32), np.uint8)
    assert L.svk_poseidon_squeeze(c, 1, ptr(bad), 1, schedule, ptr(out)) != 0
