"""The C restatement of the reference CPU path (oracle/c) is bit-identical to the Python oracle:
accumulators, statuses, folds (flat + tree), decide."""
import numpy as np
import pytest

from oracle import api, bn254, forge
from oracle.c import cref

from .util import acc_bytes, g1_from


@pytest.fixture(scope="module")
def S():
    return forge.Setup(0)


@pytest.mark.parametrize("scheme", ["bdfg21", "gwc19"])
def test_replay_matches_python_oracle(S, scheme):
    tr = cref.Trace(S, scheme)
    if scheme == "bdfg21":
        assert tr.n_scalar_muls == 21 and tr.counts["inv"] == 19 and tr.counts["t_read_point"] == 11
    n = 6
    insts, proofs = forge.forge_batch(S, scheme, n, seed0=70)
    proofs = [bytearray(p) for p in proofs]
    proofs[1][9 * 32 + 7] ^= 2           # still decodes: different accumulator
    proofs[2][9 * 32 : 10 * 32] = b"\xff" * 32  # scalar >= r
    proofs[3][0:32] = bytes(32)          # identity point
    proofs = [bytes(p) for p in proofs]
    accs, st = cref.replay(tr, proofs, insts, threads=2)
    pairs = []
    for i in range(n):
        want = api.status_of(api.succinct_verify, S.dk.svk, S.protocol, insts[i], proofs[i], scheme)
        assert (st[i] & 0xFF) == want
        if want == 0:
            a = api.succinct_verify(S.dk.svk, S.protocol, insts[i], proofs[i], scheme)[0]
            assert (g1_from(accs[i, :64].tobytes()), g1_from(accs[i, 64:].tobytes())) == (a.lhs.pt, a.rhs.pt)
            pairs.append((a.lhs.pt, a.rhs.pt))
    good = np.stack([accs[i] for i in range(n) if st[i] == 0])
    for m in (0, 2):
        acc, r, fst = cref.fold(good, m)
        (el, er), rs = api.fold(pairs, m)
        assert fst == 0 and (g1_from(acc[:64].tobytes()), g1_from(acc[64:].tobytes())) == (el, er) and r == rs[-1]
    acc, r, _ = cref.fold(np.stack([accs[0], accs[4], accs[5]]), 0)
    assert cref.decide(acc, S.dk) is True
    acc_bad, _, _ = cref.fold(good, 0)  # contains the mutated proof's accumulator
    assert cref.decide(acc_bad, S.dk) is False
    assert cref.decide(np.frombuffer(acc_bytes(None, None), dtype=np.uint8), S.dk) is True


def test_c_pippenger_matches_naive_msm():
    """oracle/c `cref_msm` (C restatement of util/msm.rs:238-317, the reference's Pippenger incl. its chunking over threads) against
    the naive sum of scalar multiplications (loader/native.rs:61-71) on the Python oracle."""
    import random

    rng = random.Random(3)
    n = 150
    pts = [bn254.g1_mul(bn254.G1_GEN, rng.randrange(1, bn254.R)) for _ in range(n)]
    pts[5] = None
    pts[9] = pts[8]
    sc = [rng.randrange(bn254.R) for _ in range(n)]
    sc[0], sc[1], sc[9] = 0, bn254.R - 1, (bn254.R - sc[8]) % bn254.R
    S_ = np.frombuffer(b"".join(x.to_bytes(32, "little") for x in sc), np.uint8).reshape(n, 32)
    P_ = np.frombuffer(b"".join(acc_bytes(p, None)[:64] for p in pts), np.uint8).reshape(n, 64)
    exp = None
    for s_, p in zip(sc, pts):
        exp = bn254.g1_add(exp, bn254.g1_mul(p, s_))
    for th in (1, 4):
        assert g1_from(cref.msm(S_, P_, th).tobytes()) == exp
    assert g1_from(cref.msm(S_[:1], P_[:1], 8).tobytes()) is None  # 0 * P


def test_public_precompile_vectors_through_the_c_oracle():
    """EIP-196 / EIP-197 client test vectors (tests/golden/eip196_197_vectors.py): the C restatement's scalar multiplication, Pippenger
    and pairing against values from outside this repository.  e(P1, Q1) e(P2, Q2) = 1 is `decide((P1, P2))` under the key
    (g2, s_g2) = (Q1, -Q2) (decider.rs:60-68)."""
    from types import SimpleNamespace

    from oracle import bn254

    from .golden import eip196_197_vectors as E

    le = lambda v: v.to_bytes(32, "little")  # noqa: E731
    acc = lambda l, r: np.frombuffer(le(l[0]) + le(l[1]) + le(r[0]) + le(r[1]), np.uint8).copy()  # noqa: E731
    dk = SimpleNamespace(g2=E.PAIR_Q1, s_g2=bn254.g2_neg(E.PAIR_Q2))
    assert cref.decide(acc(E.PAIR_P1, E.PAIR_P2), dk) is True
    assert cref.decide(acc(E.PAIR_P1, bn254.g1_neg(E.PAIR_P2)), dk) is False
    assert cref.decide(acc(E.ADD_A, E.PAIR_P2), dk) is False
    sc = lambda ks: np.frombuffer(b"".join(le(k) for k in ks), np.uint8).copy()  # noqa: E731
    pt = lambda ps: np.frombuffer(b"".join(le(x) + le(y) for x, y in ps), np.uint8).copy()  # noqa: E731
    xy = lambda o: (int.from_bytes(o[:32].tobytes(), "little"), int.from_bytes(o[32:].tobytes(), "little"))  # noqa: E731
    assert xy(cref.msm(sc([1, 1]), pt([E.ADD_A, E.ADD_B]))) == E.ADD_C
    assert xy(cref.msm(sc([E.MUL_K]), pt([E.MUL_P]))) == E.MUL_Q
