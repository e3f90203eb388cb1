"""Exact-integer self-checks of the oracle's BN254 restatement: constants of halo2curves,
group law, compressed encoding round trip, pairing bilinearity / non-degeneracy, KZG decide."""
import random

from oracle import bn254 as B


def test_constants():
    assert B.FR_ROOT_OF_UNITY == 0x03DDB9F5166D18B798865EA93DD31F743215CF6DD39329C8D34F1ED960C37C9C
    assert B.FR_DELTA == 0x09226B6E22C6F0CA64EC26AAD4C86E715B5F898E5E963F25870E56BBE533E9A2
    assert pow(B.FR_ROOT_OF_UNITY, 1 << 28, B.R) == 1 and pow(B.FR_ROOT_OF_UNITY, 1 << 27, B.R) != 1
    assert B.ATE_LOOP == 0x19D797039BE763BA8
    assert B.g1_is_on_curve(B.G1_GEN) and B.g2_is_on_curve(B.G2_GEN)
    assert B.g2_mul(B.G2_GEN, B.R) is None and B.g1_mul(B.G1_GEN, B.R - 1) == B.g1_neg(B.G1_GEN)


def test_public_known_answers():
    """Values that are public knowledge about alt_bn128 / BN254 (EIP-196/197 precompile documentation): they pin the
    curve equation, the generator and the group law of the restatement independently of this repository."""
    two_g = (1368015179489954701390400359078579693043519447331113978918064868415326638035,
             9918110051302171585080402603319702774565515993150576347155970296011118125764)
    assert B.g1_mul(B.G1_GEN, 2) == two_g and B.g1_add(B.G1_GEN, B.G1_GEN) == two_g
    assert B.P == 21888242871839275222246405745257275088696311157297823662689037894645226208583
    assert B.R == 21888242871839275222246405745257275088548364400416034343698204186575808495617
    # EIP-197 G2 generator (x = x_c0 + x_c1 u as listed there: imaginary part first in the ABI)
    assert B.G2_GEN[0][1] == 11559732032986387107991004021392285783925812861821192530917403151452391805634
    assert B.G2_GEN[0][0] == 10857046999023057135944570762232829481370756359578518086990519993285655852781


def test_public_precompile_vectors():
    """EIP-196 / EIP-197 client test vectors (tests/golden/eip196_197_vectors.py): external pins of the group law and the pairing."""
    from .golden import eip196_197_vectors as E

    for pt in (E.ADD_A, E.ADD_B, E.ADD_C, E.MUL_P, E.MUL_Q, E.PAIR_P1, E.PAIR_P2):
        assert B.g1_is_on_curve(pt)
    assert B.g1_add(E.ADD_A, E.ADD_B) == E.ADD_C
    assert B.g1_mul(E.MUL_P, E.MUL_K) == E.MUL_Q and B.g1_msm_naive([(E.MUL_K, E.MUL_P)]) == E.MUL_Q
    assert B.g2_is_on_curve(E.PAIR_Q1) and E.PAIR_Q2 == B.G2_GEN and B.g2_mul(E.PAIR_Q1, B.R) is None
    assert B.pairing_check([(E.PAIR_P1, E.PAIR_Q1), (E.PAIR_P2, E.PAIR_Q2)])
    assert not B.pairing_check([(E.PAIR_P1, E.PAIR_Q1), (B.g1_neg(E.PAIR_P2), E.PAIR_Q2)])
    assert not B.pairing_check([(E.ADD_A, E.PAIR_Q1), (E.PAIR_P2, E.PAIR_Q2)])


def test_g1_encoding_roundtrip():
    rng = random.Random(1)
    for _ in range(20):
        p = B.g1_mul(B.G1_GEN, rng.randrange(1, B.R))
        b = B.g1_to_bytes(p)
        assert B.g1_from_bytes(b) == (True, p)
        nb = bytearray(b)
        nb[31] ^= 0x80
        assert B.g1_from_bytes(bytes(nb)) == (True, B.g1_neg(p))
    assert B.g1_from_bytes(bytes(32)) == (True, None)
    assert B.g1_from_bytes(B.P.to_bytes(32, "little"))[0] is False
    assert B.fr_from_bytes(B.R.to_bytes(32, "little")) is None and B.fr_from_bytes((B.R - 1).to_bytes(32, "little")) == B.R - 1


def test_msm_naive_matches_sum():
    rng = random.Random(2)
    pairs = [(rng.randrange(B.R), B.g1_mul(B.G1_GEN, rng.randrange(1, B.R))) for _ in range(5)]
    acc = None
    for s, p in pairs:
        acc = B.g1_add(acc, B.g1_mul(p, s))
    assert B.g1_msm_naive(pairs) == acc


def test_pairing_bilinear_and_decide():
    rng = random.Random(3)
    a, b = rng.randrange(1, B.R), rng.randrange(1, B.R)
    e = B.pairing(B.G1_GEN, B.G2_GEN)
    assert e != B.F12_ONE and B.f12_pow(e, B.R) == B.F12_ONE
    assert B.pairing(B.g1_mul(B.G1_GEN, a), B.g2_mul(B.G2_GEN, b)) == B.f12_pow(e, a * b % B.R)
    s, d = rng.randrange(1, B.R), rng.randrange(1, B.R)
    sg2 = B.g2_mul(B.G2_GEN, s)
    assert B.pairing_check([(B.g1_mul(B.G1_GEN, s * d % B.R), B.G2_GEN), (B.g1_mul(B.G1_GEN, d), B.g2_neg(sg2))])
    assert not B.pairing_check([(B.g1_mul(B.G1_GEN, s * d % B.R + 1), B.G2_GEN), (B.g1_mul(B.G1_GEN, d), B.g2_neg(sg2))])
