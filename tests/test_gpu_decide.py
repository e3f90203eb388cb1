"""GPU parity: svk_kzg_decide_batch vs the oracle's `KzgAs::decide`
(snark-verifier/src/pcs/kzg/decider.rs:60-81)."""
import ctypes
import random

import numpy as np
import pytest

from oracle import bn254
from oracle.forge import g_mul
from oracle.kzg import KzgDecidingKey

from .util import acc_bytes, dk_bytes, np_u8, ptr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from snark_verifier_axiom_b200._lib import lib

    L = lib()
    c = ctypes.c_void_p()
    rc = L.svk_create(0, ctypes.byref(c))
    assert rc == 0, L.svk_last_error(None)
    yield L, c
    L.svk_destroy(c)


def test_modmul_peak(ctx):
    L, c = ctx
    rate, ms = ctypes.c_double(), ctypes.c_double()
    assert L.svk_bench_modmul_peak(c, 2000, ctypes.byref(rate), ctypes.byref(ms)) == 0
    print(f"modmul peak: {rate.value/1e9:.2f} G modmul/s in {ms.value:.3f} ms")
    assert rate.value > 1e9


def test_decide_matches_oracle(ctx):
    L, c = ctx
    rng = random.Random(5)
    s = rng.randrange(1, bn254.R)
    dk = KzgDecidingKey.new(bn254.G1_GEN, bn254.G2_GEN, bn254.g2_mul(bn254.G2_GEN, s))
    kid = L.svk_dk_load(c, dk_bytes(dk))
    assert kid >= 0, L.svk_last_error(c)
    accs, expect = [], []
    n = 150
    for i in range(n):
        d = rng.randrange(1, bn254.R)
        lhs, rhs = g_mul(s * d % bn254.R), g_mul(d)
        kind = i % 6
        if kind == 1:
            lhs = g_mul((s * d + 1) % bn254.R)
        elif kind == 2:
            rhs = g_mul(d + 1)
        elif kind == 3:
            lhs, rhs = None, None  # e(O,.)e(O,.) = 1
        elif kind == 4:
            lhs = None
        accs.append((lhs, rhs))
        expect.append(1 if kind in (0, 3, 5) else 0)
    # spot-check the constructed expectations against the oracle pairing itself
    for i in range(0, 12):
        ok = bn254.pairing_check([(accs[i][0], dk.g2), (accs[i][1], bn254.g2_neg(dk.s_g2))])
        assert int(ok) == expect[i]
    buf = np_u8(b"".join(acc_bytes(l, r) for l, r in accs))
    out = np.zeros(n, dtype=np.uint8)
    assert L.svk_kzg_decide_batch(c, kid, n, ptr(buf), ptr(out)) == 0, L.svk_last_error(c)
    assert out.tolist() == expect
    # off-curve / non-canonical inputs are rejected, never crash
    bad = bytearray(acc_bytes(*accs[0]))
    bad[0] ^= 1
    bad2 = bytes([0xFF] * 128)
    buf = np_u8(bytes(bad) + bad2)
    out = np.ones(2, dtype=np.uint8)
    assert L.svk_kzg_decide_batch(c, kid, 2, ptr(buf), ptr(out)) == 0
    assert out.tolist() == [0, 0]


def test_public_precompile_vectors_on_the_gpu(ctx):
    """EIP-196 / EIP-197 client test vectors (tests/golden/eip196_197_vectors.py) through the C ABI: values from outside this
    repository pin the kernels' group law (batched `base * scalar`, Pippenger) and BOTH pairing kernels -- e(P1, Q1) e(P2, Q2) = 1
    is `decide((P1, P2))` under the key (g2, s_g2) = (Q1, -Q2) (decider.rs:60-68)."""
    from .golden import eip196_197_vectors as E

    L, c = ctx
    le = lambda v: v.to_bytes(32, "little")  # noqa: E731
    xy = lambda o: (int.from_bytes(o[:32].tobytes(), "little"), int.from_bytes(o[32:64].tobytes(), "little"))  # noqa: E731
    # ecMul through svk_g1_mul_batch, ecAdd / ecMul through svk_msm_g1
    out = np.zeros(64, dtype=np.uint8)
    sc, pt = np_u8(le(E.MUL_K)), np_u8(le(E.MUL_P[0]) + le(E.MUL_P[1]))
    assert L.svk_g1_mul_batch(c, 1, ptr(sc), ptr(pt), 1, ptr(out)) == 0, L.svk_last_error(c)
    assert xy(out) == E.MUL_Q
    st = np.zeros(1, dtype=np.int32)
    assert L.svk_msm_g1(c, 1, ptr(sc), ptr(pt), ptr(out), ptr(st)) == 0 and st[0] == 0 and xy(out) == E.MUL_Q
    sc2 = np_u8(le(1) + le(1))
    pt2 = np_u8(b"".join(le(v) for v in (*E.ADD_A, *E.ADD_B)))
    assert L.svk_msm_g1(c, 2, ptr(sc2), ptr(pt2), ptr(out), ptr(st)) == 0 and st[0] == 0 and xy(out) == E.ADD_C
    # ecPairing through k_decide_coop (few accumulators) and the batched k_decide (more than decide_coop_max = 512)
    dk = KzgDecidingKey.new(bn254.G1_GEN, E.PAIR_Q1, bn254.g2_neg(E.PAIR_Q2))
    kid = L.svk_dk_load(c, dk_bytes(dk))
    assert kid >= 0, L.svk_last_error(c)
    good = acc_bytes(E.PAIR_P1, E.PAIR_P2)
    bad1 = acc_bytes(E.PAIR_P1, bn254.g1_neg(E.PAIR_P2))
    bad2 = acc_bytes(E.ADD_A, E.PAIR_P2)
    for reps in (1, 300):
        buf = np_u8((good + bad1 + bad2) * reps)
        got = np.full(3 * reps, 7, dtype=np.uint8)
        assert L.svk_kzg_decide_batch(c, kid, 3 * reps, ptr(buf), ptr(got)) == 0, L.svk_last_error(c)
        assert got.tolist() == [1, 0, 0] * reps
