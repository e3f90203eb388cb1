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
