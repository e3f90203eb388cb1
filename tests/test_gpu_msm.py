"""GPU parity: svk_msm_g1 (Pippenger) and svk_g1_mul_batch vs the oracle's naive
`NativeLoader::multi_scalar_multiplication` (snark-verifier/src/loader/native.rs:61-71) on small sizes, and
size-independent properties (linearity, split-and-add) at 2^16 points."""
import random

import numpy as np
import pytest

from oracle import bn254
from oracle.forge import g_mul

pytestmark = pytest.mark.gpu
R = bn254.R


@pytest.fixture(scope="module")
def ctx():
    from snark_verifier_axiom_b200 import verifier as V

    c = V.Context(0)
    yield V, c
    c.close()


def test_g1_mul_batch(ctx):
    V, c = ctx
    rng = random.Random(1)
    ks = [0, 1, 2, R - 1, R - 2] + [rng.randrange(R) for _ in range(40)]
    base = g_mul(rng.randrange(1, R))
    got = V.g1_mul_batch(c, ks, [base])
    assert got == [bn254.g1_mul(base, k) for k in ks]
    bases = [g_mul(rng.randrange(1, R)) for _ in range(5)] + [None]
    ks = [rng.randrange(R) for _ in range(12)]
    assert V.g1_mul_batch(c, ks, bases) == [bn254.g1_mul(bases[i % 6], k) for i, k in enumerate(ks)]


@pytest.mark.parametrize("n", [0, 1, 2, 33, 257, 1500])
def test_msm_matches_naive(ctx, n):
    V, c = ctx
    rng = random.Random(100 + n)
    pts = V.g1_mul_batch(c, [rng.randrange(1, R) for _ in range(n)], [bn254.G1_GEN]) if n else []
    sc = [rng.randrange(R) for _ in range(n)]
    for i, s in enumerate([0, 1, R - 1, (1 << 253) + 5, (1 << 16) - 1, 1 << 15]):
        if i < n:
            sc[i] = s
    if n > 10:
        pts[7] = None  # identity base
        pts[9] = pts[8]  # repeated base
        sc[9] = (R - sc[8]) % R  # ... cancelling
    exp = None
    for s, p in zip(sc, pts):
        exp = bn254.g1_add(exp, bn254.g1_mul(p, s))
    assert V.multi_scalar_multiplication(c, sc, pts) == exp


def test_msm_properties_large(ctx):
    V, c = ctx
    n = 1 << 16
    rng = random.Random(7)
    dl = [rng.randrange(1, R) for _ in range(n)]
    pts = V.g1_mul_batch(c, dl, [bn254.G1_GEN])
    a = [rng.randrange(R) for _ in range(n)]
    b = [rng.randrange(R) for _ in range(n)]
    # in the exponent: MSM(a) = (sum a_i d_i) G
    exp_a = g_mul(sum(x * d for x, d in zip(a, dl)) % R)
    ra = V.multi_scalar_multiplication(c, a, pts)
    assert ra == exp_a
    rb = V.multi_scalar_multiplication(c, b, pts)
    rab = V.multi_scalar_multiplication(c, [(x + y) % R for x, y in zip(a, b)], pts)
    assert bn254.g1_add(ra, rb) == rab  # linearity
    h = n // 2
    assert bn254.g1_add(V.multi_scalar_multiplication(c, a[:h], pts[:h]), V.multi_scalar_multiplication(c, a[h:], pts[h:])) == ra
    # the fold's scalar distribution: powers of r
    r = rng.randrange(R)
    pw, cur = [], 1
    for _ in range(4096):
        pw.append(cur)
        cur = cur * r % R
    assert V.multi_scalar_multiplication(c, pw, pts[:4096]) == g_mul(sum(x * d for x, d in zip(pw, dl)) % R)


def test_msm_rejects_bad_input(ctx):
    V, c = ctx
    with pytest.raises(ValueError):
        V.multi_scalar_multiplication(c, [1, 2], [bn254.G1_GEN, (1, 3)])
    with pytest.raises(ValueError):
        V.multi_scalar_multiplication(c, [R, 2], [bn254.G1_GEN, bn254.G1_GEN])


def test_msm_skewed_scalars_heavy_buckets(ctx):
    """Every point in ONE bucket (all scalars 1), in a handful (scalars < 8), and one huge repeated scalar: the heavy-bucket path
    (k_msm_heavy: chunks of a bucket summed by warps, msm.cu) gives the naive sum (native.rs:61-71)."""
    V, c = ctx
    n = 1 << 16
    rng = random.Random(21)
    dl = [rng.randrange(1, R) for _ in range(n)]
    pts = V.g1_mul_batch(c, dl, [bn254.G1_GEN])
    assert V.multi_scalar_multiplication(c, [1] * n, pts) == g_mul(sum(dl) % R)
    small = [rng.randrange(8) for _ in range(n)]
    assert V.multi_scalar_multiplication(c, small, pts) == g_mul(sum(x * d for x, d in zip(small, dl)) % R)
    k = rng.randrange(R)
    assert V.multi_scalar_multiplication(c, [k] * n, pts) == g_mul(k * sum(dl) % R)
    assert V.multi_scalar_multiplication(c, [R - 1] * 5000, pts[:5000]) == g_mul((R - 1) * sum(dl[:5000]) % R)
