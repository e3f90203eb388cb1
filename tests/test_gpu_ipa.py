"""GPU parity: the Pippenger kernels instantiated over the Pasta curves vs the oracle's naive
`multi_scalar_multiplication` and its restatement of util/msm.rs:238-317, and `IpaAs::decide` / `decide_all`
(snark-verifier/src/pcs/ipa/decider.rs:47-67) vs oracle/ipa.py -- accumulators produced by an honest prover's base folding
(pcs/ipa.rs:78-118), corrupted ones, and inputs the reference's types cannot hold.  All through the C ABI."""
import random

import numpy as np
import pytest

from oracle import ipa, pasta

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from snark_verifier_axiom_b200 import verifier as V

    c = V.Context(0)
    yield V, c
    c.close()


CURVES = [pasta.PALLAS, pasta.VESTA]


@pytest.mark.parametrize("C", CURVES, ids=["pallas", "vesta"])
@pytest.mark.parametrize("n", [0, 1, 2, 37, 300])
def test_msm_on_pasta_matches_naive(ctx, C, n):
    V, c = ctx
    rng = random.Random(1000 * C.id + n)
    pts = [C.mul(C.gen, rng.randrange(1, C.n)) for _ in range(n)]
    ks = [rng.randrange(C.n) for _ in range(n)]
    if n >= 37:
        ks[0], ks[1], ks[2], ks[3] = 0, 1, C.n - 1, C.n - 2  # 255-bit scalars: the top signed window must not overflow
        pts[5] = None
        pts[7] = pts[6]
        pts[9] = C.neg(pts[8])
        ks[9] = ks[8]
    assert V.multi_scalar_multiplication_on(c, C.id, ks, pts) == C.msm_naive(ks, pts)


@pytest.mark.parametrize("C", CURVES, ids=["pallas", "vesta"])
@pytest.mark.parametrize("n", [(1 << 13) + 3, (1 << 18) + 1])
def test_msm_on_pasta_large_windows(ctx, C, n):
    """c = 15 and c = 16 window plans: n points drawn from a pool of 48 distinct ones => expected = sum_pool (sum of its scalars) * point."""
    V, c = ctx
    rng = random.Random(77 + n + C.id)
    pool = [C.mul(C.gen, rng.randrange(1, C.n)) for _ in range(48)]
    idx = [rng.randrange(48) for _ in range(n)]
    ks = [rng.randrange(C.n) for _ in range(n)]
    ks[:6] = [C.n - 1, C.n - 1, 0, 1, (1 << 254) - 1, 1 << 254]
    tot = [0] * 48
    for i, k in zip(idx, ks):
        tot[i] = (tot[i] + k) % C.n
    assert V.multi_scalar_multiplication_on(c, C.id, ks, [pool[i] for i in idx]) == C.msm_naive(tot, pool)


def test_msm_curve_bn254_equals_msm_g1(ctx):
    V, c = ctx
    C = pasta.bn254_g1()
    rng = random.Random(9)
    pts = [C.mul(C.gen, rng.randrange(1, C.n)) for _ in range(200)]
    ks = [rng.randrange(C.n) for _ in range(200)]
    assert V.multi_scalar_multiplication_on(c, V.CURVE_BN254_G1, ks, pts) == V.multi_scalar_multiplication(c, ks, pts) == C.msm_naive(ks, pts)


def test_msm_on_pasta_rejects_unrepresentable_inputs(ctx):
    V, c = ctx
    C = pasta.PALLAS
    with pytest.raises(ValueError):
        V.multi_scalar_multiplication_on(c, C.id, [C.n], [C.gen])  # scalar >= modulus
    with pytest.raises(ValueError):
        V.multi_scalar_multiplication_on(c, C.id, [3], [(C.gen[0], 3)])  # off the curve
    # a Vesta point is not a Pallas point
    with pytest.raises(ValueError):
        V.multi_scalar_multiplication_on(c, C.id, [3], [pasta.VESTA.mul(pasta.VESTA.gen, 7)])


@pytest.mark.parametrize("C", CURVES, ids=["pallas", "vesta"])
@pytest.mark.parametrize("k", [1, 2, 6])
def test_ipa_decide_matches_oracle(ctx, C, k):
    V, c = ctx
    rng = random.Random(31 * k + C.id)
    g = [C.mul(C.gen, rng.randrange(1, C.n)) for _ in range(1 << k)]
    accs, want = [], []
    for i in range(7):
        xi = [rng.randrange(1, C.n) for _ in range(k)]
        u = ipa.fold_bases(C, g, xi)  # the honest prover's final base == commit(G, h_coeffs(xi))
        if i == 2:
            u = C.add(u, C.gen)
        if i == 3 and k > 1:
            xi = xi[::-1]
        if i == 4:
            u = None
        if i == 5:
            xi[0] = 0  # h has zeros: commit = sum over the coefficients without that factor
            u = C.msm_naive(ipa.h_coeffs(xi, 1, C.n), g)
        accs.append(V.IpaAccumulator(xi, u))
        want.append(ipa.decide(C, g, ipa.IpaAccumulator(xi, u)))
    dk = V.IpaDecidingKey(g, C.id)
    st = V.IpaAs.decide_batch(c, dk, accs)
    assert st.tolist() == want
    assert want[0] == 0 and want[2] == 3 and want[4] == 3 and want[5] == 0
    V.IpaAs.decide(c, dk, accs[0])
    with pytest.raises(V.Error) as e:
        V.IpaAs.decide(c, dk, accs[2])
    assert e.value.kind == "AssertionFailure"
    V.IpaAs.decide_all(c, dk, [accs[0], accs[1]])
    with pytest.raises(V.Error):
        V.IpaAs.decide_all(c, dk, accs)
    assert ipa.decide_all(C, g, [ipa.IpaAccumulator(a.xi, a.u) for a in accs]) == 3


def test_ipa_decide_k10_like_the_reference_test(ctx):
    """pcs/ipa.rs:407-446 runs k = 10 over pallas; the oracle side uses its Pippenger restatement (1024 bases)."""
    V, c = ctx
    C = pasta.PALLAS
    rng = random.Random(10)
    k = 10
    # bases d_i * G with known d_i => commit(G, h) = (sum h_i d_i) * G, an exact cross-check of both MSMs
    g, dl = [], []
    # build the bases incrementally (cheaper than 1024 independent scalar multiplications): g_i = g_{i-1} + delta * G
    step = C.mul(C.gen, 0x9E3779B97F4A7C15F39CC0605CEDC834)
    dcur = rng.randrange(1, C.n)
    cur = C.mul(C.gen, dcur)
    for i in range(1 << k):
        g.append(cur)
        dl.append(dcur)
        cur = C.add(cur, step)
        dcur = (dcur + 0x9E3779B97F4A7C15F39CC0605CEDC834) % C.n
    xi = [rng.randrange(1, C.n) for _ in range(k)]
    h = ipa.h_coeffs(xi, 1, C.n)
    u = C.mul(C.gen, sum(a * b for a, b in zip(h, dl)) % C.n)
    assert C.msm_pippenger(h, g) == u
    dk = V.IpaDecidingKey(g, C.id)
    st = V.IpaAs.decide_batch(c, dk, [V.IpaAccumulator(xi, u), V.IpaAccumulator(xi, C.neg(u)), V.IpaAccumulator(xi[1:] + xi[:1], u)])
    assert st.tolist() == [0, 3, 3]
    assert V.multi_scalar_multiplication_on(c, C.id, h, g) == u


def test_ipa_decide_unrepresentable_inputs_fail(ctx):
    V, c = ctx
    C = pasta.PALLAS
    rng = random.Random(4)
    k = 3
    g = [C.mul(C.gen, rng.randrange(1, C.n)) for _ in range(1 << k)]
    xi = [rng.randrange(1, C.n) for _ in range(k)]
    u = ipa.fold_bases(C, g, xi)
    dk = V.IpaDecidingKey(g, C.id)
    assert V.IpaAs.decide_batch(c, dk, [V.IpaAccumulator(xi, u)]).tolist() == [0]
    # xi >= modulus (congruent to a valid challenge): halo2curves could not even construct it
    assert V.IpaAs.decide_batch(c, dk, [V.IpaAccumulator([xi[0] + C.n] + xi[1:], u)]).tolist() == [3]
    g_bad = list(g)
    g_bad[3] = (g[3][0], (g[3][1] + 1) % C.p)
    assert V.IpaAs.decide_batch(c, V.IpaDecidingKey(g_bad, C.id), [V.IpaAccumulator(xi, u)]).tolist() == [3]
    # the same key/accumulator under the wrong curve id
    assert V.IpaAs.decide_batch(c, V.IpaDecidingKey(g, pasta.VESTA.id), [V.IpaAccumulator(xi, u)]).tolist() == [3]


@pytest.mark.parametrize("zk", [False, True])
def test_ipa_and_ipa_as_like_the_reference_tests(ctx, zk):
    """`test_ipa` (pcs/ipa.rs:407-446) and `test_ipa_as` (pcs/ipa/accumulation.rs:212-280) with the NativeLoader pieces on the GPU:
    commit / the two `Msm::evaluate`s of `succinct_verify` through svk_msm_curve, `IpaAs::decide` through svk_ipa_decide_batch;
    proving and the transcript stay in the oracle (prover-side code is out of scope)."""
    V, c = ctx
    C = pasta.PALLAS
    rng = random.Random(2024 + zk)
    k = 4
    pk = ipa.IpaProvingKey.rand(C, k, zk, rng)
    gpu_msm = lambda scalars, points: V.multi_scalar_multiplication_on(c, C.id, [s % C.n for s in scalars], points)  # noqa: E731
    dk = V.IpaDecidingKey(pk.g, C.id)
    accs = []
    for _ in range(3):
        p = [rng.randrange(C.n) for _ in range(1 << k)]
        omega = rng.randrange(C.n) if zk else None
        com = pk.commit(p, omega, msm=gpu_msm)
        assert com == pk.commit(p, omega)
        z = rng.randrange(C.n)
        v = ipa.poly_eval(p, z, C.n)
        tw = ipa.HashTranscript(C)
        ipa.ipa_create_proof(pk, p, z, omega, tw, rng)
        proof = ipa.IpaProof.read(zk, k, ipa.HashTranscript(C, tw.finalize()))
        acc = ipa.ipa_succinct_verify(pk, [(1, com)], z, v, proof, msm=gpu_msm)  # asserts C_k == c[U] + v'[H'] on device results
        V.IpaAs.decide(c, dk, V.IpaAccumulator(acc.xi, acc.u))  # assert!(IpaAs::decide(&dk, accumulator).is_ok())
        accs.append(acc)
    tw = ipa.HashTranscript(C)
    ipa.ipa_as_create_proof(pk, accs, tw, rng)
    folded = ipa.ipa_as_verify(pk, accs, ipa.HashTranscript(C, tw.finalize()), msm=gpu_msm)
    V.IpaAs.decide(c, dk, V.IpaAccumulator(folded.xi, folded.u))
    V.IpaAs.decide_all(c, dk, [V.IpaAccumulator(a.xi, a.u) for a in accs + [folded]])
    with pytest.raises(V.Error):
        V.IpaAs.decide(c, dk, V.IpaAccumulator(folded.xi, C.add(folded.u, C.gen)))


def test_ipa_decide_against_committed_golden(ctx):
    """Device vs tests/golden/ipa_golden.json: decide statuses and the commitment MSM for every fixture accumulator."""
    from .test_oracle_ipa import _load_ipa_golden

    V, c = ctx
    for C, k, g, accs in _load_ipa_golden():
        dk = V.IpaDecidingKey(g, C.id)
        st = V.IpaAs.decide_batch(c, dk, [V.IpaAccumulator(xi, u) for xi, u, _, _ in accs])
        assert st.tolist() == [s for _, _, _, s in accs]
        for xi, _, commit, _ in accs[:2]:
            assert V.multi_scalar_multiplication_on(c, C.id, ipa.h_coeffs(xi, 1, C.n), g) == commit
