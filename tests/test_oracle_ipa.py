"""IPA decider on CPU: the Pasta oracle's self-checks, the restatement of `h_coeffs` / `decide` against an honest
prover's base folding (pcs/ipa.rs:78-118), and the DEVICE field / point templates built for the host over the Pasta fields."""
import ctypes
import os
import random
import subprocess

import pytest

from oracle import ipa, pasta

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "host", "_hostlib.so")
SRC = os.path.join(HERE, "host", "hostlib.cpp")


@pytest.fixture(scope="module")
def lib():
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", SO, SRC], check=True, timeout=600)
    return ctypes.CDLL(SO)


def limbs(vals):
    out = []
    for v in vals:
        out += [(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)]
    return (ctypes.c_uint32 * len(out))(*out)


def rd(a, n):
    return [sum(int(a[8 * j + i]) << (32 * i) for i in range(8)) for j in range(n)]


def test_pasta_curves_exact_algebra():
    for C in (pasta.PALLAS, pasta.VESTA):
        assert C.is_on_curve(C.gen)
        assert C.mul(C.gen, C.n) is None and C.mul(C.gen, C.n - 1) == C.neg(C.gen)  # prime group order = the other field
        a, b = 0x1234567 << 200, 0xABCDEF << 100
        assert C.add(C.mul(C.gen, a), C.mul(C.gen, b)) == C.mul(C.gen, a + b)
    assert pasta.PALLAS.p == pasta.VESTA.n and pasta.VESTA.p == pasta.PALLAS.n
    bn = pasta.bn254_g1()
    assert bn.mul(bn.gen, bn.n) is None


@pytest.mark.parametrize("C", [pasta.PALLAS, pasta.VESTA], ids=["pallas", "vesta"])
def test_h_coeffs_decide_vs_honest_folding(C):
    rng = random.Random(11 + C.id)
    for k in (1, 3, 5):
        g = [C.mul(C.gen, rng.randrange(1, C.n)) for _ in range(1 << k)]
        xi = [rng.randrange(1, C.n) for _ in range(k)]
        u = ipa.fold_bases(C, g, xi)  # what Ipa::create_proof leaves in bases[0]
        h = ipa.h_coeffs(xi, 1, C.n)
        assert C.msm_naive(h, g) == u == C.msm_pippenger(h, g)
        z = rng.randrange(C.n)
        assert ipa.h_eval(xi, z, C.n) == sum(c * pow(z, i, C.n) for i, c in enumerate(h)) % C.n
        assert ipa.decide(C, g, ipa.IpaAccumulator(xi, u)) == 0
        assert ipa.decide(C, g, ipa.IpaAccumulator(xi, C.add(u, C.gen))) == 3
        if k > 1:
            assert ipa.decide(C, g, ipa.IpaAccumulator(xi[::-1], u)) == 3
        assert ipa.decide_all(C, g, [ipa.IpaAccumulator(xi, u)] * 2) == 0
        assert ipa.decide_all(C, g, [ipa.IpaAccumulator(xi, u), ipa.IpaAccumulator(xi, None)]) == 3


def test_device_field_templates_over_pasta_fields(lib):
    rng = random.Random(2)

    def op(f, o, a, b=0):
        out = (ctypes.c_uint32 * 8)()
        lib.host_pasta_fe_op(f, o, limbs([a]), limbs([b]), out)
        return rd(out, 1)[0]

    for f, m in ((2, pasta.PALLAS_P), (3, pasta.VESTA_P)):
        Rm = (1 << 256) % m
        Ri = pow(Rm, -1, m)
        edge = [0, 1, 2, m - 1, m - 2, Rm, (1 << 254) % m, m >> 1, (1 << 254) - 1, (1 << 254), m - (1 << 32), (1 << 224) - 1]
        vals = edge + [rng.randrange(m) for _ in range(80)]
        for a in vals:
            for b in rng.sample(vals, 6) + edge:
                assert op(f, 0, a, b) == a * b * Ri % m
                assert op(f, 1, a, b) == (a + b) % m and op(f, 2, a, b) == (a - b) % m
            assert op(f, 3, a) == (-a) % m and op(f, 6, a) == a * Rm % m and op(f, 7, a) == a * Ri % m
            assert op(f, 8, a) == a * a * Ri % m
        for a in [rng.randrange(m) for _ in range(1500)] + [m - 1 - rng.randrange(1 << 40) for _ in range(200)]:
            assert op(f, 8, a) == a * a * Ri % m
            b = rng.choice(vals)
            assert op(f, 0, a, b) == a * b * Ri % m
        for a in vals[:14] + [rng.randrange(m) for _ in range(200)]:
            assert op(f, 4, a * Rm % m) == (pow(a, -1, m) * Rm % m if a else 0)  # binary extended Euclid over a 255-bit modulus
        for a in vals[:14]:
            assert op(f, 11, a * Rm % m) == (pow(a, -1, m) * Rm % m if a else 0)  # Fermat: p - 2 needs a borrow across limb 0 here


def test_fused_dot_products_over_pasta_fields(lib):
    """The point formulas use dot2 (Y3 = r (V - X3) - 2 Y J with one reduction); its bound for 255-bit moduli: N <= 3."""
    rng = random.Random(8)
    for f, m in ((2, pasta.PALLAS_P), (3, pasta.VESTA_P)):
        Ri = pow(1 << 256, -1, m)
        edge = [0, 1, m - 1, m, m - 2, (1 << 254) + 5, (1 << 254) - 1]
        for n in (2, 3):
            for trial in range(400):
                src = edge if trial < 80 else None
                a = [rng.choice(src) if src else rng.randrange(m) for _ in range(n)]
                b = [rng.choice(src) if src else rng.randrange(m) for _ in range(n)]
                neg = rng.randrange(1 << n)
                out = (ctypes.c_uint32 * 8)()
                lib.host_pasta_fe_dot(f, n, neg, limbs(a), limbs(b), out)
                want = sum((-x if neg >> k & 1 else x) * y for k, (x, y) in enumerate(zip(a, b))) * Ri % m
                assert rd(out, 1)[0] == want, (f, n, a, b, neg)


def test_device_point_templates_over_pasta_curves(lib):
    rng = random.Random(3)
    for C in (pasta.bn254_g1(), pasta.PALLAS, pasta.VESTA):
        for trial in range(8):
            P = C.mul(C.gen, rng.randrange(1, C.n))
            Q = C.mul(C.gen, rng.randrange(1, C.n)) if trial != 1 else None
            k = rng.randrange(C.n) if trial not in (2, 3) else (0 if trial == 2 else C.n - 1)
            if trial == 4:
                Q = C.neg(C.mul(P, k))  # k P + Q = identity
            if trial == 5:
                Q = C.mul(P, k)  # doubling inside add
            out = (ctypes.c_uint32 * 16)()
            rc = lib.host_curve_muladd(C.id, limbs(P), limbs([k]), limbs(Q if Q else (0, 0)), out)
            assert rc == 0
            x, y = rd(out, 2)
            got = None if x == 0 and y == 0 else (x, y)
            assert got == C.add(C.mul(P, k), Q), (C.name, trial)
        # off-curve input is rejected
        bad = (C.gen[0], (C.gen[1] + 1) % C.p)
        assert lib.host_curve_muladd(C.id, limbs(bad), limbs([5]), limbs((0, 0)), (ctypes.c_uint32 * 16)()) == 1


@pytest.mark.parametrize("zk", [False, True])
def test_ipa_prove_verify_decide_like_the_reference(zk):
    """The reference's `test_ipa` (pcs/ipa.rs:407-446) and `test_ipa_as` (pcs/ipa/accumulation.rs:212-280) over the oracle:
    random polynomial -> create_proof -> read_proof -> succinct_verify -> (accumulate) -> decide."""
    C = pasta.PALLAS
    rng = random.Random(99 + zk)
    k = 3
    pk = ipa.IpaProvingKey.rand(C, k, zk, rng)
    accs = []
    for _ in range(3):
        p = [rng.randrange(C.n) for _ in range(1 << k)]
        omega = rng.randrange(C.n) if zk else None
        c = pk.commit(p, omega)
        z = rng.randrange(C.n)
        v = ipa.poly_eval(p, z, C.n)
        tw = ipa.HashTranscript(C)
        made = ipa.ipa_create_proof(pk, p, z, omega, tw, rng)
        tr = ipa.HashTranscript(C, tw.finalize())
        proof = ipa.IpaProof.read(zk, k, tr)
        acc = ipa.ipa_succinct_verify(pk, [(1, c)], z, v, proof)
        assert (acc.xi, acc.u) == (made.xi, made.u)
        assert ipa.decide(C, pk.g, acc) == 0
        with pytest.raises(AssertionError):
            ipa.ipa_succinct_verify(pk, [(1, c)], z, (v + 1) % C.n, proof)
        accs.append(acc)
    tw = ipa.HashTranscript(C)
    ipa.ipa_as_create_proof(pk, accs, tw, rng)
    folded = ipa.ipa_as_verify(pk, accs, ipa.HashTranscript(C, tw.finalize()))
    assert ipa.decide(C, pk.g, folded) == 0
    # a wrong accumulator among the inputs: the honest prover's L/R no longer match the verifier's commitment, so the succinct
    # check fails (NativeLoader's `ec_point_assert_eq` panics there, loader/native.rs:81-85) -- before `decide` is ever reached
    bad = [accs[0], ipa.IpaAccumulator(accs[1].xi, C.add(accs[1].u, C.gen)), accs[2]]
    tw = ipa.HashTranscript(C)
    ipa.ipa_as_create_proof(pk, bad, tw, rng)
    with pytest.raises(AssertionError):
        ipa.ipa_as_verify(pk, bad, ipa.HashTranscript(C, tw.finalize()))
    # `decide` is what catches a folded accumulator whose U is not commit(G, h)
    assert ipa.decide(C, pk.g, ipa.IpaAccumulator(folded.xi, C.add(folded.u, C.gen))) == 3


def _load_ipa_golden():
    import json

    with open(os.path.join(HERE, "golden", "ipa_golden.json")) as f:
        gold = json.load(f)
    iv = lambda s: int(s, 16)  # noqa: E731
    pt = lambda p: None if p is None else (iv(p[0]), iv(p[1]))  # noqa: E731
    for case in gold["cases"]:
        C = {"pallas": pasta.PALLAS, "vesta": pasta.VESTA}[case["curve"]]
        g = [pt(p) for p in case["g"]]
        accs = [([iv(x) for x in a["xi"]], pt(a["u"]), pt(a["commit"]), a["status"]) for a in case["accumulators"]]
        yield C, case["k"], g, accs


def test_oracle_against_committed_ipa_golden():
    """tests/golden/ipa_golden.json (made by tests/golden/make_ipa_golden.py) pins the oracle's h_coeffs / MSM / decide."""
    n = 0
    for C, k, g, accs in _load_ipa_golden():
        assert len(g) == 1 << k and all(C.is_on_curve(p) for p in g)
        for xi, u, commit, status in accs:
            assert C.msm_pippenger(ipa.h_coeffs(xi, 1, C.n), g) == commit
            assert ipa.decide(C, g, ipa.IpaAccumulator(xi, u)) == status == (0 if u == commit else 3)
            n += 1
    assert n == 30
