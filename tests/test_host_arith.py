"""Builds the DEVICE arithmetic headers for the host (PTX carry primitives emulated bit-exactly,
csrc/ptx_arith.cuh) and checks the limb algorithms against the oracle: Montgomery field ops, G1 group
law + decompression, optimal-ate pairing (exact Miller and Gt values), Poseidon spec + permutation
(reference KAT), and the protocol compiler + tape VM (challenges and accumulators).  No GPU needed."""
import ctypes
import os
import random
import subprocess

import pytest

from oracle import api, bn254, forge, poseidon
from oracle.bn254 import P, R

from .util import to_product_protocol

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "host", "_hostlib.so")
SRC = os.path.join(HERE, "host", "hostlib.cpp")


@pytest.fixture(scope="module")
def lib():
    if os.environ.get("SVK_HOSTLIB_SANITIZE"):
        # tests/test_host_asan.py re-runs this file with the address + undefined-behaviour sanitizers compiled into the host build
        so = os.path.join(HERE, "host", "_hostlib_asan.so")
        subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-shared", "-fPIC", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                        "-o", so, SRC], check=True, timeout=900)
        return ctypes.CDLL(so)
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", SO, SRC], check=True, timeout=600)
    return ctypes.CDLL(SO)


def limbs(vals):
    out = []
    for v in vals:
        out += [(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)]
    return (ctypes.c_uint32 * len(out))(*out)


def rd(a, n):
    return [sum(int(a[8 * j + i]) << (32 * i) for i in range(8)) for j in range(n)]


def g1l(pt):
    return limbs(pt if pt else (0, 0))


def rd_g1(a):
    x, y = rd(a, 2)
    return None if x == 0 and y == 0 else (x, y)


def test_field_ops(lib):
    rng = random.Random(1)

    def op(f, o, a, b=0):
        out = (ctypes.c_uint32 * 8)()
        lib.host_fe_op(f, o, limbs([a]), limbs([b]), out)
        return rd(out, 1)[0]

    for f, m in ((0, P), (1, R)):
        Rm = (1 << 256) % m
        Ri = pow(Rm, -1, m)
        edge = [0, 1, 2, m - 1, m - 2, Rm, (1 << 254) % m, m >> 1]
        vals = edge + [rng.randrange(m) for _ in range(60)]
        for a in vals:
            for b in rng.sample(vals, 6) + edge:
                assert op(f, 0, a, b) == a * b * Ri % m
                assert op(f, 1, a, b) == (a + b) % m and op(f, 2, a, b) == (a - b) % m
            assert op(f, 3, a) == (-a) % m and op(f, 6, a) == a * Rm % m and op(f, 7, a) == a * Ri % m
            assert op(f, 8, a) == a * a * Ri % m  # dedicated Montgomery squaring
        for a in [m - 1 - rng.randrange(1 << 40) for _ in range(100)] + [rng.randrange(m) for _ in range(2000)] + [
                (1 << 254) - 1 - rng.randrange(1 << 33) for _ in range(50)] + [sum(0xFFFFFFFF << (32 * i) for i in rng.sample(range(8), 5)) % m for _ in range(100)]:
            assert op(f, 8, a) == a * a * Ri % m
        for a in vals[:12] + [rng.randrange(m) for _ in range(300)] + [m - 1, m - 2, 3, (m + 1) // 2, 1 << 253]:
            want = pow(a, -1, m) * Rm % m if a else 0
            assert op(f, 4, a * Rm % m) == want  # binary extended Euclid
        for a in vals[:12]:
            assert op(f, 11, a * Rm % m) == (pow(a, -1, m) * Rm % m if a else 0)  # Fermat chain
        assert lib.host_is_canonical(f, limbs([m - 1])) == 1 and lib.host_is_canonical(f, limbs([m])) == 0
        # to_mont_wide: any 256-bit input (the Keccak challenge, evm.rs:172-182 `u256_to_fe`), worst cases included
        wide = [(1 << 256) - 1, (1 << 256) - 2, 0xFFFFFFFF << 224, m, 2 * m, 5 * m + 3, (1 << 255) + 12345]
        for a in wide + [rng.randrange(1 << 256) for _ in range(200)]:
            assert op(f, 10, a) == a * Rm % m
    # fused dot products (lazy reduction): operands up to p itself, optional p - a operands
    for f, m in ((0, P), (1, R)):
        Ri = pow(1 << 256, -1, m)
        edge = [0, 1, m - 1, m, m - 2, (1 << 253) + 5]
        for n in (2, 3):
            for trial in range(300):
                src = edge if trial < 40 else None
                a = [rng.choice(src) if src else rng.randrange(m) for _ in range(n)]
                b = [rng.choice(src) if src else rng.randrange(m) for _ in range(n)]
                neg = rng.randrange(1 << n)
                out = (ctypes.c_uint32 * 8)()
                lib.host_fe_dot(f, n, neg, limbs(a), limbs(b), out)
                want = sum((-x if neg >> k & 1 else x) * y for k, (x, y) in enumerate(zip(a, b))) * Ri % m
                assert rd(out, 1)[0] == want, (f, n, a, b, neg)
    for a in [rng.randrange(P) for _ in range(8)]:
        Rm = (1 << 256) % P
        assert op(0, 5, a * Rm % P) == pow(a, (P + 1) // 4, P) * Rm % P


def test_g1_ops(lib):
    rng = random.Random(2)
    Pt = bn254.g1_mul(bn254.G1_GEN, rng.randrange(bn254.R))
    cases = [(0, None), (1, None), (1, Pt), (2, bn254.g1_neg(bn254.g1_mul(Pt, 2))), (R - 1, Pt), (5, bn254.g1_mul(Pt, 5)),
             (rng.randrange(R), bn254.g1_mul(bn254.G1_GEN, 77))]
    for k, q in cases:
        exp = bn254.g1_add(bn254.g1_mul(Pt, k), q)
        for repr_ in (0, 1):
            out = (ctypes.c_uint32 * 16)()
            lib.host_g1_muladd(repr_, g1l(Pt), limbs([k]), g1l(q), out)
            assert rd_g1(out) == exp
    for _ in range(8):
        pt = bn254.g1_mul(bn254.G1_GEN, rng.randrange(1, R))
        out = (ctypes.c_uint32 * 16)()
        assert lib.host_g1_decompress(bn254.g1_to_bytes(pt), out) == 0 and rd_g1(out) == pt
    assert lib.host_g1_decompress(bytes(32), (ctypes.c_uint32 * 16)()) == 2
    assert lib.host_g1_decompress(P.to_bytes(32, "little"), (ctypes.c_uint32 * 16)()) == 1
    for x in range(1, 30):
        ok, pt = bn254.g1_from_bytes(x.to_bytes(32, "little"))
        out = (ctypes.c_uint32 * 16)()
        rc = lib.host_g1_decompress(x.to_bytes(32, "little"), out)
        assert (rc == 0) == ok and (not ok or rd_g1(out) == pt)


def test_straus_signed_windows(lib):
    """csrc/straus.cuh (k_msm_var / k_group_var core): recoding k + C, 16-entry tables, signed digits -- vs the naive MSM."""
    rng = random.Random(21)
    for trial in range(10):
        nt = [1, 2, 5, 11, 16, 3, 7, 1, 4, 2][trial]
        pts = [bn254.g1_mul(bn254.G1_GEN, rng.randrange(1, R)) for _ in range(nt)]
        ks = [rng.randrange(R) for _ in range(nt)]
        if trial == 1:
            ks = [0, R - 1]
        if trial == 2:
            ks[:5] = [1, 16, 17, (1 << 253) + 16, sum(16 << (5 * i) for i in range(50))]  # digits at the -16 / +16 extremes
            pts[1] = None
        if trial == 5:
            pts[1] = pts[0]
            pts[2] = bn254.g1_neg(pts[0])
            ks[2] = (ks[0] + ks[1]) % R  # total = identity
        if trial == 7:
            ks = [sum(15 << (5 * i) for i in range(51)) % R]
        out = (ctypes.c_uint32 * 16)()
        flat = []
        for p_ in pts:
            flat += list(p_ if p_ else (0, 0))
        assert lib.host_straus(nt, limbs(flat), limbs(ks), out) == 0
        assert rd_g1(out) == bn254.g1_msm_naive(list(zip(ks, pts))), trial


def test_pairing_exact(lib):
    rng = random.Random(3)
    s, d = rng.randrange(1, R), rng.randrange(1, R)
    sg2 = bn254.g2_mul(bn254.G2_GEN, s)

    def g2l(q):
        return limbs([q[0][0], q[0][1], q[1][0], q[1][1]])

    def rd12(a):
        v = rd(a, 12)
        f2 = [(v[2 * i], v[2 * i + 1]) for i in range(6)]
        return ((f2[0], f2[1], f2[2]), (f2[3], f2[4], f2[5]))

    G = bn254.G1_GEN
    cases = [(bn254.g1_mul(G, s * d % R), bn254.g1_mul(G, d), 1), (bn254.g1_mul(G, (s * d + 1) % R), bn254.g1_mul(G, d), 0), (None, None, 1)]
    for lhs, rhs, exp in cases:
        gt, ml = (ctypes.c_uint32 * 96)(), (ctypes.c_uint32 * 96)()
        rc = lib.host_pairing(g1l(lhs), g2l(bn254.G2_GEN), g1l(rhs), g2l(bn254.g2_neg(sg2)), gt, ml)
        mlo = bn254.multi_miller_loop([(lhs, bn254.G2_GEN), (rhs, bn254.g2_neg(sg2))])
        assert rd12(ml) == mlo and rd12(gt) == bn254.final_exponentiation(mlo) and rc == exp


def test_public_precompile_vectors_on_device_headers(lib):
    """EIP-196 / EIP-197 client test vectors (tests/golden/eip196_197_vectors.py) through the DEVICE headers built for the host: the
    Jacobian and XYZZ group laws, the signed-window scalar multiplication, both pairing programs (tower and block-cooperative)."""
    from .golden import eip196_197_vectors as E

    def g2l(q):
        return limbs([q[0][0], q[0][1], q[1][0], q[1][1]])

    for repr_ in (0, 1):
        out = (ctypes.c_uint32 * 16)()
        lib.host_g1_muladd(repr_, g1l(E.ADD_A), limbs([1]), g1l(E.ADD_B), out)   # 1 * A + B
        assert rd_g1(out) == E.ADD_C
        lib.host_g1_muladd(repr_, g1l(E.MUL_P), limbs([E.MUL_K]), g1l(None), out)  # K * P + O
        assert rd_g1(out) == E.MUL_Q
    for fn in (lib.host_pairing, lib.host_pairing_coop):
        gt, ml = (ctypes.c_uint32 * 96)(), (ctypes.c_uint32 * 96)()
        assert fn(g1l(E.PAIR_P1), g2l(E.PAIR_Q1), g1l(E.PAIR_P2), g2l(E.PAIR_Q2), gt, ml) == 1
        assert fn(g1l(E.PAIR_P1), g2l(E.PAIR_Q1), g1l(bn254.g1_neg(E.PAIR_P2)), g2l(E.PAIR_Q2), gt, ml) == 0
        assert fn(g1l(E.ADD_A), g2l(E.PAIR_Q1), g1l(E.PAIR_P2), g2l(E.PAIR_Q2), gt, ml) == 0


def test_coop_pairing_matches_tower(lib):
    """The block-cooperative pairing (csrc/coop_pairing.cuh: 48 dot3 lanes + 12 combine lanes per Fq12 product), executed
    lane by lane on the host, gives the oracle's exact Miller and GT values (decider.rs:60-68)."""
    rng = random.Random(5)
    s, d = rng.randrange(1, R), rng.randrange(1, R)
    sg2 = bn254.g2_mul(bn254.G2_GEN, s)

    def g2l(q):
        return limbs([q[0][0], q[0][1], q[1][0], q[1][1]])

    def rd12(a):
        v = rd(a, 12)
        f2 = [(v[2 * i], v[2 * i + 1]) for i in range(6)]
        return ((f2[0], f2[1], f2[2]), (f2[3], f2[4], f2[5]))

    G = bn254.G1_GEN
    cases = [(bn254.g1_mul(G, s * d % R), bn254.g1_mul(G, d), 1), (bn254.g1_mul(G, (s * d + 1) % R), bn254.g1_mul(G, d), 0), (None, None, 1),
             (None, bn254.g1_mul(G, d), 0)]
    for lhs, rhs, exp in cases:
        gt, ml = (ctypes.c_uint32 * 96)(), (ctypes.c_uint32 * 96)()
        rc = lib.host_pairing_coop(g1l(lhs), g2l(bn254.G2_GEN), g1l(rhs), g2l(bn254.g2_neg(sg2)), gt, ml)
        mlo = bn254.multi_miller_loop([(lhs, bn254.G2_GEN), (rhs, bn254.g2_neg(sg2))])
        assert rd12(ml) == mlo and rd12(gt) == bn254.final_exponentiation(mlo) and rc == exp


def test_glv_decompose(lib):
    """csrc/glv.cuh: k = (+-k1) + (+-k2) lambda (mod r) with |k1|, |k2| < 2^128, and phi(P) = (beta x, y) = lambda P
    (constants derived by tools/glv_constants.py)."""
    lam = 4407920970296243842393367215006156084916469457145843978461
    out = (ctypes.c_uint32 * 8)()
    lib.host_glv_beta(out)
    beta = rd(out, 1)[0]
    G = bn254.G1_GEN
    assert pow(lam, 3, R) == 1 and pow(beta, 3, bn254.P) == 1 and bn254.g1_mul(G, lam) == (beta * G[0] % bn254.P, G[1])
    rng = random.Random(9)
    ks = [0, 1, 2, R - 1, R - 2, lam, R - lam, (lam * lam) % R, 1 << 127, 1 << 128, (1 << 253) + 1, R // 2, R // 3] + [rng.randrange(R) for _ in range(3000)]
    buf = (ctypes.c_uint32 * 10)()
    worst = 0
    for k in ks:
        assert lib.host_glv_decompose(limbs([k]), buf) == 1
        v = list(buf)
        k1 = sum(v[i] << (32 * i) for i in range(4)) * (-1 if v[4] else 1)
        k2 = sum(v[5 + i] << (32 * i) for i in range(4)) * (-1 if v[9] else 1)
        assert (k1 + k2 * lam - k) % R == 0, k
        worst = max(worst, abs(k1), abs(k2))
    assert worst < 1 << 128


def test_straus_glv_single_term(lib):
    """straus.cuh straus_run_glv1 (the latency schedule of the per-proof MSM): k P through the GLV split and signed 5-bit windows
    equals the oracle's scalar multiplication, incl. 0, 1, r - 1, lambda and the window-boundary patterns."""
    rng = random.Random(31)
    lam = 4407920970296243842393367215006156084916469457145843978461
    P_ = bn254.g1_mul(bn254.G1_GEN, rng.randrange(1, R))
    ks = [0, 1, 2, 15, 16, 17, 31, 32, R - 1, R - 2, lam, R - lam, (1 << 128) - 1, 1 << 127, sum(16 << (5 * i) for i in range(25)),
          sum(15 << (5 * i) for i in range(50)) % R] + [rng.randrange(R) for _ in range(40)]
    out = (ctypes.c_uint32 * 16)()
    for k in ks:
        assert lib.host_straus_glv1(limbs(list(P_)), limbs([k]), out) == 0
        assert rd_g1(out) == bn254.g1_mul(P_, k), k
    assert lib.host_straus_glv1(limbs([0, 0]), limbs([5]), out) == 0 and rd_g1(out) is None


def test_poseidon_kat_and_random(lib):
    out = (ctypes.c_uint32 * 24)()
    assert lib.host_poseidon_permute(limbs([0, 1, 2]), 2, limbs([0]), limbs([0]), out) == 0
    assert rd(out, 3)[0] == 7853200120776062878684798364095072458815029376092732009249414926327459813530  # tests.rs:51
    sp = poseidon.spec()
    rng = random.Random(4)
    for _ in range(4):
        st = [rng.randrange(R) for _ in range(3)]
        a, b = rng.randrange(R), rng.randrange(R)
        for n in (0, 1, 2):
            lib.host_poseidon_permute(limbs(st), n, limbs([a]), limbs([b]), out)
            assert rd(out, 3) == poseidon.permutation_optimized(sp, st, [a, b][:n])


@pytest.mark.parametrize("scheme,mos", [("bdfg21", 0), ("gwc19", 1)])
def test_compiler_and_tape(lib, scheme, mos):
    S = forge.Setup(0)
    blob = to_product_protocol(S.protocol).to_bytes()
    inst, pf = forge.forge_proof(S, scheme, 7)
    accs, proof = api.succinct_verify(S.dk.svk, S.protocol, inst, pf, scheme, want_proof=True)
    ch, sc = (ctypes.c_uint32 * (8 * 32))(), (ctypes.c_uint32 * (8 * 64))()
    terms, info, err = (ctypes.c_int * 300)(), (ctypes.c_longlong * 10)(), ctypes.create_string_buffer(256)
    instb = b"".join(int(x).to_bytes(32, "little") for col in inst for x in col)
    nt = lib.host_compile_run(blob, len(blob), mos, pf, len(pf), instb, len(instb) // 32, ch, sc, terms, 100, info, err, 256)
    assert nt > 0, err.value
    assert info[2] == (27 if scheme == "bdfg21" else 28) and info[5] == len(pf) and info[6] == 0xFFFFFFFF and info[7] == 1
    exp_ch = [c.v for c in proof.challenges] + [proof.z.v]
    exp_ch += [proof.pcs.mu.v, proof.pcs.gamma.v, proof.pcs.z_prime.v] if scheme == "bdfg21" else [proof.pcs.v.v, proof.pcs.u.v]
    assert rd(ch, info[3]) == exp_ch
    npre = len(S.protocol.preprocessed)
    pts = [bn254.g1_from_bytes(pf[32 * i : 32 * i + 32])[1] for i in range(9)]
    off = 26 * 32
    pts += [bn254.g1_from_bytes(pf[off + 32 * i : off + 32 * i + 32])[1] for i in range(2 if scheme == "bdfg21" else 3)]
    res = [None, None]
    scal = rd(sc, info[4])
    for t in range(nt):
        which, base, slot = terms[3 * t], terms[3 * t + 1], terms[3 * t + 2]
        b = bn254.G1_GEN if base == -1 else (S.protocol.preprocessed[base] if base < npre else pts[base - npre])
        res[which] = bn254.g1_add(res[which], bn254.g1_mul(b, 1 if slot < 0 else scal[slot]))
    assert res[0] == accs[0].lhs.pt and res[1] == accs[0].rhs.pt
    # error paths through the same tape: scalar out of range, short proof
    bad = bytearray(pf)
    bad[9 * 32 : 10 * 32] = b"\xff" * 32
    lib.host_compile_run(blob, len(blob), mos, bytes(bad), len(bad), instb, 1, ch, sc, terms, 100, info, err, 256)
    assert info[6] == ((9 * 32) << 8 | 2)
    lib.host_compile_run(blob, len(blob), mos, pf, 100, instb, 1, ch, sc, terms, 100, info, err, 256)
    assert info[6] & 0xFF == 1


@pytest.mark.parametrize("scheme,mos", [("bdfg21", 0), ("gwc19", 1)])
def test_compiler_and_tape_evm_transcript(lib, scheme, mos):
    """Keccak `EvmTranscript` tape (transcript/evm.rs:152-243): challenges equal the oracle's; Keccak itself vs the
    Python restatement on multi-block inputs."""
    from oracle.keccak import keccak256

    rng = random.Random(9)
    for n in [0, 1, 31, 32, 33, 135, 136, 137, 500]:
        d = bytes(rng.getrandbits(8) for _ in range(n))
        out = ctypes.create_string_buffer(32)
        lib.host_keccak256(d, n, out)
        assert out.raw == keccak256(d)
    assert keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"  # public Keccak-256 KAT
    S = forge.Setup(0)
    blob = to_product_protocol(S.protocol).to_bytes()
    inst, pf = forge.forge_proof(S, scheme, 9, transcript="evm")
    accs, proof = api.succinct_verify(S.dk.svk, S.protocol, inst, pf, scheme, want_proof=True, transcript="evm")
    ch, sc = (ctypes.c_uint32 * (8 * 32))(), (ctypes.c_uint32 * (8 * 64))()
    terms, info, err = (ctypes.c_int * 300)(), (ctypes.c_longlong * 10)(), ctypes.create_string_buffer(256)
    instb = b"".join(int(x).to_bytes(32, "little") for col in inst for x in col)
    nt = lib.host_compile_run_ex(blob, len(blob), mos, 1, pf, len(pf), instb, 1, ch, sc, terms, 100, info, err, 256)
    assert nt > 0, err.value
    exp = [c.v for c in proof.challenges] + [proof.z.v]
    exp += [proof.pcs.mu.v, proof.pcs.gamma.v, proof.pcs.z_prime.v] if scheme == "bdfg21" else [proof.pcs.v.v, proof.pcs.u.v]
    assert rd(ch, info[3]) == exp and info[5] == len(pf) and info[6] == 0xFFFFFFFF


@pytest.mark.parametrize("scheme,mos", [("bdfg21", 0), ("gwc19", 1)])
def test_compiler_general_protocol(lib, scheme, mos):
    """The tape compiler on a protocol beyond StandardPlonk: lookup argument, two advice phases with a user
    challenge, rotations, two permutation grand products, num_proof = 2 (system/halo2.rs:199-243, 372-408, 593-668;
    verifier/plonk/proof.rs:179-318).  Challenges and the evaluated accumulator Msm equal the oracle's."""
    from .util import lookup_two_phase_shape

    S = forge.Setup(3, shape=lookup_two_phase_shape(), num_instance=[2], num_proof=2)
    P_ = S.protocol
    blob = to_product_protocol(P_).to_bytes()
    inst, pf = forge.forge_proof(S, scheme, 9)
    accs, proof = api.succinct_verify(S.dk.svk, P_, inst, pf, scheme, want_proof=True)
    ch, sc = (ctypes.c_uint32 * (8 * 64))(), (ctypes.c_uint32 * (8 * 512))()
    terms, info, err = (ctypes.c_int * 1500)(), (ctypes.c_longlong * 10)(), ctypes.create_string_buffer(256)
    instb = b"".join(int(x).to_bytes(32, "little") for col in inst for x in col)
    nt = lib.host_compile_run(blob, len(blob), mos, pf, len(pf), instb, len(instb) // 32, ch, sc, terms, 500, info, err, 256)
    assert nt > 0, err.value
    assert info[5] == len(pf) and info[6] == 0xFFFFFFFF and info[7] == 1 and info[4] <= 512 and info[3] <= 64
    exp_ch = [c.v for c in proof.challenges] + [proof.z.v]
    exp_ch += [proof.pcs.mu.v, proof.pcs.gamma.v, proof.pcs.z_prime.v] if scheme == "bdfg21" else [proof.pcs.v.v, proof.pcs.u.v]
    assert rd(ch, info[3]) == exp_ch
    npre = len(P_.preprocessed)
    n_first = sum(P_.num_witness) + P_.quotient.num_chunk()
    offs = [32 * i for i in range(n_first)]
    tail = 32 * (n_first + len(P_.evaluations))
    offs += [tail + 32 * i for i in range((len(pf) - tail) // 32)]
    pts = [bn254.g1_from_bytes(pf[o : o + 32])[1] for o in offs]
    assert info[9] == len(pts)
    res = [None, None]
    scal = rd(sc, info[4])
    for t in range(nt):
        which, base, slot = terms[3 * t], terms[3 * t + 1], terms[3 * t + 2]
        b = bn254.G1_GEN if base == -1 else (P_.preprocessed[base] if base < npre else pts[base - npre])
        res[which] = bn254.g1_add(res[which], bn254.g1_mul(b, 1 if slot < 0 else scal[slot]))
    assert res[0] == accs[0].lhs.pt and res[1] == accs[0].rhs.pt


@pytest.mark.parametrize("fe,fe_id", [("montgomery", 1), ("canonical", 2)])
def test_protocol_bincode_ingestion(lib, fe, fe_id):
    """`bincode::serialize(&PlonkProtocol)` (the head of a `Snark` file, sdk/src/lib.rs:44-50, sdk/src/halo2.rs:262-269) compiles
    to the same tape as the library's own blob; AUTO detects the field-element encoding; truncation is an error."""
    from oracle import bincode as obc

    from .util import lookup_two_phase_shape

    cases = [forge.Setup(0), forge.Setup(0, num_instance=14, accumulator_indices=[[(0, 1 + i) for i in range(12)]]),
             forge.Setup(3, shape=lookup_two_phase_shape(), num_instance=[2], num_proof=2)]
    for S in cases:
        blob = to_product_protocol(S.protocol).to_bytes()
        data = obc.serialize_protocol(S.protocol, fe)
        tail = b"trailing snark fields"
        for enc in (0, fe_id):
            for mos in (0, 1):
                consumed, used, err = ctypes.c_size_t(0), ctypes.c_int(0), ctypes.create_string_buffer(256)
                rc = lib.host_compile_compare_bincode(blob, len(blob), data + tail, len(data) + len(tail), enc, mos, 0,
                                                      ctypes.byref(consumed), ctypes.byref(used), err, 256)
                assert rc == 1, err.value
                assert consumed.value == len(data) and used.value == fe_id
        consumed, used, err = ctypes.c_size_t(0), ctypes.c_int(0), ctypes.create_string_buffer(256)
        assert lib.host_compile_compare_bincode(blob, len(blob), data, len(data) - 9, 0, 0, 0, ctypes.byref(consumed), ctypes.byref(used), err, 256) == -1
        wrong = 3 - fe_id  # forcing the other encoding: inconsistent Domain (or an out-of-range element)
        assert lib.host_compile_compare_bincode(blob, len(blob), data, len(data), wrong, 0, 0, ctypes.byref(consumed), ctypes.byref(used), err, 256) == -1
