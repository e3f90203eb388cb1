"""TEST INFRASTRUCTURE ONLY.  Restatement of the reference's IPA accumulator decider over exact integers:

  h_coeffs / h_eval                 snark-verifier/src/pcs/ipa.rs:366-395
  IpaAccumulator { xi, u }          snark-verifier/src/pcs/ipa/accumulator.rs:5-25
  IpaDecidingKey { svk, g }         snark-verifier/src/pcs/ipa/decider.rs:5-16
  IpaAs::decide / decide_all        snark-verifier/src/pcs/ipa/decider.rs:47-67
  fold_bases                        the base-folding loop of Ipa::create_proof, pcs/ipa.rs:78-118 (what produces `u` for an
                                    honest prover): pins the order in which `decide` must pair xi with coefficient bits.
"""
from .pasta import Curve


def h_coeffs(xi, scalar, n):
    """pcs/ipa.rs:379-395: coeffs[0] = scalar; for (i, x) in xi.rev().enumerate(): right block of len 2^i = left block * x."""
    assert len(xi) > 0
    coeffs = [0] * (1 << len(xi))
    coeffs[0] = scalar % n
    for i, x in enumerate(reversed(xi)):
        ln = 1 << i
        for j in range(ln):
            coeffs[ln + j] = coeffs[j] * x % n
    return coeffs


def h_eval(xi, z, n):
    """pcs/ipa.rs:366-377: prod_i (z^(2^i) * xi.rev()[i] + 1)."""
    out = 1
    zp = z % n
    for x in reversed(xi):
        out = out * (zp * x + 1) % n
        zp = zp * zp % n
    return out


def fold_bases(curve: Curve, g, xi):
    """pcs/ipa.rs:78-118, bases only: round i halves `bases`: bases_l[j] += bases_r[j] * xi_i."""
    bases = list(g)
    for x in xi:
        half = len(bases) // 2
        bases = [curve.add(bases[j], curve.mul(bases[half + j], x)) for j in range(half)]
    assert len(bases) == 1
    return bases[0]


class IpaAccumulator:
    def __init__(self, xi, u):
        self.xi, self.u = list(xi), u


def decide(curve: Curve, g, acc: IpaAccumulator, msm=None):
    """decider.rs:47-56 -> status 0 ok / 3 Error::AssertionFailure("U == commit(G, h)")."""
    h = h_coeffs(acc.xi, 1, curve.n)
    commit = (msm or curve.msm_naive)(h, g)
    return 0 if commit == acc.u else 3


def decide_all(curve: Curve, g, accs):
    """decider.rs:58-67 (fail-fast `try_collect`): first failing status, else 0."""
    assert len(accs) > 0
    for a in accs:
        st = decide(curve, g, a)
        if st:
            return st
    return 0


# ------------------------------------------------------------------------------------------------------------------
# The rest of the IPA scheme, restated so that the parity tests can read like the reference's own `test_ipa`
# (pcs/ipa.rs:407-446) and `test_ipa_as` (pcs/ipa/accumulation.rs:212-280): prove with a seeded rng, read the proof back,
# `succinct_verify`, accumulate, and hand the accumulator to `decide` (the part libsvk runs on the GPU).
import hashlib


class HashTranscript:
    """In-memory stand-in for halo2_proofs' Blake2b transcript that the reference tests use (`Blake2bWrite` / `Blake2bRead`):
    prefix-tagged absorption into BLAKE2b-512 personalised "Halo2-Transcript", challenge = digest mod n.  Only its
    Fiat-Shamir role matters here (accumulators do not depend on the byte format); written items are kept as values."""

    def __init__(self, curve: Curve, stream=None):
        self.curve = curve
        self.h = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")
        self.stream = list(stream) if stream is not None else []
        self.pos = 0

    def common_scalar(self, s):
        self.h.update(b"\x02" + int(s).to_bytes(32, "little"))

    def common_ec_point(self, p):
        assert p is not None, "cannot write points at infinity to the transcript"
        self.h.update(b"\x01" + p[0].to_bytes(32, "little") + p[1].to_bytes(32, "little"))

    def squeeze_challenge(self):
        self.h.update(b"\x00")
        return int.from_bytes(self.h.copy().digest(), "little") % self.curve.n

    def write_scalar(self, s):
        self.common_scalar(s)
        self.stream.append(("scalar", s))

    def write_ec_point(self, p):
        self.common_ec_point(p)
        self.stream.append(("point", p))

    def _next(self, kind):
        assert self.pos < len(self.stream) and self.stream[self.pos][0] == kind, "transcript: unexpected item"
        v = self.stream[self.pos][1]
        self.pos += 1
        return v

    def read_scalar(self):
        s = self._next("scalar")
        self.common_scalar(s)
        return s

    def read_ec_point(self):
        p = self._next("point")
        self.common_ec_point(p)
        return p

    def finalize(self):
        return list(self.stream)


class IpaProvingKey:
    """pcs/ipa.rs:184-246: domain (k), g (2^k points), h, optional s (zero-knowledge)."""

    def __init__(self, curve: Curve, k, g, h, s=None):
        self.curve, self.k, self.g, self.h, self.s = curve, k, list(g), h, s

    @staticmethod
    def rand(curve: Curve, k, zk, rng):
        pt = lambda: curve.mul(curve.gen, rng.randrange(1, curve.n))  # noqa: E731
        return IpaProvingKey(curve, k, [pt() for _ in range(1 << k)], pt(), pt() if zk else None)

    def zk(self):
        return self.s is not None

    def commit(self, poly, omega=None, msm=None):
        c = (msm or self.curve.msm_naive)(list(poly), self.g)
        if self.s is not None:
            assert omega is not None
            c = self.curve.add(c, self.curve.mul(self.s, omega))
        else:
            assert omega is None
        return c


def poly_eval(p, z, n):
    acc = 0
    for c in reversed(p):
        acc = (acc * z + c) % n
    return acc


def ipa_create_proof(pk: IpaProvingKey, p, z, omega, transcript, rng):
    """Ipa::create_proof, pcs/ipa.rs:38-123 -> IpaAccumulator(xi, bases[0])."""
    C, n = pk.curve, pk.curve.n
    p_prime = [c % n for c in p]
    if pk.zk():
        p_bar = [rng.randrange(n) for _ in range(len(p))]
        p_bar[0] = (p_bar[0] - poly_eval(p_bar, z, n)) % n
        omega_bar = rng.randrange(n)
        transcript.write_ec_point(pk.commit(p_bar, omega_bar))
        alpha = transcript.squeeze_challenge()
        transcript.write_scalar((omega + alpha * omega_bar) % n)
        p_prime = [(a + alpha * b) % n for a, b in zip(p_prime, p_bar)]
    xi_0 = transcript.squeeze_challenge()
    h_prime = C.mul(pk.h, xi_0)
    bases, coeffs = list(pk.g), p_prime
    zs = [pow(z, i, n) for i in range(len(coeffs))]
    ip = lambda a, b: sum(x * y for x, y in zip(a, b)) % n  # noqa: E731
    xi = []
    for i in range(pk.k):
        half = 1 << (pk.k - i - 1)
        l_i = C.add(C.msm_naive(coeffs[half:], bases[:half]), C.mul(h_prime, ip(coeffs[half:], zs[:half])))
        r_i = C.add(C.msm_naive(coeffs[:half], bases[half:]), C.mul(h_prime, ip(coeffs[:half], zs[half:])))
        transcript.write_ec_point(l_i)
        transcript.write_ec_point(r_i)
        xi_i = transcript.squeeze_challenge()
        xi_i_inv = pow(xi_i, -1, n)
        bases = [C.add(bases[j], C.mul(bases[half + j], xi_i)) for j in range(half)]
        coeffs = [(coeffs[j] + xi_i_inv * coeffs[half + j]) % n for j in range(half)]
        zs = [(zs[j] + xi_i * zs[half + j]) % n for j in range(half)]
        xi.append(xi_i)
    transcript.write_ec_point(bases[0])
    transcript.write_scalar(coeffs[0])
    return IpaAccumulator(xi, bases[0])


class IpaProof:
    """pcs/ipa.rs:283-352"""

    def __init__(self, c_bar_alpha, omega_prime, xi_0, rounds, u, c):
        self.c_bar_alpha, self.omega_prime, self.xi_0, self.rounds, self.u, self.c = c_bar_alpha, omega_prime, xi_0, rounds, u, c

    @staticmethod
    def read(zk, k, transcript):
        c_bar_alpha = None
        omega_prime = None
        if zk:
            c_bar = transcript.read_ec_point()
            c_bar_alpha = (c_bar, transcript.squeeze_challenge())
            omega_prime = transcript.read_scalar()
        xi_0 = transcript.squeeze_challenge()
        rounds = []
        for _ in range(k):
            l, r = transcript.read_ec_point(), transcript.read_ec_point()
            rounds.append((l, r, transcript.squeeze_challenge()))
        u = transcript.read_ec_point()
        c = transcript.read_scalar()
        return IpaProof(c_bar_alpha, omega_prime, xi_0, rounds, u, c)

    def xi(self):
        return [r[2] for r in self.rounds]


def ipa_succinct_verify(pk_or_svk: IpaProvingKey, commitment, z, ev, proof: IpaProof, msm=None):
    """Ipa::succinct_verify, pcs/ipa.rs:137-181.  `commitment` is an Msm as [(scalar, base)]; the two `evaluate(None)` go through
    `msm` (the loader's multi_scalar_multiplication).  Raises AssertionError where the reference's `ec_point_assert_eq` panics."""
    C, n = pk_or_svk.curve, pk_or_svk.curve.n
    msm = msm or C.msm_naive
    h, s = pk_or_svk.h, pk_or_svk.s
    xi = proof.xi()
    xi_inv = [pow(x, -1, n) for x in xi]
    terms = list(commitment)
    if s is not None:
        c_bar, alpha = proof.c_bar_alpha
        terms += [(alpha, c_bar), ((-proof.omega_prime) % n, s)]
    else:
        assert proof.c_bar_alpha is None and proof.omega_prime is None
    terms.append((proof.xi_0 * ev % n, h))  # h_prime * eval
    for (l, r, x), xinv in zip(proof.rounds, xi_inv):
        terms += [(xinv, l), (x, r)]
    lhs = msm([t[0] for t in terms], [t[1] for t in terms])
    v_prime = h_eval(xi, z, n) * proof.c % n
    rhs = msm([proof.c, proof.xi_0 * v_prime % n], [proof.u, h])
    assert lhs == rhs, "C_k == c[U] + v'[H']"
    return IpaAccumulator(xi, proof.u)


def ipa_as_create_proof(pk: IpaProvingKey, instances, transcript, rng):
    """IpaAs::create_proof, pcs/ipa/accumulation.rs:145-208."""
    C, n = pk.curve, pk.curve.n
    assert len(instances) > 1
    a_b_u = omega = None
    if pk.zk():
        a, b = rng.randrange(n), rng.randrange(n)
        u = C.add(C.mul(pk.g[1], a), C.mul(pk.g[0], b))
        transcript.write_scalar(a)
        transcript.write_scalar(b)
        transcript.write_ec_point(u)
        a_b_u = (a, b, u)
        omega = rng.randrange(n)
        transcript.write_scalar(omega)
    for acc in instances:
        for x in acc.xi:
            transcript.common_scalar(x)
        transcript.common_ec_point(acc.u)
    alpha = transcript.squeeze_challenge()
    z = transcript.squeeze_challenge()
    hs = [h_coeffs(acc.xi, 1, n) for acc in instances]
    if a_b_u:
        hs.append([a_b_u[1], a_b_u[0]] + [0] * ((1 << pk.k) - 2))
    h = [0] * (1 << pk.k)
    pw = 1
    for hp in hs:
        h = [(x + pw * y) % n for x, y in zip(h, hp)]
        pw = pw * alpha % n
    return ipa_create_proof(pk, h, z, omega, transcript, rng)


def ipa_as_verify(svk: IpaProvingKey, instances, transcript, msm=None):
    """IpaAsProof::read + IpaAs::verify, pcs/ipa/accumulation.rs:45-72, 95-137."""
    C, n = svk.curve, svk.curve.n
    assert len(instances) > 1
    a_b_u = omega = None
    if svk.zk():
        a, b = transcript.read_scalar(), transcript.read_scalar()
        a_b_u = (a, b, transcript.read_ec_point())
        omega = transcript.read_scalar()
    for acc in instances:
        for x in acc.xi:
            transcript.common_scalar(x)
        transcript.common_ec_point(acc.u)
    alpha = transcript.squeeze_challenge()
    z = transcript.squeeze_challenge()
    proof = IpaProof.read(svk.zk(), svk.k, transcript)
    us = [acc.u for acc in instances]
    hs = [h_eval(acc.xi, z, n) for acc in instances]
    if a_b_u:
        us.append(a_b_u[2])
        hs.append((a_b_u[0] * z + a_b_u[1]) % n)
    pws = [pow(alpha, i, n) for i in range(len(us))]
    c = list(zip(pws, us))
    if omega is not None:
        c.append((omega, svk.s))
    v = sum(p * h for p, h in zip(pws, hs)) % n
    return ipa_succinct_verify(svk, c, z, v, proof, msm)
