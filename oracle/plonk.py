"""PLONK verifier (`PlonkProtocol`, `PlonkProof`, `PlonkSuccinctVerifier`, `PlonkVerifier`) --
oracle restatement.  TEST INFRASTRUCTURE ONLY.

Follows
  snark-verifier/src/verifier/plonk/protocol.rs:21-63, 65-99, 181-279, 282-418, 504-519
  snark-verifier/src/verifier/plonk/proof.rs:21-43, 52-153, 156-318
  snark-verifier/src/verifier/plonk.rs:32-135
  snark-verifier/src/pcs.rs:22-45 (`pcs::Query`)
"""
from dataclasses import dataclass, field
from typing import Any, List, Optional, Tuple

from .loader import Domain, Fraction, Msm, NativeLoader
from .transcript import VerifyError


# ---------------------------------------------------------------- Expression (protocol.rs:308-418)
# Tuples: ("const", v) ("identity",) ("lagrange", i) ("poly", poly, rot) ("challenge", idx)
#         ("neg", a) ("sum", a, b) ("product", a, b) ("scaled", a, v) ("distribute_powers", [exprs], base)
def Const(v):
    return ("const", v)


def Identity():
    return ("identity",)


def Lagrange(i):
    return ("lagrange", i)


def Poly(poly, rot=0):
    return ("poly", poly, rot)


def Challenge(i):
    return ("challenge", i)


def Neg(a):
    return ("neg", a)


def Sum(a, b):
    return ("sum", a, b)


def Sub(a, b):
    """`impl Sub`: Sum(a, Negated(b)) (protocol.rs:464)"""
    return ("sum", a, ("neg", b))


def Product(a, b):
    return ("product", a, b)


def Scaled(a, v):
    return ("scaled", a, v)


def DistributePowers(exprs, base):
    return ("distribute_powers", list(exprs), base)


def expr_sum(exprs):
    """`impl Sum for Expression` (protocol.rs:480-484)"""
    acc = None
    for e in exprs:
        acc = e if acc is None else Sum(acc, e)
    return acc if acc is not None else Const(0)


def expr_evaluate(e, constant, common_poly, poly, challenge, negated, sum_, product, scaled):
    """`Expression::evaluate` (protocol.rs:322-370)"""

    def ev(x):
        return expr_evaluate(x, constant, common_poly, poly, challenge, negated, sum_, product, scaled)

    tag = e[0]
    if tag == "const":
        return constant(e[1])
    if tag == "identity" or tag == "lagrange":
        return common_poly(e)
    if tag == "poly":
        return poly((e[1], e[2]))
    if tag == "challenge":
        return challenge(e[1])
    if tag == "neg":
        return negated(ev(e[1]))
    if tag == "sum":
        a = ev(e[1])
        b = ev(e[2])
        return sum_(a, b)
    if tag == "product":
        a = ev(e[1])
        b = ev(e[2])
        return product(a, b)
    if tag == "scaled":
        return scaled(ev(e[1]), e[2])
    if tag == "distribute_powers":
        exprs, base = e[1], e[2]
        assert exprs
        if len(exprs) == 1:
            return ev(exprs[0])
        first = ev(exprs[0])
        scalar = ev(base)
        acc = first
        for x in exprs[1:]:
            acc = sum_(product(acc, scalar), ev(x))
        return acc
    raise ValueError(tag)


def expr_degree(e):
    """protocol.rs:372-386"""
    tag = e[0]
    if tag in ("const", "challenge"):
        return 0
    if tag in ("identity", "lagrange", "poly"):
        return 1
    if tag in ("neg", "scaled"):
        return expr_degree(e[1])
    if tag == "sum":
        return max(expr_degree(e[1]), expr_degree(e[2]))
    if tag == "product":
        return expr_degree(e[1]) + expr_degree(e[2])
    if tag == "distribute_powers":
        return max([expr_degree(x) for x in e[1]] + [expr_degree(e[2])])
    raise ValueError(tag)


def _merge(a, b):
    if a is None:
        return b
    if b is None:
        return a
    return a | b


def expr_used_lagrange(e):
    """protocol.rs:388-403 (sorted like a BTreeSet)"""
    s = expr_evaluate(
        e,
        lambda _: None,
        lambda cp: {cp[1]} if cp[0] == "lagrange" else None,
        lambda _: None,
        lambda _: None,
        lambda a: a,
        _merge,
        _merge,
        lambda a, _: a,
    )
    return sorted(s or set())


def expr_used_query(e):
    """protocol.rs:405-417 (sorted by (poly, rotation) like BTreeSet<Query>)"""
    s = expr_evaluate(
        e,
        lambda _: None,
        lambda _: None,
        lambda q: {q},
        lambda _: None,
        lambda a: a,
        _merge,
        _merge,
        lambda a, _: a,
    )
    return sorted(s or set())


# ---------------------------------------------------------------- PlonkProtocol (protocol.rs:21-63)
@dataclass
class QuotientPolynomial:
    chunk_degree: int
    numerator: Any

    def num_chunk(self):
        """protocol.rs:288-293"""
        d = max(expr_degree(self.numerator) - 1, 0)
        return -(-d // self.chunk_degree)


@dataclass
class InstanceCommittingKey:
    bases: List[Any]
    constant: Optional[Any] = None


@dataclass
class PlonkProtocol:
    domain: Domain
    preprocessed: List[Any]  # G1 affine tuples
    num_instance: List[int]
    num_witness: List[int]
    num_challenge: List[int]
    evaluations: List[Tuple[int, int]]  # Query = (poly, rotation)
    queries: List[Tuple[int, int]]
    quotient: QuotientPolynomial
    transcript_initial_state: Optional[int] = None
    instance_committing_key: Optional[InstanceCommittingKey] = None
    linearization: Optional[str] = None  # None | "WithoutConstant" | "MinusVanishingTimesQuotient"
    accumulator_indices: List[List[Tuple[int, int]]] = field(default_factory=list)

    def langranges(self):
        """protocol.rs:70-98"""
        out = list(expr_used_lagrange(self.quotient.numerator))
        if self.instance_committing_key is None:
            offset = len(self.preprocessed)
            rng = range(offset, offset + len(self.num_instance))
            mn, mx = 0, 0
            for poly, rot in expr_used_query(self.quotient.numerator):
                if poly in rng:
                    if rot < mn:
                        mn = rot
                    elif rot > mx:
                        mx = rot
            max_instance_len = max(self.num_instance, default=0)
            out += list(range(-mx, max_instance_len + abs(mn)))
        return out


@dataclass
class PcsQuery:
    """pcs.rs:22-45"""

    poly: int
    shift: int
    eval: Any = None


# ---------------------------------------------------------------- CommonPolynomialEvaluation (protocol.rs:187-279)
class CommonPolynomialEvaluation:
    def __init__(self, domain, langranges, z):
        loader = z.loader
        self._zn = z.pow_const(domain.n)
        langranges = sorted(set(langranges))
        one = loader.load_one()
        self._zn_minus_one = self._zn - one
        self._zn_minus_one_inv = Fraction.one_over(self._zn_minus_one)
        n_inv = loader.load_const(domain.n_inv)
        numer = self._zn_minus_one * n_inv
        omegas = [loader.load_const(domain.rotate_scalar(1, i)) for i in langranges]
        evals = [Fraction(numer * omega, z - omega) for omega in omegas]
        self.identity = z
        self.lagrange = dict(zip(langranges, evals))
        self._order = langranges

    def zn(self):
        return self._zn

    def zn_minus_one(self):
        return self._zn_minus_one

    def zn_minus_one_inv(self):
        return self._zn_minus_one_inv.evaluated()

    def get(self, cp):
        if cp[0] == "identity":
            return self.identity
        return self.lagrange[cp[1]].evaluated()

    def denoms(self):
        refs = [self.lagrange[i].denom_mut() for i in self._order] + [self._zn_minus_one_inv.denom_mut()]
        return [r for r in refs if r is not None]

    def evaluate(self):
        for i in self._order:
            self.lagrange[i].evaluate()
        self._zn_minus_one_inv.evaluate()


# ---------------------------------------------------------------- PlonkProof (proof.rs)
class PlonkProof:
    def __init__(self):
        self.committed_instances = None
        self.witnesses = []
        self.challenges = []
        self.quotients = []
        self.z = None
        self.evaluations = []
        self.pcs = None
        self.old_accumulators = []

    @classmethod
    def read(cls, svk, protocol, instances, transcript, AS, AE=None):
        """proof.rs:52-153.  `instances` = list of lists of Scalar."""
        self = cls()
        loader = transcript.loader
        if protocol.transcript_initial_state is not None:
            transcript.common_scalar(loader.load_const(protocol.transcript_initial_state))
        if protocol.num_instance != [len(i) for i in instances]:
            raise VerifyError("InvalidInstances")
        if protocol.instance_committing_key is not None:
            ick = protocol.instance_committing_key
            bases = [loader.ec_point_load_const(b) for b in ick.bases]
            constant = None if ick.constant is None else loader.ec_point_load_const(ick.constant)
            committed = []
            for inst in instances:
                terms = [Msm.base(b) * s for s, b in zip(inst, bases)]
                if constant is not None:
                    terms.append(Msm.base(constant))
                committed.append(Msm.sum(terms).evaluate(None))
            for c in committed:
                transcript.common_ec_point(c)
            self.committed_instances = committed
        else:
            for inst in instances:
                for x in inst:
                    transcript.common_scalar(x)
        for n, m in zip(protocol.num_witness, protocol.num_challenge):
            self.witnesses += transcript.read_n_ec_points(n)
            self.challenges += transcript.squeeze_n_challenges(m)
        self.quotients = transcript.read_n_ec_points(protocol.quotient.num_chunk())
        self.z = transcript.squeeze_challenge()
        self.evaluations = transcript.read_n_scalars(len(protocol.evaluations))
        self.pcs = AS.read_proof(svk, cls.empty_queries(protocol), transcript)
        if AE is None and protocol.accumulator_indices:
            from .kzg import LimbsEncoding

            AE = LimbsEncoding(3, 88)  # the SDK's verifier type (snark-verifier-sdk/src/lib.rs:33-40)
        self.old_accumulators = [
            AE.from_repr([instances[i][j] for i, j in idx], loader) for idx in protocol.accumulator_indices
        ]
        return self

    @staticmethod
    def empty_queries(protocol):
        """proof.rs:156-165"""
        return [PcsQuery(poly, protocol.domain.rotate_scalar(1, rot)) for poly, rot in protocol.queries]

    def queries(self, protocol, evaluations):
        """proof.rs:167-177"""
        out = []
        for q, (poly, rot) in zip(self.empty_queries(protocol), protocol.queries):
            q.eval = evaluations.pop((poly, rot))
            out.append(q)
        return out

    def commitments(self, protocol, common_poly_eval, evaluations):
        """proof.rs:179-281"""
        loader = common_poly_eval.zn().loader
        commitments = [Msm.base(loader.ec_point_load_const(p)) for p in protocol.preprocessed]
        if self.committed_instances is not None:
            commitments += [Msm.base(c) for c in self.committed_instances]
        else:
            commitments += [Msm() for _ in protocol.num_instance]
        commitments += [Msm.base(w) for w in self.witnesses]

        def q_poly(query):
            if query in evaluations:
                return Msm.constant_(evaluations[query])
            if query[1] == 0 and query[0] < len(commitments):
                return commitments[query[0]].clone()
            raise VerifyError("InvalidProtocol", f"Missing query {query}")

        def q_challenge(i):
            if i < len(self.challenges):
                return Msm.constant_(self.challenges[i])
            raise VerifyError("InvalidProtocol", f"Missing challenge {i}")

        def q_product(a, b):
            if a.size() == 0:
                return b * a.try_into_constant()
            if b.size() == 0:
                return a * b.try_into_constant()
            raise VerifyError("InvalidProtocol", "Invalid linearization")

        numerator = expr_evaluate(
            protocol.quotient.numerator,
            lambda s: Msm.constant_(loader.load_const(s)),
            lambda cp: Msm.constant_(common_poly_eval.get(cp)),
            q_poly,
            q_challenge,
            lambda a: -a,
            lambda a, b: a + b,
            q_product,
            lambda a, s: a * loader.load_const(s),
        )

        quotient_query = (len(protocol.preprocessed) + len(protocol.num_instance) + len(self.witnesses), 0)
        coeffs = common_poly_eval.zn().pow_const(protocol.quotient.chunk_degree).powers(len(self.quotients))
        quotient = Msm.sum([Msm.base(chunk) * coeff for coeff, chunk in zip(coeffs, self.quotients)])
        lin = protocol.linearization
        if lin == "WithoutConstant":
            linearization_query = (quotient_query[0] + 1, 0)
            msm, constant = numerator.split()
            commitments.append(quotient)
            commitments.append(msm)
            c = constant if constant is not None else loader.load_zero()
            evaluations[quotient_query] = (c + evaluations[linearization_query]) * common_poly_eval.zn_minus_one_inv()
        elif lin == "MinusVanishingTimesQuotient":
            msm, constant = (numerator - quotient * common_poly_eval.zn_minus_one()).split()
            commitments.append(msm)
            evaluations[quotient_query] = constant if constant is not None else loader.load_zero()
        else:
            commitments.append(quotient)
            c = numerator.try_into_constant()
            if c is None:
                raise VerifyError("InvalidProtocol", "Invalid linearization")
            evaluations[quotient_query] = c * common_poly_eval.zn_minus_one_inv()
        return commitments

    def evaluations_map(self, protocol, instances, common_poly_eval):
        """proof.rs:283-318"""
        loader = common_poly_eval.zn().loader
        evals = {}
        if protocol.instance_committing_key is None:
            offset = len(protocol.preprocessed)
            rng = range(offset, offset + len(protocol.num_instance))
            for query in expr_used_query(protocol.quotient.numerator):
                if query[0] not in rng:
                    continue
                inst = instances[query[0] - offset]
                pairs = [
                    (x, common_poly_eval.get(("lagrange", -query[1] + k))) for k, x in enumerate(inst)
                ]
                evals[query] = loader.sum_products(pairs)
        for q, e in zip(protocol.evaluations, self.evaluations):
            evals[q] = e
        return evals


# ---------------------------------------------------------------- verifiers (plonk.rs)
class PlonkSuccinctVerifier:
    """plonk.rs:32-93; `AS` is a pcs class from oracle.kzg (KzgAsBdfg21 / KzgAsGwc19)."""

    @staticmethod
    def read_proof(svk, protocol, instances, transcript, AS, AE=None):
        return PlonkProof.read(svk, protocol, instances, transcript, AS, AE)

    @staticmethod
    def verify(svk, protocol, instances, proof, AS):
        cpe = CommonPolynomialEvaluation(protocol.domain, protocol.langranges(), proof.z)
        NativeLoader.batch_invert(cpe.denoms())
        cpe.evaluate()
        evaluations = proof.evaluations_map(protocol, instances, cpe)
        commitments = proof.commitments(protocol, cpe, evaluations)
        queries = proof.queries(protocol, evaluations)
        accumulator = AS.verify(svk, commitments, proof.z, queries, proof.pcs)
        return [accumulator] + list(proof.old_accumulators)


class PlonkVerifier:
    """plonk.rs:98-135; `dk` = oracle.kzg.KzgDecidingKey"""

    @staticmethod
    def read_proof(dk, protocol, instances, transcript, AS, AE=None):
        return PlonkProof.read(dk.svk, protocol, instances, transcript, AS, AE)

    @staticmethod
    def verify(dk, protocol, instances, proof, AS):
        accumulators = PlonkSuccinctVerifier.verify(dk.svk, protocol, instances, proof, AS)
        return AS.decide_all(dk, accumulators)
