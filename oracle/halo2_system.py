"""halo2 `VerifyingKey` -> `PlonkProtocol` front-end -- oracle restatement.  TEST INFRASTRUCTURE ONLY.

Follows snark-verifier/src/system/halo2.rs:82-156 (`compile`), :164-684 (`Polynomials`): gates, permutation argument,
lookup arguments, multi-phase advice / challenges, `num_proof` > 1, accumulator indices (zk = true, the only mode the
reference implements).  halo2's `ConstraintSystem` itself (halo2_proofs, Cargo.lock:1751-1753) is not
in the tree; its derived quantities (`degree()`, `blinding_factors()`, query lists) are passed in
as a small `ConstraintSystemShape` and were derived by hand for StandardPlonk.
"""
from dataclasses import dataclass, field
from typing import Any, List, Tuple

from .bn254 import FR_DELTA, R
from .loader import Domain
from .plonk import (
    Challenge,
    Const,
    DistributePowers,
    Identity,
    Lagrange,
    PlonkProtocol,
    Poly,
    Product,
    QuotientPolynomial,
    Scaled,
    Sub,
    Sum,
    expr_sum,
)


@dataclass
class ConstraintSystemShape:
    """The facts `Polynomials::new` reads from halo2's `ConstraintSystem` (system/halo2.rs:183-243)."""

    num_fixed: int
    num_advice: int
    num_instance_columns: int
    permutation_columns: List[Tuple[str, int]]  # ("advice"|"fixed"|"instance", index)
    advice_queries: List[Tuple[int, int]]  # (column, rotation)
    fixed_queries: List[Tuple[int, int]]
    instance_queries: List[Tuple[int, int]]
    gates: List[Any]  # expressions over ("fixed", i, rot) / ("advice", i, rot) / ("instance", i, rot) / ("challenge", i)
    degree: int
    blinding_factors: int
    advice_column_phase: List[int] = field(default_factory=list)  # per advice column; [] = all first phase
    challenge_phase: List[int] = field(default_factory=list)      # per user challenge (`cs.challenge_phase()`)
    lookups: List[Tuple[List[Any], List[Any]]] = field(default_factory=list)  # (input_expressions, table_expressions)


def _remapping(phase, num_phase):
    """system/halo2.rs:199-213"""
    num = [0] * num_phase
    index = []
    for ph in phase:
        index.append(num[ph])
        num[ph] += 1
    return num, index


class Polynomials:
    """system/halo2.rs:164-668 with zk = true (the only mode the reference implements, :189)."""

    def __init__(self, cs: ConstraintSystemShape, query_instance: bool, num_instance: List[int], num_proof: int = 1):
        self.cs = cs
        self.zk = True
        self.query_instance = query_instance
        assert num_proof > 0
        self.num_proof = num_proof
        self.num_fixed = cs.num_fixed
        self.num_permutation_fixed = len(cs.permutation_columns)
        self._num_instance = list(num_instance)
        adv_phase = list(cs.advice_column_phase) or [0] * cs.num_advice
        assert len(adv_phase) == cs.num_advice
        num_phase = max(adv_phase, default=0) + 1
        self.advice_phase = adv_phase
        self.num_advice, self.advice_index = _remapping(adv_phase, num_phase)
        self._num_challenge, self.challenge_index = _remapping(list(cs.challenge_phase), num_phase)
        self.num_lookup_permuted = 2 * len(cs.lookups)
        self.permutation_chunk_size = cs.degree - 2  # zk => degree - 2 (:190-196)
        self.num_permutation_z = -(-len(cs.permutation_columns) // self.permutation_chunk_size)
        self.num_lookup_z = len(cs.lookups)

    def num_preprocessed(self):
        return self.num_fixed + self.num_permutation_fixed

    def num_instance(self):
        return list(self._num_instance) * self.num_proof

    def num_witness(self):
        """:250-258"""
        return [self.num_proof * n for n in self.num_advice] + [
            self.num_proof * self.num_lookup_permuted,
            self.num_proof * (self.num_permutation_z + self.num_lookup_z) + 1,
        ]

    def num_challenge(self):
        """:260-270"""
        nc = list(self._num_challenge)
        nc[-1] += 1  # theta
        return nc + [2, 1]

    def instance_offset(self):
        return self.num_preprocessed()

    def witness_offset(self):
        return self.instance_offset() + len(self.num_instance())

    def cs_witness_offset(self):
        return self.witness_offset() + sum(self.num_witness()[: len(self.num_advice)])

    def query(self, column_type, column_index, rotation, t=0):
        """:284-304"""
        if column_type == "fixed":
            offset = 0
        elif column_type == "instance":
            offset = self.instance_offset() + t * len(self._num_instance)
        else:
            phase = self.advice_phase[column_index]
            column_index = self.advice_index[column_index]
            phase_offset = self.num_proof * sum(self.num_advice[:phase])
            offset = self.witness_offset() + phase_offset + t * self.num_advice[phase]
        return (offset + column_index, rotation)

    def instance_queries(self, t=0):
        if not self.query_instance:
            return []
        return [self.query("instance", c, r, t) for c, r in self.cs.instance_queries]

    def advice_queries(self, t=0):
        return [self.query("advice", c, r, t) for c, r in self.cs.advice_queries]

    def fixed_queries(self):
        return [self.query("fixed", c, r) for c, r in self.cs.fixed_queries]

    def permutation_fixed_queries(self):
        return [(self.num_fixed + i, 0) for i in range(self.num_permutation_fixed)]

    def permutation_poly(self, t, i):
        z_offset = self.cs_witness_offset() + self.num_witness()[len(self.num_advice)]
        return z_offset + t * self.num_permutation_z + i

    def rotation_last(self):
        return -(self.cs.blinding_factors + 1)

    def permutation_z_queries(self, eval_order, t=0):
        """:336-370 (zk = true)"""
        n = self.num_permutation_z
        out = []
        if eval_order:
            for i in range(n):
                z = self.permutation_poly(t, i)
                out += [(z, 0), (z, 1)]
                if i != n - 1:
                    out.append((z, self.rotation_last()))
        else:
            for i in range(n):
                z = self.permutation_poly(t, i)
                out += [(z, 0), (z, 1)]
            for i in reversed(range(n - 1)):
                out.append((self.permutation_poly(t, i), self.rotation_last()))
        return out

    def lookup_poly(self, t, i):
        """:372-381"""
        permuted_offset = self.cs_witness_offset()
        z_offset = permuted_offset + self.num_witness()[len(self.num_advice)] + self.num_proof * self.num_permutation_z
        z = z_offset + t * self.num_lookup_z + i
        permuted_input = permuted_offset + 2 * (t * self.num_lookup_z + i)
        return z, permuted_input, permuted_input + 1

    def lookup_queries(self, eval_order, t=0):
        """:383-408"""
        out = []
        for i in range(self.num_lookup_z):
            z, pi, pt = self.lookup_poly(t, i)
            if eval_order:
                out += [(z, 0), (z, 1), (pi, 0), (pi, -1), (pt, 0)]
            else:
                out += [(z, 0), (pi, 0), (pt, 0), (pi, -1), (z, 1)]
        return out

    def quotient_query(self):
        return (self.witness_offset() + sum(self.num_witness()), 0)

    def random_query(self):
        return (self.witness_offset() + sum(self.num_witness()) - 1, 0)

    def convert(self, e, t=0):
        """:419-449"""
        tag = e[0]
        if tag == "const":
            return Const(e[1] % R)
        if tag in ("fixed", "advice", "instance"):
            return Poly(*self.query(tag, e[1], e[2], t))
        if tag == "challenge":
            phase = self.cs.challenge_phase[e[1]]
            return Challenge(sum(self._num_challenge[:phase]) + self.challenge_index[e[1]])
        if tag == "neg":
            return ("neg", self.convert(e[1], t))
        if tag == "sum":
            return Sum(self.convert(e[1], t), self.convert(e[2], t))
        if tag == "product":
            return Product(self.convert(e[1], t), self.convert(e[2], t))
        if tag == "scaled":
            return Scaled(self.convert(e[1], t), e[2] % R)
        raise ValueError(tag)

    def l_last(self):
        return Lagrange(self.rotation_last())

    def l_blind(self):
        return expr_sum([Lagrange(i) for i in range(self.rotation_last() + 1, 0)])

    def l_active(self):
        return Sub(Sub(Const(1), self.l_last()), self.l_blind())

    def system_challenge_offset(self):
        nc = self.num_challenge()
        return sum(nc[: len(nc) - 3])

    def theta(self):
        return Challenge(self.system_challenge_offset())

    def beta(self):
        return Challenge(self.system_challenge_offset() + 1)

    def gamma(self):
        return Challenge(self.system_challenge_offset() + 2)

    def alpha(self):
        return Challenge(self.system_challenge_offset() + 3)

    def permutation_constraints(self, t=0):
        """:501-591 (zk = true)"""
        one = Const(1)
        l_0 = Lagrange(0)
        l_last = self.l_last()
        l_active = self.l_active()
        identity = Identity()
        beta, gamma = self.beta(), self.gamma()
        polys = [Poly(*self.query(ct, ci, 0, t)) for ct, ci in self.cs.permutation_columns]
        permutation_fixeds = [Poly(self.num_fixed + i, 0) for i in range(self.num_permutation_fixed)]
        zs = []
        for i in range(self.num_permutation_z):
            z = self.permutation_poly(t, i)
            zs.append((Poly(z, 0), Poly(z, 1), Poly(z, self.rotation_last())))
        out = []
        if zs:
            out.append(Product(l_0, Sub(one, zs[0][0])))
            z_l = zs[-1][0]
            out.append(Product(l_last, Sub(Product(z_l, z_l), z_l)))
        for (z, _, _), (_, _, z_prev_last) in zip(zs[1:], zs):
            out.append(Product(l_0, Sub(z, z_prev_last)))
        cs = self.permutation_chunk_size
        for i, (z, z_omega, _) in enumerate(zs):
            pchunk = polys[i * cs : (i + 1) * cs]
            fchunk = permutation_fixeds[i * cs : (i + 1) * cs]
            acc = None
            for poly, pf in zip(pchunk, fchunk):
                term = Sum(Sum(poly, Product(beta, pf)), gamma)
                acc = term if acc is None else Product(acc, term)
            left = Product(z_omega, acc)
            acc = None
            delta = pow(FR_DELTA, i * cs, R)
            for poly in pchunk:
                term = Sum(Sum(poly, Product(Product(beta, Const(delta)), identity)), gamma)
                acc = term if acc is None else Product(acc, term)
                delta = delta * FR_DELTA % R
            right = Product(z, acc)
            out.append(Product(l_active, Sub(left, right)))
        return out

    def lookup_constraints(self, t=0):
        """:593-655 (zk = true)"""
        one = Const(1)
        l_0 = Lagrange(0)
        l_last = self.l_last()
        l_active = self.l_active()
        beta, gamma = self.beta(), self.gamma()
        out = []
        for i, (input_exprs, table_exprs) in enumerate(self.cs.lookups):
            zq, pi, pt = self.lookup_poly(t, i)
            z, z_omega = Poly(zq, 0), Poly(zq, 1)
            permuted_input, permuted_input_omega_inv, permuted_table = Poly(pi, 0), Poly(pi, -1), Poly(pt, 0)
            inp = DistributePowers([self.convert(e, t) for e in input_exprs], self.theta())
            table = DistributePowers([self.convert(e, t) for e in table_exprs], self.theta())
            out.append(Product(l_0, Sub(one, z)))
            out.append(Product(l_last, Sub(Product(z, z), z)))
            out.append(
                Product(
                    l_active,
                    Sub(
                        Product(Product(z_omega, Sum(permuted_input, beta)), Sum(permuted_table, gamma)),
                        Product(Product(z, Sum(inp, beta)), Sum(table, gamma)),
                    ),
                )
            )
            out.append(Product(l_0, Sub(permuted_input, permuted_table)))
            out.append(
                Product(Product(l_active, Sub(permuted_input, permuted_table)), Sub(permuted_input, permuted_input_omega_inv))
            )
        return out

    def quotient(self):
        """:657-668"""
        constraints = []
        for t in range(self.num_proof):
            constraints += [self.convert(g, t) for g in self.cs.gates]
            constraints += self.permutation_constraints(t)
            constraints += self.lookup_constraints(t)
        return QuotientPolynomial(1, DistributePowers(constraints, self.alpha()))

    def accumulator_indices(self, accumulator_indices):
        """:670-684"""
        return [[(poly + t * len(self._num_instance), row) for poly, row in accumulator_indices] for t in range(self.num_proof)]


def compile_protocol(k, cs: ConstraintSystemShape, preprocessed, transcript_initial_state, num_instance,
                     query_instance=False, num_proof=1, accumulator_indices=None):
    """`compile(params, vk, Config::kzg().with_num_instance(..).with_num_proof(..).with_accumulator_indices(..))`
    (system/halo2.rs:82-156)."""
    assert len(preprocessed) == cs.num_fixed + len(cs.permutation_columns)
    assert not query_instance, "instance_committing_key needs the SRS Lagrange basis (KZG path never sets it, SURVEY App. A)"
    domain = Domain(k)
    p = Polynomials(cs, query_instance, num_instance, num_proof)
    T = range(num_proof)
    evaluations = (
        [q for t in T for q in p.instance_queries(t)]
        + [q for t in T for q in p.advice_queries(t)]
        + p.fixed_queries()
        + [p.random_query()]
        + p.permutation_fixed_queries()
        + [q for t in T for q in p.permutation_z_queries(True, t)]
        + [q for t in T for q in p.lookup_queries(True, t)]
    )
    queries = (
        [
            q
            for t in T
            for q in p.instance_queries(t) + p.advice_queries(t) + p.permutation_z_queries(False, t) + p.lookup_queries(False, t)
        ]
        + p.fixed_queries()
        + p.permutation_fixed_queries()
        + [p.quotient_query()]
        + [p.random_query()]
    )
    return PlonkProtocol(
        domain=domain,
        preprocessed=list(preprocessed),
        num_instance=p.num_instance(),
        num_witness=p.num_witness(),
        num_challenge=p.num_challenge(),
        evaluations=evaluations,
        queries=queries,
        quotient=p.quotient(),
        transcript_initial_state=transcript_initial_state,
        instance_committing_key=None,
        linearization=None,
        accumulator_indices=p.accumulator_indices(accumulator_indices) if accumulator_indices else [],
    )


def standard_plonk_shape():
    """The `StandardPlonk` circuit of snark-verifier/examples/evm-verifier.rs:47-74,102-105
    (same in snark-verifier-sdk/benches/standard_plonk.rs:48-109): advice a,b,c (equality enabled),
    fixed q_a,q_b,q_c,q_ab,constant, one instance column, `set_minimum_degree(4)`.
    halo2: degree() = 4, blinding_factors() = 5."""
    a, b, c = (("advice", i, 0) for i in range(3))
    q_a, q_b, q_c, q_ab, constant = (("fixed", i, 0) for i in range(5))
    instance = ("instance", 0, 0)

    def S(x, y):
        return ("sum", x, y)

    def M(x, y):
        return ("product", x, y)

    gate = S(S(S(S(S(M(q_a, a), M(q_b, b)), M(q_c, c)), M(M(q_ab, a), b)), constant), instance)
    return ConstraintSystemShape(
        num_fixed=5,
        num_advice=3,
        num_instance_columns=1,
        permutation_columns=[("advice", 0), ("advice", 1), ("advice", 2)],
        advice_queries=[(0, 0), (1, 0), (2, 0)],
        fixed_queries=[(i, 0) for i in range(5)],
        instance_queries=[(0, 0)],
        gates=[gate],
        degree=4,
        blinding_factors=5,
    )


def standard_plonk_protocol(k, preprocessed, transcript_initial_state):
    return compile_protocol(k, standard_plonk_shape(), preprocessed, transcript_initial_state, [1])
