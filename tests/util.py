"""Shared helpers for the test-suite: byte packing at the C ABI (include/svk.h)."""
import ctypes

import numpy as np

from oracle import bn254


def fe(v):
    return int(v).to_bytes(32, "little")


def g1_bytes(pt):
    return bytes(64) if pt is None else fe(pt[0]) + fe(pt[1])


def g1_from(b):
    x = int.from_bytes(b[:32], "little")
    y = int.from_bytes(b[32:64], "little")
    return None if x == 0 and y == 0 else (x, y)


def g2_bytes(q):
    return fe(q[0][0]) + fe(q[0][1]) + fe(q[1][0]) + fe(q[1][1])


def dk_bytes(dk):
    return g1_bytes(dk.svk.g) + g2_bytes(dk.g2) + g2_bytes(dk.s_g2)


def acc_bytes(lhs, rhs):
    return g1_bytes(lhs) + g1_bytes(rhs)


def np_u8(b):
    return np.frombuffer(bytes(b), dtype=np.uint8).copy()


def ptr(arr):
    return arr.ctypes.data_as(ctypes.c_void_p)


def to_product_protocol(op):
    """oracle.plonk.PlonkProtocol (tuple expressions) -> snark_verifier_axiom_b200.protocol.PlonkProtocol"""
    from snark_verifier_axiom_b200 import protocol as pp

    E = pp.Expression

    def conv(e):
        t = e[0]
        if t == "const":
            return E.Constant(e[1])
        if t == "identity":
            return E.CommonPolynomialIdentity()
        if t == "lagrange":
            return E.CommonPolynomialLagrange(e[1])
        if t == "poly":
            return E.Polynomial(pp.Query(e[1], e[2]))
        if t == "challenge":
            return E.Challenge(e[1])
        if t == "neg":
            return E.Negated(conv(e[1]))
        if t == "sum":
            return E.Sum(conv(e[1]), conv(e[2]))
        if t == "product":
            return E.Product(conv(e[1]), conv(e[2]))
        if t == "scaled":
            return E.Scaled(conv(e[1]), e[2])
        if t == "distribute_powers":
            return E.DistributePowers([conv(x) for x in e[1]], conv(e[2]))
        raise ValueError(t)

    lin = {None: None, "WithoutConstant": 1, "MinusVanishingTimesQuotient": 2}[op.linearization]
    return pp.PlonkProtocol(
        domain=pp.Domain(op.domain.k, op.domain.gen),
        preprocessed=list(op.preprocessed),
        num_instance=list(op.num_instance),
        num_witness=list(op.num_witness),
        num_challenge=list(op.num_challenge),
        evaluations=[pp.Query(p, r) for p, r in op.evaluations],
        queries=[pp.Query(p, r) for p, r in op.queries],
        quotient=pp.QuotientPolynomial(op.quotient.chunk_degree, conv(op.quotient.numerator)),
        transcript_initial_state=op.transcript_initial_state,
        instance_committing_key=op.instance_committing_key,
        linearization=lin,
        accumulator_indices=[list(x) for x in op.accumulator_indices],
    )


def lookup_two_phase_shape():
    """A circuit shape exercising what StandardPlonk does not (system/halo2.rs:199-243, 372-408, 419-449, 593-655):
    two advice phases with a user challenge between them, rotations -1 / +1, a lookup argument with two compressed
    expressions, an instance column inside the permutation argument, and enough permutation columns for two grand-product
    polynomials (degree 5 => chunks of 3)."""
    from oracle.halo2_system import ConstraintSystemShape

    def S(x, y):
        return ("sum", x, y)

    def M(x, y):
        return ("product", x, y)

    a, b, c, d = (("advice", i, 0) for i in range(4))
    a_next, c_prev = ("advice", 0, 1), ("advice", 2, -1)
    q, t0, q_lookup = (("fixed", i, 0) for i in range(3))
    inst = ("instance", 0, 0)
    ch = ("challenge", 0)
    gates = [
        M(q, S(M(a, b), ("neg", c))),
        M(q, S(d, ("neg", M(ch, a_next)))),
        S(("scaled", M(q, c_prev), 7), S(M(M(ch, ch), b), inst)),
    ]
    return ConstraintSystemShape(
        num_fixed=3,
        num_advice=4,
        num_instance_columns=1,
        permutation_columns=[("advice", 0), ("advice", 1), ("advice", 2), ("instance", 0)],
        advice_queries=[(0, 0), (1, 0), (2, 0), (3, 0), (0, 1), (2, -1)],
        fixed_queries=[(0, 0), (1, 0), (2, 0)],
        instance_queries=[(0, 0)],
        gates=gates,
        degree=5,
        blinding_factors=5,
        advice_column_phase=[0, 0, 1, 1],
        challenge_phase=[0],
        lookups=[([M(q_lookup, a), M(q_lookup, b)], [t0, M(t0, t0)])],
    )
