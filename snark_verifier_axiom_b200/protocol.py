"""Host-side mirror of the reference's protocol description types, plus the flat serialization
`libsvk` ingests (`svk_protocol_compile`).

Mirrors (same names, same field meaning):
  PlonkProtocol, QuotientPolynomial, Query, Expression, CommonPolynomial, LinearizationStrategy
      snark-verifier/src/verifier/plonk/protocol.rs:21-63, 181-185, 282-319, 504-513
  Domain, Rotation                snark-verifier/src/util/arithmetic.rs:100-162
Field elements are Python ints (canonical), G1 points are (x, y) int tuples.

The byte layout written by `PlonkProtocol.to_bytes()` is parsed by csrc/compiler.h:parse_protocol;
it is this project's own wire format (the reference's serde/bincode `Snark` files are a "next" row,
SURVEY 8f-2).
"""
import struct
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

FR_MODULUS = 21888242871839275222246405745257275088548364400416034343698204186575808495617
FQ_MODULUS = 21888242871839275222246405745257275088696311157297823662689037894645226208583
_ROOT_OF_UNITY = pow(7, (FR_MODULUS - 1) >> 28, FR_MODULUS)


def _fe(v: int) -> bytes:
    return int(v % FR_MODULUS).to_bytes(32, "little")


def root_of_unity(k: int) -> int:
    """util/arithmetic.rs:89-96"""
    assert k <= 28
    return pow(_ROOT_OF_UNITY, 1 << (28 - k), FR_MODULUS)


@dataclass
class Domain:
    """util/arithmetic.rs:131-162"""

    k: int
    gen: int

    @classmethod
    def new(cls, k: int, gen: Optional[int] = None) -> "Domain":
        return cls(k, root_of_unity(k) if gen is None else gen)

    @property
    def n(self) -> int:
        return 1 << self.k

    @property
    def n_inv(self) -> int:
        return pow(self.n, FR_MODULUS - 2, FR_MODULUS)

    @property
    def gen_inv(self) -> int:
        return pow(self.gen, FR_MODULUS - 2, FR_MODULUS)

    def rotate_scalar(self, scalar: int, rotation: int) -> int:
        if rotation == 0:
            return scalar % FR_MODULUS
        if rotation > 0:
            return scalar * pow(self.gen, rotation, FR_MODULUS) % FR_MODULUS
        return scalar * pow(self.gen_inv, -rotation, FR_MODULUS) % FR_MODULUS


@dataclass(frozen=True, order=True)
class Query:
    """protocol.rs:296-306"""

    poly: int
    rotation: int = 0


class Expression:
    """protocol.rs:308-319.  Build with the static constructors; `+ - *` and unary `-` compose
    like the reference's operator impls (protocol.rs:432-478)."""

    __slots__ = ("tag", "args")
    CONSTANT, IDENTITY, LAGRANGE, POLYNOMIAL, CHALLENGE, NEGATED, SUM, PRODUCT, SCALED, DISTRIBUTE_POWERS = range(10)

    def __init__(self, tag, *args):
        self.tag = tag
        self.args = args

    @staticmethod
    def Constant(v):
        return Expression(Expression.CONSTANT, int(v) % FR_MODULUS)

    @staticmethod
    def CommonPolynomialIdentity():
        return Expression(Expression.IDENTITY)

    @staticmethod
    def CommonPolynomialLagrange(i):
        return Expression(Expression.LAGRANGE, int(i))

    @staticmethod
    def Polynomial(query: Query):
        return Expression(Expression.POLYNOMIAL, query)

    @staticmethod
    def Challenge(i):
        return Expression(Expression.CHALLENGE, int(i))

    @staticmethod
    def Negated(a):
        return Expression(Expression.NEGATED, a)

    @staticmethod
    def Sum(a, b):
        return Expression(Expression.SUM, a, b)

    @staticmethod
    def Product(a, b):
        return Expression(Expression.PRODUCT, a, b)

    @staticmethod
    def Scaled(a, v):
        return Expression(Expression.SCALED, a, int(v) % FR_MODULUS)

    @staticmethod
    def DistributePowers(exprs, base):
        return Expression(Expression.DISTRIBUTE_POWERS, list(exprs), base)

    def __add__(self, o):
        return Expression.Sum(self, o)

    def __sub__(self, o):
        return Expression.Sum(self, Expression.Negated(o))

    def __mul__(self, o):
        if isinstance(o, int):
            return Expression.Scaled(self, o)
        return Expression.Product(self, o)

    def __neg__(self):
        return Expression.Negated(self)

    def degree(self) -> int:
        """protocol.rs:372-386"""
        t, a = self.tag, self.args
        if t in (self.CONSTANT, self.CHALLENGE):
            return 0
        if t in (self.IDENTITY, self.LAGRANGE, self.POLYNOMIAL):
            return 1
        if t in (self.NEGATED, self.SCALED):
            return a[0].degree()
        if t == self.SUM:
            return max(a[0].degree(), a[1].degree())
        if t == self.PRODUCT:
            return a[0].degree() + a[1].degree()
        return max([e.degree() for e in a[0]] + [a[1].degree()])

    def to_bytes(self) -> bytes:
        t, a = self.tag, self.args
        out = bytes([t])
        if t == self.CONSTANT:
            return out + _fe(a[0])
        if t == self.IDENTITY:
            return out
        if t == self.LAGRANGE:
            return out + struct.pack("<i", a[0])
        if t == self.POLYNOMIAL:
            return out + struct.pack("<Ii", a[0].poly, a[0].rotation)
        if t == self.CHALLENGE:
            return out + struct.pack("<I", a[0])
        if t == self.NEGATED:
            return out + a[0].to_bytes()
        if t in (self.SUM, self.PRODUCT):
            return out + a[0].to_bytes() + a[1].to_bytes()
        if t == self.SCALED:
            return out + a[0].to_bytes() + _fe(a[1])
        exprs, base = a
        return out + struct.pack("<I", len(exprs)) + b"".join(e.to_bytes() for e in exprs) + base.to_bytes()


@dataclass
class QuotientPolynomial:
    """protocol.rs:281-294"""

    chunk_degree: int
    numerator: Expression

    def num_chunk(self) -> int:
        d = max(self.numerator.degree() - 1, 0)
        return -(-d // self.chunk_degree)


class LinearizationStrategy:
    """protocol.rs:503-513"""

    WithoutConstant = 1
    MinusVanishingTimesQuotient = 2


@dataclass
class PlonkProtocol:
    """protocol.rs:21-63 (`PlonkProtocol<G1Affine, NativeLoader>`)"""

    domain: Domain
    preprocessed: List[Tuple[int, int]]
    num_instance: List[int]
    num_witness: List[int]
    num_challenge: List[int]
    evaluations: List[Query]
    queries: List[Query]
    quotient: QuotientPolynomial
    transcript_initial_state: Optional[int] = None
    instance_committing_key: Optional[object] = None
    linearization: Optional[int] = None
    accumulator_indices: List[List[Tuple[int, int]]] = field(default_factory=list)
    accumulator_encoding: Tuple[int, int] = (3, 88)  # `LimbsEncoding<LIMBS, BITS>` of the verifier type (sdk/lib.rs:33-40)

    def to_bytes(self) -> bytes:
        out = bytearray()
        out += struct.pack("<III", 0x504B5653, 1, self.domain.k)
        out += _fe(self.domain.gen)
        out += struct.pack("<I", len(self.preprocessed))
        for pt in self.preprocessed:
            if pt is None:
                out += bytes(64)
            else:
                out += int(pt[0]).to_bytes(32, "little") + int(pt[1]).to_bytes(32, "little")
        for lst in (self.num_instance, self.num_witness, self.num_challenge):
            out += struct.pack("<I", len(lst))
            for v in lst:
                out += struct.pack("<I", v)
        for lst in (self.evaluations, self.queries):
            out += struct.pack("<I", len(lst))
            for q in lst:
                out += struct.pack("<Ii", q.poly, q.rotation)
        out += struct.pack("<I", self.quotient.chunk_degree)
        out += self.quotient.numerator.to_bytes()
        if self.transcript_initial_state is None:
            out += b"\x00"
        else:
            out += b"\x01" + _fe(self.transcript_initial_state)
        out += b"\x01" if self.instance_committing_key is not None else b"\x00"
        out += bytes([self.linearization or 0])
        out += struct.pack("<I", len(self.accumulator_indices))
        for idx in self.accumulator_indices:
            out += struct.pack("<I", len(idx))
            for i, j in idx:
                out += struct.pack("<II", i, j)
        out += bytes(self.accumulator_encoding)
        return bytes(out)
