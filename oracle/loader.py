"""`Loader` abstraction, `NativeLoader`, `Fraction`, `Domain`, `Msm` -- oracle restatement.
TEST INFRASTRUCTURE ONLY.

Follows
  snark-verifier/src/loader.rs:22-260        (LoadedScalar / ScalarLoader / EcPointLoader defaults)
  snark-verifier/src/loader/native.rs:13-88  (NativeLoader: arithmetic = halo2curves Fr / G1Affine)
  snark-verifier/src/util/arithmetic.rs:131-234 (Domain, Fraction)
  snark-verifier/src/util/msm.rs:20-205      (Msm)

A `Scalar` carries a concrete value (Python int mod r).  If the loader has a `Tracer`, every
primitive operation NativeLoader would execute is also appended to the tracer's op list, so the C
oracle (`oracle/c`) can replay exactly the reference's operation sequence for the CPU baseline.
"""
from . import bn254
from .bn254 import R


class Tracer:
    """Records the primitive-op sequence NativeLoader performs (scalar regs `s<i>`, point regs `p<i>`)."""

    def __init__(self):
        self.ops = []
        self.n_s = 0
        self.n_p = 0
        self.consts = {}
        self.const_points = {}

    def new_s(self):
        self.n_s += 1
        return self.n_s - 1

    def new_p(self):
        self.n_p += 1
        return self.n_p - 1

    def emit(self, *op):
        self.ops.append(op)


class Scalar:
    """`LoadedScalar` of NativeLoader (= halo2curves Fr)."""

    __slots__ = ("v", "loader", "reg")

    def __init__(self, v, loader, reg=None):
        self.v = v % R
        self.loader = loader
        self.reg = reg

    def _bin(self, other, op, v):
        tr = self.loader.tracer
        reg = None
        if tr is not None:
            reg = tr.new_s()
            tr.emit(op, reg, self.reg, other.reg)
        return Scalar(v, self.loader, reg)

    def __add__(self, o):
        return self._bin(o, "add", self.v + o.v)

    def __sub__(self, o):
        return self._bin(o, "sub", self.v - o.v)

    def __mul__(self, o):
        return self._bin(o, "mul", self.v * o.v)

    def __neg__(self):
        tr = self.loader.tracer
        reg = None
        if tr is not None:
            reg = tr.new_s()
            tr.emit("neg", reg, self.reg)
        return Scalar(-self.v, self.loader, reg)

    def __eq__(self, o):
        return isinstance(o, Scalar) and self.v == o.v

    def __hash__(self):
        return hash(self.v)

    def __repr__(self):
        return f"Fr({hex(self.v)})"

    # loader.rs:32-79
    def square(self):
        return self * self

    def invert(self):
        """`FieldOps::invert` -> Option (native.rs:29-33): Fermat inversion, None for 0."""
        inv = bn254.fr_inv(self.v)
        tr = self.loader.tracer
        reg = None
        if tr is not None:
            reg = tr.new_s()
            tr.emit("inv", reg, self.reg)  # replay computes x^(r-2); 0 -> 0 == `unwrap_or_else(|| value.clone())`
        if inv is None:
            return None
        return Scalar(inv, self.loader, reg)

    def pow_const(self, exp):
        """loader.rs:49-68"""
        assert exp > 0
        base = self
        while exp & 1 == 0:
            base = base.square()
            exp >>= 1
        acc = base
        while exp > 1:
            exp >>= 1
            base = base.square()
            if exp & 1 == 1:
                acc = acc * base
        return acc

    def powers(self, n):
        """loader.rs:71-78: [1, x, x^2, .., x^(n-1)]"""
        assert n >= 1  # `take(n - 1)` panics on usize underflow for n = 0
        out = [self.loader.load_one()]
        cur = self
        for _ in range(n - 1):
            out.append(cur)
            cur = cur * self  # `iter::successors` computes the successor eagerly (one unused product)
        return out


class EcPoint:
    """`LoadedEcPoint` of NativeLoader (= G1Affine). `pt` is an affine tuple or None (identity)."""

    __slots__ = ("pt", "loader", "reg", "dlog")

    def __init__(self, pt, loader, reg=None, dlog=None):
        self.pt = pt
        self.loader = loader
        self.reg = reg
        self.dlog = dlog  # only used by oracle.forge (trapdoor fixtures)

    def __eq__(self, o):
        return isinstance(o, EcPoint) and self.pt == o.pt

    def __hash__(self):
        return hash(self.pt)

    def __repr__(self):
        return f"G1({self.pt})"


class Ref:
    """A `&mut LoadedScalar` (what `Fraction::denom_mut` hands to `batch_invert`)."""

    __slots__ = ("obj", "attr")

    def __init__(self, obj, attr):
        self.obj, self.attr = obj, attr

    def get(self):
        return getattr(self.obj, self.attr)

    def set(self, v):
        setattr(self.obj, self.attr, v)


class NativeLoader:
    def __init__(self, tracer=None):
        self.tracer = tracer

    # ---- ScalarLoader (loader.rs:116-249)
    def load_const(self, value):
        value %= R
        tr = self.tracer
        reg = None
        if tr is not None:
            if value not in tr.consts:
                tr.consts[value] = tr.new_s()
            reg = tr.consts[value]
        return Scalar(value, self, reg)

    def load_input(self, value, index):
        """A per-proof input (an instance value): a plain Fr for NativeLoader; under tracing a fresh
        register filled from input slot `index` by the replayer."""
        tr = self.tracer
        reg = None
        if tr is not None:
            reg = tr.new_s()
            tr.emit("input", reg, index)
        return Scalar(value, self, reg)

    def load_zero(self):
        return self.load_const(0)

    def load_one(self):
        return self.load_const(1)

    def sum_with_coeff_and_const(self, values, constant):
        """loader.rs:135-160; `values` = [(coeff:int, Scalar)]"""
        constant %= R
        if not values:
            return self.load_const(constant)
        terms = []
        if constant != 0:
            terms.append(self.load_const(constant))
        for coeff, value in values:
            coeff %= R
            terms.append(value if coeff == 1 else self.load_const(coeff) * value)
        acc = terms[0]
        for t in terms[1:]:
            acc = acc + t
        return acc

    def sum_products_with_coeff_and_const(self, values, constant):
        """loader.rs:163-185; `values` = [(coeff:int, Scalar, Scalar)]"""
        constant %= R
        if not values:
            return self.load_const(constant)
        terms = []
        if constant != 0:
            terms.append(self.load_const(constant))
        for coeff, lhs, rhs in values:
            coeff %= R
            terms.append(lhs * rhs if coeff == 1 else self.load_const(coeff) * lhs * rhs)
        acc = terms[0]
        for t in terms[1:]:
            acc = acc + t
        return acc

    def sum_with_coeff(self, values):
        return self.sum_with_coeff_and_const(values, 0)

    def sum_with_const(self, values, constant):
        return self.sum_with_coeff_and_const([(1, v) for v in values], constant)

    def sum(self, values):
        return self.sum_with_const(values, 0)

    def sum_products_with_coeff(self, values):
        return self.sum_products_with_coeff_and_const(values, 0)

    def sum_products_with_const(self, values, constant):
        return self.sum_products_with_coeff_and_const([(1, a, b) for a, b in values], constant)

    def sum_products(self, values):
        return self.sum_products_with_const(values, 0)

    def product(self, values):
        acc = self.load_one()
        for v in values:
            acc = acc * v
        return acc

    @staticmethod
    def batch_invert(refs):
        """loader.rs:241-248 -- NOT batched natively: one inversion per element, 0 stays 0."""
        for ref in refs:
            v = ref.get()
            inv = v.invert()
            if inv is None:
                # `unwrap_or_else(|| value.clone())`; under tracing the replayed x^(r-2) of 0 is 0 too
                tr = v.loader.tracer
                inv = Scalar(0, v.loader, tr.n_s - 1) if tr is not None else v
            ref.set(inv)

    # ---- EcPointLoader (loader.rs:82-113, native.rs:44-72)
    def ec_point_load_const(self, pt):
        tr = self.tracer
        reg = None
        if tr is not None:
            if pt not in tr.const_points:
                tr.const_points[pt] = tr.new_p()
            reg = tr.const_points[pt]
        return EcPoint(pt, self, reg)

    def multi_scalar_multiplication(self, pairs):
        """native.rs:61-71; pairs = [(Scalar, EcPoint)] -> EcPoint (affine)."""
        assert pairs, "pairs should not be empty"
        res = bn254.g1_msm_naive([(s.v, b.pt) for s, b in pairs])
        tr = self.tracer
        reg = None
        if tr is not None:
            reg = tr.new_p()
            tr.emit("msm", reg, [(s.reg, b.reg) for s, b in pairs])
        return EcPoint(res, self, reg)


# ---------------------------------------------------------------- util/arithmetic.rs
class Fraction:
    """arithmetic.rs:166-234"""

    def __init__(self, numer, denom):
        self.numer = numer
        self.denom = denom
        self.eval = None
        self.inv = False

    @classmethod
    def one_over(cls, denom):
        return cls(None, denom)

    def denom_ref(self):  # `denom()`
        return self.denom if not self.inv else None

    def denom_mut(self):
        if not self.inv:
            self.inv = True
            return Ref(self, "denom")
        return None

    def evaluate(self):
        assert self.inv
        if self.eval is None:
            if self.numer is not None:
                numer, self.numer = self.numer, None
                self.eval = numer * self.denom
            else:
                self.eval = self.denom

    def evaluated(self):
        assert self.eval is not None
        return self.eval


def root_of_unity(k):
    """arithmetic.rs:89-96"""
    assert k <= bn254.FR_S
    return pow(bn254.FR_ROOT_OF_UNITY, 1 << (bn254.FR_S - k), R)


class Domain:
    """arithmetic.rs:131-162"""

    def __init__(self, k, gen=None):
        self.k = k
        self.n = 1 << k
        self.gen = root_of_unity(k) if gen is None else gen
        self.n_inv = pow(self.n, R - 2, R)
        self.gen_inv = pow(self.gen, R - 2, R)

    def rotate_scalar(self, scalar, rotation):
        if rotation == 0:
            return scalar % R
        if rotation > 0:
            return scalar * pow(self.gen, rotation, R) % R
        return scalar * pow(self.gen_inv, -rotation, R) % R


def fe_to_fe(fq_value):
    """arithmetic.rs:256-258: Fq -> Fr by integer value mod r"""
    return fq_value % R


def fe_from_limbs(limbs, bits):
    """arithmetic.rs:262-274 (result must be a canonical element of the target field)."""
    return sum(l << (i * bits) for i, l in enumerate(limbs))


def fe_to_limbs(v, n_limbs, bits):
    """arithmetic.rs:278-290"""
    mask = (1 << bits) - 1
    return [(v >> (i * bits)) & mask for i in range(n_limbs)]


# ---------------------------------------------------------------- util/msm.rs:20-205
class Msm:
    def __init__(self, constant=None, scalars=None, bases=None):
        self.constant = constant
        self.scalars = scalars or []
        self.bases = bases or []

    @classmethod
    def constant_(cls, constant):
        return cls(constant=constant)

    @classmethod
    def base(cls, base):
        return cls(scalars=[base.loader.load_one()], bases=[base])

    def clone(self):
        return Msm(self.constant, list(self.scalars), list(self.bases))

    def size(self):
        return len(self.bases)

    def split(self):
        c, self.constant = self.constant, None
        return self, c

    def try_into_constant(self):
        return self.constant if not self.bases else None

    def evaluate(self, gen, loader=None):
        """msm.rs:70-77; `gen` = G1 affine tuple or None."""
        pairs = []
        if self.constant is not None:
            assert gen is not None
            ld = self.bases[0].loader if self.bases else loader
            pairs.append((self.constant, ld.ec_point_load_const(gen)))
        pairs += list(zip(self.scalars, self.bases))
        ld = pairs[0][1].loader
        return ld.multi_scalar_multiplication(pairs)

    def scale(self, factor):
        if self.constant is not None:
            self.constant = self.constant * factor
        self.scalars = [s * factor for s in self.scalars]

    def push(self, scalar, base):
        """msm.rs:88-95: dedup equal bases BY VALUE"""
        for i, b in enumerate(self.bases):
            if b == base:
                self.scalars[i] = self.scalars[i] + scalar
                return
        self.scalars.append(scalar)
        self.bases.append(base)

    def extend(self, other):
        if self.constant is not None and other.constant is not None:
            self.constant = self.constant + other.constant
        elif self.constant is None and other.constant is not None:
            self.constant = other.constant
        for s, b in zip(other.scalars, other.bases):
            self.push(s, b)

    def __add__(self, rhs):
        out = self.clone()
        out.extend(rhs)
        return out

    def __sub__(self, rhs):
        out = self.clone()
        out.extend(-rhs)
        return out

    def __mul__(self, scalar):
        out = self.clone()
        out.scale(scalar)
        return out

    def __neg__(self):
        return Msm(
            None if self.constant is None else -self.constant,
            [-s for s in self.scalars],
            list(self.bases),
        )

    @staticmethod
    def sum(msms):
        """`impl Sum for Msm` (msm.rs:196-205)"""
        acc = None
        for m in msms:
            acc = m if acc is None else acc + m
        return acc if acc is not None else Msm()
