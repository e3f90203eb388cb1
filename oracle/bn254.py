"""BN254 (halo2curves `bn256`) arithmetic on Python ints -- oracle ground truth.

Restates the behaviour of the un-vendored dependency `halo2curves 0.3.1`
(Cargo.lock:1803-1826) at the reference's call sites:
  * `Fr::invert`, `base * scalar`, `+`, `to_affine`   snark-verifier/src/loader/native.rs:29-33, 61-71
  * `G1Affine::from_bytes / to_bytes / coordinates`   snark-verifier/src/system/halo2/transcript/halo2.rs:216,252,296
  * `Fr::from_repr_vartime / to_repr`                 snark-verifier/src/system/halo2/transcript/halo2.rs:240,288
  * `multi_miller_loop / final_exponentiation`        snark-verifier/src/pcs/kzg/decider.rs:64-65
from the public BN254 definition.  TEST INFRASTRUCTURE ONLY.
"""

# ---------------------------------------------------------------- constants
X_BN = 4965661367192848881
P = 36 * X_BN**4 + 36 * X_BN**3 + 24 * X_BN**2 + 6 * X_BN + 1  # base field Fq
R = 36 * X_BN**4 + 36 * X_BN**3 + 18 * X_BN**2 + 6 * X_BN + 1  # scalar field Fr
assert P == 21888242871839275222246405745257275088696311157297823662689037894645226208583
assert R == 21888242871839275222246405745257275088548364400416034343698204186575808495617
ATE_LOOP = 6 * X_BN + 2
B_G1 = 3
FR_S = 28  # 2-adicity of r-1
FR_GENERATOR = 7
FR_ROOT_OF_UNITY = pow(FR_GENERATOR, (R - 1) >> FR_S, R)  # halo2curves Fr::ROOT_OF_UNITY
FR_DELTA = pow(FR_GENERATOR, 1 << FR_S, R)  # halo2curves Fr::DELTA


def fr_inv(a):
    """`Fr::invert().unwrap_or(self)` semantic of loader.rs:241-248 is applied by callers;
    here 0 has no inverse (returns None)."""
    a %= R
    return None if a == 0 else pow(a, R - 2, R)


def fq_inv(a):
    a %= P
    return None if a == 0 else pow(a, P - 2, P)


def fq_sqrt(a):
    """p = 3 mod 4 => candidate a^((p+1)/4); None if a is a non-residue."""
    a %= P
    y = pow(a, (P + 1) // 4, P)
    return y if y * y % P == a else None


def fe_to_bytes(a):
    """32-byte little-endian canonical repr (= halo2curves `to_repr`)."""
    return int(a).to_bytes(32, "little")


def fr_from_bytes(b):
    """`Fr::from_repr_vartime`: None if >= r."""
    v = int.from_bytes(b, "little")
    return v if v < R else None


def fq_from_bytes(b):
    v = int.from_bytes(b, "little")
    return v if v < P else None


# ---------------------------------------------------------------- G1 (affine tuples; None = identity)
G1_GEN = (1, 2)


def g1_is_on_curve(pt):
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - B_G1) % P == 0


def g1_neg(pt):
    if pt is None:
        return None
    return (pt[0], (-pt[1]) % P)


def g1_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, P - 2, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, P - 2, P) % P
    x3 = (lam * lam - x1 - x2) % P
    y3 = (lam * (x1 - x3) - y1) % P
    return (x3, y3)


# Jacobian for speed in the oracle's own big MSMs (value-identical after to_affine)
def _jac_dbl(p):
    X, Y, Z = p
    if Z == 0:
        return p
    A = X * X % P
    B = Y * Y % P
    C = B * B % P
    D = 2 * ((X + B) * (X + B) - A - C) % P
    E = 3 * A % P
    F = E * E % P
    X3 = (F - 2 * D) % P
    Y3 = (E * (D - X3) - 8 * C) % P
    Z3 = 2 * Y * Z % P
    return (X3, Y3, Z3)


def _jac_add_affine(p, q):
    if q is None:
        return p
    X1, Y1, Z1 = p
    x2, y2 = q
    if Z1 == 0:
        return (x2, y2, 1)
    Z1Z1 = Z1 * Z1 % P
    U2 = x2 * Z1Z1 % P
    S2 = y2 * Z1 * Z1Z1 % P
    if U2 == X1:
        if S2 == Y1:
            return _jac_dbl(p)
        return (1, 1, 0)
    H = (U2 - X1) % P
    HH = H * H % P
    I = 4 * HH % P
    J = H * I % P
    r = 2 * (S2 - Y1) % P
    V = X1 * I % P
    X3 = (r * r - J - 2 * V) % P
    Y3 = (r * (V - X3) - 2 * Y1 * J) % P
    Z3 = ((Z1 + H) * (Z1 + H) - Z1Z1 - HH) % P
    return (X3, Y3, Z3)


def _jac_to_affine(p):
    X, Y, Z = p
    if Z == 0:
        return None
    zi = pow(Z, P - 2, P)
    zi2 = zi * zi % P
    return (X * zi2 % P, Y * zi2 * zi % P)


def g1_mul(pt, k):
    """`base * scalar` of native.rs:67 (halo2curves: MSB-first double-and-add over the
    256 bits of `scalar.to_repr()`); value only -- the op count lives in the C oracle."""
    k %= R
    if pt is None or k == 0:
        return None
    acc = (1, 1, 0)
    for bit in bin(k)[2:]:
        acc = _jac_dbl(acc)
        if bit == "1":
            acc = _jac_add_affine(acc, pt)
    return _jac_to_affine(acc)


def g1_msm_naive(pairs):
    """`NativeLoader::multi_scalar_multiplication` (native.rs:61-71):
    sum of base*scalar, then to_affine.  `pairs` = [(scalar, point)]."""
    assert pairs, "pairs should not be empty"
    acc = None
    for s, b in pairs:
        acc = g1_add(acc, g1_mul(b, s))
    return acc


def g1_to_bytes(pt):
    """halo2curves 0.3.1 `G1Affine::to_bytes` (compressed, 32 B): x little-endian,
    bit 7 of byte 31 = lsb(y); identity = all zero.  (SURVEY App. C -- format restated
    from the published crate; isolated here.)"""
    if pt is None:
        return bytes(32)
    x, y = pt
    b = bytearray(x.to_bytes(32, "little"))
    b[31] |= (y & 1) << 7
    return bytes(b)


def g1_from_bytes(b):
    """halo2curves 0.3.1 `G1Affine::from_bytes`.  Returns (ok, point); ok False when the
    encoding is invalid (x >= p, or x^3+3 non-residue).  x == 0 with sign bit 0 decodes to
    the identity (which `common_ec_point` then rejects, transcript/halo2.rs:214-224)."""
    assert len(b) == 32
    t = bytearray(b)
    ysign = t[31] >> 7
    t[31] &= 0x7F
    x = int.from_bytes(t, "little")
    if x >= P:
        return False, None
    if x == 0 and ysign == 0:
        return True, None
    y = fq_sqrt(x * x * x + B_G1)
    if y is None:
        return False, None
    if (y & 1) != ysign:
        y = (-y) % P
    return True, (x, y)


# ---------------------------------------------------------------- Fq2 = Fq[u]/(u^2+1)
def f2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def f2_neg(a):
    return ((-a[0]) % P, (-a[1]) % P)


def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2_sqr(a):
    return f2_mul(a, a)


def f2_muls(a, s):
    return (a[0] * s % P, a[1] * s % P)


def f2_conj(a):
    return (a[0], (-a[1]) % P)


def f2_inv(a):
    n = pow((a[0] * a[0] + a[1] * a[1]) % P, P - 2, P)
    return (a[0] * n % P, (-a[1]) * n % P)


def f2_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_sqr(a)
        e >>= 1
    return r


F2_ZERO = (0, 0)
F2_ONE = (1, 0)
XI = (9, 1)  # non-residue for Fq6: v^3 = xi


def f2_mul_xi(a):
    return f2_mul(a, XI)


# ---------------------------------------------------------------- Fq6 = Fq2[v]/(v^3 - xi)
def f6_add(a, b):
    return tuple(f2_add(x, y) for x, y in zip(a, b))


def f6_sub(a, b):
    return tuple(f2_sub(x, y) for x, y in zip(a, b))


def f6_neg(a):
    return tuple(f2_neg(x) for x in a)


def f6_mul(a, b):
    a0, a1, a2 = a
    b0, b1, b2 = b
    c0 = f2_add(f2_mul(a0, b0), f2_mul_xi(f2_add(f2_mul(a1, b2), f2_mul(a2, b1))))
    c1 = f2_add(f2_add(f2_mul(a0, b1), f2_mul(a1, b0)), f2_mul_xi(f2_mul(a2, b2)))
    c2 = f2_add(f2_add(f2_mul(a0, b2), f2_mul(a1, b1)), f2_mul(a2, b0))
    return (c0, c1, c2)


def f6_mul_v(a):
    """multiply by v: (a0,a1,a2) -> (xi*a2, a0, a1)"""
    return (f2_mul_xi(a[2]), a[0], a[1])


def f6_inv(a):
    a0, a1, a2 = a
    t0 = f2_sub(f2_sqr(a0), f2_mul_xi(f2_mul(a1, a2)))
    t1 = f2_sub(f2_mul_xi(f2_sqr(a2)), f2_mul(a0, a1))
    t2 = f2_sub(f2_sqr(a1), f2_mul(a0, a2))
    d = f2_add(f2_mul(a0, t0), f2_mul_xi(f2_add(f2_mul(a2, t1), f2_mul(a1, t2))))
    di = f2_inv(d)
    return (f2_mul(t0, di), f2_mul(t1, di), f2_mul(t2, di))


F6_ZERO = (F2_ZERO, F2_ZERO, F2_ZERO)
F6_ONE = (F2_ONE, F2_ZERO, F2_ZERO)


# ---------------------------------------------------------------- Fq12 = Fq6[w]/(w^2 - v)
def f12_mul(a, b):
    a0, a1 = a
    b0, b1 = b
    t0 = f6_mul(a0, b0)
    t1 = f6_mul(a1, b1)
    c0 = f6_add(t0, f6_mul_v(t1))
    c1 = f6_sub(f6_sub(f6_mul(f6_add(a0, a1), f6_add(b0, b1)), t0), t1)
    return (c0, c1)


def f12_sqr(a):
    return f12_mul(a, a)


def f12_inv(a):
    a0, a1 = a
    d = f6_sub(f6_mul(a0, a0), f6_mul_v(f6_mul(a1, a1)))
    di = f6_inv(d)
    return (f6_mul(a0, di), f6_neg(f6_mul(a1, di)))


def f12_conj(a):
    return (a[0], f6_neg(a[1]))


def f12_pow(a, e):
    r = F12_ONE
    while e:
        if e & 1:
            r = f12_mul(r, a)
        a = f12_sqr(a)
        e >>= 1
    return r


F12_ONE = (F6_ONE, F6_ZERO)


# ---------------------------------------------------------------- G2 on the twist y^2 = x^3 + 3/xi (affine, None = identity)
B_G2 = f2_mul((3, 0), f2_inv(XI))
G2_GEN = (
    (
        10857046999023057135944570762232829481370756359578518086990519993285655852781,
        11559732032986387107991004021392285783925812861821192530917403151452391805634,
    ),
    (
        8495653923123431417604973247489272438418190587263600148770280649306958101930,
        4082367875863433681332203403145435568316851327593401208105741076214120093531,
    ),
)


def g2_is_on_curve(pt):
    if pt is None:
        return True
    x, y = pt
    return f2_sub(f2_sqr(y), f2_add(f2_mul(f2_sqr(x), x), B_G2)) == F2_ZERO


def g2_neg(pt):
    if pt is None:
        return None
    return (pt[0], f2_neg(pt[1]))


def g2_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if f2_add(y1, y2) == F2_ZERO:
            return None
        lam = f2_mul(f2_muls(f2_sqr(x1), 3), f2_inv(f2_muls(y1, 2)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_sqr(lam), x1), x2)
    y3 = f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1)
    return (x3, y3)


def g2_mul(pt, k):
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = g2_add(acc, acc)
        if bit == "1":
            acc = g2_add(acc, pt)
    return acc


# Frobenius constants gamma_{1,i} = xi^{i (p-1)/6}
_G12 = f2_pow(XI, (P - 1) // 3)
_G13 = f2_pow(XI, (P - 1) // 2)


def g2_frobenius(pt):
    """pi_p on the twist: (x,y) -> (conj(x)*xi^((p-1)/3), conj(y)*xi^((p-1)/2))"""
    x, y = pt
    return (f2_mul(f2_conj(x), _G12), f2_mul(f2_conj(y), _G13))


# ---------------------------------------------------------------- optimal ate pairing
def _line(T, Q2, Pt):
    """Line through T and Q2 (twist affine, T == Q2 => tangent) evaluated at Pt in E(Fq),
    as an Fq12 element via the untwist (x',y') -> (x' w^2, y' w^3):
        l = yP - lam*xP*w + (lam*xT - yT)*w^3 ,  w^3 = v*w
    Returns (l, T+Q2).  Vertical lines are dropped (killed by the final exponentiation)."""
    xP, yP = Pt
    x1, y1 = T
    x2, y2 = Q2
    if x1 == x2:
        if f2_add(y1, y2) == F2_ZERO:
            return F12_ONE, None
        lam = f2_mul(f2_muls(f2_sqr(x1), 3), f2_inv(f2_muls(y1, 2)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_sqr(lam), x1), x2)
    y3 = f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1)
    c_1 = (yP % P, 0)
    c_w = f2_neg(f2_muls(lam, xP))
    c_w3 = f2_sub(f2_mul(lam, x1), y1)
    l = ((c_1, F2_ZERO, F2_ZERO), (c_w, c_w3, F2_ZERO))
    return l, (x3, y3)


def miller_loop(Pt, Q):
    """Optimal-ate Miller loop f_{6x+2,Q}(P) * l_{T,pi(Q)}(P) * l_{T,-pi^2(Q)}(P)."""
    if Pt is None or Q is None:
        return F12_ONE
    f = F12_ONE
    T = Q
    for bit in bin(ATE_LOOP)[3:]:
        l, T = _line(T, T, Pt)
        f = f12_mul(f12_sqr(f), l)
        if bit == "1":
            l, T = _line(T, Q, Pt)
            f = f12_mul(f, l)
    Q1 = g2_frobenius(Q)
    Q2 = g2_neg(g2_frobenius(Q1))
    l, T = _line(T, Q1, Pt)
    f = f12_mul(f, l)
    l, T = _line(T, Q2, Pt)
    f = f12_mul(f, l)
    return f


FINAL_EXP = (P**12 - 1) // R


def final_exponentiation(f):
    return f12_pow(f, FINAL_EXP)


def multi_miller_loop(terms):
    """`M::multi_miller_loop(&[(G1Affine, G2Prepared)])` (decider.rs:64-65): product of Miller loops."""
    f = F12_ONE
    for Pt, Q in terms:
        f = f12_mul(f, miller_loop(Pt, Q))
    return f


def pairing(Pt, Q):
    return final_exponentiation(miller_loop(Pt, Q))


def pairing_check(terms):
    """True iff prod e(P_i, Q_i) == 1 in Gt."""
    return final_exponentiation(multi_miller_loop(terms)) == F12_ONE
