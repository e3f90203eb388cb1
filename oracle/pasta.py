"""TEST INFRASTRUCTURE ONLY (imported by tests/, smoke(), bench cpu_baseline -- never by the product path).

Pasta curves (Pallas / Vesta) in exact Python integers, for the IPA decider.  The reference takes them from
`halo2_curves::pasta` (re-export of the `pasta_curves` crate; not vendored) -- the reference test is over
`pasta::pallas` (snark-verifier/src/pcs/ipa.rs:407-446).  Restated from the public definition: y^2 = x^3 + 5 over
  Pallas base  p = 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001  (= Vesta scalar field)
  Vesta  base  q = 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001  (= Pallas scalar field)
generator (-1, 2) on both.  PARITY UNPINNED by the reference (it holds no Pasta vector); self-checked by exact algebra
(group order annihilates the generator on both curves, tests/test_oracle_ipa.py)."""

PALLAS_P = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001
VESTA_P = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001
B = 5


class Curve:
    """Short Weierstrass curve y^2 = x^3 + b over F_p with prime order n; points are (x, y) tuples, identity is None."""

    def __init__(self, name, curve_id, p, n, b, gen):
        self.name, self.id, self.p, self.n, self.b, self.gen = name, curve_id, p, n, b, gen

    def is_on_curve(self, pt):
        if pt is None:
            return True
        x, y = pt
        return 0 <= x < self.p and 0 <= y < self.p and (y * y - x * x * x - self.b) % self.p == 0

    def neg(self, pt):
        return None if pt is None else (pt[0], (-pt[1]) % self.p)

    def add(self, a, b):
        p = self.p
        if a is None:
            return b
        if b is None:
            return a
        x1, y1 = a
        x2, y2 = b
        if x1 == x2:
            if (y1 + y2) % p == 0:
                return None
            lam = 3 * x1 * x1 * pow(2 * y1, -1, p) % p
        else:
            lam = (y2 - y1) * pow(x2 - x1, -1, p) % p
        x3 = (lam * lam - x1 - x2) % p
        return (x3, (lam * (x1 - x3) - y1) % p)

    # Jacobian internals so that a scalar multiplication costs one inversion
    def _dbl(self, P):
        p = self.p
        X, Y, Z = P
        if Z == 0:
            return P
        A, Bq = X * X % p, Y * Y % p
        C = Bq * Bq % p
        D = 2 * ((X + Bq) ** 2 - A - C) % p
        E = 3 * A % p
        X3 = (E * E - 2 * D) % p
        return (X3, (E * (D - X3) - 8 * C) % p, 2 * Y * Z % p)

    def _add_aff(self, P, q):
        p = self.p
        X, Y, Z = P
        if q is None:
            return P
        if Z == 0:
            return (q[0], q[1], 1)
        ZZ = Z * Z % p
        U2, S2 = q[0] * ZZ % p, q[1] * Z * ZZ % p
        if U2 == X:
            return self._dbl(P) if S2 == Y else (1, 1, 0)
        H, r = (U2 - X) % p, (S2 - Y) % p
        HH = H * H % p
        HHH = H * HH % p
        V = X * HH % p
        X3 = (r * r - HHH - 2 * V) % p
        return (X3, (r * (V - X3) - Y * HHH) % p, Z * H % p)

    def _to_affine(self, P):
        X, Y, Z = P
        if Z == 0:
            return None
        zi = pow(Z, -1, self.p)
        return (X * zi * zi % self.p, Y * zi * zi * zi % self.p)

    def mul(self, pt, k):
        k %= self.n
        acc = (1, 1, 0)
        for bit in bin(k)[2:] if k else "":
            acc = self._dbl(acc)
            if bit == "1":
                acc = self._add_aff(acc, pt)
        return self._to_affine(acc)

    def msm_naive(self, scalars, points):
        """`Σ base * scalar` then to_affine, as NativeLoader would (loader/native.rs:61-71)."""
        acc = None
        for s, b in zip(scalars, points):
            acc = self.add(acc, self.mul(b, s))
        return acc

    def msm_pippenger(self, scalars, points):
        """util/msm.rs:238-317 restated (serial path): window ceil(ln n)+2, 2^w - 1 buckets, running sums, Horner."""
        import math

        n = len(scalars)
        if n == 0:
            return None
        w = 3 if n < 4 else (4 if n < 32 else math.ceil(math.log(n)) + 2)  # msm.rs:253-259
        nbits = 256
        num_windows = (nbits + w - 1) // w
        acc = (1, 1, 0)
        for win in range(num_windows - 1, -1, -1):
            for _ in range(w):
                acc = self._dbl(acc)
            buckets = [(1, 1, 0)] * ((1 << w) - 1)
            for s, b in zip(scalars, points):
                d = (s >> (win * w)) & ((1 << w) - 1)
                if d:
                    buckets[d - 1] = self._add_aff(buckets[d - 1], b)
            run = None
            tot = None
            for bk in reversed(buckets):
                run = self.add(run, self._to_affine(bk))
                tot = self.add(tot, run)
            acc = self._add_aff(acc, tot)
        return self._to_affine(acc)


PALLAS = Curve("pallas", 1, PALLAS_P, VESTA_P, B, (PALLAS_P - 1, 2))
VESTA = Curve("vesta", 2, VESTA_P, PALLAS_P, B, (VESTA_P - 1, 2))


def bn254_g1():
    from . import bn254

    return Curve("bn254_g1", 0, bn254.P, bn254.R, 3, (1, 2))
