#!/usr/bin/env python
"""Derives the GLV constants of BN254 G1 used by csrc/glv.cuh and checks them with the oracle's exact arithmetic.
  phi(x, y) = (beta x, y) is an endomorphism of y^2 = x^3 + 3 with phi(P) = lambda P on the r-torsion;
  a short basis (a1, b1), (a2, b2) of the lattice { (a, b) : a + b lambda = 0 mod r } splits k = k1 + k2 lambda with
  |k1|, |k2| < 2^128.  Any integers c1, c2 give a CORRECT split (the basis vectors are in the lattice); rounding only
  decides how short k1, k2 are.  Run:  python tools/glv_constants.py"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bn254  # noqa: E402

R, P = bn254.R, bn254.P


def cube_roots_of_unity(m):
    # m = 1 mod 3: g^((m-1)/3) for a non-cube g
    for g in range(2, 50):
        w = pow(g, (m - 1) // 3, m)
        if w != 1:
            return w, w * w % m
    raise AssertionError


def short_basis(lam):
    # extended Euclid on (r, lambda): remainders r_i = s_i r + t_i lambda
    seq = [(R, 0), (lam, 1)]
    while seq[-1][0] != 0:
        (r0, t0), (r1, t1) = seq[-2], seq[-1]
        q = r0 // r1
        seq.append((r0 - q * r1, t0 - q * t1))
    sq = 1 << (R.bit_length() // 2 + 1)
    i = next(k for k, (rk, _) in enumerate(seq) if rk * rk < R)
    (rl, tl), (rl1, tl1), (rl2, tl2) = seq[i - 1], seq[i], seq[i + 1]
    v1 = (rl1, -tl1)
    cands = [(rl, -tl), (rl2, -tl2)]
    v2 = min(cands, key=lambda v: v[0] * v[0] + v[1] * v[1])
    return v1, v2


def main():
    lam_a, lam_b = cube_roots_of_unity(R)
    beta_a, beta_b = cube_roots_of_unity(P)
    G = bn254.G1_GEN
    pairs = [(l, b) for l in (lam_a, lam_b) for b in (beta_a, beta_b) if bn254.g1_mul(G, l) == (b * G[0] % P, G[1])]
    lam, beta = min(pairs)  # two valid pairs (lambda, beta) and (lambda^2, beta^2): take the smaller lambda
    (a1, b1), (a2, b2) = short_basis(lam)
    for a, b in ((a1, b1), (a2, b2)):
        assert (a + b * lam) % R == 0
    det = a1 * b2 - a2 * b1
    assert abs(det) == R
    # k = k1 + k2 lambda with (k1, k2) = (k, 0) - c1 (a1, b1) - c2 (a2, b2), c1 = round(b2 k / det), c2 = round(-b1 k / det)
    # device: c_i = (k * g_i) >> 256 with g_i = round(2^256 |b_j| / r), signs applied afterwards
    sgn = 1 if det > 0 else -1
    g1 = ((abs(b2) << 256) + R // 2) // R
    g2 = ((abs(b1) << 256) + R // 2) // R
    s1 = sgn * (1 if b2 > 0 else -1)        # sign of c1
    s2 = sgn * (1 if -b1 > 0 else -1)       # sign of c2
    rng = random.Random(1)
    worst = 0
    for t in range(20000):
        k = rng.randrange(R) if t > 10 else [0, 1, R - 1, R // 2, lam, R - lam, 2**253, 2**128, 2**127, 2**254 % R, 3][t]
        c1 = s1 * ((k * g1) >> 256)
        c2 = s2 * ((k * g2) >> 256)
        k1 = k - c1 * a1 - c2 * a2
        k2 = -c1 * b1 - c2 * b2
        assert (k1 + k2 * lam - k) % R == 0
        worst = max(worst, abs(k1), abs(k2))
    assert worst < 1 << 128, worst.bit_length()

    def limbs(v, n=8):
        return ", ".join(f"0x{(v >> (32 * i)) & 0xffffffff:08x}u" for i in range(n))

    print(f"lambda = {lam}\nbeta   = {beta}")
    print(f"a1 = {a1}\nb1 = {b1}\na2 = {a2}\nb2 = {b2}\ndet sign = {sgn}, s1 = {s1}, s2 = {s2}")
    print(f"max |k_i| over the sample: {worst.bit_length()} bits")
    beta_mont = beta * (1 << 256) % P
    print("BETA_MONT  {", limbs(beta_mont), "}")
    print("G1         {", limbs(g1, 5), "}   //", g1.bit_length(), "bits")
    print("G2         {", limbs(g2, 5), "}   //", g2.bit_length(), "bits")
    for name, v in (("A1", a1), ("B1", b1), ("A2", a2), ("B2", b2)):
        print(f"{name:3s} sign {'-' if v < 0 else '+'} {{ {limbs(abs(v), 5)} }}   // {abs(v).bit_length()} bits")


if __name__ == "__main__":
    main()
