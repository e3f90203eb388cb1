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
