"""KZG PCS + accumulation scheme (`KzgAs`, `Bdfg21`, `Gwc19`, `KzgAccumulator`, `LimbsEncoding`,
`KzgDecidingKey`) -- oracle restatement.  TEST INFRASTRUCTURE ONLY.

Follows
  snark-verifier/src/pcs/kzg.rs:21-37
  snark-verifier/src/pcs/kzg/accumulator.rs:6-26, 34, 36-79
  snark-verifier/src/pcs/kzg/accumulation.rs:17-63, 97-137, 139-196
  snark-verifier/src/pcs/kzg/decider.rs:6-36, 38-82
  snark-verifier/src/pcs/kzg/multiopen/bdfg21.rs:25-80, 84-115, 117-167, 169-219, 222-260, 263-367
  snark-verifier/src/pcs/kzg/multiopen/gwc19.rs:21-81, 85-109, 111-158
"""
from dataclasses import dataclass
from typing import Any

from . import bn254
from .bn254 import R
from .loader import EcPoint, Fraction, Msm, NativeLoader, fe_from_limbs
from .transcript import VerifyError


@dataclass
class KzgSuccinctVerifyingKey:
    g: Any  # G1 affine tuple


@dataclass
class KzgDecidingKey:
    svk: KzgSuccinctVerifyingKey
    g2: Any
    s_g2: Any

    @classmethod
    def new(cls, g1, g2, s_g2):
        return cls(KzgSuccinctVerifyingKey(g1), g2, s_g2)


@dataclass
class KzgAccumulator:
    lhs: EcPoint
    rhs: EcPoint


class LimbsEncoding:
    """accumulator.rs:36-79 (native `from_repr`)"""

    def __init__(self, limbs=3, bits=88):
        self.limbs, self.bits = limbs, bits

    def from_repr(self, limbs, loader):
        from .transcript import ReferencePanic

        assert len(limbs) == 4 * self.limbs
        coords = []
        for i in range(4):
            v = fe_from_limbs([l.v for l in limbs[i * self.limbs : (i + 1) * self.limbs]], self.bits)
            if v >= bn254.P:  # `assert!(bytes.len() <= 32)` / `from_repr().unwrap()` of fe_from_big (arithmetic.rs:237-243)
                raise ReferencePanic("fe_from_big")
            coords.append(v)
        pts = []
        for x, y in ((coords[0], coords[1]), (coords[2], coords[3])):
            if (x, y) == (0, 0):  # halo2curves `from_xy`: `is_on_curve() | is_identity()` accepts (0, 0) as the identity
                pts.append(None)
            elif bn254.g1_is_on_curve((x, y)):
                pts.append((x, y))
            else:  # `C::from_xy(..).unwrap()` panics when off-curve (accumulator.rs:72-73)
                raise ReferencePanic("from_xy().unwrap()")
        return KzgAccumulator(loader.ec_point_load_const(pts[0]), loader.ec_point_load_const(pts[1]))


# ================================================================ SHPLONK (bdfg21.rs)
class Bdfg21Proof:
    def __init__(self, mu, gamma, w, z_prime, w_prime):
        self.mu, self.gamma, self.w, self.z_prime, self.w_prime = mu, gamma, w, z_prime, w_prime

    @classmethod
    def read(cls, transcript):
        """bdfg21.rs:101-114"""
        mu = transcript.squeeze_challenge()
        gamma = transcript.squeeze_challenge()
        w = transcript.read_ec_point()
        z_prime = transcript.squeeze_challenge()
        w_prime = transcript.read_ec_point()
        return cls(mu, gamma, w, z_prime, w_prime)


class _BdfgQuerySet:
    def __init__(self, shifts, polys, evals):
        self.shifts, self.polys, self.evals = shifts, polys, evals

    def msm(self, coeff, commitments, powers_of_mu):
        """bdfg21.rs:229-259"""
        terms = []
        for poly, evals, power_of_mu in zip(self.polys, self.evals, powers_of_mu):
            loader = power_of_mu.loader
            if coeff.commitment_coeff is not None:
                commitment = commitments[poly] * coeff.commitment_coeff.evaluated()
            else:
                commitment = commitments[poly].clone()
            r_eval = loader.sum_products(
                [(c.evaluated(), e) for c, e in zip(coeff.eval_coeffs, evals)]
            ) * coeff.r_eval_coeff.evaluated()
            terms.append((commitment - Msm.constant_(r_eval)) * power_of_mu)
        return Msm.sum(terms)


def bdfg21_query_sets(queries):
    """bdfg21.rs:117-167"""
    poly_shifts = []  # (poly, shifts, evals)
    for q in queries:
        pos = next((i for i, ps in enumerate(poly_shifts) if ps[0] == q.poly), None)
        if pos is not None:
            _, shifts, evals = poly_shifts[pos]
            if q.shift not in shifts:
                shifts.append(q.shift)
                evals.append(q.eval)
        else:
            poly_shifts.append((q.poly, [q.shift], [q.eval]))
    sets = []
    for poly, shifts, evals in poly_shifts:
        pos = next((i for i, s in enumerate(sets) if set(s.shifts) == set(shifts)), None)
        if pos is not None:
            s = sets[pos]
            if poly not in s.polys:
                s.polys.append(poly)
                s.evals.append([evals[shifts.index(lhs)] for lhs in s.shifts])
        else:
            sets.append(_BdfgQuerySet(shifts, [poly], [evals]))
    return sets


class _QuerySetCoeff:
    def __init__(self, shifts, powers_of_z, z_prime, z_prime_minus_z_shift_i, z_s_1):
        """bdfg21.rs:276-329"""
        loader = z_prime.loader
        nep = []
        for j, sj in enumerate(shifts):
            acc = None
            for i, si in enumerate(shifts):
                if i != j:
                    d = (sj - si) % R
                    acc = d if acc is None else acc * d % R
            nep.append(1 if acc is None else acc)
        z = powers_of_z[1]
        z_pow_k_minus_one = powers_of_z[len(shifts) - 1]
        self.eval_coeffs = [
            Fraction.one_over(
                loader.sum_products_with_coeff(
                    [(n, z_pow_k_minus_one, z_prime), ((-(n * s)) % R, z_pow_k_minus_one, z)]
                )
            )
            for s, n in zip(shifts, nep)
        ]
        self.z_s = loader.product([z_prime_minus_z_shift_i[s] for s in shifts])
        self.commitment_coeff = None if z_s_1 is None else Fraction(z_s_1, self.z_s)
        self.r_eval_coeff = None

    def denoms(self):
        """bdfg21.rs:331-362"""
        if self.eval_coeffs[0].denom_ref() is not None:
            fr = list(self.eval_coeffs) + ([self.commitment_coeff] if self.commitment_coeff is not None else [])
            return [r for r in (f.denom_mut() for f in fr) if r is not None]
        if self.r_eval_coeff is None:
            loader = self.z_s.loader
            for f in list(self.eval_coeffs) + ([self.commitment_coeff] if self.commitment_coeff is not None else []):
                f.evaluate()
            bw_sum = loader.sum([f.evaluated() for f in self.eval_coeffs])
            if self.commitment_coeff is not None:
                self.r_eval_coeff = Fraction(self.commitment_coeff.evaluated(), bw_sum)
            else:
                self.r_eval_coeff = Fraction.one_over(bw_sum)
            return [self.r_eval_coeff.denom_mut()]
        raise AssertionError("unreachable")

    def evaluate(self):
        self.r_eval_coeff.evaluate()


def bdfg21_query_set_coeffs(sets, z, z_prime):
    """bdfg21.rs:169-219"""
    loader = z.loader
    superset = sorted({s for st in sets for s in st.shifts})
    size = max([len(st.shifts) for st in sets] + [2])
    powers_of_z = z.powers(size)
    zpm = {shift: z_prime - z * loader.load_const(shift) for shift in superset}
    z_s_1 = None
    coeffs = []
    for st in sets:
        c = _QuerySetCoeff(st.shifts, powers_of_z, z_prime, zpm, z_s_1)
        if z_s_1 is None:
            z_s_1 = c.z_s
        coeffs.append(c)
    NativeLoader.batch_invert([r for c in coeffs for r in c.denoms()])
    NativeLoader.batch_invert([r for c in coeffs for r in c.denoms()])
    for c in coeffs:
        c.evaluate()
    return coeffs


# ================================================================ GWC (gwc19.rs)
class Gwc19Proof:
    def __init__(self, v, ws, u):
        self.v, self.ws, self.u = v, ws, u

    @classmethod
    def read(cls, queries, transcript):
        """gwc19.rs:100-108"""
        v = transcript.squeeze_challenge()
        ws = transcript.read_n_ec_points(len(gwc19_query_sets(queries)))
        u = transcript.squeeze_challenge()
        return cls(v, ws, u)


class _GwcQuerySet:
    def __init__(self, shift, polys, evals):
        self.shift, self.polys, self.evals = shift, polys, evals

    def msm(self, commitments, powers_of_v):
        """gwc19.rs:122-137"""
        terms = []
        for (poly, ev), pv in zip(zip(self.polys, self.evals), powers_of_v):
            terms.append((commitments[poly].clone() - Msm.constant_(ev)) * pv)
        return Msm.sum(terms)


def gwc19_query_sets(queries):
    """gwc19.rs:140-158"""
    sets = []
    for q in queries:
        pos = next((i for i, s in enumerate(sets) if s.shift == q.shift), None)
        if pos is not None:
            sets[pos].polys.append(q.poly)
            sets[pos].evals.append(q.eval)
        else:
            sets.append(_GwcQuerySet(q.shift, [q.poly], [q.eval]))
    return sets


# ================================================================ KzgAs<Bn256, MOS>
class _KzgAsBase:
    """accumulation.rs + decider.rs (shared by both multi-open schemes)"""

    # ---- AccumulationScheme (accumulation.rs:17-63, 97-137); vk_zk = KzgAsVerifyingKey.zk()
    @staticmethod
    def as_read_proof(vk_zk, instances, transcript):
        assert instances, "assert!(!instances.is_empty())"
        for acc in instances:
            transcript.common_ec_point(acc.lhs)
            transcript.common_ec_point(acc.rhs)
        blind = None
        if vk_zk:
            blind = (transcript.read_ec_point(), transcript.read_ec_point())
        r = transcript.squeeze_challenge()
        return (blind, r)

    @staticmethod
    def as_verify(vk_zk, instances, proof):
        blind, r = proof
        lhs = [a.lhs for a in instances] + ([blind[0]] if blind else [])
        rhs = [a.rhs for a in instances] + ([blind[1]] if blind else [])
        powers_of_r = r.powers(len(lhs))
        out = []
        for bases in (lhs, rhs):
            out.append(Msm.sum([Msm.base(b) * p for b, p in zip(bases, powers_of_r)]).evaluate(None))
        return KzgAccumulator(out[0], out[1])

    @classmethod
    def create_proof(cls, instances, transcript):
        """`AccumulationSchemeProver::create_proof` with `KzgAsProvingKey::default()` (zk = false):
        writes nothing to the transcript (accumulation.rs:139-196; sdk aggregation.rs:235-245)."""
        proof = cls.as_read_proof(False, instances, transcript)
        return cls.as_verify(False, instances, proof)

    # ---- AccumulationDecider (decider.rs:60-81)
    @staticmethod
    def decide(dk, acc):
        terms = [(acc.lhs.pt, dk.g2), (acc.rhs.pt, bn254.g2_neg(dk.s_g2))]
        if not bn254.pairing_check(terms):
            raise VerifyError("AssertionFailure", "e(lhs, g2)·e(rhs, -s_g2) == O")

    @classmethod
    def decide_all(cls, dk, accumulators):
        assert accumulators, "assert!(!accumulators.is_empty())"
        for a in accumulators:
            cls.decide(dk, a)


class KzgAsBdfg21(_KzgAsBase):
    """`KzgAs<Bn256, Bdfg21>` (the SDK's `SHPLONK`, snark-verifier-sdk/src/lib.rs:38-42)"""

    NAME = "bdfg21"

    @staticmethod
    def read_proof(svk, queries, transcript):
        return Bdfg21Proof.read(transcript)

    @staticmethod
    def verify(svk, commitments, z, queries, proof):
        """bdfg21.rs:47-79"""
        sets = bdfg21_query_sets(queries)
        coeffs = bdfg21_query_set_coeffs(sets, z, proof.z_prime)
        powers_of_mu = proof.mu.powers(max(len(s.polys) for s in sets))
        msms = [s.msm(c, commitments, powers_of_mu) for s, c in zip(sets, coeffs)]
        f = Msm.sum([m * g for m, g in zip(msms, proof.gamma.powers(len(sets)))]) - Msm.base(proof.w) * coeffs[0].z_s
        rhs = Msm.base(proof.w_prime)
        lhs = f + rhs * proof.z_prime
        return KzgAccumulator(lhs.evaluate(svk.g), rhs.evaluate(svk.g))


class KzgAsGwc19(_KzgAsBase):
    """`KzgAs<Bn256, Gwc19>` (the SDK's `GWC`)"""

    NAME = "gwc19"

    @staticmethod
    def read_proof(svk, queries, transcript):
        return Gwc19Proof.read(queries, transcript)

    @staticmethod
    def verify(svk, commitments, z, queries, proof):
        """gwc19.rs:43-80"""
        sets = gwc19_query_sets(queries)
        powers_of_u = proof.u.powers(len(sets))
        powers_of_v = proof.v.powers(max(len(s.polys) for s in sets))
        f = Msm.sum([s.msm(commitments, powers_of_v) * pu for s, pu in zip(sets, powers_of_u)])
        z_omegas = [z.loader.load_const(s.shift) * z for s in sets]
        rhs = [Msm.base(w) * pu for w, pu in zip(proof.ws, powers_of_u)]
        lhs = f + Msm.sum([uw * zo for uw, zo in zip(rhs, z_omegas)])
        return KzgAccumulator(lhs.evaluate(svk.g), Msm.sum(rhs).evaluate(svk.g))
