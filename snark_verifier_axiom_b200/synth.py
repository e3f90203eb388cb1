"""Synthetic workload generator: DISTINCT, VALID StandardPlonk SHPLONK proofs at bench scale, made with the
library's own GPU kernels and the trapdoor of the test SRS.

The reference produces proofs with the halo2 prover (snark-verifier-sdk/src/halo2.rs:77-146), which cannot run
here.  With the SRS secret `s` and the discrete logs of every G1 point known, the last opening point is the unique
solution of `lhs = s * rhs` (decider.rs:64-67):  W' = phi / (s - z') * G,  phi = dlog of `f` (bdfg21.rs:55-70),
which is  sum_t scalar_t * dlog(base_t)  over the terms of the final `Msm` except W' itself.  The per-proof scalars
come from `svk_plonk_msm_scalars_batch` (the compiled tape run on the GPU), the point multiplications from
`svk_g1_mul_batch`; only ~25 modular multiplications per proof run in Python.

This is data generation for benchmarks/tests; the oracle validates its output (tests/test_gpu_synth.py).
"""
import ctypes

import numpy as np

from .protocol import FR_MODULUS as R
from .verifier import SHPLONK, PlonkVerifier, Snark, _ptr

G1_GEN_BYTES = (1).to_bytes(32, "little") + (2).to_bytes(32, "little")


def _compress(points: np.ndarray) -> np.ndarray:
    """[k, 64] affine canonical (x LE, y LE) -> [k, 32] halo2curves compressed (sign of y in bit 7 of byte 31)"""
    out = points[:, :32].copy()
    out[:, 31] |= (points[:, 32] & 1) << 7
    return out


def _g_multiples(ctx, scalars: np.ndarray) -> np.ndarray:
    """[k, 32] canonical scalars -> [k, 64] affine points scalars[i] * G"""
    k = scalars.shape[0]
    out = np.zeros((k, 64), np.uint8)
    gen = np.frombuffer(G1_GEN_BYTES, np.uint8).copy()
    sc = np.ascontiguousarray(scalars)
    ctx._check(ctx._L.svk_g1_mul_batch(ctx._c, k, _ptr(sc), _ptr(gen), 1, _ptr(out)))
    return out


def _rand_scalars(rng, shape) -> np.ndarray:
    b = rng.integers(0, 256, size=tuple(shape) + (32,), dtype=np.uint8)
    b[..., 31] &= 0x1F  # < 2^253 < r: canonical
    return b


def _ints(a: np.ndarray):
    """[k, 32] LE bytes -> list of Python ints"""
    return [int.from_bytes(row.tobytes(), "little") for row in a]


def forge_shplonk_batch(pv: PlonkVerifier, trapdoor_s: int, vk_dlogs, n: int, seed: int = 1):
    """-> (instances uint8[n, 32], proofs uint8[n, 896]) of n distinct valid proofs for `pv` (a SHPLONK
    PlonkVerifier of the StandardPlonk protocol whose preprocessed commitments have the discrete logs `vk_dlogs`)."""
    assert pv.mos == SHPLONK
    ctx = pv.ctx
    info = pv.info
    n_pts = info["n_points"]  # 6 witness + 3 quotient + W + W'
    n_ev = (info["proof_len"] - 32 * n_pts) // 32
    rng = np.random.default_rng(seed)
    dl = _rand_scalars(rng, (n, n_pts - 1))  # dlogs of every proof point except W'
    pts = _compress(_g_multiples(ctx, dl.reshape(-1, 32))).reshape(n, n_pts - 1, 32)
    evals = _rand_scalars(rng, (n, n_ev))
    inst = _rand_scalars(rng, (n, 1)).reshape(n, 32)
    n_front = n_pts - 2  # witness + quotient points precede the evaluations; W and W' follow
    proofs = np.zeros((n, info["proof_len"]), np.uint8)
    proofs[:, : 32 * n_front] = pts[:, :n_front].reshape(n, -1)
    proofs[:, 32 * n_front : 32 * (n_front + n_ev)] = evals.reshape(n, -1)
    proofs[:, 32 * (n_front + n_ev) : 32 * (n_front + n_ev + 1)] = pts[:, n_front]
    placeholder = _compress(np.frombuffer(G1_GEN_BYTES, np.uint8).reshape(1, 64))[0]
    proofs[:, -32:] = placeholder
    # the final Msm with the placeholder W': terms + per-proof scalars + challenges
    terms = (ctypes.c_int32 * (3 * 64))()
    nt = ctx._check(ctx._L.svk_protocol_msm_terms(ctx._c, pv.pid, 0, terms, 64))
    ns, nch = info["n_scalar_slots"], info["n_challenges"]
    scal = np.zeros((n, ns, 32), np.uint8)
    chal = np.zeros((n, nch, 32), np.uint8)
    st = np.zeros(n, np.int32)
    ctx._check(ctx._L.svk_plonk_msm_scalars_batch(ctx._c, pv.pid, n, _ptr(inst), 1, _ptr(proofs), proofs.shape[1], None, _ptr(scal), _ptr(chal), _ptr(st)))
    assert (st == 0).all()
    n_pre = len(vk_dlogs)
    w_prime_ord = n_pts - 1
    w_new = np.zeros((n, 32), np.uint8)
    for i in range(n):
        d_i = _ints(dl[i])
        s_i = _ints(scal[i])
        phi = 0
        for t in range(nt):
            fixed, base, slot = terms[3 * t], terms[3 * t + 1], terms[3 * t + 2]
            if not fixed and base == w_prime_ord:
                continue
            d = (1 if base == n_pre else vk_dlogs[base]) if fixed else d_i[base]
            phi += (1 if slot < 0 else s_i[slot]) * d
        z_prime = int.from_bytes(chal[i, nch - 1].tobytes(), "little")
        w = phi % R * pow((trapdoor_s - z_prime) % R, R - 2, R) % R
        w_new[i] = np.frombuffer(w.to_bytes(32, "little"), np.uint8)
    proofs[:, -32:] = _compress(_g_multiples(ctx, w_new))
    return inst, proofs


def snarks_from_arrays(inst: np.ndarray, proofs: np.ndarray):
    return [Snark([[int.from_bytes(inst[i].tobytes(), "little")]], proofs[i].tobytes()) for i in range(len(proofs))]
