"""GPU parity at the sizes BASELINE.json names, through the C ABI:
  config 2  4096 distinct StandardPlonk SHPLONK proofs, 41 corrupted (evaluation flips, bad point encodings, identity points,
            scalars >= r) against the C restatement of the reference CPU path (oracle/c): every status, every accumulator,
            the folded accumulator + root challenge, the verdict, the located culprits
            (verifier/plonk.rs:98-135, pcs/kzg/accumulation.rs:29-62, pcs/kzg/decider.rs:60-81)
  config 5  2^16 accumulators, 1/64 corrupted: constructed expectation + a sample through the oracle's pairing
  config 3  BN254 MSM at 2^18 and 2^20 (the c = 16 window plan): in-the-exponent equality, uniform and powers-of-r scalars
            (util/msm.rs:238-317)
  config 4  the per-rank shard, 2048 GWC proofs with 21 corrupted, against oracle/c (the cross-rank half is
            tests/test_gpu_multi_rank.py and bench.py's pre-flight)
"""
import ctypes
import os

import numpy as np
import pytest

from oracle import bn254, forge
from oracle.c import cref
from oracle.forge import g_mul

from .util import ptr

pytestmark = pytest.mark.gpu
R = bn254.R


@pytest.fixture(scope="module")
def env():
    from snark_verifier_axiom_b200 import verifier as V
    from snark_verifier_axiom_b200.standard_plonk import load_golden

    g = load_golden()
    S = forge.Setup(0)
    ctx = V.Context(0)
    pv = V.PlonkVerifier(ctx, g["dk"], g["protocol"], V.SHPLONK)
    yield V, g, S, ctx, pv
    ctx.close()


def _inputs(inst):
    """instances u8[n, 32] (one instance per proof) -> the u64 limb layout cref.replay_packed expects"""
    return np.ascontiguousarray(inst).view(np.uint64).reshape(inst.shape[0], -1)


def test_config2_4096_proofs_41_corrupted(env):
    from snark_verifier_axiom_b200 import synth

    V, g, S, ctx, pv = env
    n, m = 4096, 8
    threads = os.cpu_count() or 1
    inst, base = synth.forge_shplonk_batch(pv, g["trapdoor_s"], g["vk_dlogs"], n, seed=20261)
    inst = np.ascontiguousarray(inst)
    assert len({p.tobytes() for p in base}) == n
    rng = np.random.default_rng(5)
    idx = rng.choice(n, 41, replace=False)
    ev, pt_bad, pt_id, sc_bad = idx[:25], idx[25:33], idx[33:37], idx[37:41]
    evalbad = base.copy()
    for k, i in enumerate(ev):  # still decodes: another accumulator, the pairing rejects it
        evalbad[i, 9 * 32 + (k % 5) * 32 + (k % 31)] ^= 1 << (k % 7)
    mixed = evalbad.copy()
    bad_x = next(x for x in range(1, 100) if not bn254.g1_from_bytes(x.to_bytes(32, "little"))[0])
    for k, i in enumerate(pt_bad):  # x^3 + 3 is a non-residue
        mixed[i, 32 * (k % 9) : 32 * (k % 9) + 32] = np.frombuffer(bad_x.to_bytes(32, "little"), np.uint8)
    for k, i in enumerate(pt_id):
        mixed[i, 32 * (2 * k) : 32 * (2 * k) + 32] = 0
    for i in sc_bad:
        mixed[i, 9 * 32 : 10 * 32] = 0xFF
    tr = cref.Trace(S, "bdfg21")
    lens = np.full(n, base.shape[1], np.int32)
    inp = _inputs(inst)
    ref_accs, ref_st = cref.replay_packed(tr, np.ascontiguousarray(mixed), lens, inp, threads)
    assert sorted(np.nonzero(ref_st)[0]) == sorted(np.concatenate([pt_bad, pt_id, sc_bad]))

    # (1) succinct verify of the mixed batch: every status, every accumulator
    L, c = ctx._L, ctx._c
    out_acc = np.zeros((n, 128), np.uint8)
    out_st = np.full(n, -7, np.int32)
    ulens = lens.astype(np.uint32)
    rc = L.svk_plonk_succinct_verify_batch(c, pv.pid, n, ptr(inst), 1, ptr(mixed), mixed.shape[1], ptr(ulens), ptr(out_acc), None, ptr(out_st))
    assert rc == 0, L.svk_last_error(c)
    assert ((out_st & 0xFF) == (ref_st & 0xFF)).all()
    assert all((out_st[i] >> 8) == 3 for i in pt_bad) and all((out_st[i] >> 8) == 4 for i in pt_id) and all((out_st[i] >> 8) == 2 for i in sc_bad)
    ok_rows = ref_st == 0
    assert (out_acc[ok_rows] == ref_accs[ok_rows]).all() and not out_acc[~ok_rows].any()

    # accumulators of the two all-decodable batches from the same replay (+ the few rows the corruption replaced)
    def with_rows(accs, rows, proofs):
        rows = np.asarray(rows)
        a, st = cref.replay_packed(tr, np.ascontiguousarray(proofs[rows]), lens[: len(rows)], np.ascontiguousarray(inp[rows]), threads)
        assert (st == 0).all()
        out = accs.copy()
        out[rows] = a
        return out

    acc_evalbad = with_rows(ref_accs, np.concatenate([pt_bad, pt_id, sc_bad]), evalbad)
    acc_valid = with_rows(acc_evalbad, ev, base)

    # (2) PlonkVerifier::verify over the batch with the 25 evaluation-corrupted proofs: fold tree bit-exact, one pairing rejects,
    # decide_all names exactly the culprits
    st2 = np.zeros(n, np.int32)
    folded = np.zeros(128, np.uint8)
    okb = np.zeros(1, np.uint8)
    rc = L.svk_plonk_verify_batch(c, pv.pid, n, ptr(inst), 1, ptr(evalbad), evalbad.shape[1], ptr(ulens), m, 1, ptr(st2), ptr(folded), ptr(okb))
    assert rc == 0, L.svk_last_error(c)
    facc, r_root, fst = cref.fold(acc_evalbad, m, threads=threads)
    assert fst == 0 and not okb[0] and cref.decide(facc, S.dk) is False
    assert sorted(np.nonzero(st2)[0]) == sorted(ev) and all(st2[i] == 3 for i in ev)
    d_accs = np.zeros((n, 128), np.uint8)
    d_rec = np.zeros(256, np.uint8)
    # the folded accumulator of a REJECTED batch through the device-pointer entry (the host call returns it only on accept)
    import torch

    dev = torch.device("cuda", 0)
    t_inst, t_pf, t_len = torch.from_numpy(inst).to(dev), torch.from_numpy(evalbad).to(dev), torch.from_numpy(ulens.astype(np.int32)).to(dev)
    t_accs, t_st, t_rec = torch.zeros(n * 128, dtype=torch.uint8, device=dev), torch.zeros(n, dtype=torch.int32, device=dev), torch.zeros(256, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    rc = L.svk_plonk_verify_multi_dev(c, pv.pid, 1, n, p(t_inst), 1, p(t_pf), evalbad.shape[1], p(t_len), m, p(t_accs), p(t_st), p(t_rec))
    assert rc == 0, L.svk_last_error(c)
    ctx.sync()
    d_accs[:] = t_accs.cpu().numpy().reshape(n, 128)
    d_rec[:] = t_rec.cpu().numpy()
    assert (d_accs == acc_evalbad).all()
    assert (d_rec[:128] == facc).all() and int.from_bytes(d_rec[128:160].tobytes(), "little") == r_root
    assert d_rec[164] == 0 and d_rec[165] == 0 and not d_rec[160:164].any()

    # (3) the all-valid batch: accepted, same fold as the oracle's
    rc = L.svk_plonk_verify_batch(c, pv.pid, n, ptr(inst), 1, ptr(base), base.shape[1], ptr(ulens), m, 1, ptr(st2), ptr(folded), ptr(okb))
    assert rc == 0, L.svk_last_error(c)
    facc, r_root, fst = cref.fold(acc_valid, m, threads=threads)
    assert fst == 0 and okb[0] == 1 and (st2 == 0).all() and (folded == facc).all() and cref.decide(facc, S.dk) is True


def test_config5_decide_65536_accumulators(env):
    V, g, S, ctx, pv = env
    n = 1 << 16
    accs, _, st = pv.succinct_verify(g["schemes"]["bdfg21"]["snarks"])
    assert (st == 0).all()
    base = np.frombuffer(b"".join(a.to_bytes() for a in accs), dtype=np.uint8).reshape(len(accs), 128)
    host = np.tile(base, (n // len(accs) + 1, 1))[:n].copy()
    bad = np.arange(0, n, 64)
    host[bad, 64:128] = host[(bad + 1) % n, 64:128]  # another valid rhs: on the curve, the pairing must reject
    host[5] = 0                                      # (identity, identity): e(O, .) e(O, .) = 1 accepts (decider.rs:60-68)
    host[70, 0] ^= 1                                 # off the curve: not a G1Affine, reported as reject
    expect = np.ones(n, np.uint8)
    expect[bad] = 0
    expect[70] = 0
    kid = ctx.load_deciding_key(g["dk"])
    got = np.zeros(n, np.uint8)
    rc = ctx._L.svk_kzg_decide_batch(ctx._c, kid, n, ptr(host), ptr(got))
    assert rc == 0, ctx._L.svk_last_error(ctx._c)
    assert (got == expect).all()
    # a sample through the oracle's pairing (incl. corrupted rows); the same rows again through the block-cooperative kernel
    sample = np.concatenate([np.arange(0, 256), np.arange(n - 64, n)])
    for i in sample:
        if i != 70:  # the C oracle takes G1Affine inputs (on the curve) like the reference
            assert cref.decide(host[i], S.dk) == bool(expect[i]), i
    sub = np.ascontiguousarray(host[sample])
    got2 = np.zeros(len(sample), np.uint8)
    assert ctx._L.svk_kzg_decide_batch(ctx._c, kid, len(sample), ptr(sub), ptr(got2)) == 0  # <= 512: k_decide_coop
    assert (got2 == expect[sample]).all()


def _scalars_le(vals):
    return np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vals), np.uint8).copy()


@pytest.mark.parametrize("log_n", [18, 20])
def test_config3_msm_in_the_exponent(env, log_n):
    V, g, S, ctx, pv = env
    n = 1 << log_n
    rng = np.random.default_rng(100 + log_n)
    L, c = ctx._L, ctx._c

    def rand_scalars(k):
        b = rng.integers(0, 256, (k, 32), dtype=np.uint8)
        b[:, 31] &= 0x1F  # < 2^253 < r
        return b

    def ints(b):
        return [int.from_bytes(row.tobytes(), "little") for row in b]

    dl = rand_scalars(n)
    gen = np.frombuffer((1).to_bytes(32, "little") + (2).to_bytes(32, "little"), np.uint8).copy()
    pts = np.zeros((n, 64), np.uint8)
    assert L.svk_g1_mul_batch(c, n, ptr(dl), ptr(gen), 1, ptr(pts)) == 0
    d = ints(dl)
    out, st = np.zeros(64, np.uint8), np.zeros(1, np.int32)

    def msm(sc):
        assert L.svk_msm_g1(c, n, ptr(sc), ptr(pts), ptr(out), ptr(st)) == 0 and st[0] == 0
        return V._g1_from(out.tobytes())

    a = rand_scalars(n)
    a[0], a[1], a[2] = 0, _scalars_le([1])[:32], _scalars_le([R - 1])[:32]
    ai = ints(a)
    assert msm(a) == g_mul(sum(x * y for x, y in zip(ai, d)) % R)
    # the fold's scalar distribution (accumulation.rs:51-59): 1, r, r^2, ...
    r = int.from_bytes(rng.integers(0, 256, 31, dtype=np.uint8).tobytes(), "little")
    pw, cur = [], 1
    for _ in range(n):
        pw.append(cur)
        cur = cur * r % R
    assert msm(_scalars_le(pw).reshape(n, 32)) == g_mul(sum(x * y for x, y in zip(pw, d)) % R)


def test_config4_gwc_2048_proofs_per_rank():
    """BASELINE config 4's per-rank shard: 2048 GWC (Gwc19) proofs -- the 64 golden ones tiled, 21 of them corrupted -- against the C
    restatement: statuses, accumulators, the fold tree, the verdict (pcs/kzg/multiopen/gwc19.rs:21-81, accumulation.rs:29-62)."""
    from snark_verifier_axiom_b200 import verifier as V
    from snark_verifier_axiom_b200.standard_plonk import load_golden

    g = load_golden()
    S = forge.Setup(0)
    ctx = V.Context(0)
    try:
        pv = V.PlonkVerifier(ctx, g["dk"], g["protocol"], V.GWC)
        n, m = 2048, 4
        threads = os.cpu_count() or 1
        sn = g["schemes"]["gwc19"]["snarks"]
        inst, n_inst, base, lens = pv.pack([sn[i % len(sn)] for i in range(n)])
        assert n_inst == 1
        rng = np.random.default_rng(11)
        idx = rng.choice(n, 21, replace=False)
        ev, pt_id, sc_bad = idx[:13], idx[13:17], idx[17:]
        evalbad = base.copy()
        for k, i in enumerate(ev):  # an evaluation flipped: still decodes, another accumulator, the pairing rejects
            evalbad[i, 9 * 32 + (k % 17) * 32 + (k % 31)] ^= 1 << (k % 7)
        mixed = evalbad.copy()
        w_points = (9 + 17) * 32                             # 9 commitments, 17 evaluations, then W_0..W_2 (gwc19.rs:62-81)
        for k, i in enumerate(pt_id):
            off = 32 * (2 * k) if k < 3 else w_points + 32   # identity among the commitments / as W_1
            mixed[i, off : off + 32] = 0
        for i in sc_bad:
            mixed[i, 9 * 32 : 10 * 32] = 0xFF              # scalar >= r
        tr = cref.Trace(S, "gwc19")
        ilens = lens.astype(np.int32)
        inp = _inputs(inst)
        ref_accs, ref_st = cref.replay_packed(tr, np.ascontiguousarray(mixed), ilens, inp, threads)
        assert sorted(np.nonzero(ref_st)[0]) == sorted(np.concatenate([pt_id, sc_bad]))
        L, c = ctx._L, ctx._c
        out_acc = np.zeros((n, 128), np.uint8)
        out_st = np.full(n, -7, np.int32)
        rc = L.svk_plonk_succinct_verify_batch(c, pv.pid, n, ptr(inst), 1, ptr(mixed), mixed.shape[1], ptr(lens), ptr(out_acc), None, ptr(out_st))
        assert rc == 0, L.svk_last_error(c)
        assert ((out_st & 0xFF) == (ref_st & 0xFF)).all()
        ok_rows = ref_st == 0
        assert (out_acc[ok_rows] == ref_accs[ok_rows]).all() and not out_acc[~ok_rows].any()
        # the batch whose corrupted proofs all decode: fold tree bit-exact, one pairing rejects, decide_all names the culprits
        acc_evalbad, st_e = cref.replay_packed(tr, np.ascontiguousarray(evalbad), ilens, inp, threads)
        assert (st_e == 0).all()
        st2, folded, okb = np.zeros(n, np.int32), np.zeros(128, np.uint8), np.zeros(1, np.uint8)
        rc = L.svk_plonk_verify_batch(c, pv.pid, n, ptr(inst), 1, ptr(evalbad), evalbad.shape[1], ptr(lens), m, 1, ptr(st2), ptr(folded), ptr(okb))
        assert rc == 0, L.svk_last_error(c)
        facc, _, fst = cref.fold(acc_evalbad, m, threads=threads)
        assert fst == 0 and not okb[0] and cref.decide(facc, S.dk) is False
        assert sorted(np.nonzero(st2)[0]) == sorted(ev) and all(st2[i] == 3 for i in ev)
        # all valid: accepted, the oracle's fold
        acc_valid, st_v = cref.replay_packed(tr, np.ascontiguousarray(base), ilens, inp, threads)
        assert (st_v == 0).all()
        rc = L.svk_plonk_verify_batch(c, pv.pid, n, ptr(inst), 1, ptr(base), base.shape[1], ptr(lens), m, 1, ptr(st2), ptr(folded), ptr(okb))
        assert rc == 0, L.svk_last_error(c)
        facc, _, fst = cref.fold(acc_valid, m, threads=threads)
        assert fst == 0 and okb[0] == 1 and (st2 == 0).all() and (folded == facc).all() and cref.decide(facc, S.dk) is True
    finally:
        ctx.close()
