"""ctypes driver of oracle/c/libcref.so (C restatement of the reference's CPU path).
TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py cpu_baseline / --impl reference).

The per-proof operation sequence is recorded ONCE by the Python oracle (`oracle.loader.Tracer` around the
literal restatement of PlonkSuccinctVerifier) and replayed in C for every proof: same primitive operations
NativeLoader executes (21 naive scalar multiplications, 19 Fermat inversions, 27 serial Poseidon
permutations, 11 decompressions per StandardPlonk SHPLONK proof)."""
import ctypes
import os
import subprocess
import time

import numpy as np

from .. import api, forge, poseidon
from ..bn254 import R
from ..loader import NativeLoader, Tracer

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libcref.so")
OPS = {"add": 0, "sub": 1, "mul": 2, "neg": 3, "inv": 4, "msm": 5, "t_squeeze": 6, "t_common_scalar": 7, "t_common_point": 8,
       "t_read_scalar": 9, "t_read_point": 10, "input": 11, "t_clear": 12}


class TraceT(ctypes.Structure):
    _fields_ = [("n_ops", ctypes.c_int), ("ops", ctypes.c_void_p), ("msm_pairs", ctypes.c_void_p), ("n_s", ctypes.c_int), ("n_p", ctypes.c_int),
                ("n_const_s", ctypes.c_int), ("const_s_reg", ctypes.c_void_p), ("const_s_val", ctypes.c_void_p),
                ("n_const_p", ctypes.c_int), ("const_p_reg", ctypes.c_void_p), ("const_p_xy", ctypes.c_void_p),
                ("out_lhs", ctypes.c_int), ("out_rhs", ctypes.c_int)]


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO):
            subprocess.run(["make", "-C", HERE], check=True, timeout=600)
        L = ctypes.CDLL(SO)
        sp = poseidon.spec()
        vals = []
        for row in sp.start:
            vals += row
        vals += sp.partial
        for row in sp.end:
            vals += row
        for m in (sp.mds, sp.pre_sparse_mds):
            for row in m:
                vals += row
        for row, _ in sp.sparse:
            vals += row
        for _, col in sp.sparse:
            vals += col
        vals.append(1 << 64)
        assert len(vals) == 385
        L.cref_set_poseidon(_limbs64(vals).ctypes.data_as(ctypes.c_void_p))
        _LIB = L
    return _LIB


def _limbs64(vals):
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        for j in range(4):
            out[i, j] = (int(v) >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


class Trace:
    """The recorded NativeLoader operation sequence for one (protocol, scheme)."""

    def __init__(self, setup, scheme):
        inst, pf = forge.forge_proof(setup, scheme, 12345)
        tr = Tracer()
        loader = NativeLoader(tr)
        api.succinct_verify(setup.dk.svk, setup.protocol, inst, pf, scheme, loader=loader)
        ops, pairs = [], []
        msm_dsts = []
        for op in tr.ops:
            code = OPS[op[0]]
            if op[0] in ("add", "sub", "mul"):
                ops.append((code, op[1], op[2], op[3]))
            elif op[0] in ("neg", "inv", "input"):
                ops.append((code, op[1], op[2], 0))
            elif op[0] == "msm":
                ops.append((code, op[1], len(pairs), len(op[2])))
                pairs += list(op[2])
                msm_dsts.append(op[1])
            elif op[0] in ("t_squeeze", "t_read_scalar", "t_read_point"):
                ops.append((code, op[1], 0, 0))
            elif op[0] in ("t_common_scalar", "t_common_point"):
                ops.append((code, 0, op[1], 0))
            else:
                ops.append((code, 0, 0, 0))
        self.counts = {k: sum(1 for o in tr.ops if o[0] == k) for k in OPS}
        self.n_scalar_muls = len(pairs)
        self.ops = np.array(ops, dtype=np.int32)
        self.pairs = np.array(pairs, dtype=np.int32).reshape(-1, 2)
        self.cs_reg = np.array(list(tr.consts.values()), dtype=np.int32)
        self.cs_val = _limbs64(list(tr.consts.keys()))
        cps = list(tr.const_points.items())
        self.cp_reg = np.array([r for _, r in cps], dtype=np.int32)
        self.cp_xy = _limbs64([c for pt, _ in cps for c in (pt if pt else (0, 0))])
        self.n_inputs = sum(setup.protocol.num_instance)
        self.proof_len = len(pf)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        self.c = TraceT(len(ops), p(self.ops), p(self.pairs), tr.n_s, tr.n_p, len(self.cs_reg), p(self.cs_reg), p(self.cs_val),
                        len(self.cp_reg), p(self.cp_reg), p(self.cp_xy), msm_dsts[-2], msm_dsts[-1])


def replay(trace, proofs, instances, threads=1):
    """-> (accs uint8[n,128], status int32[n]);  proofs: list of bytes, instances: list of [[int]]"""
    n = len(proofs)
    stride = max(len(p) for p in proofs)
    buf = np.zeros((n, stride), dtype=np.uint8)
    lens = np.zeros(n, dtype=np.int32)
    for i, p in enumerate(proofs):
        buf[i, : len(p)] = np.frombuffer(p, dtype=np.uint8)
        lens[i] = len(p)
    inp = _limbs64([x for inst in instances for col in inst for x in col]).reshape(n, -1)
    return replay_packed(trace, buf, lens, inp, threads)


def replay_packed(trace, buf, lens, inp, threads=1):
    n = buf.shape[0]
    accs = np.zeros((n, 128), dtype=np.uint8)
    st = np.zeros(n, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    lib().cref_replay(ctypes.byref(trace.c), n, p(buf), buf.shape[1], p(lens), p(inp), trace.n_inputs, p(accs), p(st), threads)
    return accs, st


def fold(accs, group_size=0, threads=1):
    """accs uint8[n,128] -> (acc uint8[128], r_root int, status); same tree as oracle.api.fold.
    Groups of one level are independent; `threads` > 1 folds them concurrently (each group itself is the
    reference's serial KzgAs::create_proof)."""
    from concurrent.futures import ThreadPoolExecutor

    L = lib()
    cur = np.ascontiguousarray(accs)
    m = group_size if group_size and group_size > 1 else len(cur)

    def one(grp):
        grp = np.ascontiguousarray(grp)
        out = np.zeros(128, dtype=np.uint8)
        rr = np.zeros(4, dtype=np.uint64)
        st = L.cref_fold_group(len(grp), grp.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), rr.ctypes.data_as(ctypes.c_void_p))
        return st, out, sum(int(rr[j]) << (64 * j) for j in range(4))

    pool = ThreadPoolExecutor(threads) if threads > 1 else None
    while True:
        groups = [cur[i : i + m] for i in range(0, len(cur), m)]
        res = list(pool.map(one, groups)) if pool else [one(x) for x in groups]
        for st, _, _ in res:
            if st:
                return None, None, st
        cur = np.stack([o for _, o, _ in res])
        if len(cur) == 1:
            return cur[0], res[-1][2], 0


def msm(scalars, points, threads=1):
    """`util::msm::multi_scalar_multiplication` (util/msm.rs:238-317) restated in C: scalars uint8[n,32] LE canonical, points
    uint8[n,64] affine canonical -> uint8[64]."""
    sc, pt = np.ascontiguousarray(scalars, dtype=np.uint8), np.ascontiguousarray(points, dtype=np.uint8)
    n = sc.size // 32
    out = np.zeros(64, dtype=np.uint8)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    lib().cref_msm(n, p(sc), p(pt), p(out), threads)
    return out


def decide(acc, dk):
    g2 = _limbs64([dk.g2[0][0], dk.g2[0][1], dk.g2[1][0], dk.g2[1][1]])
    sg2 = _limbs64([dk.s_g2[0][0], dk.s_g2[0][1], dk.s_g2[1][0], dk.s_g2[1][1]])
    a = np.ascontiguousarray(acc)
    return bool(lib().cref_decide(a.ctypes.data_as(ctypes.c_void_p), g2.ctypes.data_as(ctypes.c_void_p), sg2.ctypes.data_as(ctypes.c_void_p)))


_BENCH_CACHE = {}


def bench_baseline(g, sample, group_size, scheme="bdfg21", optimised=True):
    """cpu_baseline object for bench.py: the reference's algorithm on the host cores.  `g` = the product's
    loaded golden fixture (snarks); sample = proofs per measurement (0 = 512)."""
    if "setup" not in _BENCH_CACHE:
        S = forge.Setup(0)
        _BENCH_CACHE["setup"] = S
        _BENCH_CACHE["trace"] = Trace(S, scheme)
    S, tr = _BENCH_CACHE["setup"], _BENCH_CACHE["trace"]
    snarks = g["schemes"][scheme]["snarks"]
    n = sample or 512
    reps = [snarks[i % len(snarks)] for i in range(n)]
    proofs = [s.proof for s in reps]
    insts = [s.instances for s in reps]
    cores = os.cpu_count() or 1
    stride = max(len(p) for p in proofs)
    buf = np.zeros((n, stride), dtype=np.uint8)
    lens = np.zeros(n, dtype=np.int32)
    for i, p in enumerate(proofs):
        buf[i, : len(p)] = np.frombuffer(p, dtype=np.uint8)
        lens[i] = len(p)
    inp = _limbs64([x for inst in insts for col in inst for x in col]).reshape(n, -1)
    if "single" not in _BENCH_CACHE:  # one thread, measured once per process
        n1 = min(n, 64)
        t0 = time.perf_counter()
        replay_packed(tr, buf[:n1], lens[:n1], inp[:n1], 1)
        _BENCH_CACHE["single"] = n1 / (time.perf_counter() - t0)
    single = _BENCH_CACHE["single"]
    t0 = time.perf_counter()
    accs, st = replay_packed(tr, buf, lens, inp, cores)
    assert (st == 0).all()
    acc, r, fst = fold(accs, group_size, threads=cores)
    ok = decide(acc, S.dk)
    dt = time.perf_counter() - t0
    assert ok and fst == 0
    base = {"value": n / dt, "unit": "proofs/s", "cores": cores, "kind": "port", "single_thread_value": single, "sample_proofs": n, "sample_seconds": dt,
            "sample": f"{n} proofs: succinct verify ({cores} threads, one proof per thread) + KzgAs fold (groups of {group_size}) + one pairing; "
                      f"C restatement of the reference algorithm (oracle/c): {tr.n_scalar_muls} naive 256-step scalar muls, "
                      f"{tr.counts['inv']} Fermat inversions, serial Poseidon sponge per proof"}
    if not optimised:
        return base
    # the same sample through the windowed scalar multiplication (value-identical accumulators): what a tuned CPU verifier gains
    lib().cref_set_optimised(1)
    try:
        t0 = time.perf_counter()
        accs2, st2 = replay_packed(tr, buf, lens, inp, cores)
        acc2, r2, fst2 = fold(accs2, group_size, threads=cores)
        ok2 = decide(acc2, S.dk)
        dt2 = time.perf_counter() - t0
    finally:
        lib().cref_set_optimised(0)
    assert ok2 and bytes(acc2) == bytes(acc) and (np.asarray(accs2) == np.asarray(accs)).all()
    base["optimised_value"] = n / dt2
    base["optimised_note"] = ("same sample with 4-bit fixed-window scalar multiplications instead of the reference's 256-step double-and-add-always "
                              "(oracle/c cref_set_optimised); identical accumulators")
    return base
