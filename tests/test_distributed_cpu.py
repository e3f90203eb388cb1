"""World-size-2 gloo test of the proof-sharding / all-gather / final-fold host logic
(snark_verifier_axiom_b200/distributed.py) with the arithmetic provided by the CPU oracle instead of
libsvk (no GPU here).  Checks the sharded result equals: fold(shard 0), fold(shard 1) -> fold of the two."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

N_TOTAL = 10
GROUP = 4


class OracleOps:
    """Stand-in for LibsvkOps: same three operations on CPU tensors, computed by oracle/c."""

    def __init__(self):
        from oracle import forge
        from oracle.c import cref

        self.cref = cref
        self.S = forge.Setup(0)
        self.trace = cref.Trace(self.S, "bdfg21")

    def local_verify(self, d_inst, n_inst, d_proofs, n_batches, n, group_size, d_accs, d_status, d_record, d_lens=None, decide=True):
        assert n_batches == 1
        self.local_decides = getattr(self, "local_decides", 0) + (1 if decide else 0)
        buf = d_proofs.numpy()
        lens = np.full(n, buf.shape[1], dtype=np.int32)
        inp = np.ascontiguousarray(d_inst.numpy()).view(np.uint64).reshape(n, -1)
        accs, st = self.cref.replay_packed(self.trace, np.ascontiguousarray(buf), lens, inp, 1)
        d_accs[: n * 128] = torch.from_numpy(accs.reshape(-1))
        d_status[:n] = torch.from_numpy(st)
        acc, r, fst = self.cref.fold(accs, group_size)
        rec = np.zeros(256, dtype=np.uint8)
        rec[:128] = acc
        ok = self.cref.decide(acc, self.S.dk) if decide else True  # a rank of a sharded job does not run the pairing
        rec[164] = 1 if ok else 0
        rec[165] = 1 if (ok and (st == 0).all() and fst == 0) else 0
        d_record[:] = torch.from_numpy(rec)

    def fold(self, n_seg, n, d_accs, d_record):
        assert n_seg == 1
        acc, r, fst = self.cref.fold(d_accs.numpy().reshape(n, 128), 0)
        d_record[:128] = torch.from_numpy(acc)
        d_record[160:164] = torch.from_numpy(np.array([fst], dtype=np.int32).view(np.uint8))

    def decide(self, n_records, d_record):
        d_record[164] = 1 if self.cref.decide(d_record.numpy()[:128], self.S.dk) else 0


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from snark_verifier_axiom_b200.distributed import ShardedBatchVerifier, shard_range
    from snark_verifier_axiom_b200.standard_plonk import load_golden

    g = load_golden()
    snarks = g["schemes"]["bdfg21"]["snarks"][:N_TOTAL]
    lo, hi = shard_range(N_TOTAL, world, rank)
    mine = snarks[lo:hi]
    proofs = torch.from_numpy(np.stack([np.frombuffer(s.proof, dtype=np.uint8) for s in mine]).copy())
    inst = torch.from_numpy(np.frombuffer(b"".join(int(x).to_bytes(32, "little") for s in mine for col in s.instances for x in col), dtype=np.uint8).copy())
    sv = ShardedBatchVerifier(None, world, rank, torch.device("cpu"), None, group_size=GROUP, ops=OracleOps())
    sv.verify_dev(inst, 1, proofs, len(mine))
    assert sv.ops.local_decides == 0  # world > 1: the single pairing runs after the cross-rank fold
    q.put((rank, lo, hi, sv.last_ok(), sv.final_accumulator()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range():
    from snark_verifier_axiom_b200.distributed import shard_range

    for n, w in [(10, 2), (4096, 8), (7, 4), (3, 8)]:
        spans = [shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_sharded_verify_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1:3] == (0, 5) and res[1][1:3] == (5, 10)
    assert res[0][3] and res[1][3]
    assert res[0][4] == res[1][4]  # every rank ends with the same folded accumulator
    # expected: per-shard tree folds, then a flat fold of the two shard accumulators (oracle, in-process)
    from oracle import api, forge
    from snark_verifier_axiom_b200.standard_plonk import load_golden

    S = forge.Setup(0)
    snarks = load_golden()["schemes"]["bdfg21"]["snarks"][:N_TOTAL]
    pairs = []
    for sn in snarks:
        a = api.succinct_verify(S.dk.svk, S.protocol, sn.instances, sn.proof, "bdfg21")[0]
        pairs.append((a.lhs.pt, a.rhs.pt))
    top = [api.fold(pairs[:5], GROUP)[0], api.fold(pairs[5:], GROUP)[0]]
    (l, r), _ = api.fold(top, 0)
    exp = b"".join(int(v).to_bytes(32, "little") for v in (l[0], l[1], r[0], r[1]))
    assert res[0][4] == exp
    assert api.decide(S.dk, (l, r))


class OracleMsmOps:
    def msm(self, n, d_scalars, d_points, d_out, d_status):
        from oracle import bn254

        s = d_scalars.numpy().reshape(n, 32)
        p = d_points.numpy().reshape(n, 64)
        acc = None
        for i in range(n):
            k = int.from_bytes(s[i].tobytes(), "little")
            x, y = int.from_bytes(p[i, :32].tobytes(), "little"), int.from_bytes(p[i, 32:].tobytes(), "little")
            pt = None if x == 0 and y == 0 else (x, y)
            acc = bn254.g1_add(acc, bn254.g1_mul(pt, k))
        b = bytes(64) if acc is None else acc[0].to_bytes(32, "little") + acc[1].to_bytes(32, "little")
        d_out[:] = torch.from_numpy(np.frombuffer(b, dtype=np.uint8).copy())


def _msm_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import random

    from oracle import bn254
    from snark_verifier_axiom_b200.distributed import msm_sharded, shard_range

    rng = random.Random(11)
    n = 9
    pts = [bn254.g1_mul(bn254.G1_GEN, rng.randrange(1, bn254.R)) for _ in range(n)]
    sc = [rng.randrange(bn254.R) for _ in range(n)]
    lo, hi = shard_range(n, world, rank)
    ds = torch.from_numpy(np.frombuffer(b"".join(s.to_bytes(32, "little") for s in sc[lo:hi]), dtype=np.uint8).copy())
    dp = torch.from_numpy(np.frombuffer(b"".join(p[0].to_bytes(32, "little") + p[1].to_bytes(32, "little") for p in pts[lo:hi]), dtype=np.uint8).copy())
    out = msm_sharded(OracleMsmOps(), world, torch.device("cpu"), ds, dp, hi - lo)
    exp = bn254.g1_msm_naive(list(zip(sc, pts)))
    q.put((rank, out.numpy().tobytes() == exp[0].to_bytes(32, "little") + exp[1].to_bytes(32, "little")))
    dist.barrier()
    dist.destroy_process_group()


def test_msm_sharded_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_msm_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]
