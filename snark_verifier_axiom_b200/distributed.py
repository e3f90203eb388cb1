"""Proof-sharded batch verification over the GPUs of one box.

Proofs are independent until the fold, so rank g verifies and folds its own contiguous shard with no
data-path collective.  The only exchange is the `world` per-rank folded accumulators (256 B records):
one `all_gather` (NCCL over NVLink; EC addition is not an `ncclRedOp`, so "reduce" = all-gather +
fold), after which every rank folds the `world` accumulators with `KzgAs` once more and runs the single
final pairing.  This is the tree fold of `svk_kzg_as_fold` with the top level cut along rank
boundaries; the oracle reproduces it with `api.fold` per shard + `api.fold` of the results.

The collective plumbing is `torch.distributed`; the arithmetic is libsvk.  `ops` is injectable so the
sharding / gather / final-fold logic is testable on CPU with the gloo backend (tests/test_distributed_cpu.py).
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

RECORD = 256  # { svk_acc folded (128) ; svk_fe r (32) ; int32 fold_status ; uint8 decide_ok ; uint8 ok ; pad }
OFF_FOLD_STATUS, OFF_DECIDE_OK, OFF_OK = 160, 164, 165


def shard_range(n_total: int, world: int, rank: int):
    """contiguous index range of `rank` (keeps proofs in order; SURVEY 8e 'partitioning')"""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class LibsvkOps:
    """Device arithmetic through the C ABI (`*_dev` entry points, device pointers)."""

    def __init__(self, pv):
        self.pv = pv
        self.ctx = pv.ctx
        self.L, self.c = pv.ctx._L, pv.ctx._c

    @staticmethod
    def _p(t):
        return ctypes.c_void_p(t.data_ptr())

    def local_verify(self, d_inst, n_inst, d_proofs, n, group_size, d_accs, d_status, d_record, d_lens=None):
        rc = self.L.svk_plonk_verify_batch_dev(self.c, self.pv.pid, n, self._p(d_inst), n_inst, self._p(d_proofs), d_proofs.shape[1],
                                               self._p(d_lens) if d_lens is not None else None, group_size, self._p(d_accs), self._p(d_status),
                                               self._p(d_record))
        self.ctx._check(rc)

    def fold(self, n, d_accs, d_record):
        """flat KzgAs fold of n accumulators -> record[0:128] acc, [128:160] r, [160:164] status"""
        rc = self.L.svk_kzg_as_fold_dev(self.c, n, self._p(d_accs), 0, self._p(d_record), ctypes.c_void_p(d_record.data_ptr() + 128),
                                        ctypes.c_void_p(d_record.data_ptr() + OFF_FOLD_STATUS))
        self.ctx._check(rc)

    def decide(self, d_record):
        rc = self.L.svk_kzg_decide_batch_dev(self.c, self.pv.kzg_as.dk_id, 1, self._p(d_record), ctypes.c_void_p(d_record.data_ptr() + OFF_DECIDE_OK))
        self.ctx._check(rc)


class ShardedBatchVerifier:
    def __init__(self, pv, world, rank, device, stream=None, group_size=8, ops=None):
        self.world, self.rank, self.device, self.stream = world, rank, device, stream
        self.group_size = group_size
        self.ops = ops or LibsvkOps(pv)
        self.pv = pv
        self._n = 0
        kw = dict(dtype=torch.uint8, device=device)
        self.d_record = torch.zeros(RECORD, **kw)
        self.d_gather = torch.zeros(world * RECORD, **kw)
        self.d_final = torch.zeros(RECORD, **kw)
        self.d_accs = self.d_status = None

    def _ensure(self, n):
        if self._n < n:
            self.d_accs = torch.zeros(n * 128, dtype=torch.uint8, device=self.device)
            self.d_status = torch.zeros(n, dtype=torch.int32, device=self.device)
            self._n = n

    def _on_stream(self):
        if self.stream is not None and self.device.type == "cuda":
            return torch.cuda.stream(self.stream)
        import contextlib

        return contextlib.nullcontext()

    def verify_dev(self, d_inst, n_inst, d_proofs, n, d_lens=None):
        """Enqueue: local succinct verify + fold + decide of this rank's shard; if world > 1 all_gather the
        per-rank records, fold the `world` accumulators and decide.  No host synchronisation."""
        self._ensure(n)
        self.ops.local_verify(d_inst, n_inst, d_proofs, n, self.group_size, self.d_accs, self.d_status, self.d_record, d_lens)
        if self.world > 1:
            with self._on_stream():
                dist.all_gather_into_tensor(self.d_gather, self.d_record)
                accs = self.d_gather.view(self.world, RECORD)[:, :128].contiguous()
            self.ops.fold(self.world, accs, self.d_final)
            self.ops.decide(self.d_final)
            self._keep = accs

    def last_ok(self) -> bool:
        """Host read of the verdict (synchronises)."""
        if self.world == 1:
            rec = self.d_record.cpu().numpy()
            return bool(rec[OFF_OK])
        g = self.d_gather.cpu().numpy().reshape(self.world, RECORD)
        f = self.d_final.cpu().numpy()
        fold_status = int(np.frombuffer(f[OFF_FOLD_STATUS : OFF_FOLD_STATUS + 4].tobytes(), np.int32)[0])
        return bool(g[:, OFF_OK].all() and fold_status == 0 and f[OFF_DECIDE_OK])

    def final_accumulator(self) -> bytes:
        src = self.d_record if self.world == 1 else self.d_final
        return src.cpu().numpy()[:128].tobytes()

    def verify_host(self, h_inst, n_inst, h_proofs, h_lens, n):
        """End-to-end call with HOST buffers: H2D of instances + proofs (+ lengths), the batch verification,
        D2H of the per-proof statuses and the verdict.  -> (ok, status int32[n])"""
        if self.world == 1 and isinstance(self.ops, LibsvkOps):
            if getattr(self, "_h_n", 0) < n:  # pinned result buffers, reused: pageable targets make the copies synchronous
                pin = self.device.type == "cuda"
                self._h_st = torch.zeros(n, dtype=torch.int32, pin_memory=pin).numpy()
                self._h_folded = torch.zeros(128, dtype=torch.uint8, pin_memory=pin).numpy()
                self._h_ok = torch.zeros(8, dtype=torch.uint8, pin_memory=pin).numpy()
                self._h_n = n
            st, folded, ok = self._h_st[:n], self._h_folded, self._h_ok
            o = self.ops
            rc = o.L.svk_plonk_verify_batch(o.c, self.pv.pid, n, h_inst.ctypes.data_as(ctypes.c_void_p), n_inst,
                                            h_proofs.ctypes.data_as(ctypes.c_void_p), h_proofs.shape[1],
                                            h_lens.ctypes.data_as(ctypes.c_void_p), self.group_size, 0, st.ctypes.data_as(ctypes.c_void_p),
                                            folded.ctypes.data_as(ctypes.c_void_p), ok.ctypes.data_as(ctypes.c_void_p))
            o.ctx._check(rc)
            return bool(ok[0]), st
        with self._on_stream():
            d_inst = torch.from_numpy(h_inst).to(self.device, non_blocking=True)
            d_proofs = torch.from_numpy(h_proofs).to(self.device, non_blocking=True)
            d_lens = torch.from_numpy(h_lens.astype(np.int32)).to(self.device, non_blocking=True)
        self.verify_dev(d_inst, n_inst, d_proofs, n, d_lens)
        with self._on_stream():
            st = self.d_status[:n].cpu().numpy()
        return self.last_ok(), st
