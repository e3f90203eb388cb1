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
        self.sharded = False  # True once nccl_init gave the context its own communicator

    @staticmethod
    def _p(t):
        return ctypes.c_void_p(t.data_ptr())

    def local_verify(self, d_inst, n_inst, d_proofs, n_batches, batch, group_size, d_accs, d_status, d_records, d_lens=None, decide=True):
        """n_batches batches of `batch` proofs: succinct verify + fold (+ decide) + verdict, one record per batch.  decide=False
        is the per-rank half of a sharded job: the single pairing runs after the cross-rank fold."""
        fn = self.L.svk_plonk_verify_multi_dev if decide else self.L.svk_plonk_fold_multi_dev
        rc = fn(self.c, self.pv.pid, n_batches, batch, self._p(d_inst), n_inst, self._p(d_proofs), d_proofs.shape[1],
                                               self._p(d_lens) if d_lens is not None else None, group_size, self._p(d_accs), self._p(d_status),
                                               self._p(d_records))
        self.ctx._check(rc)

    def nccl_init(self, world, rank, group=None):
        """Creates this context's own NCCL communicator (C ABI: svk_nccl_unique_id / svk_nccl_init).  The 128-byte unique id travels
        from rank 0 through `torch.distributed` (any side channel would do); collective over the job."""
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (ctypes.c_uint8 * 128)()
            if self.L.svk_nccl_unique_id(buf) != 0:
                raise RuntimeError("libsvk: NCCL is not available")
            idt = torch.tensor(list(buf), dtype=torch.uint8)
        dev = torch.device("cuda", self.ctx.device)
        idd = idt.to(dev)
        dist.broadcast(idd, src=0, group=group)
        ids = bytes(idd.cpu().tolist())
        self.ctx._check(self.L.svk_nccl_init(self.c, world, rank, ids))
        self.sharded = True

    def sharded_verify(self, d_inst, n_inst, d_proofs, n_batches, batch, group_size, d_accs, d_status, d_records, d_gather, d_final, d_lens=None):
        """The whole sharded call inside the library: local verify + fold, ncclAllGather, cross-rank fold, one pairing per batch."""
        rc = self.L.svk_plonk_verify_sharded_dev(self.c, self.pv.pid, n_batches, batch, self._p(d_inst), n_inst, self._p(d_proofs), d_proofs.shape[1],
                                                 self._p(d_lens) if d_lens is not None else None, group_size, self._p(d_accs), self._p(d_status),
                                                 self._p(d_records), self._p(d_gather), self._p(d_final))
        self.ctx._check(rc)

    def fold(self, n_seg, n, d_accs, d_records):
        """flat KzgAs fold of n_seg x n accumulators -> one record per segment"""
        self.ctx._check(self.L.svk_kzg_as_fold_multi_dev(self.c, n_seg, n, self._p(d_accs), 0, self._p(d_records)))

    def decide(self, n_records, d_records):
        self.ctx._check(self.L.svk_kzg_decide_records_dev(self.c, self.pv.kzg_as.dk_id, n_records, self._p(d_records)))


class ShardedBatchVerifier:
    """Verifies `n_batches` batches per call on this rank's shard; with world > 1 the per-rank batch accumulators are
    all-gathered and batch b of every rank is folded into the global accumulator of batch b, then decided."""

    def __init__(self, pv, world, rank, device, stream=None, group_size=8, ops=None, max_batches=1, own_nccl=True):
        self.world, self.rank, self.device, self.stream = world, rank, device, stream
        self.group_size = group_size
        self.ops = ops or LibsvkOps(pv)
        self.pv = pv
        self._n = 0
        self.max_batches = max_batches
        # libsvk must run on the stream the collectives and the buffer initialisation are ordered on
        if stream is not None and device.type == "cuda" and isinstance(self.ops, LibsvkOps):
            pv.ctx.set_stream(stream.cuda_stream)
        kw = dict(dtype=torch.uint8, device=device)
        with self._on_stream():
            self.d_records = torch.zeros(max_batches * RECORD, **kw)
            self.d_gather = torch.zeros(world * max_batches * RECORD, **kw)
            self.d_final = torch.zeros(max_batches * RECORD, **kw)
        self.d_accs = self.d_status = None
        self.nb = 1
        # world > 1 on GPUs: the context gets its own communicator and the exchange runs inside libsvk; `own_nccl=False` (or an
        # injected `ops`, as in the gloo CPU test) keeps the torch.distributed all_gather below
        if own_nccl and world > 1 and device.type == "cuda" and isinstance(self.ops, LibsvkOps):
            self.ops.nccl_init(world, rank)

    def _ensure(self, n):
        if self._n < n:
            apk = 1 + getattr(self.pv, "info", {}).get("n_old_accumulators", 0)  # [new, old...] per proof (verifier/plonk.rs:86-91)
            self._sync()  # kernels of an earlier call may still use the buffers being replaced
            with self._on_stream():
                self.d_accs = torch.zeros(n * apk * 128, dtype=torch.uint8, device=self.device)
                self.d_status = torch.zeros(n, dtype=torch.int32, device=self.device)
            self._n = n

    def _sync(self):
        if self.stream is not None and self.device.type == "cuda":
            self.stream.synchronize()

    def _on_stream(self):
        if self.stream is not None and self.device.type == "cuda":
            return torch.cuda.stream(self.stream)
        import contextlib

        return contextlib.nullcontext()

    def verify_dev(self, d_inst, n_inst, d_proofs, n, d_lens=None, n_batches=1):
        """Enqueue (no host synchronisation): `n_batches` batches of n // n_batches proofs each -- local succinct verify
        + fold + decide per batch; if world > 1 all_gather the records, fold batch b over the ranks and decide."""
        assert n % n_batches == 0 and n_batches <= self.max_batches
        self._ensure(n)
        self.nb = n_batches
        nb = n_batches
        if self.world > 1 and getattr(self.ops, "sharded", False):
            # everything behind the C ABI (csrc/sharded.cu): ncclAllGather on the context stream, cross-rank fold, one pairing
            self.ops.sharded_verify(d_inst, n_inst, d_proofs, nb, n // nb, self.group_size, self.d_accs, self.d_status, self.d_records,
                                    self.d_gather, self.d_final, d_lens)
            return
        self.ops.local_verify(d_inst, n_inst, d_proofs, nb, n // nb, self.group_size, self.d_accs, self.d_status, self.d_records, d_lens,
                              decide=self.world == 1)
        if self.world > 1:
            with self._on_stream():
                dist.all_gather_into_tensor(self.d_gather[: self.world * nb * RECORD], self.d_records[: nb * RECORD])
                # [rank][batch][256] -> accumulators [batch][rank][128]
                accs = self.d_gather[: self.world * nb * RECORD].view(self.world, nb, RECORD)[:, :, :128].permute(1, 0, 2).contiguous()
            self.ops.fold(nb, self.world, accs, self.d_final)
            self.ops.decide(nb, self.d_final)
            self._keep = accs

    def last_ok(self):
        """Host read of the verdicts (synchronises) -> bool (all batches of the last call accepted)."""
        return all(self.last_verdicts())

    def last_verdicts(self):
        nb = self.nb
        self._sync()  # `.cpu()` below only orders against torch's current stream
        if self.world == 1:
            rec = self.d_records[: nb * RECORD].cpu().numpy().reshape(nb, RECORD)
            return [bool(x) for x in rec[:, OFF_OK]]
        g = self.d_gather[: self.world * nb * RECORD].cpu().numpy().reshape(self.world, nb, RECORD)
        f = self.d_final[: nb * RECORD].cpu().numpy().reshape(nb, RECORD)
        out = []
        for b in range(nb):
            fold_status = int(np.frombuffer(f[b, OFF_FOLD_STATUS : OFF_FOLD_STATUS + 4].tobytes(), np.int32)[0])
            ok = bool(g[:, b, OFF_OK].all() and fold_status == 0 and f[b, OFF_DECIDE_OK])
            if getattr(self.ops, "sharded", False):
                assert ok == bool(f[b, OFF_OK]), "k_sharded_verdict disagrees with the host-side combination"
            out.append(ok)
        return out

    def final_accumulator(self, b=0) -> bytes:
        self._sync()
        src = self.d_records if self.world == 1 else self.d_final
        return src.cpu().numpy()[b * RECORD : b * RECORD + 128].tobytes()

    def verify_host(self, h_inst, n_inst, h_proofs, h_lens, n, n_batches=1):
        """End-to-end call with HOST buffers (single GPU): H2D of instances + proofs (+ lengths), the verification of
        `n_batches` batches, D2H of the per-proof statuses and the per-batch records.  -> (ok, status int32[n])"""
        assert self.world == 1 and isinstance(self.ops, LibsvkOps)
        if getattr(self, "_h_n", 0) < n or getattr(self, "_h_nb", 0) < n_batches:
            pin = self.device.type == "cuda"  # pinned result buffers, reused: pageable targets make the copies synchronous
            self._h_st = torch.zeros(n, dtype=torch.int32, pin_memory=pin).numpy()
            self._h_rec = torch.zeros(n_batches * RECORD, dtype=torch.uint8, pin_memory=pin).numpy()
            self._h_n, self._h_nb = n, n_batches
        st, rec = self._h_st[:n], self._h_rec[: n_batches * RECORD]
        o = self.ops
        rc = o.L.svk_plonk_verify_multi(o.c, self.pv.pid, n_batches, n // n_batches, h_inst.ctypes.data_as(ctypes.c_void_p), n_inst,
                                        h_proofs.ctypes.data_as(ctypes.c_void_p), h_proofs.shape[1], h_lens.ctypes.data_as(ctypes.c_void_p),
                                        self.group_size, 0, st.ctypes.data_as(ctypes.c_void_p), rec.ctypes.data_as(ctypes.c_void_p))
        o.ctx._check(rc)
        return bool(rec.reshape(n_batches, RECORD)[:, OFF_OK].all()), st


# ---------------------------------------------------------------------------------------------- sharded G1 MSM (config 3)
class LibsvkMsmOps:
    def __init__(self, ctx):
        self.ctx = ctx

    def msm(self, n, d_scalars, d_points, d_out, d_status):
        c = self.ctx
        c._check(c._L.svk_msm_g1_dev(c._c, n, ctypes.c_void_p(d_scalars.data_ptr()), ctypes.c_void_p(d_points.data_ptr()),
                                     ctypes.c_void_p(d_out.data_ptr()), ctypes.c_void_p(d_status.data_ptr())))


def msm_sharded(ops, world, device, d_scalars, d_points, n_local, stream=None):
    """`util::msm::multi_scalar_multiplication` over points sharded across ranks (contiguous n/world shards):
    every rank reduces its shard with the Pippenger kernel, the `world` partial results (64 B each) are all-gathered,
    and each rank adds them -- EC addition is not an NCCL reduce op, so the "reduce" is an all-gather followed by an MSM
    with unit scalars over `world` points.  Returns a uint8[64] device tensor (canonical affine point)."""
    kw = dict(dtype=torch.uint8, device=device)
    part = torch.zeros(64, **kw)
    st = torch.zeros(2, dtype=torch.int32, device=device)
    ops.msm(n_local, d_scalars, d_points, part, st)
    if world == 1:
        return part
    import contextlib

    cm = torch.cuda.stream(stream) if (stream is not None and device.type == "cuda") else contextlib.nullcontext()
    with cm:
        gathered = torch.zeros(world * 64, **kw)
        dist.all_gather_into_tensor(gathered, part)
        ones = torch.zeros(world, 32, **kw)
        ones[:, 0] = 1
    out = torch.zeros(64, **kw)
    ops.msm(world, ones.view(-1), gathered, out, st[1:])
    return out
