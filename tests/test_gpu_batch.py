"""GPU parity: KzgAs fold (flat + tree) and the composed PlonkVerifier::verify over a batch, through
the host-side mirror (snark_verifier_axiom_b200.verifier) -- reads like the reference's own use at
snark-verifier-sdk/src/halo2/aggregation.rs:216-245 + decider.rs:60-68."""
import pytest

from oracle import api, forge

from .util import to_product_protocol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from snark_verifier_axiom_b200 import verifier as V

    S = forge.Setup(0)
    ctx = V.Context(0)
    dk = V.KzgDecidingKey.new(S.dk.svk.g, S.dk.g2, S.dk.s_g2)
    AS = V.KzgAs(ctx, dk)
    proto = to_product_protocol(S.protocol)
    pv = {m: V.PlonkVerifier(ctx, dk, proto, m, kzg_as=AS) for m in (V.SHPLONK, V.GWC)}
    yield V, S, ctx, AS, pv
    ctx.close()


@pytest.mark.parametrize("scheme,mos", [("bdfg21", 0), ("gwc19", 1)])
def test_batch_verify_and_fold(env, scheme, mos):
    V, S, ctx, AS, pv = env
    n = 21
    insts, proofs = forge.forge_batch(S, scheme, n, seed0=500)
    snarks = [V.Snark(i, p) for i, p in zip(insts, proofs)]
    accs, chals, st = pv[mos].succinct_verify(snarks)
    assert (st == 0).all()
    oracle_accs = [api.succinct_verify(S.dk.svk, S.protocol, i, p, scheme)[0] for i, p in zip(insts, proofs)]
    pairs = [(a.lhs.pt, a.rhs.pt) for a in oracle_accs]
    assert [(a.lhs, a.rhs) for a in accs] == pairs
    # fold: flat (the reference's aggregation.rs:235-245) and trees, root challenge + folded accumulator
    for m in (0, 2, 4, 8):
        (elhs, erhs), rs = api.fold(pairs, m)
        got, r = AS.create_proof(accs, m)
        assert (got.lhs, got.rhs) == (elhs, erhs), m
        assert r == rs[-1], m
        AS.decide(got)
    # the composed path: one pairing for the batch
    for m in (0, 4):
        res = pv[mos].verify(snarks, group_size=m)
        assert res.ok and (res.status == 0).all()
        assert (res.folded.lhs, res.folded.rhs) == api.fold(pairs, m)[0]
    pv[mos].verify_one(snarks[0])
    # one corrupted evaluation: batch rejects and the offender is located
    bad = bytearray(proofs[5])
    bad[9 * 32 + 40] ^= 4
    snarks2 = list(snarks)
    snarks2[5] = V.Snark(insts[5], bytes(bad))
    res = pv[mos].verify(snarks2, group_size=4)
    assert not res.ok
    assert [int(s) for s in res.status] == [3 if i == 5 else 0 for i in range(n)]
    with pytest.raises(V.Error) as e:
        pv[mos].verify_one(snarks2[5])
    assert e.value.kind == "AssertionFailure"
    assert api.status_of(api.verify, S.dk, S.protocol, insts[5], bytes(bad), scheme) == 3
    # an undecodable proof: Transcript error, no fold
    bad = bytearray(proofs[2])
    bad[9 * 32 : 10 * 32] = b"\xff" * 32
    snarks3 = list(snarks)
    snarks3[2] = V.Snark(insts[2], bytes(bad))
    res = pv[mos].verify(snarks3)
    assert not res.ok and (res.status[2] & 0xFF) == 4 and res.status[0] == 0


def test_fold_rejects_identity(env):
    V, S, ctx, AS, pv = env
    from oracle.forge import g_mul

    a = V.KzgAccumulator(g_mul(5), g_mul(7))
    with pytest.raises(V.Error) as e:
        AS.create_proof([a, V.KzgAccumulator(None, g_mul(3))])
    assert e.value.kind == "Transcript"


def test_verify_batches_equals_separate_calls(env):
    """svk_plonk_verify_multi: 3 batches in one call == 3 separate PlonkVerifier::verify batches (folded accumulator,
    verdict, statuses), including one batch with a corrupted proof."""
    V, S, ctx, AS, pv = env
    insts, proofs = forge.forge_batch(S, "bdfg21", 15, seed0=800)
    snarks = [V.Snark(i, p) for i, p in zip(insts, proofs)]
    bad = bytearray(proofs[7])
    bad[9 * 32 + 11] ^= 8
    snarks[7] = V.Snark(insts[7], bytes(bad))
    batches = [snarks[0:5], snarks[5:10], snarks[10:15]]
    for m in (0, 2):
        multi = pv[0].verify_batches(batches, group_size=m)
        for b, res in zip(batches, multi):
            single = pv[0].verify(b, group_size=m)
            assert res.ok == single.ok
            assert [int(x) for x in res.status] == [int(x) for x in single.status]
            if res.ok:
                assert (res.folded.lhs, res.folded.rhs) == (single.folded.lhs, single.folded.rhs)
        assert [r.ok for r in multi] == [True, False, True]
        assert [int(x) for x in multi[1].status] == [0, 0, 3, 0, 0]


def test_sharded_fold_then_single_pairing(env):
    """SURVEY 8e on one GPU: two 'ranks' fold their shards WITHOUT deciding (svk_plonk_fold_multi_dev), the per-rank accumulators
    are folded once more and decided once (svk_kzg_as_fold_multi_dev + svk_kzg_decide_records_dev) -- equals the oracle's
    fold-of-folds and its decision; a corrupted proof in one shard fails the verdict of that shard and the final pairing."""
    import numpy as np
    import torch

    from snark_verifier_axiom_b200.distributed import OFF_DECIDE_OK, OFF_FOLD_STATUS, OFF_OK, RECORD, LibsvkOps

    V, S, ctx, AS, pv = env
    dev = torch.device("cuda", 0)
    n, world, m = 12, 2, 2
    insts, proofs = forge.forge_batch(S, "bdfg21", n, seed0=1300)
    for corrupt in (False, True):
        pr = list(proofs)
        if corrupt:
            bad = bytearray(pr[8])
            bad[9 * 32 + 5] ^= 4  # an evaluation: the proof still reads, its accumulator is wrong
            pr[8] = bytes(bad)
        snarks = [V.Snark(i, p) for i, p in zip(insts, pr)]
        ops = LibsvkOps(pv[0])
        recs = []
        for r in range(world):
            mine = snarks[r * n // world : (r + 1) * n // world]
            inst, n_inst, pb, lens = pv[0].pack(mine)
            d_inst, d_pb = torch.from_numpy(inst).to(dev), torch.from_numpy(pb).to(dev)
            d_accs = torch.zeros(len(mine) * 128, dtype=torch.uint8, device=dev)
            d_st = torch.zeros(len(mine), dtype=torch.int32, device=dev)
            d_rec = torch.zeros(RECORD, dtype=torch.uint8, device=dev)
            ops.local_verify(d_inst, n_inst, d_pb, 1, len(mine), m, d_accs, d_st, d_rec, decide=False)
            ctx.sync()
            rec = d_rec.cpu().numpy()
            assert rec[OFF_DECIDE_OK] == 1 and (d_st.cpu().numpy() == 0).all()  # not decided here
            assert rec[OFF_OK] == 1
            recs.append(rec)
        gathered = torch.from_numpy(np.stack([r[:128] for r in recs]).reshape(-1).copy()).to(dev)
        d_final = torch.zeros(RECORD, dtype=torch.uint8, device=dev)
        ops.fold(1, world, gathered, d_final)
        ops.decide(1, d_final)
        ctx.sync()
        f = d_final.cpu().numpy()
        assert int(np.frombuffer(f[OFF_FOLD_STATUS : OFF_FOLD_STATUS + 4].tobytes(), np.int32)[0]) == 0
        # oracle: fold each shard (groups of m), then fold the two results flat, then decide
        pairs = []
        for i, p in zip(insts, pr):
            a = api.succinct_verify(S.dk.svk, S.protocol, i, p, "bdfg21")[0]
            pairs.append((a.lhs.pt, a.rhs.pt))
        shard_accs = [api.fold(pairs[r * n // world : (r + 1) * n // world], m)[0] for r in range(world)]
        for r in range(world):
            got = V.KzgAccumulator.from_bytes(recs[r][:128].tobytes())
            assert (got.lhs, got.rhs) == shard_accs[r]
        final = api.fold(shard_accs, 0)[0]
        got = V.KzgAccumulator.from_bytes(f[:128].tobytes())
        assert (got.lhs, got.rhs) == final
        assert bool(f[OFF_DECIDE_OK]) == api.decide(S.dk, final) == (not corrupt)


def test_kzg_as_zero_knowledge_accumulation_proof(env):
    """`KzgAs` with `KzgAsVerifyingKey(true)` (pcs/kzg/accumulation.rs:29-62, 113-136): the verifier reads two blind points from the
    accumulation proof, absorbs them after the instances and folds them in with r^n.  The blind pair is what a zk
    `create_proof` writes (:153-170): (s * G, G) for a random s -- itself a valid accumulator."""
    import random

    from oracle import bn254
    from oracle.kzg import KzgAccumulator as OAcc, KzgAsBdfg21
    from oracle.loader import NativeLoader
    from oracle.transcript import PoseidonTranscript, VerifyError

    V, S, ctx, AS, pv = env
    rng = random.Random(77)
    n = 5
    insts, proofs = forge.forge_batch(S, "bdfg21", n, seed0=3100)
    accs, _, st = pv[0].succinct_verify([V.Snark(i, p) for i, p in zip(insts, proofs)])
    assert (st == 0).all()
    k = rng.randrange(1, bn254.R)
    blind = (bn254.g1_mul(bn254.G1_GEN, k * S.s % bn254.R), bn254.g1_mul(bn254.G1_GEN, k))
    as_proof = bn254.g1_to_bytes(blind[0]) + bn254.g1_to_bytes(blind[1])
    loader = NativeLoader()
    o_accs = [OAcc(loader.ec_point_load_const(a.lhs), loader.ec_point_load_const(a.rhs)) for a in accs]
    tr = PoseidonTranscript(loader, as_proof)
    proof = KzgAsBdfg21.as_read_proof(True, o_accs, tr)
    want = KzgAsBdfg21.as_verify(True, o_accs, proof)
    got, r = AS.verify_zk(accs, as_proof)
    assert (got.lhs, got.rhs) == (want.lhs.pt, want.rhs.pt) and r == proof[1].v
    assert AS.decide_batch([got]) == [True] and api.decide(S.dk, (got.lhs, got.rhs))
    # error paths of `read_ec_point` (transcript/halo2.rs:235-260): short stream, bad encoding, identity
    bad_x = next(x for x in range(1, 100) if not bn254.g1_from_bytes(x.to_bytes(32, "little"))[0])
    for bad_proof, sub in ((as_proof[:40], 1), (bad_x.to_bytes(32, "little") + as_proof[32:], 3), (bytes(32) + as_proof[32:], 4)):
        with pytest.raises(V.Error) as e:
            AS.verify_zk(accs, bad_proof)
        assert e.value.status & 0xFF == 4 and (sub is None or e.value.status >> 8 == sub)
        with pytest.raises(VerifyError):
            KzgAsBdfg21.as_read_proof(True, o_accs, PoseidonTranscript(loader, bad_proof))
