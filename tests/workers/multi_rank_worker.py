"""Worker of tests/test_gpu_multi_rank.py and of bench.py's N > 1 pre-flight: one process per GPU (torch.distributed.run).
Every rank verifies its shard through libsvk, the per-rank folded accumulators are all-gathered over NCCL, folded across
ranks and decided once (snark_verifier_axiom_b200/distributed.py).  Rank 0 recomputes the whole job with the C restatement
of the reference CPU path (oracle/c): per-rank fold trees, the fold of folds (pcs/kzg/accumulation.rs:29-62), the pairing
(decider.rs:60-68) and compares accumulators, root challenge and verdict -- once for an all-valid job and once with one
corrupted proof on the LAST rank.  Prints one JSON line on rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def check(per_rank=256, group_size=4, sv=None, pv=None, g=None, dev=None, seed0=900):
    """Runs on every rank (collective).  Returns a dict on rank 0, None elsewhere."""
    from oracle import forge
    from oracle.c import cref
    from snark_verifier_axiom_b200 import synth
    from snark_verifier_axiom_b200.distributed import RECORD

    world, rank = dist.get_world_size(), dist.get_rank()
    inst, proofs = synth.forge_shplonk_batch(pv, g["trapdoor_s"], g["vk_dlogs"], per_rank, seed=seed0 + rank)
    inst = np.ascontiguousarray(inst)
    out = {}
    for case in ("valid", "corrupted"):
        pf = proofs.copy()
        if case == "corrupted" and rank == world - 1:
            pf[per_rank // 2, 9 * 32 + 5] ^= 4  # an evaluation: still decodes, the global pairing must reject
        d_inst, d_pf = torch.from_numpy(inst).to(dev), torch.from_numpy(pf).to(dev)
        torch.cuda.synchronize()
        sv.verify_dev(d_inst, 1, d_pf, per_rank)
        verdict = sv.last_ok()
        final = np.frombuffer(sv.final_accumulator(), np.uint8)
        rec = sv.d_final[:RECORD].cpu().numpy()
        # all proofs of the job to rank 0 (NCCL all_gather of the byte tensors)
        all_pf = [torch.empty_like(d_pf) for _ in range(world)]
        all_in = [torch.empty_like(d_inst) for _ in range(world)]
        dist.all_gather(all_pf, d_pf)
        dist.all_gather(all_in, d_inst)
        if rank == 0:
            S = forge.Setup(0)
            tr = cref.Trace(S, "bdfg21")
            threads = os.cpu_count() or 1
            shard_accs = []
            for r in range(world):
                p_, i_ = all_pf[r].cpu().numpy(), all_in[r].cpu().numpy()
                accs, st = cref.replay_packed(tr, np.ascontiguousarray(p_), np.full(per_rank, p_.shape[1], np.int32),
                                              np.ascontiguousarray(i_).view(np.uint64).reshape(per_rank, -1), threads)
                assert (st == 0).all()
                a, _, fst = cref.fold(accs, group_size, threads=threads)
                assert fst == 0
                shard_accs.append(a)
            want, r_root, fst = cref.fold(np.stack(shard_accs), 0)
            want_ok = cref.decide(want, S.dk)
            got_r = int.from_bytes(rec[128:160].tobytes(), "little")
            out[case] = {"accumulator_equal": bool((final == want).all()), "root_challenge_equal": got_r == r_root, "verdict": bool(verdict),
                         "oracle_verdict": bool(want_ok)}
    return out if rank == 0 else None


def main():
    from snark_verifier_axiom_b200 import verifier as V
    from snark_verifier_axiom_b200.distributed import ShardedBatchVerifier
    from snark_verifier_axiom_b200.standard_plonk import load_golden

    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    g = load_golden()
    ctx = V.Context(local)
    stream = torch.cuda.Stream(device=dev)
    pv = V.PlonkVerifier(ctx, g["dk"], g["protocol"], V.SHPLONK)
    sv = ShardedBatchVerifier(pv, dist.get_world_size(), dist.get_rank(), dev, stream, group_size=4)
    res = check(256, 4, sv, pv, g, dev)
    if res is not None:
        print(json.dumps({"world": dist.get_world_size(), **res}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
