#!/usr/bin/env python
"""BASELINE config 3 across GPUs: G1 MSM with points sharded over the ranks (weak scaling: 2^LOG_N points per rank).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_msm_multi.py --log-n 22
One JSON line on rank 0; device time (CUDA events), max over ranks."""
import argparse, ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from snark_verifier_axiom_b200 import verifier as V
from snark_verifier_axiom_b200.distributed import LibsvkMsmOps, msm_sharded


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=22)
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ctx = V.Context(local)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    n = 1 << a.log_n
    g = torch.Generator(device=dev)
    g.manual_seed(42 + rank)
    with torch.cuda.stream(stream):
        dl = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
        dl[:, 31] &= 0x1F
        sc = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
        sc[:, 31] &= 0x1F
        gen = torch.zeros(64, dtype=torch.uint8, device=dev)
        gen[0], gen[32] = 1, 2
        pts = torch.empty(n * 64, dtype=torch.uint8, device=dev)
    ctx._check(ctx._L.svk_g1_mul_batch_dev(ctx._c, n, ctypes.c_void_p(dl.data_ptr()), ctypes.c_void_p(gen.data_ptr()), 1, ctypes.c_void_p(pts.data_ptr())))
    stream.synchronize()
    ops = LibsvkMsmOps(ctx)
    out = msm_sharded(ops, world, dev, sc.view(-1), pts, n, stream)  # warm
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.iters):
        out = msm_sharded(ops, world, dev, sc.view(-1), pts, n, stream)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.iters], dtype=torch.float64, device=dev)
    res = out.clone()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ref = res.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref, res), "ranks disagree on the MSM result"
    if rank == 0:
        print(json.dumps({"config": "msm_g1_sharded", "n_gpus": world, "points_per_rank": n, "total_points": n * world, "ms": float(ms.item()),
                          "points_per_s": n * world / (float(ms.item()) * 1e-3), "scaling": "weak", "result_x_prefix": res[:8].cpu().numpy().tobytes().hex()}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


main()
