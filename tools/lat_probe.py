#!/usr/bin/env python
"""One 4096-proof batch at a time (nothing else in flight): the latency kernels under ncu / CUDA events.
Usage: python tools/lat_probe.py [--iters 3] [--batch 4096] [--group-size 8]"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from snark_verifier_axiom_b200 import synth, verifier as V  # noqa: E402
from snark_verifier_axiom_b200.distributed import ShardedBatchVerifier  # noqa: E402
from snark_verifier_axiom_b200.standard_plonk import load_golden  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--group-size", type=int, default=8)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = load_golden()
    ctx = V.Context(0)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    pv = V.PlonkVerifier(ctx, g["dk"], g["protocol"], V.SHPLONK)
    sv = ShardedBatchVerifier(pv, 1, 0, dev, stream, group_size=args.group_size, max_batches=1)
    inst, proofs = synth.forge_shplonk_batch(pv, g["trapdoor_s"], g["vk_dlogs"], args.batch, seed=7)
    d_inst = torch.from_numpy(np.ascontiguousarray(inst)).to(dev)
    d_proofs = torch.from_numpy(proofs).to(dev)
    torch.cuda.synchronize()
    sv.verify_dev(d_inst, 1, d_proofs, args.batch)
    if not sv.last_ok():
        st = sv.d_status[: args.batch].cpu().numpy()
        rec = sv.d_records[:256].cpu().numpy()
        raise SystemExit(f"batch rejected: {int((st != 0).sum())} proof statuses != 0 (first {st[st != 0][:4]}), fold_status {rec[160:164]}, decide_ok {rec[164]}, ok {rec[165]}")
    L, c = ctx._L, ctx._c
    L.svk_profile_enable(c, 1)
    lat = []
    for _ in range(args.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sv.verify_dev(d_inst, 1, d_proofs, args.batch)
        e1.record(stream)
        stream.synchronize()
        lat.append(e0.elapsed_time(e1))
    buf = ctypes.create_string_buffer(1 << 16)
    L.svk_profile_report(c, buf, len(buf))
    prof = {k: round(v["ms"] / args.iters, 4) for k, v in json.loads(buf.value.decode()).items()}
    assert sv.last_ok()
    print(json.dumps({"batch": args.batch, "group_size": args.group_size, "latency_ms_with_event_profiling": min(lat), "kernels_ms": prof}))


if __name__ == "__main__":
    main()
