#!/usr/bin/env python
"""Two (or S) calls of B batches side by side, as bench.py --steps 20 runs them, with the per-kernel timeline of every stream
(svk_profile_timeline).  Usage: python tools/timeline_probe.py [--slots 2] [--batches 10]"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np  # noqa: E402
import torch  # noqa: E402

from snark_verifier_axiom_b200 import synth, verifier as V  # noqa: E402
from snark_verifier_axiom_b200.distributed import ShardedBatchVerifier  # noqa: E402
from snark_verifier_axiom_b200.standard_plonk import load_golden  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--slots", type=int, default=2)
    ap.add_argument("--batches", type=int, default=10)
    ap.add_argument("--group-size", type=int, default=4)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = load_golden()
    n = 4096 * a.batches
    slots = []
    for _ in range(a.slots):
        ctx = V.Context(0)
        st = torch.cuda.Stream(device=dev)
        ctx.set_stream(st.cuda_stream)
        pv = V.PlonkVerifier(ctx, g["dk"], g["protocol"], V.SHPLONK)
        slots.append((ctx, st, pv, ShardedBatchVerifier(pv, 1, 0, dev, st, group_size=a.group_size, max_batches=a.batches)))
    inst, proofs = synth.forge_shplonk_batch(slots[0][2], g["trapdoor_s"], g["vk_dlogs"], n, seed=3)
    d_inst, d_pf = torch.from_numpy(np.ascontiguousarray(inst)).to(dev), torch.from_numpy(proofs).to(dev)
    torch.cuda.synchronize()
    for rep in range(2):  # warm-up, then the recorded pass
        if rep == 1:
            for ctx, *_ in slots:
                ctx._L.svk_profile_enable(ctx._c, 1)
        for ctx, st, pv, sv in slots:
            sv.verify_dev(d_inst, 1, d_pf, n, n_batches=a.batches)
        torch.cuda.synchronize()
    rows = []
    for si, (ctx, *_rest) in enumerate(slots):
        buf = ctypes.create_string_buffer(1 << 20)
        ctx._L.svk_profile_timeline(ctx._c, buf, len(buf))
        rows += [(t0, t1, si, name) for name, t0, t1 in json.loads(buf.value.decode())]
    rows.sort()
    end = max(r[1] for r in rows)
    print(f"# {a.slots} calls x {a.batches} batches: {end:.2f} ms = {a.slots * n / end / 1e3:.3f} M proofs/s")
    for t0, t1, si, name in rows:
        print(f"{t0:8.3f} {t1:8.3f} {t1 - t0:7.3f}  s{si} {name}")


if __name__ == "__main__":
    main()
