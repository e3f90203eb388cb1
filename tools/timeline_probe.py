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
    ap.add_argument("--plan", default="", help="batches per call, comma separated (overrides --slots/--batches), e.g. 9,9,1,1")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-timeline", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = load_golden()
    plan = [int(x) for x in a.plan.split(",")] if a.plan else [a.batches] * a.slots
    a.batches = max(plan)
    n = 4096 * a.batches
    slots = []
    for _ in plan:
        ctx = V.Context(0)
        st = torch.cuda.Stream(device=dev)
        ctx.set_stream(st.cuda_stream)
        pv = V.PlonkVerifier(ctx, g["dk"], g["protocol"], V.SHPLONK)
        slots.append((ctx, st, pv, ShardedBatchVerifier(pv, 1, 0, dev, st, group_size=a.group_size, max_batches=a.batches)))
    inst, proofs = synth.forge_shplonk_batch(slots[0][2], g["trapdoor_s"], g["vk_dlogs"], n, seed=3)
    d_inst, d_pf = torch.from_numpy(np.ascontiguousarray(inst)).to(dev), torch.from_numpy(proofs).to(dev)
    torch.cuda.synchronize()

    def go():
        for (ctx, st, pv, sv), nb in zip(slots, plan):
            sv.verify_dev(d_inst[: nb * 4096], 1, d_pf[: nb * 4096], nb * 4096, n_batches=nb)

    go()
    torch.cuda.synchronize()
    times = []
    for rep in range(a.reps):  # device time from a common start event to the last stream's end
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream(dev))
        for _, st, *_r in slots:
            st.wait_event(e0)
        go()
        ends = []
        for _, st, *_r in slots:
            e = torch.cuda.Event(enable_timing=True)
            e.record(st)
            ends.append(e)
        torch.cuda.synchronize()
        times.append(max(e0.elapsed_time(e) for e in ends))
        assert all(sv.last_ok() for *_r, sv in slots)
    tot = sum(plan) * 4096
    print(f"# plan {plan}: min {min(times):.2f} ms median {sorted(times)[len(times) // 2]:.2f} ms = {tot / min(times) / 1e3:.3f} M proofs/s (best)")
    if a.no_timeline:
        return
    for ctx, *_ in slots:
        ctx._L.svk_profile_enable(ctx._c, 1)
    go()
    torch.cuda.synchronize()
    rows = []
    for si, (ctx, *_rest) in enumerate(slots):
        buf = ctypes.create_string_buffer(1 << 20)
        ctx._L.svk_profile_timeline(ctx._c, buf, len(buf))
        rows += [(t0, t1, si, name) for name, t0, t1 in json.loads(buf.value.decode())]
    rows.sort()
    end = max(r[1] for r in rows)
    print(f"# recorded pass (events around every kernel): {end:.2f} ms")
    for t0, t1, si, name in rows:
        print(f"{t0:8.3f} {t1:8.3f} {t1 - t0:7.3f}  s{si} {name}")


if __name__ == "__main__":
    main()
