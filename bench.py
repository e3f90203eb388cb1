#!/usr/bin/env python
"""bench.py -- BASELINE.json metric "BN254 KZG proofs verified/sec" on the config-2 workload:
a batch of 4096 StandardPlonk k=8 SHPLONK proofs natively verified with accumulator folding + one
pairing (per GPU; weak scaling for --gpus N: every rank verifies its own 4096-proof shard, the
per-rank folded accumulators are all-gathered over NCCL, folded once more and decided).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 4096] [--group-size 8]
  python bench.py --impl reference ...      # the reference's CPU algorithm (oracle) on the host cores

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "bn254_kzg_proofs_verified_per_sec"
UNIT = "proofs/s"
WORKLOAD = "standard_plonk_k8_shplonk_poseidon: succinct verify each + KzgAs fold + one pairing"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4096, help="proofs per GPU per step")
    ap.add_argument("--group-size", type=int, default=8, help="KzgAs fold group size (0 = the reference's flat fold)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="proofs in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ workload
def make_workload(batch):
    """`batch` proofs from the committed fixture (64 distinct trapdoor-forged proofs per scheme, generated
    by tests/golden/make_golden.py), tiled.  Verification time does not depend on the proof bytes."""
    import numpy as np

    from snark_verifier_axiom_b200.standard_plonk import load_golden

    g = load_golden()
    snarks = g["schemes"]["bdfg21"]["snarks"]
    reps = [snarks[i % len(snarks)] for i in range(batch)]
    return g, reps, np


def cpu_baseline(g, sample, group_size):
    """The reference's algorithm on the host: oracle/c (`kind: port`, C restatement: naive per-pair scalar
    multiplication, per-element Fermat inversion, serial sponge, one pairing) when built, else the Python
    oracle.  Bounded sample of the same workload; returns the cpu_baseline object."""
    try:
        from oracle.c import cref

        return cref.bench_baseline(g, sample, group_size)
    except Exception as e:  # C oracle not built: Python oracle, tiny sample
        note = f"python oracle (C oracle unavailable: {type(e).__name__})"
    from oracle import api, forge

    S = forge.Setup(0)
    n = min(sample or 4, 4)
    snarks = g["schemes"]["bdfg21"]["snarks"][:n]
    t0 = time.perf_counter()
    pairs = []
    for s in snarks:
        a = api.succinct_verify(S.dk.svk, S.protocol, s.instances, s.proof, "bdfg21")[0]
        pairs.append((a.lhs.pt, a.rhs.pt))
    acc, _ = api.fold(pairs, group_size)
    ok = api.decide(S.dk, acc)
    dt = time.perf_counter() - t0
    assert ok
    return {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"{n} proofs: succinct verify + fold + one pairing, {note}"}


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    g, reps, np = make_workload(64)
    vals = []
    base = None
    for _ in range(args.warmup + args.steps):
        base = cpu_baseline(g, args.cpu_sample, args.group_size)
        vals.append(base["value"])
    vals = vals[args.warmup:] or vals
    v = sum(vals) / len(vals)
    base["value"] = v
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * args.batch / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32x8 (254-bit Fq/Fr)",
        "data": "synthetic: trapdoor-forged StandardPlonk k=8 SHPLONK proofs (tests/golden), CPU sample",
        "config": {"workload": WORKLOAD, "batch_per_gpu": args.batch, "fold_group_size": args.group_size},
        "cpu_baseline": base, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    from snark_verifier_axiom_b200 import verifier as V
    from snark_verifier_axiom_b200.distributed import ShardedBatchVerifier

    g, reps, np = make_workload(args.batch)
    ctx = V.Context(local)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    pv = V.PlonkVerifier(ctx, g["dk"], g["protocol"], V.SHPLONK)
    sv = ShardedBatchVerifier(pv, world, rank, dev, stream, group_size=args.group_size)
    inst, n_inst, proofs, lens = pv.pack(reps)
    n = args.batch
    h_inst = torch.from_numpy(inst).pin_memory()
    h_proofs = torch.from_numpy(proofs).pin_memory()
    with torch.cuda.stream(stream):
        d_inst = h_inst.to(dev, non_blocking=True)
        d_proofs = h_proofs.to(dev, non_blocking=True)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev():
        return sv.verify_dev(d_inst, n_inst, d_proofs, n)

    # ---- warm-up
    for _ in range(args.warmup):
        with torch.cuda.stream(stream):
            flush.fill_(1)
        step_dev()
    stream.synchronize()
    assert sv.last_ok(), "warm-up batch did not verify"

    # ---- timed: K steps, device time per step by CUDA events on the launching stream, L2 flushed between steps
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    l0 = ctx.launch_count
    evs = []
    for _ in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(1)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        step_dev()
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count - l0
    ms_total = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    assert sv.last_ok(), "timed batch did not verify"
    value = world * n / (ms_per_step * 1e-3)

    # ---- e2e: the public host-buffer call (H2D of proofs + instances, D2H of statuses + verdict inside)
    e2e_steps = max(3, min(args.steps, 10))
    h_i, h_p, h_l = h_inst.numpy(), h_proofs.numpy(), lens
    sv.verify_host(h_i, n_inst, h_p, h_l, n)  # warm
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ok, status = sv.verify_host(h_i, n_inst, h_p, h_l, n)
    barrier()
    dt = time.perf_counter() - t0
    assert ok and (status == 0).all()
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * n * e2e_steps / float(t.item())
    h2d = int(h_i.nbytes + h_p.nbytes + h_l.nbytes)
    d2h = int(n * 4 + 256)

    if rank == 0:
        # ---- per-kernel device time (CUDA events around every launch) -> dominant kernel -> roofline
        L, c = ctx._L, ctx._c
        L.svk_profile_enable(c, 1)
        prof_steps = 3
        for _ in range(prof_steps):
            with torch.cuda.stream(stream):
                flush.fill_(1)
            step_dev()
        buf = ctypes.create_string_buffer(1 << 16)
        L.svk_profile_report(c, buf, len(buf))
        L.svk_profile_enable(c, 0)
        prof = json.loads(buf.value.decode())
        peak, peak_ms = ctx.modmul_peak(4000)
        info = pv.info
        # algorithmic Fq/Fr multiplications per launch (DESIGN.md "work model")
        work = {
            "k_tape": n * info["n_fr_mul"],
            "k_decompress": n * info["n_points"] * 372,
            "k_proof_msm": n * ((info["n_lhs_terms"] - 1) * 3020 + 16 * 4 + 400),
        }
        top = max(prof.items(), key=lambda kv: kv[1]["ms"])
        kernels = {k: {"launches": v["count"] // prof_steps if v["count"] >= prof_steps else v["count"], "ms_per_step": v["ms"] / prof_steps} for k, v in prof.items()}
        name = top[0]
        ms_launch = top[1]["ms"] / top[1]["count"]
        ach = work.get(name, 0) / (ms_launch * 1e-3) if name in work else None
        roofline = {
            "bound": "imad", "kernel": name, "achieved": (ach / 1e9) if ach else None, "peak": peak / 1e9, "unit": "Gmodmul/s (1 modmul = 8x32-bit-limb Montgomery mul = 139 IMAD)",
            "frac": (ach / peak) if ach else None, "traffic": None, "peak_source": "svk_bench_modmul_peak measured in this run (independent Montgomery-mul chains on all SMs)",
            "kernel_ms_per_launch": ms_launch, "kernels": kernels,
            "hbm_gbs_algorithmic": (h2d + d2h) / (ms_per_step * 1e-3) / 1e9,
        }
        base = None if args.no_cpu_baseline else cpu_baseline(g, args.cpu_sample, args.group_size)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32x8 (254-bit Fq/Fr Montgomery, integer pipe)",
            "data": "synthetic: 64 distinct trapdoor-forged StandardPlonk k=8 SHPLONK proofs (tests/golden, oracle-generated) tiled to the batch",
            "config": {"workload": WORKLOAD, "batch_per_gpu": n, "global_batch": world * n, "proof_bytes": info["proof_len"], "fold_group_size": args.group_size,
                       "fold": "flat (reference aggregation.rs:235-245)" if args.group_size in (0, 1) else f"tree, groups of {args.group_size}",
                       "l2": "flushed between steps (256 MiB fill)", "parallelism": f"proof-sharded x{world}, NCCL all_gather of {world} folded accumulators" if world > 1 else "single GPU"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "roofline": roofline, "cpu_baseline": base,
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
