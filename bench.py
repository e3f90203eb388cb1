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
# One hardware work queue per in-flight batch: with the default of 8 connections, streams alias onto 8 queues
# and the long serial kernels of a batch (fold sponge, the single pairing) block unrelated batches behind them
# (measured: 310 k -> 465 k proofs/s at 32 in flight, profiles/r1_notes.md).  Must be set before CUDA initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

METRIC = "bn254_kzg_proofs_verified_per_sec"
UNIT = "proofs/s"
WORKLOAD = "standard_plonk_k8_{scheme}_poseidon: succinct verify each + KzgAs fold + one pairing"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1024, help="timed steps; one step = one 4096-proof batch (succinct verify each, fold, one pairing)")
    ap.add_argument("--inflight", type=int, default=16, help="launches in flight (one libsvk context + stream each)")
    ap.add_argument("--batches-per-launch", type=int, default=0, help="4096-proof batches verified by one call (each folded + decided on its own); "
                    "0 = auto: the largest divisor B <= 32 of --steps that leaves at least two launches (choose_batches_per_launch)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4096, help="proofs per GPU per step")
    ap.add_argument("--group-size", type=int, default=4, help="KzgAs fold group size (0 = the reference's flat fold); 4 minimises the serial "
                    "permutations of the tree over 4096 accumulators (6 levels x 9 = 54; groups of 8: 4 x 17 = 68)")
    ap.add_argument("--scheme", default="bdfg21", choices=["bdfg21", "gwc19"], help="multi-open scheme of the proofs (BASELINE config 2: SHPLONK; config 4: GWC)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="proofs in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tiled", action="store_true", help="tile the 64 fixture proofs instead of forging distinct ones")
    ap.add_argument("--no-preflight", action="store_true", help="N > 1: skip the parity pre-flight (256 proofs per rank, one corrupted on the last rank; "
                    "cross-rank accumulator, root challenge and verdict against the oracle's fold of folds -- the checker, never the timed path)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary configs (MSM 2^20 / 2^24, decide 2^16, single-proof latency)")
    ap.add_argument("--stream-priorities", type=int, default=0, help="1: slot i gets a higher stream priority than slot i + 1, so that concurrent launches finish "
                    "one after the other and the narrow tail of one (fold levels, pairing) runs beside the wide kernels of the next")
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
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        """Samples taken inside [t0, t1] (host clock around the timed region); the sampler itself starts earlier, because
        nvidia-smi needs longer to start than a short timed region lasts.  With no sample inside the window (a region shorter than
        the polling period) the two samples around it are used."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = list(self.lines)
        if t0 is not None:
            inside = [ln for t, ln in rows if t0 <= t <= t1 + 0.05]
            if not inside:
                before, after = [ln for t, ln in rows if t < t0][-1:], [ln for t, ln in rows if t > t1][:1]
                inside = before + after
            rows = inside
        else:
            rows = [ln for _, ln in rows]
        for ln in rows:
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
def make_workload(batch, scheme="bdfg21"):
    """`batch` proofs from the committed fixture (64 distinct trapdoor-forged proofs per scheme, generated
    by tests/golden/make_golden.py), tiled.  Verification time does not depend on the proof bytes."""
    import numpy as np

    from snark_verifier_axiom_b200.standard_plonk import load_golden

    g = load_golden()
    snarks = g["schemes"][scheme]["snarks"]
    reps = [snarks[i % len(snarks)] for i in range(batch)]
    return g, reps, np


def cpu_baseline(g, sample, group_size, scheme="bdfg21", optimised=True):
    """The reference's algorithm on the host: oracle/c (`kind: port`, C restatement: naive per-pair scalar
    multiplication, per-element Fermat inversion, serial sponge, one pairing) when built, else the Python
    oracle.  Bounded sample of the same workload; returns the cpu_baseline object."""
    try:
        from oracle.c import cref

        return cref.bench_baseline(g, sample, group_size, scheme, optimised=optimised)
    except Exception as e:  # C oracle not built: Python oracle, tiny sample
        note = f"python oracle (C oracle unavailable: {type(e).__name__})"
    from oracle import api, forge

    S = forge.Setup(0)
    n = min(sample or 4, 4)
    snarks = g["schemes"][scheme]["snarks"][:n]
    t0 = time.perf_counter()
    pairs = []
    for s in snarks:
        a = api.succinct_verify(S.dk.svk, S.protocol, s.instances, s.proof, scheme)[0]
        pairs.append((a.lhs.pt, a.rhs.pt))
    acc, _ = api.fold(pairs, group_size)
    ok = api.decide(S.dk, acc)
    dt = time.perf_counter() - t0
    assert ok
    return {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"{n} proofs: succinct verify + fold + one pairing, {note}"}


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's CPU algorithm (oracle/c, all host threads) on the same workload.  A step of this arm is a BOUNDED SAMPLE of
    the 4096-proof batch (`config.step_sample_proofs` proofs: succinct verify each, fold, one pairing), so that K + W steps end
    within minutes; `ms_per_step` is the measured wall time of such a step, `value` = proofs per second over the timed steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    g, reps, np = make_workload(64, args.scheme)
    sample = args.cpu_sample or 512
    times, base = [], None
    t_begin = time.time()
    for it in range(args.warmup + args.steps):
        base = cpu_baseline(g, sample, args.group_size, args.scheme, optimised=False)
        times.append(base.get("sample_seconds") or sample / base["value"])  # the timed region of the step: replay + fold + pairing
        # keep the whole run within a few minutes whatever K is
        if it >= args.warmup and time.time() - t_begin > 150:
            break
    timed = times[min(args.warmup, len(times) - 1):]
    step_s = sum(timed) / len(timed)
    v = sample / step_s
    base["value"] = v
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(timed), "steps_requested": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * step_s, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32x8 (254-bit Fq/Fr)",
        "data": "synthetic: trapdoor-forged StandardPlonk k=8 SHPLONK proofs (tests/golden), CPU sample",
        "config": {"workload": WORKLOAD.format(scheme="shplonk" if args.scheme == "bdfg21" else "gwc"), "batch_per_gpu": args.batch, "fold_group_size": args.group_size,
                   "step_sample_proofs": sample, "note": "a step of this arm is a bounded sample of the 4096-proof batch; proofs/s does not depend on the sample size"},
        "cpu_baseline": base, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------ our arm
class Slot:
    """One in-flight batch: its own libsvk context (stream + scratch) and output buffers.  Several slots
    keep several independent 4096-proof batches in flight, which is how a throughput device hides the
    latency-bound tail of a batch (serial fold sponge, the single pairing)."""

    def __init__(self, torch, V, ShardedBatchVerifier, g, local, dev, world, rank, group_size, max_batches, mos, priority=0):
        self.ctx = V.Context(local)
        self.stream = torch.cuda.Stream(device=dev, priority=priority)
        self.ctx.set_stream(self.stream.cuda_stream)
        self.pv = V.PlonkVerifier(self.ctx, g["dk"], g["protocol"], mos)
        self.sv = ShardedBatchVerifier(self.pv, world, rank, dev, self.stream, group_size=group_size, max_batches=max_batches)


def choose_batches_per_launch(steps, slots, requested=0, cap=32):
    """Batches carried by one call.  Auto (requested <= 0): the LARGEST divisor B <= cap of `steps` that still leaves two launches,
    so exactly `steps` batches are timed, the wide kernels of a launch fill the machine, and the narrow tail of one launch (fold
    levels, the pairing) runs beside the wide kernels of another.  Measured on B200 at --steps 20 (profiles/r2_notes.md):
    2 x 10 batches 1.30 M proofs/s, 1 x 20 1.25 M, 4 x 5 1.16 M, 20 x 1 0.70 M; at --steps 1024: 32 x 32.  A prime `steps` <= cap
    goes out as ONE launch rather than `steps` single-batch launches."""
    if requested > 0:
        return requested
    divs = [b for b in range(1, min(cap, steps) + 1) if steps % b == 0]
    two = [b for b in divs if steps // b >= 2]
    b = max(two) if two else max(divs)
    if b == 1 and steps <= cap:
        b = steps
    return b


def run_secondary(torch, dist, V, ctx, stream, dev, g, peak, world, rank, args):
    """The other BASELINE configs, bounded (< 60 s): config 3 G1 MSM at 2^20 and 2^24 points per GPU (uniform scalars; powers of r at
    2^20; sharded over the ranks with an all-gather of the partial sums when N > 1), config 5 batched decide of 2^16 accumulators,
    config 1 single-proof latency.  Each entry carries its own roofline (canonical work: 176 M per MSM point = 16 windows x 11 M,
    20 k M per decide) and a CPU figure from the C restatement of the reference (oracle/c), timed on a bounded sample.
    Collective when N > 1 (every rank calls it); returns the dict on rank 0."""
    import numpy as np

    from snark_verifier_axiom_b200.distributed import LibsvkMsmOps, msm_sharded

    L, c = ctx._L, ctx._c
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    R_ = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    out = {}

    def rand_scalars(n, seed):
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        t = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=gen)
        t[:, 31] &= 0x1F  # < 2^253 < r
        return t.contiguous()

    def timed(fn, iters):
        fn()
        stream.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(iters):
            fn()
        e1.record(stream)
        stream.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    try:
        from oracle.c import cref
    except Exception:
        cref = None
    cores = os.cpu_count() or 1

    # ---- config 3
    nmax = 1 << 24
    with torch.cuda.stream(stream):
        dl = rand_scalars(nmax, 42 + rank)
        gen1 = torch.zeros(64, dtype=torch.uint8, device=dev)
        gen1[0], gen1[32] = 1, 2
        pts = torch.empty(nmax * 64, dtype=torch.uint8, device=dev)
    ctx._check(L.svk_g1_mul_batch_dev(c, nmax, p(dl), p(gen1), 1, p(pts)))
    stream.synchronize()
    ops = LibsvkMsmOps(ctx)
    msm_cpu = None
    if rank == 0 and cref is not None:
        ns = 1 << 16
        h_sc, h_pt = rand_scalars(ns, 7).cpu().numpy(), pts[: ns * 64].cpu().numpy().reshape(ns, 64)
        t0 = time.perf_counter()
        cref.msm(h_sc, h_pt, cores)
        msm_cpu = {"value": ns / (time.perf_counter() - t0), "unit": "points/s", "cores": cores, "kind": "port",
                   "sample": f"2^16 points: util::msm::multi_scalar_multiplication (util/msm.rs:238-317: window ln n + 2, chunked over {cores} threads), C restatement (oracle/c cref_msm)"}
    msm = []
    for lg, kind in ((20, "uniform"), (20, "powers_of_r"), (24, "uniform")):
        n = 1 << lg
        if kind == "uniform":
            with torch.cuda.stream(stream):
                sc = rand_scalars(n, 1000 + lg + rank)
        else:  # the fold's distribution: 1, r, r^2, ...  (accumulation.rs:51-59)
            r, cur, rows = 0x1F2E3D4C5B6A79880123456789ABCDEF0FEDCBA9876543211122334455667788 % R_, 1, bytearray()
            for _ in range(n):
                rows += cur.to_bytes(32, "little")
                cur = cur * r % R_
            sc = torch.frombuffer(rows, dtype=torch.uint8).to(dev).view(n, 32)
        ms = timed(lambda: msm_sharded(ops, world, dev, sc.view(-1), pts, n, stream), 3 if lg <= 22 else 2)
        pps = n * world / (ms * 1e-3)
        msm.append({"log_n_per_gpu": lg, "scalars": kind, "n_gpus": world, "ms": ms, "value": pps, "unit": "points/s",
                    "roofline": {"bound": "imad", "achieved": n * 176 / (ms * 1e-3) / 1e9, "peak": peak / 1e9, "unit": "Gmodmul/s canonical (176 M per point)",
                                 "frac": n * 176 / (ms * 1e-3) / peak, "hbm_gbs_algorithmic": n * 96 / (ms * 1e-3) / 1e9},
                    "cpu_baseline": msm_cpu})
        del sc
    del pts, dl
    out["config3_msm_g1"] = msm

    # ---- config 4: GWC19 proofs sharded over the ranks (BASELINE: 2^14 proofs over 8 GPUs = 2048 per rank), accumulators
    # all-gathered over NCCL, folded and decided once; the 64 committed GWC fixture proofs tiled (verification time does not
    # depend on the proof bytes).  Runs at every N; collective.
    try:
        from snark_verifier_axiom_b200.distributed import ShardedBatchVerifier

        per_rank = 2048
        pv4 = V.PlonkVerifier(ctx, g["dk"], g["protocol"], V.GWC)
        sn4 = g["schemes"]["gwc19"]["snarks"]
        inst4, n_inst4, pf4, _ = pv4.pack([sn4[i % len(sn4)] for i in range(per_rank)])
        d_i4, d_p4 = torch.from_numpy(inst4).to(dev), torch.from_numpy(pf4).to(dev)
        sv4 = ShardedBatchVerifier(pv4, world, rank, dev, stream, group_size=args.group_size, max_batches=1)
        torch.cuda.synchronize()
        ms = timed(lambda: sv4.verify_dev(d_i4, n_inst4, d_p4, per_rank), 4)
        assert sv4.last_ok(), "config 4 batch did not verify"
        out["config4_gwc_sharded"] = {"n_gpus": world, "proofs_per_rank": per_rank, "total_proofs": per_rank * world, "ms": ms,
                                      "value": per_rank * world / (ms * 1e-3), "unit": "proofs/s",
                                      "note": "one call per rank, nothing else in flight (latency schedules): succinct verify + fold per rank, "
                                              "ncclAllGather of the accumulators, cross-rank fold, one pairing"}
        del sv4, d_i4, d_p4
    except Exception as e:  # the headline must not depend on a secondary config
        out["config4_gwc_sharded"] = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        # ---- config 5: 2^16 accumulators, 1/64 corrupted (valid ones: the oracle-checked golden accumulators, tiled)
        kid = ctx.load_deciding_key(g["dk"])
        pv = V.PlonkVerifier(ctx, g["dk"], g["protocol"], V.SHPLONK)
        accs, _, stt = pv.succinct_verify(g["schemes"]["bdfg21"]["snarks"])
        assert (stt == 0).all()
        n = 1 << 16
        base = np.frombuffer(b"".join(a.to_bytes() for a in accs), dtype=np.uint8).reshape(len(accs), 128)
        host = np.tile(base, (n // len(accs) + 1, 1))[:n].copy()
        bad = np.arange(0, n, 64)
        host[bad, 64:128] = host[(bad + 1) % n, 64:128]
        expect = np.ones(n, dtype=np.uint8)
        expect[bad] = 0
        d_accs, d_ok = torch.from_numpy(host).to(dev), torch.zeros(n, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        world1 = world
        world = 1  # rank-0 only from here on: no collectives inside `timed`
        ms = timed(lambda: ctx._check(L.svk_kzg_decide_batch_dev(c, kid, n, p(d_accs), p(d_ok))), 2)
        assert (d_ok.cpu().numpy() == expect).all(), "decide mismatch"
        dec_cpu = None
        if cref is not None:
            S_dk = __import__("oracle.forge", fromlist=["Setup"]).Setup(0).dk
            t0 = time.perf_counter()
            for i in range(32):
                assert cref.decide(host[i], S_dk) == bool(expect[i])
            dec_cpu = {"value": 32 / (time.perf_counter() - t0), "unit": "decides/s", "cores": 1, "kind": "port",
                       "sample": "32 accumulators, one thread: 2-pair Miller loop + final exponentiation (decider.rs:60-68), C restatement (oracle/c)"}
        out["config5_kzg_decide"] = {"n": n, "corrupted": int(len(bad)), "ms": ms, "value": n / (ms * 1e-3), "unit": "decides/s",
                                     "roofline": {"bound": "imad", "achieved": n * 20000 / (ms * 1e-3) / 1e9, "peak": peak / 1e9,
                                                  "unit": "Gmodmul/s canonical (20 k M per decide; 16.4 k executed)", "frac": n * 20000 / (ms * 1e-3) / peak,
                                                  "frac_executed": n * 16400 / (ms * 1e-3) / peak},
                                     "cpu_baseline": dec_cpu}
        # ---- config 1: one SHPLONK proof through the host call (succinct verify + fold of one + the pairing)
        sn = g["schemes"]["bdfg21"]["snarks"][0]
        pv.verify_one(sn)
        t0 = time.perf_counter()
        for _ in range(5):
            pv.verify_one(sn)
        lat = 1e3 * (time.perf_counter() - t0) / 5
        out["config1_single_proof"] = {"gpu_ms_host_call": lat, "value": 1e3 / lat, "unit": "proofs/s (one at a time)",
                                       "note": "latency of PlonkVerifier::verify for ONE proof (27 serial Poseidon permutations, one Straus MSM, one pairing): the "
                                               "B200 path is a throughput device; the one-thread CPU figure is cpu_baseline.single_thread_value"}
        world = world1
    if world > 1:
        dist.barrier()
    return out if rank == 0 else None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from concurrent.futures import ThreadPoolExecutor

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: native libraries (NCCL's version banner on the first communicator) write to fd 1
    # directly, so fd 1 is pointed at stderr for the run and the line goes to a duplicate of the original stdout
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world > 1:
        # NCCL writes its debug output (the version banner under NCCL_DEBUG=VERSION, NVLS / ring info under INFO) to stdout by
        # default; stdout is reserved for the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    from snark_verifier_axiom_b200 import verifier as V
    from snark_verifier_axiom_b200.distributed import ShardedBatchVerifier

    # K timed steps = K batches.  A launch carries B batches; no more slots (contexts, streams) than launches.
    B = choose_batches_per_launch(args.steps, max(1, args.inflight), args.batches_per_launch)
    S = max(1, min(args.inflight, -(-args.steps // B)))
    g, reps, np = make_workload(args.batch * B, args.scheme)
    mos = V.SHPLONK if args.scheme == "bdfg21" else V.GWC
    nb1 = args.batch           # proofs per batch (= per step)
    n = args.batch * B         # proofs per launch
    steps = -(-args.steps // B) * B  # whole launches
    prio = [max(-5, -(S - 1 - i)) if args.stream_priorities else 0 for i in range(S)]
    slots = [Slot(torch, V, ShardedBatchVerifier, g, local, dev, world, rank, args.group_size, B, mos, prio[i]) for i in range(S)]
    pv = slots[0].pv
    data = "synthetic: 64 distinct trapdoor-forged StandardPlonk k=8 proofs (tests/golden, oracle-generated) tiled to the batch"
    if args.scheme == "bdfg21" and not args.tiled:
        # n DISTINCT valid proofs per launch, forged with the test SRS trapdoor on the GPU (snark_verifier_axiom_b200/synth.py)
        from snark_verifier_axiom_b200 import synth

        inst, proofs = synth.forge_shplonk_batch(pv, g["trapdoor_s"], g["vk_dlogs"], n, seed=1 + rank)
        inst = np.ascontiguousarray(inst)
        n_inst, lens = 1, np.full(n, proofs.shape[1], dtype=np.uint32)
        data = f"synthetic: {n} distinct valid StandardPlonk k=8 SHPLONK proofs per launch, forged at start-up with the test-SRS trapdoor (synth.py; oracle-validated in tests/test_gpu_synth.py)"
    else:
        inst, n_inst, proofs, lens = pv.pack(reps)
    h_inst = torch.from_numpy(inst).pin_memory()
    h_proofs = torch.from_numpy(proofs).pin_memory()
    # distinct device copies of the inputs, rotated per step, together larger than the 126 MB L2
    n_copies = max(2, -(-(160 << 20) // int(h_proofs.numel() + h_inst.numel())))
    d_inputs = []
    for k in range(n_copies):
        d_inputs.append((h_inst.to(dev), h_proofs.to(dev)))
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def launch(k, nb=B):
        sl = slots[k % S]
        di, dp = d_inputs[k % n_copies]
        if nb == B:
            sl.sv.verify_dev(di, n_inst, dp, n, n_batches=B)
        else:
            sl.sv.verify_dev(di.view(n, -1)[: nb * nb1], n_inst, dp[: nb * nb1], nb * nb1, n_batches=nb)

    # ---- N > 1 pre-flight: the sharded job against the oracle's fold of folds (tests/workers/multi_rank_worker.py), untimed
    preflight = None
    if world > 1 and not args.no_preflight and args.scheme == "bdfg21":
        import importlib.util

        spec = importlib.util.spec_from_file_location("multi_rank_worker", os.path.join(ROOT, "tests", "workers", "multi_rank_worker.py"))
        mrw = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mrw)
        preflight = mrw.check(256, args.group_size, slots[0].sv, pv, g, dev)
        if rank == 0:
            assert preflight["valid"] == {"accumulator_equal": True, "root_challenge_equal": True, "verdict": True, "oracle_verdict": True}, preflight
            assert preflight["corrupted"]["accumulator_equal"] and not preflight["corrupted"]["verdict"] and not preflight["corrupted"]["oracle_verdict"], preflight
        barrier()

    # ---- warm-up (every slot).  The clock sampler starts here: nvidia-smi takes longer to start than a short timed region lasts.
    sampler = ClockSampler(local)
    sampler.start()
    for k in range(max(args.warmup, 1) * S):
        launch(k)
    torch.cuda.synchronize()
    for sl in slots:
        assert sl.sv.last_ok(), "warm-up batch did not verify"

    # ---- latency of ONE 4096-proof batch and of one launch of B batches (one slot, nothing else in flight)
    def one_latency(nb):
        lat = []
        for k in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(slots[0].stream)
            launch(0, nb)
            e1.record(slots[0].stream)
            slots[0].stream.synchronize()
            lat.append(e0.elapsed_time(e1))
        return min(lat)

    latency_ms = one_latency(1)
    launch_latency_ms = one_latency(B)

    # ---- timed: exactly K steps (batches), round-robin over the in-flight slots; device time by CUDA
    # events on the launching streams: from a common start event to the last slot's end event
    barrier()
    t_wall0 = time.time()
    l0 = sum(sl.ctx.launch_count for sl in slots)
    master = torch.cuda.current_stream(dev)
    e_start = torch.cuda.Event(enable_timing=True)
    e_start.record(master)
    for sl in slots:
        sl.stream.wait_event(e_start)
    n_launches = steps // B
    for k in range(n_launches):
        launch(k)
    e_ends = []
    for sl in slots:
        e = torch.cuda.Event(enable_timing=True)
        e.record(sl.stream)
        e_ends.append(e)
    barrier()
    t_wall1 = time.time()
    time.sleep(0.06)  # one more polling period, so that a sample lands right behind a short region
    clocks = sampler.stop(t_wall0, t_wall1)
    launches = sum(sl.ctx.launch_count for sl in slots) - l0
    ms_total = max(e_start.elapsed_time(e) for e in e_ends)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / steps
    for sl in slots:
        assert sl.sv.last_ok(), "timed batch did not verify"
    value = world * nb1 / (ms_per_step * 1e-3)

    # ---- e2e: the public host-buffer call (pinned host buffers -> H2D of proofs + instances, verification,
    # D2H of statuses + verdict inside every call), same number of batches in flight (one host thread per slot)
    e2e_launches = max(S, min(n_launches, 4 * S))
    e2e_steps = e2e_launches * B
    h_i, h_p, h_l = h_inst.numpy(), h_proofs.numpy(), torch.from_numpy(lens.astype(np.int32)).pin_memory().numpy().view(np.uint32)

    def e2e_worker(si):
        torch.cuda.set_device(local)
        res = None
        for k in range(si, e2e_launches, S):
            res = slots[si].sv.verify_host(h_i, n_inst, h_p, h_l, n, n_batches=B)
        return res

    pool = ThreadPoolExecutor(S) if world == 1 else None

    # world > 1: the collectives of all slots must be issued in the same order on every rank, so one host thread
    # pipelines the slots asynchronously: pinned H2D copy -> verify (C ABI, device pointers) -> D2H into pinned buffers.
    e2e_bufs = []
    if world > 1:
        for sl in slots:
            e2e_bufs.append(dict(d_inst=torch.empty_like(d_inputs[0][0]), d_proofs=torch.empty_like(d_inputs[0][1]),
                                 h_status=torch.zeros(n, dtype=torch.int32).pin_memory(),
                                 h_gather=torch.zeros(world * B * 256, dtype=torch.uint8).pin_memory(), h_final=torch.zeros(B * 256, dtype=torch.uint8).pin_memory()))

    def e2e_async(steps):
        for k in range(steps):
            sl, b = slots[k % S], e2e_bufs[k % S]
            with torch.cuda.stream(sl.stream):
                b["d_inst"].copy_(h_inst, non_blocking=True)
                b["d_proofs"].copy_(h_proofs, non_blocking=True)
            sl.sv.verify_dev(b["d_inst"], n_inst, b["d_proofs"], n, n_batches=B)
            with torch.cuda.stream(sl.stream):
                b["h_status"].copy_(sl.sv.d_status[:n], non_blocking=True)
                b["h_gather"].copy_(sl.sv.d_gather, non_blocking=True)
                b["h_final"].copy_(sl.sv.d_final, non_blocking=True)
        torch.cuda.synchronize()
        out = []
        for sl, b in zip(slots, e2e_bufs):
            g_ = b["h_gather"].numpy().reshape(world, B, 256)
            f_ = b["h_final"].numpy().reshape(B, 256)
            ok_ = bool(g_[:, :, 165].all() and f_[:, 164].all() and not f_[:, 160:164].any())
            out.append((ok_, b["h_status"].numpy()))
        return out

    if pool:
        list(pool.map(e2e_worker, range(S)))  # warm
    else:
        e2e_async(S)
    # the e2e region is short (tens of ms of host threads, pinned copies and kernels): it is run three times, every pass
    # max-reduced over the ranks, and the MEDIAN pass is reported (all three are in the line)
    e2e_passes = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        if pool:
            results = list(pool.map(e2e_worker, range(S)))
        else:
            results = e2e_async(e2e_launches)
        barrier()
        dt = time.perf_counter() - t0
        for ok, status in results:
            assert ok and (status == 0).all()
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_passes.append(world * nb1 * e2e_steps / float(t.item()))
    e2e_val = sorted(e2e_passes)[1]
    h2d = int(h_i.nbytes + h_p.nbytes + h_l.nbytes) // B   # per step (= per 4096-proof batch)
    d2h = int(nb1 * 4 + 256)

    # ---- per-kernel device time (CUDA events around every launch, one batch in flight) -> roofline.
    # Runs on EVERY rank: verify_dev contains the all_gather when world > 1.
    sl = slots[0]
    L, c = sl.ctx._L, sl.ctx._c
    L.svk_profile_enable(c, 1)
    prof_steps = 3
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for k in range(prof_steps):
        with torch.cuda.stream(sl.stream):
            flush.fill_(1)
        sl.sv.verify_dev(d_inputs[k % n_copies][0], n_inst, d_inputs[k % n_copies][1], n, n_batches=B)
    buf = ctypes.create_string_buffer(1 << 16)
    L.svk_profile_report(c, buf, len(buf))
    L.svk_profile_enable(c, 0)
    barrier()
    peak, peak_ms = sl.ctx.modmul_peak(4000)
    secondary = None
    if not args.no_secondary and args.scheme == "bdfg21":
        for sl_ in slots[1:]:
            sl_.ctx.close()  # the secondary configs need the memory (2^24-point MSM: ~6 GB of scratch)
        torch.cuda.empty_cache()
        secondary = run_secondary(torch, dist, V, sl.ctx, sl.stream, dev, g, peak, world, rank, args)
    if rank == 0:
        prof = json.loads(buf.value.decode())
        info = pv.info
        # Work per launch, two ways (DESIGN.md "work model"):
        #  canonical -- SURVEY 8d's algorithmic count (a squaring = 1 M, a Poseidon permutation = 600 M, a Fermat chain = 380 M);
        #  executed  -- integer-pipe time of what the kernels actually issue, in units of one plain Montgomery product
        #               (136 IMAD-pipe instructions): squaring 0.86 (117), dot2 1.47 (200), dot3 1.94 (264), measured by
        #               cuobjdump / tools/sqr_probe.cu.  roofline.frac uses EXECUTED work, so it is a utilisation (<= 1).
        S_, D2, D3 = 117 / 136, 200 / 136, 264 / 136
        perm_exec = (8 * (3 * (2 * S_ + 1) + 3 * D3) + 28 * (2 * (2 * S_ + 1) + 3 * D2 + D3)   # optimised Poseidon, T = 3; scaled partial
                     + ((2 * S_ + 1) + D2 + 2) + 1)                                          # rounds two at a time (poseidon.cuh): ~447
        chain_exec = 253 * S_ + 58                                                      # sliding-window square-root chain: 276
        inv_exec = 80                                                                   # binary-Euclid inversion: ALU work; ~55 IMAD.MOV/IMAD.X per step on the multiply pipe
        n_perm = info["n_poseidon_perms"]
        tape_other = info["n_fr_mul"] - 600 * n_perm                                    # scalar algebra incl. 3 Fermat chains at 380
        dbl_e, madd_e, add_e = 2 + 5 * S_, 5 + 4 * S_ + D2, 9 + 5 * S_ + D2             # Jacobian dbl / mixed add / add with dot2
        var_canon = info["msm_var_modmul_per_proof"]
        norm_e = 6 + S_                                                                # per table entry: prefix product, z^-1, z^-2, x, y
        var_exec = info["n_var_terms"] * (8 * dbl_e + 7 * madd_e + 50 * madd_e + 16 * norm_e) + info["var_lanes"] * (255 * dbl_e + inv_exec)
        work = {
            "k_tape": n * info["n_fr_mul"],
            "k_decompress": n * info["n_points"] * 372,
            "k_msm_var": n * var_canon,
        }
        work_exec = {
            "k_tape": n * (n_perm * perm_exec + tape_other - 3 * (380 - inv_exec)),
            "k_decompress": n * info["n_points"] * (chain_exec + 12),
            "k_msm_var": n * var_exec,
        }
        work_other = n * (info["msm_modmul_per_proof"] - info["msm_var_modmul_per_proof"])  # k_msm_sum + k_to_affine
        work_other_exec = work_other * (madd_e / 11)
        kernels = {k: {"launches": v["count"] / prof_steps, "ms_per_launch_of_B_batches": v["ms"] / prof_steps} for k, v in prof.items()}
        name = max((k for k in prof if k in work), key=lambda k: prof[k]["ms"])
        ms_launch = prof[name]["ms"] / prof[name]["count"]
        ach = work_exec[name] / (ms_launch * 1e-3)
        ach_canon = work[name] / (ms_launch * 1e-3)
        # DRAM traffic of the dominant kernel from the committed ncu --set full capture (profiles/r2_traffic.json, derived from
        # profiles/r2_ncu_full_summary.json), per proof x proofs per launch
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r2_traffic.json")))
            if name in tr:
                traffic = tr[name]["dram_bytes_per_launch"] / tr[name]["proofs_per_launch"] * n
                traffic_src = tr[name].get("source")
        except (OSError, ValueError, KeyError):
            pass
        hbm_peak = None  # driver-measured copy bandwidth (MEASURED_PEAKS.json); this path is not bound by it
        try:
            hbm_peak = float(json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")))["hbm_gbs"])
        except (OSError, ValueError, KeyError):
            pass
        total_work = (sum(work.values()) + work_other) / B  # per 4096-proof batch (fold and pairing not counted: < 8 %)
        total_work_exec = (sum(work_exec.values()) + work_other_exec) / B
        roofline = {
            "bound": "imad", "kernel": name, "achieved": ach / 1e9, "peak": peak / 1e9,
            "unit": "Gmodmul/s EXECUTED (integer-pipe time in units of one 8x32-bit-limb Montgomery product = 136 IMAD-pipe instructions; "
                    "squaring 0.86, dot2 1.47, dot3 1.94 -- cuobjdump, tools/sqr_probe.cu)",
            "achieved_canonical": ach_canon / 1e9, "frac_canonical": ach_canon / peak,
            "canonical_note": "SURVEY 8d algorithmic count (squaring = 1 M, Poseidon permutation = 600 M): exceeds the executed work "
                              "because the kernels issue fewer multiply instructions than the canonical algorithm, so it may pass 1",
            "frac": ach / peak, "traffic": traffic, "traffic_unit": "bytes of DRAM traffic per launch of this kernel (ncu dram__bytes_read.sum + dram__bytes_write.sum, "
            "scaled by proofs per launch)", "traffic_source": traffic_src,
            "peak_source": "measured in this run: svk_bench_modmul_peak (independent Montgomery-mul chains, 8 warps/SMSP on all SMs)",
            "kernel_ms_per_launch": ms_launch, "batches_per_launch": B, "kernels_one_launch_in_flight": kernels,
            # per GPU: every rank does `total_work` per step (weak scaling) against its own multiply peak
            "whole_step_frac": (total_work_exec / (ms_per_step * 1e-3)) / peak,
            "whole_step_frac_canonical": (total_work / (ms_per_step * 1e-3)) / peak,
            "hbm_gbs_algorithmic": (h2d + d2h) / (ms_per_step * 1e-3) / 1e9,
            "hbm_gbs_dominant_kernel": (traffic / (ms_launch * 1e-3) / 1e9) if traffic else None,
            "hbm_peak_gbs": hbm_peak, "hbm_frac_dominant_kernel": (traffic / (ms_launch * 1e-3) / 1e9 / hbm_peak) if (traffic and hbm_peak) else None,
        }
        base = None if args.no_cpu_baseline else cpu_baseline(g, args.cpu_sample, args.group_size, args.scheme)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32x8 (254-bit Fq/Fr Montgomery, integer pipe)",
            "data": data,
            "config": {"workload": WORKLOAD.format(scheme="shplonk" if args.scheme == "bdfg21" else "gwc"), "batch_per_gpu": nb1, "global_batch": world * nb1, "proof_bytes": info["proof_len"], "fold_group_size": args.group_size,
                       "batches_per_launch": B, "launches_in_flight": S, "launch_latency_ms": launch_latency_ms,
                       "fold": "flat (reference aggregation.rs:235-245)" if args.group_size in (0, 1) else f"tree, groups of {args.group_size}",
                       "batches_in_flight": S * B, "single_batch_latency_ms": latency_ms,
                       "l2": f"inputs rotate over {n_copies} distinct device copies ({n_copies * (h2d >> 20)} MiB > L2)",
                       "parallelism": f"proof-sharded x{world}, NCCL all_gather of {world} folded accumulators" if world > 1 else "single GPU",
                       "preflight_vs_oracle": preflight},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "passes": e2e_passes, "reported": "median of three passes of the same region"},
            "roofline": roofline, "cpu_baseline": base, "secondary": secondary,
        }
        json_out.write(json.dumps(out) + "\n")
        json_out.flush()
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
