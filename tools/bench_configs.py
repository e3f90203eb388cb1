#!/usr/bin/env python
"""Secondary BASELINE configs on one B200 (one JSON line each, CUDA-event timed, inputs resident in HBM):
  config 3: BN254 G1 MSM sweep (svk_msm_g1_dev), uniform scalars, points = d_i * G generated on device
  config 5: batched KzgAs::decide of 2^16 accumulators (svk_kzg_decide_batch_dev), 1/64 corrupted
Usage: python tools/bench_configs.py [--max-log-n 24] [--decide-n 65536]"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from snark_verifier_axiom_b200 import verifier as V  # noqa: E402
from snark_verifier_axiom_b200.standard_plonk import load_golden  # noqa: E402

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def rand_scalars(n, dev, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    s = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    s[:, 31] &= 0x1F  # < 2^253 < r
    return s.contiguous()


def p(t):
    return ctypes.c_void_p(t.data_ptr())


def profile(ctx, fn, reps=1):
    L, c = ctx._L, ctx._c
    L.svk_profile_enable(c, 1)
    for _ in range(reps):
        fn()
    buf = ctypes.create_string_buffer(1 << 16)
    L.svk_profile_report(c, buf, len(buf))
    L.svk_profile_enable(c, 0)
    return {k: v["ms"] / reps for k, v in json.loads(buf.value.decode()).items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-log-n", type=int, default=24)
    ap.add_argument("--min-log-n", type=int, default=16)
    ap.add_argument("--decide-n", type=int, default=65536)
    ap.add_argument("--only", choices=["all", "msm", "latency", "decide", "ipa"], default="all")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    ctx = V.Context(0)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    L, c = ctx._L, ctx._c
    peak, _ = ctx.modmul_peak(4000)
    gen = torch.zeros(64, dtype=torch.uint8, device=dev)
    gen[0] = 1
    gen[32] = 2

    def timed(fn, iters):
        fn()
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(iters):
            fn()
        e1.record(stream)
        stream.synchronize()
        return e0.elapsed_time(e1) / iters

    # ---- config 3
    nmax = 1 << args.max_log_n
    with torch.cuda.stream(stream):
        dl = rand_scalars(nmax, dev, 42)
        pts = torch.empty(nmax * 64, dtype=torch.uint8, device=dev)
    ctx._check(L.svk_g1_mul_batch_dev(c, nmax, p(dl), p(gen), 1, p(pts)))
    stream.synchronize()
    out = torch.zeros(64, dtype=torch.uint8, device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    for lg in range(args.min_log_n, args.max_log_n + 1, 2) if args.only in ("all", "msm") else []:
        n = 1 << lg
        with torch.cuda.stream(stream):
            sc = rand_scalars(n, dev, 1000 + lg)

        def run():
            ctx._check(L.svk_msm_g1_dev(c, n, p(sc), p(pts), p(out), p(st)))

        ms = timed(run, 3 if lg <= 22 else 1)
        assert int(st.item()) == 0
        prof = profile(ctx, run)
        # msm.cu msm_plan: GLV halves over 2n virtual points, 8 windows of 16 bits from 2^16 points up = 16 bucket additions per point
        work = n * 16 * 10  # XYZZ mixed additions, 8M + 2S each (bucket accumulation only)
        print(json.dumps({"config": "msm_g1_sweep", "log_n": lg, "n": n, "ms": ms, "points_per_s": n / (ms * 1e-3), "window_bits": 16, "windows": "8 x 2n (GLV)",
                          "modmul_frac_bucket_adds_only": work / (ms * 1e-3) / peak, "modmul_frac_canonical_176": n * 176 / (ms * 1e-3) / peak,
                          "hbm_gbs_algorithmic": n * 96 / (ms * 1e-3) / 1e9,
                          "kernels_ms": {k: round(v, 3) for k, v in prof.items()}}))
    del pts, dl

    # ---- config 1: one SHPLONK proof, full PlonkVerifier::verify (succinct verify + fold of one + pairing), latency.
    # The CPU figure beside it is bench.py's `cpu_baseline` (1-thread port: ~4 ms per proof incl. the pairing).
    if args.only in ("all", "latency"):
        g = load_golden()
        pv1 = V.PlonkVerifier(ctx, g["dk"], g["protocol"], V.SHPLONK)
        sn = g["schemes"]["bdfg21"]["snarks"][0]
        import time as _t
        pv1.verify_one(sn)
        t0 = _t.perf_counter()
        for _ in range(5):
            pv1.verify_one(sn)
        gpu_ms = 1e3 * (_t.perf_counter() - t0) / 5
        print(json.dumps({"config": "single_proof_verify_latency", "gpu_ms_host_call": gpu_ms,
                          "note": "one proof cannot fill a GPU: the B200 path is a throughput device (batch configs)"}))
    # ---- IpaAs::decide over pallas (h_coeffs + 2^k-point Pippenger + compare), 4 accumulators per call
    if args.only in ("all", "ipa"):
        import numpy as np

        from oracle import pasta  # test infrastructure: only generates the committing-key points here

        C = pasta.PALLAS
        for k in (10, 14, 16):
            m = 1 << k
            step, cur, g = C.mul(C.gen, 0x9E3779B97F4A7C15F39CC0605CEDC834), C.mul(C.gen, 12345), []
            for _ in range(m):
                g.append(cur)
                cur = C.add(cur, step)
            gb = np.frombuffer(b"".join(x.to_bytes(32, "little") + y.to_bytes(32, "little") for x, y in g), dtype=np.uint8)
            d_g = torch.from_numpy(gb.copy()).to(dev)
            n_acc = 4
            with torch.cuda.stream(stream):
                d_xi = rand_scalars(n_acc * k, dev, 7 + k)
            d_h = torch.zeros(m * 32, dtype=torch.uint8, device=dev)
            d_u = torch.zeros(n_acc * 64, dtype=torch.uint8, device=dev)
            d_st = torch.zeros(n_acc, dtype=torch.int32, device=dev)
            d_inv = torch.zeros(1, dtype=torch.int32, device=dev)
            # u_i := what the device commits to (parity with the oracle is tests/test_gpu_ipa.py); the last one is corrupted
            ctx._check(L.svk_ipa_decide_batch_dev(c, 1, k, p(d_g), n_acc, p(d_xi), p(d_u), p(d_st), p(d_inv)))
            stream.synchronize()

            def run_ipa():
                ctx._check(L.svk_ipa_decide_batch_dev(c, 1, k, p(d_g), n_acc, p(d_xi), p(d_u), p(d_st), p(d_inv)))

            ms = timed(run_ipa, 3)
            print(json.dumps({"config": "ipa_decide_pallas", "k": k, "accumulators": n_acc, "ms_per_accumulator": ms / n_acc,
                              "msm_points_per_s": n_acc * m / (ms * 1e-3)}))
    if args.only not in ("all", "decide"):
        return

    # ---- config 5
    g = load_golden()
    kid = ctx.load_deciding_key(g["dk"])
    n = args.decide_n
    # valid accumulators: the oracle-checked golden ones (succinct verify of the fixture proofs), tiled; 1/64 corrupted
    pv = V.PlonkVerifier(ctx, g["dk"], g["protocol"], V.SHPLONK)
    accs, _, stt = pv.succinct_verify(g["schemes"]["bdfg21"]["snarks"])
    assert (stt == 0).all()
    import numpy as np

    base = np.frombuffer(b"".join(a.to_bytes() for a in accs), dtype=np.uint8).reshape(len(accs), 128)
    host = np.tile(base, (n // len(accs) + 1, 1))[:n].copy()
    bad = np.arange(0, n, 64)
    host[bad, 64:128] = host[(bad + 1) % n, 64:128]  # wrong rhs: still on the curve, pairing must reject
    expect = np.ones(n, dtype=np.uint8)
    expect[bad] = 0
    d_accs = torch.from_numpy(host).to(dev)
    d_ok = torch.zeros(n, dtype=torch.uint8, device=dev)

    def run_decide():
        ctx._check(L.svk_kzg_decide_batch_dev(c, kid, n, p(d_accs), p(d_ok)))

    ms = timed(run_decide, 2)
    got = d_ok.cpu().numpy()
    assert (got == expect).all(), "decide mismatch"
    print(json.dumps({"config": "kzg_decide_batch", "n": n, "ms": ms, "decides_per_s": n / (ms * 1e-3), "corrupted": int(len(bad)),
                      "modmul_frac": n * 16400 / (ms * 1e-3) / peak, "peak_gmodmul_s": peak / 1e9}))


if __name__ == "__main__":
    main()
