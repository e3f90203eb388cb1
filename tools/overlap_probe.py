#!/usr/bin/env python
"""Diagnostic: throughput of the pipeline STAGES with S batches in flight (one context + stream each).
stage 'succinct' = decompress + tape + per-proof MSM; 'fold' = KzgAs tree fold of 4096 accumulators;
'decide' = one pairing; 'full' = svk_plonk_verify_batch_dev."""
import ctypes, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from snark_verifier_axiom_b200 import verifier as V
from snark_verifier_axiom_b200.standard_plonk import load_golden

def p(t): return ctypes.c_void_p(t.data_ptr())

def main():
    n = 4096
    dev = torch.device("cuda", 0)
    g = load_golden()
    snarks = [g["schemes"]["bdfg21"]["snarks"][i % 64] for i in range(n)]
    for S in (1, 8):
        slots = []
        for _ in range(S):
            ctx = V.Context(0); st = torch.cuda.Stream(device=dev); ctx.set_stream(st.cuda_stream)
            pv = V.PlonkVerifier(ctx, g["dk"], g["protocol"], V.SHPLONK)
            slots.append((ctx, st, pv))
        inst, n_inst, proofs, lens = slots[0][2].pack(snarks)
        d_inst = torch.from_numpy(inst).to(dev); d_proofs = torch.from_numpy(proofs).to(dev)
        bufs = [dict(acc=torch.zeros(n*128, dtype=torch.uint8, device=dev), st=torch.zeros(n, dtype=torch.int32, device=dev), rec=torch.zeros(256, dtype=torch.uint8, device=dev)) for _ in range(S)]
        def run(stage, k):
            ctx, st, pv = slots[k % S]; b = bufs[k % S]; L, c = ctx._L, ctx._c
            if stage == "succinct":
                ctx._check(L.svk_plonk_succinct_verify_batch_dev(c, pv.pid, n, p(d_inst), n_inst, p(d_proofs), proofs.shape[1], None, p(b["acc"]), None, p(b["st"])))
            elif stage == "fold":
                ctx._check(L.svk_kzg_as_fold_dev(c, n, p(b["acc"]), 8, p(b["rec"]), ctypes.c_void_p(b["rec"].data_ptr()+128), ctypes.c_void_p(b["rec"].data_ptr()+160)))
            elif stage == "decide":
                ctx._check(L.svk_kzg_decide_batch_dev(c, pv.kzg_as.dk_id, 1, p(b["rec"]), ctypes.c_void_p(b["rec"].data_ptr()+164)))
            else:
                ctx._check(L.svk_plonk_verify_batch_dev(c, pv.pid, n, p(d_inst), n_inst, p(d_proofs), proofs.shape[1], None, 8, p(b["acc"]), p(b["st"]), p(b["rec"])))
        for stage in ("succinct", "fold", "decide", "full"):
            for k in range(2 * S): run(stage, k)
            torch.cuda.synchronize()
            K = 8 * S
            t0 = time.perf_counter()
            for k in range(K): run(stage, k)
            t_launch = time.perf_counter() - t0
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(json.dumps({"inflight": S, "stage": stage, "ms_per_batch": 1e3 * dt / K, "host_launch_ms_per_batch": 1e3 * t_launch / K}))
        for ctx, _, _ in slots: ctx.close()

main()
