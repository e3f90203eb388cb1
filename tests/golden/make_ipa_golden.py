#!/usr/bin/env python
"""Generates tests/golden/ipa_golden.json with the ORACLE (oracle/pasta.py, oracle/ipa.py): IPA deciding keys, accumulators produced by
an honest prover's base folding (pcs/ipa.rs:78-118), tampered ones, the expected `IpaAs::decide` statuses and the commitment
`multi_scalar_multiplication(h_coeffs(xi), g)` (pcs/ipa/decider.rs:47-56).  Run from the repo root:  python tests/golden/make_ipa_golden.py"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ipa, pasta  # noqa: E402


def main():
    out = {"note": "hex big-endian integers; points [x, y], identity = null", "cases": []}
    hx = lambda v: format(v, "x")  # noqa: E731
    pt = lambda p: None if p is None else [hx(p[0]), hx(p[1])]  # noqa: E731
    for C in (pasta.PALLAS, pasta.VESTA):
        for k in (1, 3, 4):
            rng = random.Random(7000 + 10 * C.id + k)
            g = [C.mul(C.gen, rng.randrange(1, C.n)) for _ in range(1 << k)]
            accs = []
            for i in range(5):
                xi = [rng.randrange(1, C.n) for _ in range(k)]
                u = ipa.fold_bases(C, g, xi)
                if i == 1:
                    u = C.add(u, C.gen)
                if i == 2 and k > 1:
                    xi = xi[1:] + xi[:1]
                if i == 3:
                    xi[-1] = 0
                    u = C.msm_naive(ipa.h_coeffs(xi, 1, C.n), g)
                if i == 4:
                    u = None
                commit = C.msm_naive(ipa.h_coeffs(xi, 1, C.n), g)
                accs.append({"xi": [hx(x) for x in xi], "u": pt(u), "commit": pt(commit),
                             "status": ipa.decide(C, g, ipa.IpaAccumulator(xi, u))})
            out["cases"].append({"curve": C.name, "curve_id": C.id, "k": k, "g": [pt(p) for p in g], "accumulators": accs})
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ipa_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("cases:", len(out["cases"]))


if __name__ == "__main__":
    main()
