"""Generates tests/golden/standard_plonk_k8.json with the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  Deterministic (seeded); the file is committed so that the GPU
box (which has neither /root/reference nor needs the slow Python forger at bench time) can load it.

Content: trapdoor SRS + StandardPlonk k=8 verifying key, N forged-valid proofs per
multi-open scheme with the oracle's expected challenges / accumulators for the first few, expected
folds (flat + trees) and decide bits."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import api, bn254, forge  # noqa: E402

N = 64
N_EXPECT = 8


def hx(v):
    return "%064x" % v


def pt(p):
    return None if p is None else [hx(p[0]), hx(p[1])]


def main():
    S = forge.Setup(0)
    out = {
        "k": 8,
        "g1": pt(S.g1),
        "g2": [[hx(S.g2[0][0]), hx(S.g2[0][1])], [hx(S.g2[1][0]), hx(S.g2[1][1])]],
        "s_g2": [[hx(S.s_g2[0][0]), hx(S.s_g2[0][1])], [hx(S.s_g2[1][0]), hx(S.s_g2[1][1])]],
        "preprocessed": [pt(p) for p in S.preprocessed],
        "transcript_initial_state": hx(S.transcript_initial_state),
        # trapdoor of this TEST SRS and discrete logs of the vk commitments (what makes forging possible)
        "trapdoor_s": hx(S.s),
        "vk_dlogs": [hx(d) for d in S.vk_dlogs],
        "schemes": {},
    }
    for scheme in ("bdfg21", "gwc19"):
        insts, proofs = forge.forge_batch(S, scheme, N, seed0=1)
        expect = []
        pairs = []
        for i in range(N):
            accs, proof = api.succinct_verify(S.dk.svk, S.protocol, insts[i], proofs[i], scheme, want_proof=True)
            pairs.append((accs[0].lhs.pt, accs[0].rhs.pt))
            if i < N_EXPECT:
                ch = [c.v for c in proof.challenges] + [proof.z.v]
                if scheme == "bdfg21":
                    ch += [proof.pcs.mu.v, proof.pcs.gamma.v, proof.pcs.z_prime.v]
                else:
                    ch += [proof.pcs.v.v, proof.pcs.u.v]
                expect.append({"challenges": [hx(c) for c in ch], "lhs": pt(pairs[i][0]), "rhs": pt(pairs[i][1])})
        folds = {}
        for m in (0, 4, 8):
            (l, r), rs = api.fold(pairs, m)
            folds[str(m)] = {"lhs": pt(l), "rhs": pt(r), "r_root": hx(rs[-1])}
            assert api.decide(S.dk, (l, r))
        out["schemes"][scheme] = {
            "instances": [[[hx(x) for x in col] for col in inst] for inst in insts],
            "proofs": [p.hex() for p in proofs],
            "expect": expect,
            "folds": folds,
        }
        print(scheme, "done")
    path = os.path.join(ROOT, "tests", "golden", "standard_plonk_k8.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
