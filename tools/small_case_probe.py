#!/usr/bin/env python
"""One small pass over every kernel family through the host-buffer C ABI (ctypes + numpy only, no torch), in either schedule of
every kernel family (tests/test_gpu_schedules.py): `python tools/small_case_probe.py latency|throughput [n_msm]`.
Results are checked against the committed golden fixture (tests/golden, oracle-generated) and against single multiplications,
so a pass is a parity pass on small, ragged cases.  (compute-sanitizer is not available on this pool: memory safety of the
shared arithmetic is covered on the CPU by tests/host built with -fsanitize=address,undefined, see tests/test_host_asan.py.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

which = sys.argv[1] if len(sys.argv) > 1 else "latency"
n_msm = int(sys.argv[2]) if len(sys.argv) > 2 else 300
for k in ("SVK_TAPE_COOP_MAX", "SVK_MSM_LATENCY_THREADS_MAX", "SVK_FOLD_DBL_THREADS_MAX", "SVK_DECIDE_COOP_MAX"):
    os.environ[k] = "0" if which == "throughput" else "1000000"

import numpy as np  # noqa: E402

from snark_verifier_axiom_b200 import verifier as V  # noqa: E402
from snark_verifier_axiom_b200.standard_plonk import load_golden  # noqa: E402


def main():
    g = load_golden()
    ctx = V.Context(0)
    AS = V.KzgAs(ctx, g["dk"])
    done = []
    for name, mos in (("bdfg21", V.SHPLONK), ("gwc19", V.GWC)):
        sc = g["schemes"][name]
        pv = V.PlonkVerifier(ctx, g["dk"], g["protocol"], mos, kzg_as=AS)
        snarks = sc["snarks"][:40]  # more than a warp, ragged
        accs, chals, st = pv.succinct_verify(snarks)
        assert (st == 0).all()
        for a, c, e in zip(accs, chals, sc["expect"]):
            assert (a.lhs, a.rhs) == (e["lhs"], e["rhs"]) and list(c)[: len(e["challenges"])] == e["challenges"], name
        full, _, st = pv.succinct_verify(sc["snarks"])
        for m, f in sc["folds"].items():
            got, r = AS.create_proof(full, m)
            assert (got.lhs, got.rhs) == (f["lhs"], f["rhs"]) and r == f["r_root"], (name, m)
        res = pv.verify(snarks, group_size=4)
        assert res.ok
        bad = bytearray(snarks[3].proof)
        bad[9 * 32 + 5] ^= 2
        res = pv.verify(snarks[:8] + [V.Snark(snarks[3].instances, bytes(bad))], group_size=4)
        assert not res.ok and int(res.status[-1]) == 3 and (res.status[:-1] == 0).all()
        assert AS.decide_batch(accs[:5]) == [True] * 5
        assert AS.decide_batch([V.KzgAccumulator(accs[0].lhs, accs[1].rhs)]) == [False]
        done.append(name)
    # Pippenger (all its kernels) against a sum of single multiplications
    rng = np.random.default_rng(1)
    R = V.FR_MODULUS
    scal = [int.from_bytes(rng.integers(0, 256, 32, dtype=np.uint8).tobytes(), "little") % R for _ in range(n_msm)]
    base = V.g1_mul_batch(ctx, scal[::-1], [(1, 2)])
    got = V.multi_scalar_multiplication(ctx, scal, base)
    want = V.g1_mul_batch(ctx, [sum(a * b for a, b in zip(scal, scal[::-1])) % R], [(1, 2)])[0]
    assert got == want
    done.append(f"msm{n_msm}")
    ctx.close()
    print("small_case_probe", which, "ok:", " ".join(done), flush=True)


if __name__ == "__main__":
    main()
