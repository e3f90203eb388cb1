"""Every kernel family has two schedules chosen by the size of a launch (csrc/svk_ctx.h): the THROUGHPUT form (one proof / group /
accumulator per thread: k_tape, k_fold_sponge, k_msm_var<true> + k_msm_sum<1>, k_group_var, k_decide) and the LATENCY form
(warp-cooperative Poseidon k_tape_coop / k_fold_sponge_dbl, per-term GLV MSM lanes, doublings beside the sponge + k_fold_add,
block-cooperative pairing k_decide_coop).  The small batches of the other GPU tests only reach the latency forms; here the same
proofs go through BOTH (forced with the SVK_* environment knobs read by svk_create) and must give the oracle's values bit for bit:
accumulators, challenges, statuses, fold trees, verdicts (verifier/plonk.rs:58-135, accumulation.rs:29-62, decider.rs:60-81)."""
import os

import pytest

from oracle import api, forge

from .util import to_product_protocol

pytestmark = pytest.mark.gpu

KNOBS = ("SVK_TAPE_COOP_MAX", "SVK_MSM_LATENCY_THREADS_MAX", "SVK_FOLD_DBL_THREADS_MAX", "SVK_DECIDE_COOP_MAX")
# "throughput-2-lanes": the throughput forms with two threads per proof in k_msm_var and k_msm_sum (SVK_VAR_LANES, SVK_MSM_LANES)
LANES = ("SVK_VAR_LANES", "SVK_MSM_LANES")


@pytest.fixture(scope="module", params=["throughput", "latency", "throughput-2-lanes"])
def env(request):
    from snark_verifier_axiom_b200 import verifier as V

    saved = {k: os.environ.get(k) for k in KNOBS + LANES}
    for k in KNOBS:
        os.environ[k] = "1000000" if request.param == "latency" else "0"
    for k in LANES:
        if request.param == "throughput-2-lanes":
            os.environ[k] = "2"
        else:
            os.environ.pop(k, None)
    try:
        S = forge.Setup(0)
        ctx = V.Context(0)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    dk = V.KzgDecidingKey.new(S.dk.svk.g, S.dk.g2, S.dk.s_g2)
    AS = V.KzgAs(ctx, dk)
    proto = to_product_protocol(S.protocol)
    pv = {m: V.PlonkVerifier(ctx, dk, proto, m, kzg_as=AS) for m in (V.SHPLONK, V.GWC)}
    yield request.param, V, S, ctx, AS, pv
    ctx.close()


@pytest.mark.parametrize("scheme,mos", [("bdfg21", 0), ("gwc19", 1)])
def test_both_schedules_match_oracle(env, scheme, mos):
    which, V, S, ctx, AS, pv = env
    n = 37  # more than a warp, ragged
    insts, proofs = forge.forge_batch(S, scheme, n, seed0=7000)
    proofs = [bytearray(p) for p in proofs]
    proofs[3][9 * 32 + 3] ^= 1          # evaluation: decodes, pairing rejects
    proofs[11][0:32] = bytes(32)        # identity point
    proofs[20][9 * 32 : 10 * 32] = b"\xff" * 32  # scalar >= r
    snarks = [V.Snark(i, bytes(p)) for i, p in zip(insts, proofs)]
    l0 = ctx.launch_count
    accs, chals, st = pv[mos].succinct_verify(snarks)
    assert ctx.launch_count > l0
    good = []
    for i in range(n):
        want = api.status_of(api.succinct_verify, S.dk.svk, S.protocol, insts[i], bytes(proofs[i]), scheme)
        assert (int(st[i]) & 0xFF) == want, (which, i)
        if want == 0:
            oa, proof = api.succinct_verify(S.dk.svk, S.protocol, insts[i], bytes(proofs[i]), scheme, want_proof=True)
            ch = [c.v for c in proof.challenges] + [proof.z.v]
            ch += [proof.pcs.mu.v, proof.pcs.gamma.v, proof.pcs.z_prime.v] if scheme == "bdfg21" else [proof.pcs.v.v, proof.pcs.u.v]
            assert (accs[i].lhs, accs[i].rhs) == (oa[0].lhs.pt, oa[0].rhs.pt), (which, i)
            assert chals[i] == ch, (which, i)
            if i != 3:
                good.append(i)
    pairs = [(accs[i].lhs, accs[i].rhs) for i in good]
    for m in (0, 2, 4, 8):
        (el, er), rs = api.fold(pairs, m)
        got, r = AS.create_proof([accs[i] for i in good], m)
        assert (got.lhs, got.rhs) == (el, er) and r == rs[-1], (which, m)
    dec = AS.decide_batch([accs[i] for i in good] + [accs[3]])
    assert dec == [True] * len(good) + [False]
    res = pv[mos].verify([snarks[i] for i in good], group_size=4)
    assert res.ok and (res.folded.lhs, res.folded.rhs) == api.fold(pairs, 4)[0]
    res = pv[mos].verify([snarks[i] for i in good] + [snarks[3]], group_size=4)
    assert not res.ok and int(res.status[-1]) == 3 and (res.status[:-1] == 0).all()


def test_schedules_really_differ(env):
    """The knobs select different kernels: per-kernel profile of one succinct verify + fold + decide."""
    import ctypes
    import json

    which, V, S, ctx, AS, pv = env
    insts, proofs = forge.forge_batch(S, "bdfg21", 9, seed0=7100)
    L, c = ctx._L, ctx._c
    L.svk_profile_enable(c, 1)
    assert pv[0].verify([V.Snark(i, p) for i, p in zip(insts, proofs)], group_size=4).ok
    buf = ctypes.create_string_buffer(1 << 16)
    L.svk_profile_report(c, buf, len(buf))
    L.svk_profile_enable(c, 0)
    names = set(json.loads(buf.value.decode()))
    if which.startswith("throughput"):
        assert {"k_tape", "k_fold_sponge", "k_group_var", "k_decide"} <= names and not names & {"k_tape_coop", "k_decide_coop", "k_fold_sponge_dbl"}
    else:
        assert {"k_tape_coop", "k_fold_sponge_dbl", "k_fold_add", "k_decide_coop"} <= names and not names & {"k_tape", "k_decide", "k_fold_sponge"}
