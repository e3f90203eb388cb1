"""The GPU-assisted synthetic-workload generator produces distinct proofs that the ORACLE accepts
(snark_verifier_axiom_b200/synth.py; trapdoor forging)."""
import json
import os

import numpy as np
import pytest

from oracle import api, forge

pytestmark = pytest.mark.gpu


def test_synth_proofs_accepted_by_oracle_and_gpu():
    from snark_verifier_axiom_b200 import synth, verifier as V
    from snark_verifier_axiom_b200.standard_plonk import GOLDEN_PATH, load_golden

    g = load_golden()
    raw = json.load(open(GOLDEN_PATH))
    s, vk_dlogs = int(raw["trapdoor_s"], 16), [int(x, 16) for x in raw["vk_dlogs"]]
    S = forge.Setup(0)
    assert s == S.s and vk_dlogs == S.vk_dlogs
    ctx = V.Context(0)
    pv = V.PlonkVerifier(ctx, g["dk"], g["protocol"], V.SHPLONK)
    n = 40
    inst, proofs = synth.forge_shplonk_batch(pv, s, vk_dlogs, n, seed=3)
    assert len({p.tobytes() for p in proofs}) == n and len({p[:32].tobytes() for p in proofs}) == n  # distinct
    snarks = synth.snarks_from_arrays(inst, proofs)
    for sn in snarks[:6]:
        assert api.status_of(api.verify, S.dk, S.protocol, sn.instances, sn.proof, "bdfg21") == 0
    res = pv.verify(snarks, group_size=8)
    assert res.ok and (res.status == 0).all()
    bad = bytearray(snarks[3].proof)
    bad[-1] ^= 0x80  # flip the sign of W': still a curve point, pairing must fail
    snarks[3] = V.Snark(snarks[3].instances, bytes(bad))
    res = pv.verify(snarks, group_size=8)
    assert not res.ok and int(res.status[3]) == 3
    ctx.close()
