"""CPU-side checks of the product: libsvk loads and exports every symbol include/svk.h declares, it
fails loudly without a GPU (no CPU fallback), and the host-side protocol mirror serializes the
StandardPlonk protocol exactly like the oracle's restatement of `compile()`."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "svk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(svk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from snark_verifier_axiom_b200._lib import LIB_PATH

    if not os.path.exists(LIB_PATH):
        import __graft_entry__ as g

        g.build()
    L = ctypes.CDLL(LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(L, s), f"libsvk.so does not export {s}"


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from snark_verifier_axiom_b200 import SvkError
    from snark_verifier_axiom_b200.verifier import Context

    with pytest.raises(SvkError):
        Context(0)


def test_product_protocol_bytes_match_oracle_restatement():
    from oracle import forge
    from snark_verifier_axiom_b200.standard_plonk import load_golden, standard_plonk_protocol

    from .util import to_product_protocol

    S = forge.Setup(0)
    a = to_product_protocol(S.protocol).to_bytes()
    b = standard_plonk_protocol(8, S.preprocessed, S.transcript_initial_state).to_bytes()
    assert a == b
    g = load_golden()
    assert g["protocol"].to_bytes() == a
    assert len(g["schemes"]["bdfg21"]["snarks"]) == 64 and len(g["schemes"]["bdfg21"]["snarks"][0].proof) == 896
    assert g["protocol"].quotient.num_chunk() == 3
