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


def test_rust_ffi_declares_the_header():
    """rust/snark-verifier-cuda/src/ffi.rs (the binding a snark-verifier maintainer links) declares every entry point
    of include/svk.h with the same number of parameters, and nothing the header does not have."""
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "svk.h")).read(), flags=re.S)
    header = {}
    for _ret, name, params in re.findall(r"^(int|void|const char\*|uint64_t)\s+(svk_[a-z0-9_]+)\s*\((.*?)\);", src, flags=re.S | re.M):
        params = params.strip()
        header[name] = 0 if params in ("", "void") else params.count(",") + 1
    assert sorted(header) == declared_symbols()
    rs = open(os.path.join(ROOT, "rust", "snark-verifier-cuda", "src", "ffi.rs")).read()
    rs = re.sub(r"/\*.*?\*/", "", rs, flags=re.S)
    rs = re.sub(r"//[^\n]*", "", rs)
    rust = {}
    for name, params in re.findall(r"pub fn\s+(svk_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->[^;]*)?;", rs, flags=re.S):
        assert name not in rust, f"{name} declared twice in ffi.rs"
        rust[name] = len([x for x in params.split(",") if x.strip()])
    assert sorted(rust) == sorted(header), (sorted(set(header) - set(rust)), sorted(set(rust) - set(header)))
    for name, n in header.items():
        assert rust[name] == n, f"{name}: {n} parameters in svk.h, {rust[name]} in ffi.rs"


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
