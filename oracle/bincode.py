"""What `bincode::serialize` writes for the reference's `PlonkProtocol<G1Affine>` / `Snark`.  TEST INFRASTRUCTURE ONLY.

bincode 1.3.3, default options (snark-verifier-sdk/Cargo.toml:18; `bincode::deserialize_from` at sdk/src/halo2.rs:262-269):
little-endian fixed-width integers, `usize` as u64, enum variant index as u32, `Option` as one tag byte, `Vec` as u64 length +
elements, structs / tuples / `Box` as the concatenation of their fields in declaration order.  Field order from the serde
derives: `PlonkProtocol` verifier/plonk/protocol.rs:20-63, `Domain` util/arithmetic.rs:130-142, `Query` protocol.rs:296-300,
`Rotation(i32)` arithmetic.rs:99-100, `QuotientPolynomial` :281-285, `Expression` :308-319, `CommonPolynomial` :180-185,
`LinearizationStrategy` :503-513, `Snark` sdk/src/lib.rs:44-50.

PARITY UNPINNED: halo2curves' serde impl for Fr / Fq / G1Affine is not in the tree (Cargo.lock:1803-1826) and no
reference-written file exists to compare with.  `fe="montgomery"` restates the `derive(Serialize)` of halo2curves 0.3.x on
`Fr([u64; 4])` (raw Montgomery limbs, R = 2^256) as recalled; `fe="canonical"` the 32-byte `to_repr` of later versions.
"""
import struct

from .bn254 import P, R

_RMONT = 1 << 256


def _u64(v):
    return struct.pack("<Q", v)


def _fe(v, modulus, fe):
    if fe == "montgomery":
        v = v * _RMONT % modulus
    return int(v).to_bytes(32, "little")


def _query(q):
    return _u64(q[0]) + struct.pack("<i", q[1])


def _expr(e, fe):
    t = e[0]
    if t == "const":
        return struct.pack("<I", 0) + _fe(e[1], R, fe)
    if t == "identity":
        return struct.pack("<II", 1, 0)
    if t == "lagrange":
        return struct.pack("<IIi", 1, 1, e[1])
    if t == "poly":
        return struct.pack("<I", 2) + _query((e[1], e[2]))
    if t == "challenge":
        return struct.pack("<I", 3) + _u64(e[1])
    if t == "neg":
        return struct.pack("<I", 4) + _expr(e[1], fe)
    if t == "sum":
        return struct.pack("<I", 5) + _expr(e[1], fe) + _expr(e[2], fe)
    if t == "product":
        return struct.pack("<I", 6) + _expr(e[1], fe) + _expr(e[2], fe)
    if t == "scaled":
        return struct.pack("<I", 7) + _expr(e[1], fe) + _fe(e[2], R, fe)
    if t == "distribute_powers":
        return struct.pack("<I", 8) + _u64(len(e[1])) + b"".join(_expr(x, fe) for x in e[1]) + _expr(e[2], fe)
    raise ValueError(t)


def serialize_protocol(p, fe="montgomery"):
    d = p.domain
    out = _u64(d.k) + _u64(d.n) + _fe(d.n_inv, R, fe) + _fe(d.gen, R, fe) + _fe(d.gen_inv, R, fe)
    out += _u64(len(p.preprocessed))
    for pt in p.preprocessed:
        x, y = pt if pt is not None else (0, 0)
        out += _fe(x, P, fe) + _fe(y, P, fe)
    for lst in (p.num_instance, p.num_witness, p.num_challenge):
        out += _u64(len(lst)) + b"".join(_u64(v) for v in lst)
    for lst in (p.evaluations, p.queries):
        out += _u64(len(lst)) + b"".join(_query(q) for q in lst)
    out += _u64(p.quotient.chunk_degree) + _expr(p.quotient.numerator, fe)
    out += b"\x00" if p.transcript_initial_state is None else b"\x01" + _fe(p.transcript_initial_state, R, fe)
    assert p.instance_committing_key is None
    out += b"\x00"
    lin = {None: None, "WithoutConstant": 0, "MinusVanishingTimesQuotient": 1}[p.linearization]
    out += b"\x00" if lin is None else b"\x01" + struct.pack("<I", lin)
    out += _u64(len(p.accumulator_indices))
    for idx in p.accumulator_indices:
        out += _u64(len(idx)) + b"".join(_u64(i) + _u64(j) for i, j in idx)
    return out


def serialize_snark(p, instances, proof, fe="montgomery"):
    """`Snark { protocol, instances: Vec<Vec<Fr>>, proof: Vec<u8> }` (sdk/src/lib.rs:44-50)"""
    out = serialize_protocol(p, fe)
    out += _u64(len(instances))
    for col in instances:
        out += _u64(len(col)) + b"".join(_fe(x, R, fe) for x in col)
    return out + _u64(len(proof)) + bytes(proof)
