"""CPU oracle for the NativeLoader KZG/PLONK verification path of snark-verifier.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  The product (`snark_verifier_axiom_b200`) never does.

It is an exact-integer (Python `int`) restatement of the reference's algorithm
for the hot path; every function cites the reference `file:line` it follows
(paths relative to the reference root, e.g. `snark-verifier/src/...`).

Parity pinning status
---------------------
* Poseidon (constants, MDS, optimised permutation): PINNED by the reference's two
  known-answer tests (`snark-verifier/src/util/hash/poseidon/tests.rs:7-32`, `:35-85`),
  checked in `tests/test_oracle_poseidon.py`.
* Everything that bottoms out in the un-vendored `halo2curves 0.3.1`
  (Cargo.lock:1803-1826; BN254 Fr/Fq/G1/G2/pairing, point encodings): the
  reference ships no golden vectors and cannot be built here (no Rust toolchain)
  => "parity unpinned" BY THE REFERENCE.  The restatement follows the published
  BN254 definition and is self-checked by exact algebra (bilinearity,
  trapdoor-forged proofs accept, mutations reject).  The group law and the
  pairing are additionally pinned from OUTSIDE the repository by the public
  EIP-196 / EIP-197 precompile test vectors (ecAdd, ecMul, ecPairing:
  `tests/golden/eip196_197_vectors.py`, checked at every layer -- this oracle, oracle/c,
  the device headers built for the host, the CUDA kernels).  What stays unpinned:
  halo2curves' compressed-G1 byte format (isolated in `bn254.g1_from_bytes /
  g1_to_bytes`) and the bincode layout of `Snark` files.
"""
