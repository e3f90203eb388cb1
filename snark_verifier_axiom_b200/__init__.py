"""B200-native batch verifier for the NativeLoader KZG/PLONK path of snark-verifier.

Host-side mirror of the reference's call surface over the C ABI of `libsvk.so`
(`include/svk.h`).  All verification arithmetic runs in hand-written sm_100a CUDA kernels."""
from ._lib import SvkError, lib  # noqa: F401
