"""Memory safety of the arithmetic the kernels share with the host build (field, curve, pairing incl. the block-cooperative
program, Poseidon, GLV, Straus, the protocol compiler + tape VM, bincode ingestion): tests/test_host_arith.py once more with
-fsanitize=address,undefined compiled into tests/host/hostlib.cpp.  Any out-of-bounds access, use of a dead stack slot, signed
overflow or misaligned / oversized shift in those headers aborts the run.  (compute-sanitizer is not available on the GPU pool.)"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_arith_under_asan_ubsan():
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan not installed")
    # libstdc++ is preloaded with it: python does not link it, and ASan's __cxa_throw interceptor needs the real one at start-up
    std = subprocess.run(["gcc", "-print-file-name=libstdc++.so"], capture_output=True, text=True).stdout.strip()
    env = dict(os.environ, SVK_HOSTLIB_SANITIZE="1", LD_PRELOAD=f"{asan} {std}", ASAN_OPTIONS="detect_leaks=0:abort_on_error=0",
               UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_host_arith.py", "-x", "-q", "-p", "no:cacheprovider"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0, tail
    assert "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr, tail
