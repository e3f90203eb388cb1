"""Shared helpers for the test-suite: byte packing at the C ABI (include/svk.h)."""
import ctypes

import numpy as np

from oracle import bn254


def fe(v):
    return int(v).to_bytes(32, "little")


def g1_bytes(pt):
    return bytes(64) if pt is None else fe(pt[0]) + fe(pt[1])


def g1_from(b):
    x = int.from_bytes(b[:32], "little")
    y = int.from_bytes(b[32:64], "little")
    return None if x == 0 and y == 0 else (x, y)


def g2_bytes(q):
    return fe(q[0][0]) + fe(q[0][1]) + fe(q[1][0]) + fe(q[1][1])


def dk_bytes(dk):
    return g1_bytes(dk.svk.g) + g2_bytes(dk.g2) + g2_bytes(dk.s_g2)


def acc_bytes(lhs, rhs):
    return g1_bytes(lhs) + g1_bytes(rhs)


def np_u8(b):
    return np.frombuffer(bytes(b), dtype=np.uint8).copy()


def ptr(arr):
    return arr.ctypes.data_as(ctypes.c_void_p)
