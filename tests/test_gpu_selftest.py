"""Exhaustive device self-test: the 3-instruction division by 10 and the branch-free lroundf used by
the kernels equal the reference's expressions for every float bit pattern in their domains."""
import ctypes as C

import pytest

from vcfgl_b200 import capi

pytestmark = pytest.mark.gpu


def test_arithmetic_shortcuts_exhaustive():
    lib = capi.load()
    n, first = C.c_int64(-1), C.c_uint32(0)
    assert lib.vgl_selftest(0, C.byref(n), C.byref(first)) == 0
    assert n.value == 0, "first mismatching float bits: 0x%08x (%d mismatches)" % (first.value, n.value)
