"""ctypes access to oracle/libvcf_in_oracle.so (CPU restatement of the input path) -- test infrastructure."""
import ctypes as C
import gzip
import json
import os
import subprocess

import numpy as np

from vcfgl_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "libvcf_in_oracle.so")
INPUTS = os.path.join(ROOT, "tests", "golden", "inputs")

_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(ORACLE_DIR, "vcf_in_oracle.c")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "libvcf_in_oracle.so"], stdout=subprocess.DEVNULL)
        L = C.CDLL(LIB)
        L.vin_oracle_parse.argtypes = [C.c_void_p, C.c_size_t, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                       C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]
        L.vin_oracle_parse.restype = C.c_int64
        _lib = L
    return _lib


def parse(body: bytes, S: int, gt_source: int, rm_invar: int = 0, final: bool = True, max_records: int = None):
    """-> (sites [n] structured like vgl_in_site, rows uint8 [n, S], bytes consumed)"""
    if max_records is None:
        max_records = body.count(b"\n") + 1
    sites = np.zeros(max_records, capi.IN_SITE_DTYPE)
    rows = np.zeros((max_records, S), np.uint8)
    used = C.c_size_t()
    buf = np.frombuffer(body, np.uint8)
    n = lib().vin_oracle_parse(buf.ctypes.data if len(buf) else None, len(buf), S, gt_source, rm_invar, int(final), max_records,
                               sites.ctypes.data, rows.ctypes.data, C.byref(used))
    return sites[:n], rows[:n], int(used.value)


def unpack_row(row: np.ndarray) -> np.ndarray:
    """packed bytes [S] -> int8 [2S] like true_gts_acgt_int (-1 missing)"""
    lo = (row & 0xF).astype(np.int8)
    hi = (row >> 4).astype(np.int8)
    out = np.empty(2 * len(row), np.int8)
    out[0::2] = np.where(lo == 0xF, -1, lo)
    out[1::2] = np.where(hi == 0xF, -1, hi)
    return out


def load_input(name: str) -> bytes:
    return gzip.open(os.path.join(INPUTS, name + ".gz"), "rb").read()


def in_cases() -> dict:
    return json.load(open(os.path.join(INPUTS, "in_cases.json")))
