"""Reader for the replay-capture files written by oracle/ref_dump_hooks.h
(instrumented reference, oracle/_ref/vcfgl_ref_dump) -- test infrastructure."""
import struct
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

MAGIC = 0x444C4756


@dataclass
class SiteDump:
    ret: int
    pos: int
    rid: int
    S: int
    stale_base: int
    site_eprob: Optional[float]
    n_alleles: int
    n_alleles_observed: int
    n_genotypes: int
    allele_unobserved: int
    alleles2acgt: np.ndarray
    acgt2alleles: np.ndarray
    flags: int
    gts: np.ndarray          # int8 [2S]
    depths: np.ndarray       # int32 [S] as drawn
    fmt_dp: np.ndarray       # int32 [S]
    info_dp: int
    r_sample: np.ndarray
    r_base: np.ndarray
    r_strand: np.ndarray
    r_qs: np.ndarray
    r_adjqs: np.ndarray
    r_eprob: np.ndarray
    tails: np.ndarray
    em_sample: np.ndarray
    em_n: np.ndarray
    em_codes: np.ndarray
    out: dict = field(default_factory=dict)   # name -> ndarray (bit-exact reference outputs)


class _R:
    def __init__(self, buf):
        self.b = buf
        self.o = 0

    def s(self, fmt):
        v = struct.unpack_from("<" + fmt, self.b, self.o)
        self.o += struct.calcsize("<" + fmt)
        return v if len(v) > 1 else v[0]

    def a(self, dtype, n):
        dt = np.dtype(dtype)
        v = np.frombuffer(self.b, dtype=dt, count=n, offset=self.o).copy()
        self.o += dt.itemsize * n
        return v


def read_dump(path) -> List[SiteDump]:
    return list(iter_dump(path))


def iter_dump(path):
    """the sites of a capture one at a time (large captures: tools/make_stats_golden.py)"""
    if str(path).endswith(".gz"):
        import gzip
        buf = gzip.open(path, "rb").read()
    else:
        buf = np.memmap(path, dtype=np.uint8, mode="r") if __import__("os").path.getsize(path) > 0 else b""
    r = _R(buf)
    while r.o < len(buf):
        magic = r.s("I")
        assert magic == MAGIC, "bad magic at %d" % r.o
        ret = r.s("i")
        pos = r.s("q")
        rid, S, n_reads, n_tails, stale, has_e = r.s("iiiiii")
        site_e = r.s("d")
        nA, nAo, nG, a_un = r.s("iiii")
        a2b = r.a("<i4", 5)
        b2a = r.a("<i4", 5)
        sizeG, sizeR, flags = r.s("iiI")
        gts = r.a("i1", 2 * S)
        depths = r.a("<i4", S)
        fmt_dp = r.a("<i4", S)
        info_dp = r.s("i")
        rs = r.a("<i4", n_reads)
        rb = r.a("u1", n_reads)
        rst = r.a("u1", n_reads)
        rq = r.a("<i4", n_reads)
        raq = r.a("<i4", n_reads)
        re = r.a("<f8", n_reads)
        tails = r.a("<i4", n_tails)
        n_em = r.s("i")
        ems = r.a("<i4", n_em)
        emn = r.a("<i4", n_em)
        n_codes = r.s("i")
        emc = r.a("<u2", n_codes)
        out = {}
        nAf = nA if sizeG else 0
        spec = [(0, "gl", "<f4", sizeG), (1, "pl", "<i4", sizeG), (2, "gp", "<f4", sizeG),
                (3, "qs", "<f4", nAf), (4, "i16", "<f4", 16),
                (5, "fmt_ad", "<i4", sizeR), (6, "fmt_adf", "<i4", sizeR), (7, "fmt_adr", "<i4", sizeR),
                (8, "info_ad", "<i4", nAf), (9, "info_adf", "<i4", nAf), (10, "info_adr", "<i4", nAf)]
        for bit, name, dt, n in spec:
            if flags & (1 << bit):
                out[name] = r.a(dt, n)
        yield SiteDump(ret, pos, rid, S, stale, site_e if has_e else None, nA, nAo, nG, a_un,
                       a2b, b2a, flags, gts, depths, fmt_dp, info_dp, rs, rb, rst, rq, raq, re,
                       tails, ems, emn, emc, out)
