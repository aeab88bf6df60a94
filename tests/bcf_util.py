"""Helpers of the BCF output-path tests -- test infrastructure (imports oracle/)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bcf_oracle as bo  # noqa: E402

import golden_cases as gc  # noqa: E402
from vcfgl_b200 import args as vargs  # noqa: E402

BCF_DIR = os.path.join(gc.GOLD, "bcf")

# cases the device serialiser does not take: -doGVCF merges records in an order-dependent host state machine
# (bcf_utils.cpp:662-942), SURVEY.md 8(f) row 3
def is_gvcf(cid):
    return gc.case_args(cid).do_gvcf != 0


BCF_CASES = [c for c in gc.CASE_IDS if not is_gvcf(c)]


def reference_bcf(cid):
    return bo.read_bcf(os.path.join(gc.FUZZ if cid in gc.FUZZ_MANIFEST else BCF_DIR, cid + ".bcf.gz"))


def any_args(cid):
    return gc.fuzz_args(cid) if cid in gc.FUZZ_MANIFEST else gc.case_args(cid)


def any_sites(cid):
    return gc.fuzz_sites(cid) if cid in gc.FUZZ_MANIFEST else gc.case_sites(cid)


def enabled_tags(a):
    """(FORMAT tags, INFO tags) add_tags() writes for these arguments, each in update order"""
    m = a.tag_mask
    f = [t for t, bit in (("DP", vargs.TAG_FMT_DP), ("GL", vargs.TAG_GL), ("PL", vargs.TAG_PL), ("GP", vargs.TAG_GP),
                          ("AD", vargs.TAG_FMT_AD), ("ADF", vargs.TAG_FMT_ADF), ("ADR", vargs.TAG_FMT_ADR)) if m & bit]
    i = [t for t, bit in (("DP", vargs.TAG_INFO_DP), ("QS", vargs.TAG_QS), ("I16", vargs.TAG_I16), ("AD", vargs.TAG_INFO_AD),
                          ("ADF", vargs.TAG_INFO_ADF), ("ADR", vargs.TAG_INFO_ADR)) if m & bit]
    return f, i


def site_arrays(a, d):
    """the arrays add_tags() passes to htslib for one captured site (tests/vgl_dump.SiteDump)"""
    ftags, itags = enabled_tags(a)
    S, A = d.S, d.n_alleles
    fmt, info = {}, {}
    for t in ftags:
        if t == "DP":
            fmt[t] = d.fmt_dp
        elif t in ("GL", "PL", "GP"):
            fmt[t] = d.out[t.lower()]
        else:
            fmt[t] = d.out["fmt_" + t.lower()]
    for t in itags:
        if t == "DP":
            info[t] = np.array([d.info_dp], np.int32)
        elif t == "QS":
            info[t] = d.out["qs"] if d.out["qs"].size else np.zeros(A, np.float32)
        elif t == "I16":
            info[t] = d.out["i16"]
        else:
            v = d.out["info_" + t.lower()]
            info[t] = v if v.size else np.zeros(A, np.int32)
    return fmt, info
