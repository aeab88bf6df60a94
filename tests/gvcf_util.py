"""Helpers of the gVCF block-merger tests -- test infrastructure (imports oracle/)."""
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bcf_oracle as bo  # noqa: E402
import gvcf_oracle as go  # noqa: E402

import golden_cases as gc  # noqa: E402
import vgl_dump  # noqa: E402
from vcfgl_b200 import args as vargs  # noqa: E402

GVCF_DIR = os.path.join(gc.GOLD, "gvcf")
GVCF_MANIFEST = json.load(open(os.path.join(GVCF_DIR, "manifest.json")))
MAIN_GVCF = [c for c in gc.CASE_IDS if gc.case_args(c).do_gvcf]
CASES = sorted(GVCF_MANIFEST) + MAIN_GVCF


def load(cid):
    """-> (SimArgs, written sites of the capture, (header text, ids, records) of the reference's BCF)"""
    if cid in GVCF_MANIFEST:
        a = vargs.parse_args(list(GVCF_MANIFEST[cid]["argv"]))
        sites = vgl_dump.read_dump(os.path.join(GVCF_DIR, cid + ".vgld.gz"))
        bcf = bo.read_bcf(os.path.join(GVCF_DIR, cid + ".bcf.gz"))
    else:
        a = gc.case_args(cid)
        sites = gc.case_sites(cid)
        bcf = bo.read_bcf(os.path.join(gc.GOLD, "bcf", cid + ".bcf.gz"))
    return a, [d for d in sites if d.ret == 0], bcf


def dps_of(a):
    return [int(x) for x in a.gvcf_dps.split(",")]


def oracle_input(kept):
    return [dict(rid=d.rid, pos=d.pos, n_alleles_observed=d.n_alleles_observed, fmt_dp=d.fmt_dp,
                 pl=d.out["pl"] if d.out.get("pl") is not None and d.out["pl"].size else None) for d in kept]


_NP = {bo.BT_INT8: "<i1", bo.BT_INT16: "<i2", bo.BT_INT32: "<i4", bo.BT_FLOAT: "<f4"}
_MISS = {bo.BT_INT8: -128, bo.BT_INT16: -32768, bo.BT_INT32: -2 ** 31}


def typed_values(buf):
    """decode `typed key + typed vector` bytes (an INFO pair) -> numpy array (int missing -> INT32_MIN)"""
    n, t, o = bo._typed_size(buf, 0)
    o += bo._WIDTH[t]
    n, t, o = bo._typed_size(buf, o)
    v = np.frombuffer(buf, _NP[t], n, o)
    if t in _MISS:
        v = np.where(v == _MISS[t], -2 ** 31, v.astype(np.int64)).astype(np.int64)
    return v


def fmt_values(block, n_sample):
    n, t, o = bo._typed_size(block, 0)
    o += bo._WIDTH[t]
    n, t, o = bo._typed_size(block, o)
    v = np.frombuffer(block, _NP[t], n * n_sample, o)
    if t in _MISS:
        v = np.where(v == _MISS[t], -2 ** 31, v.astype(np.int64)).astype(np.int64)
    return v


def decode(rec, ids):
    """reference BCF record -> dict(rid, pos, rlen, alleles, info{name: values}, fmt{name: values})"""
    r = bo.split_record(rec)
    inames = {v: k.split("/", 1)[1] for k, v in ids.items() if k.startswith("INFO/")}
    fnames = {v: k.split("/", 1)[1] for k, v in ids.items() if k.startswith("FORMAT/")}
    info = {inames[k]: typed_values(b) for k, b in r["infos"]}
    fmt = {fnames[k]: fmt_values(b, r["n_sample"]) for k, n, t, b in r["fmts"]}
    return dict(rid=r["rid"], pos=r["pos"], rlen=r["rlen"], alleles=r["alleles"], info=info, fmt=fmt)
