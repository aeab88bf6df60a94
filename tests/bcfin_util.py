"""Helpers of the BCF-input tests -- test infrastructure (imports oracle/)."""
import gzip
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bcf_in_oracle as bio  # noqa: E402

INPUTS = os.path.join(ROOT, "tests", "golden", "inputs")
MANIFEST = json.load(open(os.path.join(INPUTS, "bcf_inputs.json")))


def load(name):
    """-> (record bytes after the header, offsets [n + 1], manifest entry)"""
    m = MANIFEST[name]
    data = gzip.open(os.path.join(INPUTS, name + ".gz"), "rb").read()
    body = data[m["first_record"]:]
    return body, record_offsets(body), m


def record_offsets(body: bytes):
    """what the host does: hop l_shared + l_indiv + 8 from record to record"""
    off, o = [0], 0
    while o < len(body):
        l_shared, l_indiv = struct.unpack_from("<II", body, o)
        o += 8 + l_shared + l_indiv
        off.append(o)
    assert o == len(body)
    return np.array(off, np.uint32)


def oracle(body, off, S, gt_source, gt_key, rm_invar=0):
    return [bio.record(body[off[i]:off[i + 1]], S, gt_source, gt_key, rm_invar) for i in range(len(off) - 1)]
