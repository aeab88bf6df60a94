"""Shared access to tests/golden (manifest + dumps) -- test infrastructure."""
import json
import os

from vcfgl_b200 import args as vargs

import vgl_dump

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))
CASE_IDS = sorted(MANIFEST)


def case_args(cid) -> vargs.SimArgs:
    m = MANIFEST[cid]
    # file arguments (--qs-bins, --depths-file) were inlined by tools/make_golden.py
    return vargs.parse_args(m["argv"], qs_bins=m.get("qs_bins"), depths=m.get("depths"))


_cache = {}


def case_sites(cid):
    if cid not in _cache:
        _cache[cid] = vgl_dump.read_dump(os.path.join(GOLD, cid + ".vgld.gz"))
    return _cache[cid]


# ---- tests/golden/fuzz: captures of seeded random configurations (tools/make_golden_fuzz.py, tests/fuzz_cases.py)
FUZZ = os.path.join(GOLD, "fuzz")
FUZZ_MANIFEST = json.load(open(os.path.join(FUZZ, "manifest.json"))) if os.path.exists(os.path.join(FUZZ, "manifest.json")) else {}
FUZZ_IDS = sorted(FUZZ_MANIFEST)


def fuzz_args(cid) -> vargs.SimArgs:
    m = FUZZ_MANIFEST[cid]
    return vargs.parse_args(m["argv"], qs_bins=m.get("qs_bins"), depths=m.get("depths"))


def fuzz_sites(cid):
    if cid not in _cache:
        _cache[cid] = vgl_dump.read_dump(os.path.join(FUZZ, cid + ".vgld.gz"))
    return _cache[cid]
