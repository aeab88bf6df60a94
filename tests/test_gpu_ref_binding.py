"""INTEGRATION.md compiled: oracle/_ref/vcfgl_ref_vgl is the reference's own main(), argument parser, htslib record handling,
add_tags() and writer with the hot path -- simulate_record_values(), vcfgl.cpp:327-1087 -- answered by libvgl.so
(oracle/ref_vgl_binding.h: replay of the reference's own draws; every array handed to htslib is poisoned, then filled from
libvgl's answer, allele string included).  The BCF it writes must equal, byte for byte after the header text, the file the
UNMODIFIED reference wrote for the same command line (tests/golden/bcf, made by tools/make_golden_bcf.py; the reference's
own 17 golden tests of test/runTests.sh are among the cases).  File in -> file out through the reference's host code.
"""
import gzip
import os
import struct
import subprocess

import numpy as np
import pytest

import golden_cases as gc

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "vcfgl_ref_vgl")


def records(buf):
    assert buf[:5] == b"BCF\x02\x02"
    l_text = struct.unpack_from("<I", buf, 5)[0]
    return buf[9 + l_text:]


def with_bcf_output(argv):
    out = list(argv)
    for i, x in enumerate(out[:-1]):
        if x in ("-O", "--output-mode"):
            out[i + 1] = "u"
            return out
    return out + ["-O", "u"]


@pytest.mark.parametrize("cid", gc.CASE_IDS)
def test_reference_host_with_libvgl_hot_path_writes_the_reference_file(cid, tmp_path):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/vcfgl_ref_vgl not built (needs /root/reference at build time)")
    m = gc.MANIFEST[cid]
    src = os.path.join(gc.GOLD, "inputs", m["input"] + ".gz")
    gold = os.path.join(gc.GOLD, "bcf", cid + ".bcf.gz")
    if not (os.path.exists(src) and os.path.exists(gold)):
        pytest.skip("no input / output fixture for this case")
    infile = str(tmp_path / m["input"])
    open(infile, "wb").write(gzip.open(src, "rb").read())
    argv = with_bcf_output(m["argv"])
    if m.get("qs_bins"):
        p = str(tmp_path / "bins.csv")
        open(p, "w").write("".join("%d,%d,%d\n" % tuple(b) for b in m["qs_bins"]))
        argv += ["--qs-bins", p]
    if m.get("depths"):
        p = str(tmp_path / "depths.txt")
        open(p, "w").write("".join("%r\n" % d for d in m["depths"]))
        argv += ["--depths-file", p]
    pref = str(tmp_path / "out")
    r = subprocess.run([EXE, "-i", infile, "-o", pref] + argv, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    tail = [l for l in r.stderr.splitlines() if l.startswith("[vcfgl_ref_vgl]")]
    assert tail and "sites through libvgl" in tail[-1], r.stderr[-500:]
    n_values = int(tail[-1].split("replay), ")[1].split()[0])
    got = records(open(pref + ".bcf", "rb").read())
    want = records(gzip.open(gold, "rb").read())
    assert len(got) == len(want) and len(got) > 0
    a = gc.case_args(cid)
    if a.precise_gl and a.error_qs == 2:
        # --precise-gl 1 takes log10 of per-read error probabilities: CUDA's log10 and glibc's differ in the last place for some
        # arguments (GL within 1e-6 relative), so a few float bytes may differ; everything else is identical
        g, w = np.frombuffer(got, np.uint8), np.frombuffer(want, np.uint8)
        assert (g != w).mean() < 0.01
    else:
        assert got == want, cid
    assert n_values > 0 or len(want) < 200
