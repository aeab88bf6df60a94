"""Pins oracle/bcf_oracle.py (CPU restatement of the reference's BCF output path) against the bytes the
UNMODIFIED reference binary wrote with `-O u` for every non-gVCF golden case (tests/golden/bcf/, written by
tools/make_golden_bcf.py): each record is rebuilt from the replay capture of the same run and must be identical."""
import numpy as np
import pytest

import bcf_util as bu
import golden_cases as gc

bo = bu.bo


@pytest.mark.parametrize("cid", bu.BCF_CASES + gc.FUZZ_IDS)
def test_oracle_rebuilds_reference_records(cid):
    a = bu.any_args(cid)
    text, ids, recs = bu.reference_bcf(cid)
    kept = [d for d in bu.any_sites(cid) if d.ret == 0]
    assert len(kept) == len(recs) and (len(recs) > 0 or cid in gc.FUZZ_MANIFEST)
    ftags, itags = bu.enabled_tags(a)
    for d, rec in zip(kept, recs):
        r = bo.split_record(rec)
        assert (r["rid"], r["pos"], r["n_sample"]) == (d.rid, d.pos, d.S)
        assert r["n_fmt"] == len(ftags)            # the input's GT is gone (vcfgl.cpp:793)
        n_in = r["n_info"] - len(itags)            # INFO fields the input record carried
        assert n_in >= 0
        passthrough = r["filter_bytes"] + b"".join(b for _, b in r["infos"][:n_in])
        alleles = bo.alleles_of_site(d.n_alleles, d.alleles2acgt, d.info_dp, a.do_unobserved, a.do_gvcf)
        assert alleles == r["alleles"]
        fmt, info = bu.site_arrays(a, d)
        got = bo.encode_record(d.rid, d.pos, r["qual_bits"], r["id_bytes"], passthrough, n_in, alleles, d.S, ids, fmt, info)
        assert got == rec, (cid, d.pos)


def test_encoders_match_htslib_rules():
    # htslib/vcf.h:1392-1446, vcf.c:2249-2294
    assert bo.enc_size(3, bo.BT_FLOAT) == bytes([0x35])
    assert bo.enc_size(15, bo.BT_FLOAT) == bytes([0xF5, 0x11, 15])
    assert bo.enc_size(200, bo.BT_INT8) == bytes([0xF1, 0x12, 200, 0])
    assert bo.enc_int1(5) == bytes([0x11, 5]) and bo.enc_int1(128) == bytes([0x12, 128, 0])
    assert bo.enc_int1(bo.INT32_MISSING) == bytes([0x11, 0x80])
    assert bo.enc_vint([]) == bytes([0x00])
    assert bo.enc_vint([0, 127, bo.INT32_MISSING], 3) == bytes([0x31, 0, 127, 0x80])
    assert bo.enc_vint([0, 128], 2) == bytes([0x22, 0, 0, 128, 0])
    assert bo.enc_vint([bo.INT32_MISSING] * 2, 1) == bytes([0x11, 0x80, 0x80])   # all missing -> int8
    assert bo.enc_vint([-121, 0], 2)[0] == 0x22                                   # -121 is reserved in int8
    assert bo.enc_vint([40000, 1], 2)[0] == 0x23
    assert bo.enc_vfloat(np.array([1.0], np.float32)) == bytes([0x15, 0, 0, 0x80, 0x3F])
    assert bo.enc_vchar("<*>") == b"\x37<*>" and bo.enc_vchar("") == b"\x07"


@pytest.mark.parametrize("cid", bu.BCF_CASES[:6])
def test_hts_reader_tool_sees_the_reference_records(cid, tmp_path):
    """oracle/_ref/hts_read_bcf (oracle/hts_read_bcf.c on the reference's htslib; the GPU tests use it to read device-compressed
    BGZF files) against the oracle's own parse of the reference's -O u files"""
    import gzip
    import os
    import struct
    import subprocess
    reader = os.path.join(bu.ROOT, "oracle", "_ref", "hts_read_bcf")
    if not os.path.exists(reader):
        pytest.skip("oracle/_ref not built")
    path = str(tmp_path / "x.bcf")
    src = os.path.join(bu.BCF_DIR, cid + ".bcf.gz")
    with open(path, "wb") as fh:
        fh.write(gzip.open(src, "rb").read())
    lines = subprocess.check_output([reader, path]).decode().splitlines()
    _, _, recs = bu.reference_bcf(cid)
    assert lines[0].split()[:4] == ["format", "9", "compression", "0"] and lines[-1] == "records %d" % len(recs)
    for line, rec in zip(lines[1:-1], recs):
        l_shared, l_indiv, rid, pos = struct.unpack_from("<IIii", rec, 0)
        f = line.split()
        assert (int(f[0]), int(f[1])) == (rid, pos)
        assert (int(f[7]), int(f[8])) == (l_shared - 24, l_indiv)      # htslib keeps the 24 fixed bytes out of shared.s
