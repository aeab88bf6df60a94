"""Pins the CPU oracle (oracle/vgl_oracle.c) bit-for-bit against the reference:
every golden dump holds the reference's own draws and its bit-exact outputs."""
import ctypes as C
import os

import numpy as np
import pytest

import golden_cases as gc
import oracle_lib

OUT_KEYS = ["gl", "pl", "gp", "qs", "i16", "fmt_ad", "fmt_adf", "fmt_adr", "info_ad", "info_adf", "info_adr"]


def bits(x):
    x = np.ascontiguousarray(x)
    return x.view(np.uint32) if x.dtype == np.float32 else x


@pytest.mark.parametrize("cid", gc.CASE_IDS)
def test_oracle_reproduces_reference_dump(cid):
    check_oracle_on_dump(cid, gc.case_args(cid), gc.case_sites(cid))


@pytest.mark.parametrize("cid", gc.FUZZ_IDS)
def test_oracle_reproduces_fuzz_dump(cid):
    """32 seeded random configurations captured from the reference (tests/golden/fuzz, tools/make_golden_fuzz.py)"""
    check_oracle_on_dump(cid, gc.fuzz_args(cid), gc.fuzz_sites(cid), need_values=False)


def check_oracle_on_dump(cid, a, sites, need_values=True):
    assert sites, "empty dump"
    orc = oracle_lib.Oracle(a, sites[0].S)
    n_checked = 0
    for k, d in enumerate(sites):
        o = orc.site_from_dump(d)
        assert o["ret"] == d.ret, (cid, k, o["ret"], d.ret)
        assert np.array_equal(o["fmt_dp"], d.fmt_dp)
        assert o["info_dp"] == d.info_dp
        if d.ret != 0 or not d.out:
            continue
        assert (o["n_alleles"], o["n_alleles_observed"], o["n_genotypes"]) == \
            (d.n_alleles, d.n_alleles_observed, d.n_genotypes), (cid, k)
        if d.info_dp > 0:
            assert np.array_equal(o["alleles2acgt"], d.alleles2acgt), (cid, k)
            assert np.array_equal(o["acgt2alleles"], d.acgt2alleles), (cid, k)
        for key in OUT_KEYS:
            if key in d.out:
                got, want = bits(o[key]), bits(d.out[key])
                assert got.shape == want.shape, (cid, k, key)
                assert np.array_equal(got, want), (cid, k, key, o[key], d.out[key])
                n_checked += got.size
    assert n_checked > 0 or not need_values


def test_precalc_anchor_values():
    """SURVEY 8(c) anchors: e=0.2 -> qs 6, adjusted 7 (test1); e=0.01 -> qs 20 and the model-2
    constants of shared.cpp:111-113."""
    from vcfgl_b200 import args as vargs
    a = vargs.parse_args("-d 1 -e 0.2 -GL 1 --adjust-qs 3 -addQS 1".split())
    assert oracle_lib.Oracle(a, 1).precalc()[:2] == (6, 7)
    a = vargs.parse_args("-d 1 -e 0.01 -GL 2".split())
    qs, adj, g = oracle_lib.Oracle(a, 1).precalc()
    assert qs == 20 and adj == -1
    assert g == [-0.004364805, -0.303935, -2.477121]


def test_errmod_count_kats():
    """count -> GL/PL known answers from test/reference/test1/test1.vcf (qs 7, theta 0.83)."""
    from vcfgl_b200 import args as vargs
    a = vargs.parse_args("-d 1 -e 0.2 -GL 1 --adjust-qs 3 -addQS 1 -addPL 1 -doUnobserved 1".split())
    orc = oracle_lib.Oracle(a, 1)
    kats = [((1, 0, 0, 0), "0,-0.301034,-0.7", [0, 3, 7]),
            ((2, 0, 0, 0), "0,-0.602068,-1.24246", [0, 6, 12]),
            ((1, 1, 0, 0), "-0.143579,0,-0.143579,-0.444613,-0.444613,-0.588193", [1, 0, 1, 4, 4, 6]),
            ((1, 1, 0, 1), "-0.0113571,0,-0.0113571,0,0,-0.0113571,-0.312391,-0.312391,-0.312391,-0.323748",
             [0, 0, 0, 0, 0, 0, 3, 3, 3, 3])]
    for counts, gl_txt, pl in kats:
        bases = np.repeat(np.arange(4), counts).astype(np.uint8)
        o = orc.site(np.array([0, 0], np.int8), np.array([len(bases)], np.int32), bases)
        assert ",".join("%g" % x for x in o["gl"]) == gl_txt
        assert list(o["pl"]) == pl


@pytest.mark.container
@pytest.mark.skipif(not os.path.exists(os.path.join(oracle_lib.ORACLE_DIR, "_ref/libref_shared.so")),
                    reason="oracle/_ref not built (needs /root/reference)")
def test_lut_matches_reference():
    ref = C.CDLL(os.path.join(oracle_lib.ORACLE_DIR, "_ref/libref_shared.so"))
    want = np.array((C.c_double * (3 * 257)).in_dll(ref, "qScore_to_log10_gl")[:])
    got = np.ctypeslib.as_array(oracle_lib.lib().vgo_lut_log10_gl(), shape=(3 * 257,))
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


@pytest.mark.container
@pytest.mark.skipif(not os.path.exists(os.path.join(oracle_lib.ORACLE_DIR, "_ref/libref_errmod.so")),
                    reason="oracle/_ref not built (needs /root/reference)")
def test_errmod_tables_match_htslib():
    """our cal_coef restatement vs htslib's errmod_init, table by table, bit by bit"""
    from vcfgl_b200 import args as vargs
    ref = C.CDLL(os.path.join(oracle_lib.ORACLE_DIR, "_ref/libref_errmod.so"))

    class Em(C.Structure):
        _fields_ = [("depcorr", C.c_double), ("fk", C.POINTER(C.c_double)),
                    ("beta", C.POINTER(C.c_double)), ("lhet", C.POINTER(C.c_double))]
    ref.errmod_init.restype = C.POINTER(Em)
    ref.errmod_init.argtypes = [C.c_double]
    ref.errmod_cal.argtypes = [C.POINTER(Em), C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    for theta in (0.83, 0.5):
        a = vargs.parse_args(("-d 1 -e 0.01 -GL 1 --gl1-theta %g" % theta).split())
        orc = oracle_lib.Oracle(a, 1)
        em = ref.errmod_init(1.0 - theta)
        L = oracle_lib.lib()
        for name, n in (("fk", 256), ("beta", 64 * 256 * 256), ("lhet", 256 * 256)):
            want = np.ctypeslib.as_array(getattr(em.contents, name), shape=(n,))
            got = np.ctypeslib.as_array(getattr(L, "vgo_errmod_" + name)(orc.ctx), shape=(n,))
            assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), name
        rng = np.random.default_rng(5)
        for _ in range(300):
            n = int(rng.integers(1, 256))
            codes = (rng.integers(0, 64, n) << 5 | rng.integers(0, 4, n)).astype(np.uint16)
            q1 = np.zeros(25, np.float32)
            q2 = np.zeros(25, np.float32)
            c2 = codes.copy()
            L.vgo_errmod_cal(orc.ctx, n, codes.ctypes.data, q1.ctypes.data)
            ref.errmod_cal(em, n, 5, c2.ctypes.data, q2.ctypes.data)
            assert np.array_equal(q1.view(np.uint32), q2.view(np.uint32))
