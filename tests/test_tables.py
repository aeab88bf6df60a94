"""The product's host-side constant tables (vcfgl_b200/csrc/tables.cpp) against the oracle and, in
the build container, against the reference's own tables -- bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
from vcfgl_b200 import args as vargs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tests", "_tables_shim.so")


@pytest.fixture(scope="module")
def shim():
    srcs = [os.path.join(ROOT, "tests", "tables_shim.cpp"), os.path.join(ROOT, "vcfgl_b200", "csrc", "tables.cpp")]
    if not os.path.exists(SHIM) or os.path.getmtime(SHIM) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-o", SHIM] + srcs)
    L = C.CDLL(SHIM)
    for f in ("shim_lut", "shim_fk", "shim_beta", "shim_lhet"):
        getattr(L, f).restype = C.POINTER(C.c_double)
    L.shim_errmod.restype = C.c_void_p
    L.shim_errmod.argtypes = [C.c_double]
    for f in ("shim_fk", "shim_beta", "shim_lhet", "shim_free"):
        getattr(L, f).argtypes = [C.c_void_p]
    L.shim_fixed_bsum.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.shim_precalc.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_int),
                               C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.shim_inc_beta.restype = C.c_double
    L.shim_inc_beta.argtypes = [C.c_double] * 3
    L.shim_qs_classes.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    return L


def u64(x):
    return np.ascontiguousarray(x).view(np.uint64)


def test_lut_equals_oracle_lut(shim):
    got = np.ctypeslib.as_array(shim.shim_lut(), shape=(771,))
    want = np.ctypeslib.as_array(oracle_lib.lib().vgo_lut_log10_gl(), shape=(771,))
    assert np.array_equal(u64(got), u64(want))
    # SURVEY 8(c) anchor values (shared.cpp:111-113, qs 20)
    assert got[20] == -0.004364805 and got[257 + 20] == -0.303935 and got[514 + 20] == -2.477121


@pytest.mark.parametrize("theta", [0.83, 0.6])
def test_errmod_tables_equal_oracle(shim, theta):
    a = vargs.parse_args(("-d 1 -e 0.01 -GL 1 --gl1-theta %g" % theta).split())
    orc = oracle_lib.Oracle(a, 1)
    t = shim.shim_errmod(1.0 - theta)
    L = oracle_lib.lib()
    for name, n in (("fk", 256), ("beta", 64 * 256 * 256), ("lhet", 256 * 256)):
        got = np.ctypeslib.as_array(getattr(shim, "shim_" + name)(t), shape=(n,))
        want = np.ctypeslib.as_array(getattr(L, "vgo_errmod_" + name)(orc.ctx), shape=(n,))
        assert np.array_equal(u64(got), u64(want)), name
    # fixed-qs running sums: bsum[n][c] = sequential sum_{i<c} fk[i]*beta[q][n][i]
    fk = np.ctypeslib.as_array(shim.shim_fk(t), shape=(256,))
    beta = np.ctypeslib.as_array(shim.shim_beta(t), shape=(64, 256, 256))
    for q in (7, 20, 2, 63):
        out = np.zeros(65536)
        shim.shim_fixed_bsum(t, q, out.ctypes.data)
        out = out.reshape(256, 256)
        qq = min(max(q, 4), 63)
        for n in (1, 2, 17, 255):
            acc = 0.0
            for c in range(n):
                acc = acc + fk[c] * beta[qq, n, c]
                assert out[n, c + 1] == acc
    shim.shim_free(t)


def test_precalc_equals_oracle(shim):
    for argv in ("-d 1 -e 0.2 -GL 1 --adjust-qs 3 -addQS 1", "-d 1 -e 0.01 -GL 2", "-d 1 -e 0.013 -GL 2 --precise-gl 1",
                 "-d 1 -e 0 -GL 2", "-d 1 -e 0 -GL 2 --precise-gl 1", "-d 1 -e 0.000001 -GL 2 --adjust-qs 1"):
        a = vargs.parse_args(argv.split())
        qs, adj, g = oracle_lib.Oracle(a, 1).precalc()
        q1, q2, g3 = C.c_int(), C.c_int(), (C.c_double * 3)()
        assert shim.shim_precalc(a.error_rate, a.error_qs, a.gl_model, a.precise_gl, a.adjust_qs, a.adjust_by,
                                 C.byref(q1), C.byref(q2), g3) == 0
        assert (q1.value, q2.value) == (qs, adj), argv
        assert np.array_equal(u64(np.array(g3[:])), u64(np.array(g))), argv


@pytest.mark.container
@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle/_ref/libref_shared.so")), reason="needs oracle/_ref")
def test_lut_equals_reference_table(shim):
    ref = C.CDLL(os.path.join(ROOT, "oracle/_ref/libref_shared.so"))
    want = np.array((C.c_double * 771).in_dll(ref, "qScore_to_log10_gl")[:])
    got = np.ctypeslib.as_array(shim.shim_lut(), shape=(771,))
    assert np.array_equal(u64(got), u64(want))


def beta_shapes(mean, var):
    a = (((1.0 - mean) / var) - 1.0 / mean) * mean ** 2   # rng.h:370-371
    return a, a * (1.0 / mean - 1.0)


def test_incomplete_beta_matches_scipy(shim):
    from scipy import special
    for mean, var in ((0.01, 1e-5), (0.02, 1e-4), (0.05, 1e-3), (0.3, 0.01)):
        a, b = beta_shapes(mean, var)
        for x in (1e-8, 1e-4, 0.003, 0.01, 0.02, 0.1, 0.5, 0.79, 0.999):
            for aa in (a, a + 1.0):
                got, want = shim.shim_inc_beta(aa, b, x), special.betainc(aa, b, x)
                assert abs(got - want) <= 1e-12 + 1e-10 * want, (mean, var, x, got, want)


@pytest.mark.parametrize("mean,var,shift,bins", [(0.01, 1e-5, 0.0, None), (0.01, 1e-5, 0.499, None), (0.02, 1e-4, 0.0, "rta3"),
                                                  (0.02, 1e-4, 0.0, "rta3_40"), (0.3, 0.02, 0.0, None)])
def test_qs_class_table_is_the_beta_law(shim, mean, var, shift, bins):
    """P(q) from the Beta CDF (scipy); the alias table reproduces the quantised probabilities exactly (integer arithmetic)."""
    from scipy import stats
    a, b = beta_shapes(mean, var)
    lut = np.zeros(256, np.uint8)
    bin_max = -1
    if bins:
        for lo, hi, val in ((0, 2, 2), (3, 14, 12), (15, 30, 23), (31, 40 if bins == "rta3_40" else 63, 37)):
            lut[lo:hi + 1] = val
            bin_max = hi
    words = np.zeros(768, np.uint32)
    prob = np.zeros(257, np.float64)
    n = shim.shim_qs_classes(a, b, shift, int(bins is not None), lut.ctypes.data, bin_max, words.ctypes.data, prob.ctypes.data)
    assert n == 768
    assert abs(prob.sum() - 1.0) < 1e-9
    want = np.zeros(257)
    d = stats.beta(a, b)
    for k in range(0, 421):
        lo = 0.0 if k == 0 else k - shift
        hi = k + 1.0 - shift
        if hi <= 0:
            continue
        lo = max(lo, 0.0)
        p_hi, p_lo = 10 ** (-lo / 10.0), (0.0 if k == 420 else 10 ** (-hi / 10.0))
        q = (256 if k > bin_max else int(lut[k])) if bins else min(k, 63)
        want[q] += d.cdf(p_hi) - d.cdf(p_lo)
    assert np.abs(prob - want).max() < 2e-9, np.abs(prob - want).max()
    if bins == "rta3_40":
        assert prob[256] > 0        # the reference would exit on such a read; the kernel raises VGL_ERANGE
    # the alias table: exact column arithmetic
    alias, thr = words[:256] & 0xFF, (words[:256] >> 8).astype(np.int64)
    info = words[256:512]
    mass = np.zeros(256, np.int64)
    for k in range(256):
        if alias[k] == k:
            mass[k] += 1 << 24
        else:
            mass[k] += thr[k]
            mass[alias[k]] += (1 << 24) - thr[k]
    for c in range(256):
        if mass[c]:
            q, oob = int(info[c] & 0xFF), int((info[c] >> 9) & 1)
            assert mass[c] == round(prob[256 if oob else q] * 2 ** 32), (c, q)
    # words 512..767: the conditional law of the classes other than the heaviest one (the tile kernel draws the number of
    # such reads per cell, then their classes from this table)
    dom = int(np.argmax(mass))
    alias2, thr2 = words[512:] & 0xFF, (words[512:] >> 8).astype(np.int64)
    m2 = np.zeros(256, np.int64)
    for k in range(256):
        if alias2[k] == k:
            m2[k] += 1 << 24
        else:
            m2[k] += thr2[k]
            m2[alias2[k]] += (1 << 24) - thr2[k]
    assert m2.sum() == 2 ** 32
    rest = 2 ** 32 - mass[dom]
    if rest == 0:
        assert m2[dom] == 2 ** 32
    else:
        assert m2[dom] == 0
        for c in range(256):
            if c != dom:
                assert abs(m2[c] / 2 ** 32 - mass[c] / rest) < 1e-9, (c, m2[c], mass[c])


# ---- the prefix codes of the device's BGZF compressor (csrc/bgzf.cu): a stream put together in Python from the tables
# (header bits, literal / length / distance codes, end of block) must inflate with zlib to the bytes it encodes
def _bgzf_encode(code, tokens):
    lit, ln, dist, eob, hdr_bits, hdr = code[:256], code[256:512], code[512:544], int(code[544]), int(code[545]), code[546:]
    acc, n = 0, 0
    for i in range((hdr_bits + 31) // 32):
        acc |= int(hdr[i]) << (32 * i)
    n = hdr_bits
    for t in tokens:
        if isinstance(t, int):
            e = int(lit[t])
            acc |= (e & 0xFFFF) << n
            n += e >> 16
        else:
            L, D = t
            e = int(ln[L - 3])
            acc |= (e & 0xFFFFFF) << n
            n += e >> 24
            if D <= 4:
                dc, deb, dev = D - 1, 0, 0
            else:
                u = D - 1
                hb = u.bit_length() - 1
                deb = hb - 1
                dc = 2 * hb + ((u >> deb) & 1)
                dev = u & ((1 << deb) - 1)
            e = int(dist[dc])
            acc |= (e & 0xFFFF) << n
            n += e >> 16
            acc |= dev << n
            n += deb
    acc |= (eob & 0xFFFF) << n
    n += eob >> 16
    return acc.to_bytes((n + 7) // 8, "little")


@pytest.mark.parametrize("kind", ["fixed", "flat", "skewed", "one_symbol", "huge_counts"])
def test_bgzf_prefix_codes_inflate_with_zlib(shim, kind):
    import zlib
    rng = np.random.default_rng(5)
    hist = np.zeros(320, np.uint32)
    if kind == "skewed":
        hist[:256] = (1e6 * rng.random(256) ** 8).astype(np.uint32)
        hist[0] = 5_000_000
        hist[257:286] = rng.integers(0, 1000, 29)
        hist[288:318] = rng.integers(0, 100000, 30)
    elif kind == "one_symbol":
        hist[7] = 123456
    elif kind == "huge_counts":     # Fibonacci-like counts: the unlimited Huffman tree would be far deeper than 15
        f = [1, 1]
        while len(f) < 46:
            f.append(f[-1] + f[-2])
        hist[:46] = np.minimum(np.array(f), 2**32 - 1)
        hist[288:318] = np.array(f[:30])
    n_words = shim.shim_bgzf_code_words()
    assert n_words == 640
    code = np.zeros(n_words, np.uint32)
    shim.shim_bgzf_code(hist.ctypes.data_as(C.c_void_p), int(kind == "fixed"), code.ctypes.data_as(C.c_void_p))
    assert ((code[:256] >> 16) >= 1).all() and ((code[:256] >> 16) <= 15).all()      # every byte value codable
    assert (code[256:512] >> 24).max() <= 20 and (code[512:542] >> 16).max() <= 15
    if kind == "fixed":
        assert int(code[545]) == 3 and int(code[546]) == 3 and int(code[544]) == 7 << 16
    # a token stream with every literal, every match length and distances over the whole window
    data, tokens = bytearray(), []
    for v in list(range(256)) + [int(x) for x in rng.integers(0, 256, 3000)]:
        data.append(v)
        tokens.append(v)
    for L in list(range(3, 259)) + [int(x) for x in rng.integers(3, 259, 300)]:
        D = int(rng.integers(1, min(len(data), 32768) + 1))
        for k in range(L):
            data.append(data[len(data) - D])
        tokens.append((L, D))
        data.append(L & 0xFF)
        tokens.append(L & 0xFF)
    while len(data) < 40000:
        data.append(0)
        tokens.append(0)
    for D in (1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 13, 16, 17, 24, 25, 32, 33, 48, 49, 64, 65, 96, 97, 128, 129, 192, 193, 256, 257, 384, 385, 512, 513,
              768, 769, 1024, 1025, 1536, 1537, 2048, 2049, 3072, 3073, 4096, 4097, 6144, 6145, 8192, 8193, 12288, 12289, 16384, 16385,
              24576, 24577, 32768):
        for k in range(5):
            data.append(data[len(data) - D])
        tokens.append((5, D))
    z = _bgzf_encode(code, tokens)
    d = zlib.decompressobj(-15)
    out = d.decompress(z)
    assert d.eof and out == bytes(data)
