"""oracle/bcf_oracle.py -- TEST INFRASTRUCTURE, not product code (only tests/, smoke() and bench.py's
reference legs may import it).

CPU restatement of the reference's OUTPUT path for one simulated record: what simRecord::add_tags()
(bcf_utils.cpp:426-507) hands to htslib and what htslib 1.15.1 then writes for `-O u`:

    bcf_update_format / bcf_update_info     htslib/vcf.c:4418-4470, 4256-4390
    bcf_enc_vint / vfloat / vchar           htslib/vcf.c:2249-2294, 2337-2350
    bcf_enc_size / bcf_enc_int1             htslib/htslib/vcf.h:1392-1446
    bcf1_sync (ID, alleles, FILTER, INFO, FORMAT order; removed GT)   htslib/vcf.c:1773-1917
    bcf_write (the 32 fixed bytes)          htslib/vcf.c:1951-2001
    _bcf1_sync_alleles (rlen)               htslib/vcf.c:4583-4614

Parity PINNED: tests/test_bcf_oracle.py rebuilds every record of tests/golden/bcf/<id>.bcf.gz (written by
the unmodified reference binary, tools/make_golden_bcf.py) from the replay capture of the same run and
requires identical bytes.
"""
import gzip
import re
import struct
import zlib

import numpy as np

BT_NULL, BT_INT8, BT_INT16, BT_INT32, BT_FLOAT, BT_CHAR = 0, 1, 2, 3, 5, 7
INT32_MISSING = -2147483648      # htslib/vcf.h:1325
INT32_VECTOR_END = -2147483647
MAX_INT8, MIN_INT8 = 127, -120   # htslib/vcf.h BCF_MAX_BT_INT8 / BCF_MIN_BT_INT8
MAX_INT16, MIN_INT16 = 32767, -32760

# order in which add_tags() updates the record (bcf_utils.cpp:426-507)
FORMAT_ORDER = ["DP", "GL", "PL", "GP", "AD", "ADF", "ADR"]
INFO_ORDER = ["DP", "QS", "I16", "AD", "ADF", "ADR"]


# ---------------------------------------------------------------- encoders (htslib/vcf.h:1392-1446)
def enc_size(size, typ):
    if size >= 15:
        out = bytes([15 << 4 | typ])
        if size >= 128:
            if size >= 32768:
                return out + bytes([1 << 4 | BT_INT32]) + struct.pack("<i", size)
            return out + bytes([1 << 4 | BT_INT16]) + struct.pack("<h", size)
        return out + bytes([1 << 4 | BT_INT8, size])
    return bytes([size << 4 | typ])


def enc_int1(x):
    if x == INT32_VECTOR_END:
        return enc_size(1, BT_INT8) + b"\x81"
    if x == INT32_MISSING:
        return enc_size(1, BT_INT8) + b"\x80"
    if MIN_INT8 <= x <= MAX_INT8:
        return enc_size(1, BT_INT8) + struct.pack("<b", x)
    if MIN_INT16 <= x <= MAX_INT16:
        return enc_size(1, BT_INT16) + struct.pack("<h", x)
    return enc_size(1, BT_INT32) + struct.pack("<i", x)


def enc_vint(a, wsize=-1):
    """bcf_enc_vint(s, n, a, wsize), htslib/vcf.c:2249-2294"""
    a = np.asarray(a, dtype=np.int64)
    n = a.size
    if n <= 0:
        return enc_size(0, BT_NULL)
    if n == 1:
        return enc_int1(int(a[0]))
    if wsize <= 0:
        wsize = n
    real = a[(a != INT32_MISSING) & (a != INT32_VECTOR_END)]
    mx = int(real.max()) if real.size else -2147483648
    mn = int(real.min()) if real.size else 2147483647
    if mx <= MAX_INT8 and mn >= MIN_INT8:
        v = a.copy()
        v[a == INT32_VECTOR_END] = -127
        v[a == INT32_MISSING] = -128
        return enc_size(wsize, BT_INT8) + v.astype("<i1").tobytes()
    if mx <= MAX_INT16 and mn >= MIN_INT16:
        v = a.copy()
        v[a == INT32_VECTOR_END] = -32767
        v[a == INT32_MISSING] = -32768
        return enc_size(wsize, BT_INT16) + v.astype("<i2").tobytes()
    return enc_size(wsize, BT_INT32) + a.astype("<i4").tobytes()


def enc_vfloat(a):
    a = np.ascontiguousarray(a, dtype="<f4")
    return enc_size(a.size, BT_FLOAT) + a.tobytes()


def enc_vchar(s):
    b = s if isinstance(s, bytes) else s.encode()
    return enc_size(len(b), BT_CHAR) + b


# ---------------------------------------------------------------- one record
def alleles_of_site(n_alleles, alleles2acgt, info_dp, do_unobserved, do_gvcf):
    """allele list as passed to bcf_update_alleles_str (vcfgl.cpp:739-782; no-reads sites :228-277)"""
    nonref = "<NON_REF>" if do_unobserved in (2, 5) else "<*>"
    if info_dp == 0:
        if do_gvcf:
            return ["<NON_REF>"]
        return {0: ["."], 1: ["<*>"], 2: ["<NON_REF>"], 3: list("ACGT"), 4: list("ACGT") + ["<*>"],
                5: list("ACGT") + ["<NON_REF>"]}[do_unobserved]
    out = []
    for a in range(n_alleles):
        b = int(alleles2acgt[a])
        out.append(nonref if b == 4 else "ACGT"[b])
    return out


def encode_record(rid, pos, qual_bits, id_bytes, filter_info_bytes, n_info_in, alleles, n_samples, dict_ids, fmt, info, in_fmt=()):
    """One BCF record exactly as bcf_write() emits it after the reference's edits.

    id_bytes / filter_info_bytes: the input record's typed ID string and its FILTER vector followed by its own INFO
    pairs -- bytes bcf1_sync copies unchanged (vcf.c:1802-1838).  fmt / info: dicts tag -> array for the enabled
    tags (FORMAT arrays hold n_samples * k values).  dict_ids: {"FORMAT/GL": id, "INFO/DP": id, ...}.
    in_fmt: the FORMAT blocks the INPUT record carried besides GT, in its order, as (dictionary id, block bytes): the
    reference removes only GT (bcf_update_genotypes(NULL), vcfgl.cpp:793); bcf_update_format replaces a block whose key a
    simulated tag has IN PLACE (vcf.c: the fmt slot is kept) and appends the other simulated tags behind the input's.
    """
    shared = bytearray(id_bytes)
    for al in alleles:
        shared += enc_vchar(al)
    shared += filter_info_bytes
    n_info = n_info_in
    for tag in INFO_ORDER:
        if tag not in info:
            continue
        v = info[tag]
        shared += enc_int1(dict_ids["INFO/" + tag])
        shared += enc_vfloat(v) if tag in ("QS", "I16") else enc_vint(v, -1)
        n_info += 1
    indiv = bytearray()
    n_fmt = 0

    def put(tag):
        v = np.asarray(fmt[tag])
        nps = v.size // n_samples
        assert nps * n_samples == v.size and nps > 0
        out = enc_int1(dict_ids["FORMAT/" + tag])
        if tag in ("GL", "GP"):
            return out + enc_size(nps, BT_FLOAT) + np.ascontiguousarray(v, dtype="<f4").tobytes()
        return out + enc_vint(v, nps)

    by_id = {dict_ids["FORMAT/" + tag]: tag for tag in FORMAT_ORDER if tag in fmt}
    placed = set()
    for key, block in in_fmt:
        if key in by_id:            # the simulated tag takes the input block's slot
            indiv += put(by_id[key])
            placed.add(by_id[key])
        else:
            indiv += block
        n_fmt += 1
    for tag in FORMAT_ORDER:
        if tag not in fmt or tag in placed:
            continue
        indiv += put(tag)
        n_fmt += 1
    rlen = len(alleles[0]) if alleles else 0  # _bcf1_sync_alleles, no INFO/END on these records
    head = struct.pack("<IIiiiIHHI", len(shared) + 24, len(indiv), rid, pos, rlen, qual_bits, n_info, len(alleles),
                       (n_fmt << 24) | (n_samples & 0xFFFFFF))
    return bytes(head) + bytes(shared) + bytes(indiv)


# ---------------------------------------------------------------- reading the reference's files
def bgzf_decompress(raw):
    out = bytearray()
    while raw:
        d = zlib.decompressobj(31)
        out += d.decompress(raw)
        raw = d.unused_data
    return bytes(out)


def read_bcf(path):
    """-> (header_text, dict_ids {"FORMAT/GL": id, "INFO/DP": id, "FILTER/PASS": 0 ...}, [record bytes])"""
    raw = open(path, "rb").read()
    if path.endswith(".gz"):
        raw = gzip.decompress(raw)  # the fixture's outer wrapper
    data = bgzf_decompress(raw) if raw[:2] == b"\x1f\x8b" else raw  # `-O u` is written without BGZF framing, `-O b` with
    assert data[:5] == b"BCF\x02\x02", data[:5]
    (l_text,) = struct.unpack_from("<I", data, 5)
    text = data[9:9 + l_text].rstrip(b"\0").decode()
    ids, nxt = {}, {}
    order = {"PASS": 0}
    for line in text.splitlines():
        m = re.match(r"##(FILTER|INFO|FORMAT)=<ID=([^,>]+)", line)
        if not m:
            continue
        kind, name = m.groups()
        idx = re.search(r"IDX=(\d+)", line)
        if idx:
            order[name] = int(idx.group(1))
        elif name not in order:
            order[name] = max(order.values()) + 1
        ids[kind + "/" + name] = order[name]
    recs = []
    o = 9 + l_text
    while o < len(data):
        l_shared, l_indiv = struct.unpack_from("<II", data, o)
        n = 8 + l_shared + l_indiv
        recs.append(data[o:o + n])
        o += n
    return text, ids, recs


def _typed_size(buf, o):
    """descriptor at o -> (n, type, offset after the descriptor)"""
    b = buf[o]
    n, t = b >> 4, b & 0xF
    o += 1
    if n == 15:
        t2 = buf[o] & 0xF
        o += 1
        n = struct.unpack_from({BT_INT8: "<b", BT_INT16: "<h", BT_INT32: "<i"}[t2], buf, o)[0]
        o += {BT_INT8: 1, BT_INT16: 2, BT_INT32: 4}[t2]
    return n, t, o


_WIDTH = {BT_NULL: 0, BT_INT8: 1, BT_INT16: 2, BT_INT32: 4, BT_FLOAT: 4, BT_CHAR: 1}


def split_record(rec):
    """the pieces of a record: fixed fields, the pass-through bytes, allele strings, INFO pairs, FORMAT blocks"""
    l_shared, l_indiv, rid, pos, rlen, qual_bits, n_info, n_allele, ns = struct.unpack_from("<IIiiiIHHI", rec, 0)
    n_sample, n_fmt = ns & 0xFFFFFF, ns >> 24
    o = 32
    n, t, o2 = _typed_size(rec, o)
    id_bytes = rec[o:o2 + n]
    o = o2 + n
    alleles = []
    for _ in range(n_allele):
        n, t, o2 = _typed_size(rec, o)
        alleles.append(rec[o2:o2 + n].decode())
        o = o2 + n
    flt0 = o
    n, t, o2 = _typed_size(rec, o)
    o = o2 + n * _WIDTH[t]
    filter_bytes = rec[flt0:o]
    infos = []
    for _ in range(n_info):
        k0 = o
        n, t, o2 = _typed_size(rec, o)          # key: typed int
        key = int.from_bytes(rec[o2:o2 + _WIDTH[t]], "little", signed=True)
        o = o2 + _WIDTH[t]
        n, t, o2 = _typed_size(rec, o)
        o = o2 + n * _WIDTH[t]
        infos.append((key, rec[k0:o]))
    assert o == 8 + l_shared, (o, l_shared)
    fmts = []
    end = 8 + l_shared + l_indiv
    while o < end:
        k0 = o
        n, t, o2 = _typed_size(rec, o)
        key = int.from_bytes(rec[o2:o2 + _WIDTH[t]], "little", signed=True)
        o = o2 + _WIDTH[t]
        n, t, o2 = _typed_size(rec, o)
        o = o2 + n * _WIDTH[t] * n_sample
        fmts.append((key, n, t, rec[k0:o]))
    assert o == end and len(fmts) == n_fmt
    return dict(rid=rid, pos=pos, rlen=rlen, qual_bits=qual_bits, n_info=n_info, n_allele=n_allele, n_sample=n_sample,
                n_fmt=n_fmt, id_bytes=id_bytes, alleles=alleles, filter_bytes=filter_bytes, infos=infos, fmts=fmts)
