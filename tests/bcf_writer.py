"""VCF text -> uncompressed BCF, for input-path fixtures -- test infrastructure.

Encodes exactly what htslib's VCF parser would (vcf_parse / vcf_parse_format, htslib/vcf.c:2425-3300) for the small subset
the fixtures use: ID, REF/ALT, QUAL, FILTER, Integer / Float / Flag INFO, GT and Integer FORMAT fields.  tools/make_bcf_inputs.py
validates every file it writes by running the UNMODIFIED reference on it: the output must be identical to the run on the VCF."""
import re
import struct
import sys
import os

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import bcf_oracle as bo  # noqa: E402

INT8_END, INT8_MISSING = -127, -128


def parse_header(text):
    """-> (dict name -> id, types {('INFO'|'FORMAT', name): type}, contigs [names], samples)"""
    ids, types, contigs, samples = {"PASS": 0}, {}, [], []
    for line in text.splitlines():
        m = re.match(r"##(FILTER|INFO|FORMAT)=<ID=([^,>]+)(.*)>", line)
        if m:
            kind, name, rest = m.groups()
            if name not in ids:
                ids[name] = max(ids.values()) + 1
            t = re.search(r"Type=(\w+)", rest)
            if t:
                types[(kind, name)] = t.group(1)
        m = re.match(r"##contig=<ID=([^,>]+)", line)
        if m:
            contigs.append(m.group(1))
        if line.startswith("#CHROM"):
            samples = line.split("\t")[9:]
    return ids, types, contigs, samples


def enc_typed_ints(vals):
    return bo.enc_vint(np.asarray(vals, np.int64), -1)


def gt_values(field, ploidy):
    """one sample's GT sub-field -> int values as vcf_parse_format stores them (htslib/vcf.c:2643-2673)"""
    # the phase bit belongs to the allele AFTER the separator; the first allele has none
    vals, ph = [], 0
    toks = re.split(r"([|/])", field)
    for i in range(0, len(toks), 2):
        a = toks[i]
        vals.append(ph if a == "." else ((int(a) + 1) << 1) | ph)
        ph = 1 if i + 1 < len(toks) and toks[i + 1] == "|" else 0
    return vals + [None] * (ploidy - len(vals))      # None = vector_end


def encode_record(line, ids, types, contigs, n_samples, gt_width=1):
    f = line.rstrip("\r\n").split("\t")
    chrom, pos, vid, ref, alt, qual, flt, info, fmt = f[:9]
    cols = f[9:9 + n_samples]
    alleles = [ref] + (alt.split(",") if alt != "." else [])
    shared = bytearray()
    shared += bo.enc_vchar("" if vid == "." else vid)
    for a in alleles:
        shared += bo.enc_vchar(a)
    if flt == ".":
        shared += bo.enc_size(0, bo.BT_NULL)
    else:
        shared += enc_typed_ints([ids[x] for x in flt.split(";")])
    n_info = 0
    if info != ".":
        for kv in info.split(";"):
            k, _, v = kv.partition("=")
            shared += bo.enc_int1(ids[k])
            t = types[("INFO", k)]
            if t == "Flag":
                shared += bo.enc_size(0, bo.BT_NULL)
            elif t == "Integer":
                shared += enc_typed_ints([int(x) for x in v.split(",")])
            elif t == "Float":
                shared += bo.enc_vfloat(np.array([float(x) for x in v.split(",")], np.float32))
            else:
                shared += bo.enc_vchar(v)
            n_info += 1
    indiv = bytearray()
    keys = fmt.split(":")
    subs = [c.split(":") for c in cols]
    for j, k in enumerate(keys):
        vals = [s[j] if j < len(s) else "." for s in subs]
        indiv += bo.enc_int1(ids[k])
        if k == "GT":
            ploidy = max(len(re.split(r"[|/]", v)) for v in vals)
            m = np.array([[INT8_END if x is None else x for x in gt_values(v, ploidy)] for v in vals], np.int64)
            assert m.max() <= 127
            if gt_width == 1:
                indiv += bo.enc_size(ploidy, bo.BT_INT8) + m.astype("<i1").tobytes()
            elif gt_width == 2:   # what htslib writes when an allele index needs 16 bits; vector_end = int16 min + 1
                indiv += bo.enc_size(ploidy, bo.BT_INT16) + np.where(m == INT8_END, -32767, m).astype("<i2").tobytes()
            else:
                indiv += bo.enc_size(ploidy, bo.BT_INT32) + np.where(m == INT8_END, -2147483647, m).astype("<i4").tobytes()
        else:
            assert types[("FORMAT", k)] == "Integer", k
            m = np.array([INT8_MISSING if v == "." else int(v) for v in vals], np.int64)
            assert m.max() <= 127 and (m[m != INT8_MISSING] > -120).all()
            indiv += bo.enc_size(1, bo.BT_INT8) + m.astype("<i1").tobytes()
    qual_bits = 0x7F800001 if qual == "." else struct.unpack("<I", struct.pack("<f", float(qual)))[0]
    rlen = len(ref)
    head = struct.pack("<IIiiiIII", len(shared) + 24, len(indiv), contigs.index(chrom), int(pos) - 1, rlen, qual_bits,
                       (len(alleles) << 16) | n_info, (len(keys) << 24) | n_samples)
    return head + bytes(shared) + bytes(indiv)


def vcf_to_bcf(buf: bytes, gt_width=1):
    """VCF bytes -> (BCF bytes, byte offset of the first record, [record offsets relative to it] + [total])"""
    text = buf.decode()
    body_at = text.index("#CHROM")
    body_at = text.index("\n", body_at) + 1
    header = text[:body_at]
    ids, types, contigs, samples = parse_header(header)
    hdr = header.encode() + b"\0"
    out = bytearray(b"BCF\x02\x02" + struct.pack("<I", len(hdr)) + hdr)
    first = len(out)
    offs = []
    for line in text[body_at:].splitlines():
        if not line:
            continue
        offs.append(len(out) - first)
        out += encode_record(line, ids, types, contigs, len(samples), gt_width)
    offs.append(len(out) - first)
    return bytes(out), first, offs, ids
