"""oracle/bcf_in_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

What the reference does to turn one BCF record into the true genotypes of its hot path: bcf_read / bcf_unpack (fixed fields,
ID, allele strings: htslib/vcf.c:1535-1600), bcf_get_genotypes (the FORMAT block keyed GT as int32 values: missing = (v >> 1) == 0,
allele = (v >> 1) - 1, vector_end padding), then check_rec_alleles (vcfgl.cpp:75-163) exactly as for VCF text.  Status codes
as oracle/vcf_in_oracle.c / include/vgl.h vgl_in_status (smallest code wins).

Parity status: PINNED through two links -- (1) tests/bcf_writer.py's BCF encoding of every fixture VCF is accepted by the
UNMODIFIED reference, which produces from it the same output as from the VCF (tools/make_bcf_inputs.py); (2) on those files
this oracle must equal oracle/vcf_in_oracle.c on the VCF (itself pinned on the reference's captures) record for record
(tests/test_bcfin_oracle.py)."""
import struct

import numpy as np

import bcf_oracle as bo

OK, ENALLELE, EALLELE, ENOGT, ENSAMPLES, EPLOIDY, EALLELEIDX, ESYMBOLIC = 0, 3, 4, 5, 6, 8, 9, 10
_W = {bo.BT_INT8: 1, bo.BT_INT16: 2, bo.BT_INT32: 4}
_END = {1: -127, 2: -32767, 4: -2147483647}
_MISS = {1: -128, 2: -32768, 4: -2147483648}


def allele_code(a: str, gt_source: int) -> int:
    if gt_source == 0:
        return {"0": 0, "1": 1}.get(a[:1], -1)
    if len(a) == 1:
        return "ACGT".find(a)
    return 4 if a in ("<*>", "<NON_REF>") else -1


def record(rec: bytes, S: int, gt_source: int, gt_key: int, rm_invar: int = 0):
    """-> dict(status, skip_code, pos, n_allele, allele_acgt[8], allele_sum, row uint8[S])"""
    r = bo.split_record(rec)
    st = []
    n_allele = r["n_allele"]
    amap = [allele_code(a, gt_source) for a in r["alleles"][:5]]
    if any(c < 0 for c in amap):
        st.append(EALLELE)
    if n_allele > 5 or (gt_source == 0 and n_allele > 2):
        st.append(ENALLELE)
    if r["n_sample"] != S:
        st.append(ENSAMPLES)
    row = np.full(S, 0xFF, np.uint8)
    asum = 0
    gt = [f for f in r["fmts"] if f[0] == gt_key and f[2] in _W]
    if not gt:
        st.append(ENOGT)
    else:
        key, n, t, block = gt[0]
        w = _W[t]
        if n != 2:
            st.append(EPLOIDY)
        elif r["n_sample"] == S:
            o = len(block) - n * w * S
            v = np.frombuffer(block, {1: "<i1", 2: "<i2", 4: "<i4"}[w], 2 * S, o).astype(np.int64).reshape(S, 2)
            for s in range(S):
                b = 0
                for h in range(2):
                    x = int(v[s, h])
                    nib = 0xF
                    if x == _END[w]:
                        st.append(EPLOIDY)
                    elif x == _MISS[w]:
                        st.append(EALLELEIDX)
                    elif (x >> 1) != 0:
                        a = (x >> 1) - 1
                        if a < 0 or a >= n_allele:
                            st.append(EALLELEIDX)
                        else:
                            asum += a
                            m = amap[a] if a < 5 else -1
                            if m == 4:
                                st.append(ESYMBOLIC)
                            elif m >= 0:
                                nib = m
                    b |= nib << (4 * h)
                row[s] = b
    status = min(st) if st else OK
    skip = 0
    if status == OK:
        if (rm_invar & 1) and asum == 0:
            skip = -1
        elif rm_invar & 2:
            for a in range(1, n_allele):
                if a * S * 2 == asum:
                    skip = -2
    acgt = [(amap[i] if i < len(amap) and i < n_allele else -1) for i in range(8)]
    return dict(status=status, skip_code=skip, pos=r["pos"], n_allele=n_allele, allele_acgt=acgt, allele_sum=asum, row=row)
