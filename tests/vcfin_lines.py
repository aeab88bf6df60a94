"""Single-record cases of the input path shared by the CPU tests, the GPU tests and tools/make_line_cases.py."""
from vcfgl_b200 import capi

BAD = [
    # (record line, S, source, expected status)
    (b"1\t5\t.\t0\t1\t.\tPASS\t.", 2, 0, capi.IN_ENCOLS),
    (b"1\t5\t.\t0\t1\t.\tPASS\t.\tGT", 2, 0, capi.IN_ENCOLS),
    (b"", 2, 0, capi.IN_ENCOLS),
    (b"1\t99999999999\t.\t0\t1\t.\tPASS\t.\tGT\t0|0\t0|1", 2, 0, capi.IN_EPOS),
    (b"1\t5\t.\t0\t1,1\t.\tPASS\t.\tGT\t0|0\t0|1", 2, 0, capi.IN_ENALLELE),
    (b"1\t5\t.\tA\tC,G,T,<*>,<NON_REF>\t.\tPASS\t.\tGT\t0|0\t0|1", 2, 1, capi.IN_ENALLELE),
    (b"1\t5\t.\t0\t2\t.\tPASS\t.\tGT\t0|0\t0|1", 2, 0, capi.IN_EALLELE),
    (b"1\t5\t.\tA\tN\t.\tPASS\t.\tGT\t0|0\t0|1", 2, 1, capi.IN_EALLELE),
    (b"1\t5\t.\tA\tCT\t.\tPASS\t.\tGT\t0|0\t0|1", 2, 1, capi.IN_EALLELE),
    (b"1\t5\t.\tA\t\t.\tPASS\t.\tGT\t0|0\t0|1", 2, 1, capi.IN_EALLELE),
    (b"1\t5\t.\t0\t1\t.\tPASS\t.\tDP:GQ\t3:4\t5:6", 2, 0, capi.IN_ENOGT),
    (b"1\t5\t.\t0\t1\t.\tPASS\t.\tGT\t0|0", 2, 0, capi.IN_ENSAMPLES),
    (b"1\t5\t.\t0\t1\t.\tPASS\t.\tGT\t0|0\t0|x", 2, 0, capi.IN_EGTCHAR),
    (b"1\t5\t.\t0\t1\t.\tPASS\t.\tGT\t0|0\t0|", 2, 0, capi.IN_EGTCHAR),
    (b"1\t5\t.\t0\t1\t.\tPASS\t.\tGT\t0|0\t", 2, 0, capi.IN_EGTCHAR),
    (b"1\t5\t.\t0\t1\t.\tPASS\t.\tGT\t0|0\t0|1;", 2, 0, capi.IN_EGTCHAR),
    (b"1\t5\t.\t0\t1\t.\tPASS\t.\tGT:DP\t0|0:3\t:3", 2, 0, capi.IN_EGTCHAR),
    (b"1\t5\t.\t0\t1\t.\tPASS\t.\tGT\t0|0\t1", 2, 0, capi.IN_EPLOIDY),
    (b"1\t5\t.\t0\t1\t.\tPASS\t.\tGT\t0|0\t.", 2, 0, capi.IN_EPLOIDY),
    (b"1\t5\t.\t0\t1\t.\tPASS\t.\tGT\t0|0|1\t0|1", 2, 0, capi.IN_EPLOIDY),
    (b"1\t5\t.\t0\t1\t.\tPASS\t.\tDP:GT\t3:0|0\t4", 2, 0, capi.IN_EPLOIDY),
    (b"1\t5\t.\t0\t1\t.\tPASS\t.\tGT\t0|0\t0|2", 2, 0, capi.IN_EALLELEIDX),
    (b"1\t5\t.\t0\t.\t.\tPASS\t.\tGT\t0|0\t0|1", 2, 0, capi.IN_EALLELEIDX),
    (b"1\t5\t.\tA\tC\t.\tPASS\t.\tGT\t0|0\t0|10", 2, 1, capi.IN_EALLELEIDX),
    (b"1\t5\t.\tA\t<*>\t.\tPASS\t.\tGT\t0|0\t0|1", 2, 1, capi.IN_ESYMBOLIC),
    (b"1\t5\t.\tA\t<NON_REF>\t.\tPASS\t.\tGT\t1/1\t0|0", 2, 1, capi.IN_ESYMBOLIC),
]

GOOD = [
    # (record line, S, source, expected packed genotypes, pos, n_allele)
    (b"1\t5\t.\t0\t1\t.\tPASS\t.\tGT\t0|0\t0|1\r", 2, 0, [0x00, 0x10], 4, 2),
    (b"1\t+7\t.\t1\t0\t.\tPASS\t.\tGT\t0|1\t1/1\textra", 2, 0, [0x01, 0x00], 6, 2),
    (b"1\t0\t.\tT\tG,A\t.\tPASS\t.\tGT:DP\t2|1:3\t./.:.", 2, 1, [0x20, 0xFF], -1, 3),
    (b"c\t12\tid\tG\t.\t9\tq\tX=1\tDP:GT:GQ\t1:0/0:2\t3:.|0", 2, 1, [0x22, 0x2F], 11, 1),
    (b"c\t12\tid\tG\tA,<*>\t9\tq\tX=1\tGT\t00|1\t+|01", 2, 1, [0x02, 0x02], 11, 3),
    (b"1\t5\t.\t0x\t1y\t.\tPASS\t.\tGT\t1|1\t1|0", 2, 0, [0x11, 0x01], 4, 2),
]


