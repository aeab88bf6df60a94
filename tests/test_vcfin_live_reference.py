"""Live pin of the input-path oracle (oracle/vcf_in_oracle.c) and the host site planner (vcfgl_b200/vcfinput.py) on the
instrumented reference binary: seeded random VCF texts (ACGT and binary alleles, 1-3 ALT alleles and <*>, phased /
unphased / missing genotypes, FORMAT with extra keys, IDs / QUAL / FILTER / INFO content, gaps between positions) with
random --source / -explode / --rm-invar-sites; the (position, true genotypes) sequence the reference handed to
simulate_record_values (vcfgl.cpp:1469-1538) must be the sequence oracle parse -> SitePlanner gives.

Container only: skipped where oracle/_ref does not exist (the GPU box uses tests/golden/inputs/ instead)."""
import os
import random
import subprocess

import numpy as np
import pytest

import vgl_dump
from test_vcfin_oracle import planned_sequence
from vcfgl_b200 import vcfinput

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN_DUMP = os.path.join(ROOT, "oracle", "_ref", "vcfgl_ref_dump")

pytestmark = pytest.mark.skipif(not os.path.exists(BIN_DUMP), reason="oracle/_ref not built (needs /root/reference)")

HDR = ("##fileformat=VCFv4.2\n##FILTER=<ID=PASS,Description=\"All filters passed\">\n##FILTER=<ID=q10,Description=\"low\">\n"
       "%s"
       "##INFO=<ID=NS,Number=1,Type=Integer,Description=\"n\">\n##INFO=<ID=AF,Number=A,Type=Float,Description=\"af\">\n"
       "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
       "##FORMAT=<ID=DP,Number=1,Type=Integer,Description=\"d\">\n##FORMAT=<ID=GQ,Number=1,Type=Integer,Description=\"q\">\n"
       "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t%s\n")


def random_vcf(rnd, acgt, extra_values=(".", "3", "17", "250")):
    S = rnd.choice([1, 2, 3, 5, 8])
    n_rec = rnd.randrange(4, 40)
    p_alt = rnd.choice([0.0, 0.05, 0.3, 0.6, 1.0])     # 0 / 1: runs of invariant records for --rm-invar-sites
    p_miss = rnd.choice([0.0, 0.0, 0.1, 0.5])
    pos = 0
    recs = []
    # one, two or three contigs; the records move on to the next contig at a random record (possibly never: a contig without records)
    contigs = ["chrA", "chrB", "chrC"][:rnd.choice([1, 1, 2, 3])]
    lengths = []
    ci = 0
    for _ in range(n_rec):
        if ci + 1 < len(contigs) and rnd.random() < 0.08:
            lengths.append(pos + rnd.choice([0, 0, 3, 11]))
            ci += 1
            pos = 0
        pos += rnd.choice([1, 1, 1, 2, 3, 7])
        if acgt:
            ref = rnd.choice("ACGT")
            alts = rnd.sample([b for b in "ACGT" if b != ref], rnd.choice([1, 1, 1, 2, 3]))
            n_real = len(alts)
            if rnd.random() < 0.1:
                alts.append("<*>")
        else:
            ref, alts = ("0", ["1"]) if rnd.random() < 0.9 else ("1", ["0"])
            n_real = 1
        n_extra = rnd.choice([0, 0, 1, 2])
        cols = []
        for _s in range(S):
            sep = rnd.choice("||/")
            if rnd.random() < p_miss:
                g = "." + sep + "."
            elif p_miss and rnd.random() < 0.1:         # half-missing
                h = str(rnd.randrange(0, n_real + 1))
                g = rnd.choice(["." + sep + h, h + sep + "."])
            else:
                g = sep.join(str(rnd.randrange(1, n_real + 1) if rnd.random() < p_alt else 0) for _h in range(2))
            cols.append(":".join([g] + [rnd.choice(extra_values) for _x in range(n_extra)]))
        recs.append([contigs[ci], str(pos), rnd.choice([".", ".", "rs%d" % pos, "a;b"]), ref, ",".join(alts),
                     rnd.choice([".", "30", "12.5"]), rnd.choice([".", "PASS", "q10"]),
                     rnd.choice([".", "NS=3"]), ":".join(["GT", "DP", "GQ"][:1 + n_extra])] + cols)
    lengths.append(pos + rnd.choice([0, 0, 3, 11]))
    while len(lengths) < len(contigs):
        lengths.append(rnd.choice([1, 4, 9]))
    ctg = "".join("##contig=<ID=%s,length=%d>\n" % cl for cl in zip(contigs, lengths))
    text = HDR % (ctg, "\t".join("s%d" % i for i in range(S))) + "".join("\t".join(r) + "\n" for r in recs)
    return S, text.encode()


@pytest.mark.parametrize("block", range(4))
def test_planned_sequence_equals_live_reference(block, tmp_path):
    rnd = random.Random(9300 + block)
    n_sites = n_runs = 0
    for k in range(15):
        acgt = rnd.random() < 0.6
        S, buf = random_vcf(rnd, acgt)
        explode, rm_invar = rnd.choice([0, 0, 1]), rnd.choice([0, 0, 1, 2, 3])
        vcf = str(tmp_path / ("in%d.vcf" % k))
        open(vcf, "wb").write(buf)
        dump = str(tmp_path / ("c%d.vgld" % k))
        argv = ["--seed", "3", "-O", "v", "--source", str(int(acgt)), "-explode", str(explode), "--rm-invar-sites", str(rm_invar),
                "-d", "1", "-e", "0.01", "-GL", "2"]
        r = subprocess.run([BIN_DUMP, "-i", vcf, "-o", str(tmp_path / ("o%d" % k))] + argv, capture_output=True, text=True,
                           env=dict(os.environ, VGL_DUMP_PATH=dump))
        assert r.returncode == 0, (argv, buf.decode(), r.stderr[-1500:])
        want = vgl_dump.read_dump(dump) if os.path.exists(dump) and os.path.getsize(dump) else []
        hdr = vcfinput.read_header(buf)
        assert len(hdr.samples) == S
        for chunk_records in (None, 4):
            seq = planned_sequence(buf[hdr.body_offset:], S, int(acgt), explode, rm_invar & 3, hdr.contigs, chunk_records=chunk_records)
            where = (argv, buf.decode())
            assert [p for p, _ in seq] == [d.pos for d in want], where
            for (p, g), d in zip(seq, want):
                assert np.array_equal(g, d.gts), (p, g, d.gts, where)
        n_sites += len(want)
        n_runs += 1
    assert n_runs == 15 and n_sites > 100
