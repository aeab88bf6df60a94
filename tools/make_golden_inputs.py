#!/usr/bin/env python3
"""tools/make_golden_inputs.py -- fixtures of the INPUT path (container only; needs /root/reference and oracle/_ref).

tests/golden/inputs/ receives

  <input>.vcf.gz     the input VCF of every case in tests/golden/manifest.json (the reference's own test/data files and the
                     synthetic inputs of tools/make_golden.py, regenerated from their seeds), so that the (pos, true
                     genotypes) sequence the instrumented reference dumped for that case (tests/golden/<id>.vgld.gz) pins
                     the input-path oracle on the GPU box, where /root/reference does not exist;
  in_cases.json      extra input-path cases the golden set does not reach: --rm-invar-sites 1 / 2 / 3 (records dropped by
                     check_rec_alleles, vcfgl.cpp:150-160), -explode 1 with --source 1, unphased and missing genotypes,
                     FORMAT with more keys than GT, five alleles.  Each is the UNMODIFIED instrumented reference run on a
                     hand-written VCF (stored next to it); the capture is reduced to [pos, true_gts_acgt_int[2S]] per site
                     that reached simulate_record_values.
"""
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_golden  # noqa: E402
import vgl_dump  # noqa: E402

REF = make_golden.REF
OUT = os.path.join(ROOT, "tests/golden/inputs")

HDR = ("##fileformat=VCFv4.2\n##FILTER=<ID=PASS,Description=\"All filters passed\">\n##FILTER=<ID=q10,Description=\"low\">\n"
       "##contig=<ID=chrA,length=%d>\n"
       "##INFO=<ID=NS,Number=1,Type=Integer,Description=\"n\">\n##INFO=<ID=AF,Number=A,Type=Float,Description=\"af\">\n"
       "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
       "##FORMAT=<ID=DP,Number=1,Type=Integer,Description=\"d\">\n##FORMAT=<ID=GQ,Number=1,Type=Integer,Description=\"q\">\n"
       "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t%s\n")


def vcf(length, samples, recs):
    return HDR % (length, "\t".join(samples)) + "".join("\t".join(r) + "\n" for r in recs)


def hand_written():
    s3 = ["i1", "i2", "i3"]
    acgt = vcf(14, s3, [
        ["chrA", "2", "rs1", "A", "C", "30", "PASS", "NS=3", "GT:DP:GQ", "0|1:3:40", "1/1:5:50", "./.:.:."],
        ["chrA", "3", ".", "G", "T,A", "12.5", "q10", "NS=3;AF=0.1,0.2", "GT:DP", "2|0:1", "0/2:7", ".|.:2"],
        ["chrA", "5", "x;y", "T", "A,C,G,<*>", ".", ".", ".", "GT", "3|2", "1|0", "0/3"],
        ["chrA", "6", ".", "C", ".", ".", "PASS", ".", "GT:GQ", "0|0:9", "0/0:9", "0|0:1"],
        ["chrA", "9", ".", "C", "G", ".", "PASS", ".", "GT", "1|1", "1|1", "1/1"],
        ["chrA", "12", ".", "G", "A", ".", "PASS", ".", "GT", "0|0", "0|0", "0|0"],
    ])
    s4 = ["tsk_0", "tsk_1", "tsk_2", "tsk_3"]
    binary = vcf(12, s4, [
        ["chrA", "1", ".", "0", "1", ".", "PASS", ".", "GT", "0|0", "0|0", "0|0", "0|0"],
        ["chrA", "3", ".", "0", "1", ".", "PASS", ".", "GT", "1|1", "1|1", "1|1", "1|1"],
        ["chrA", "4", ".", "0", "1", ".", "PASS", ".", "GT", "0|1", "1|0", "0|0", "1|1"],
        ["chrA", "7", ".", "1", "0", ".", "PASS", ".", "GT", "0|1", "0/0", "1/1", ".|."],
        ["chrA", "8", ".", "0", "1", ".", "PASS", ".", "GT", "0|0", "0|0", ".|.", "0|0"],
        ["chrA", "11", ".", "0", "1", ".", "PASS", ".", "GT", "1|1", ".|.", "1|1", "1|1"],
    ])
    return {"in_acgt.vcf": acgt, "in_binary.vcf": binary}


CASES = [
    # id, input, reference arguments
    ("in_acgt_plain", "in_acgt.vcf", "--seed 3 -O v --source 1 -d 2 -e 0.01 -GL 2"),
    ("in_acgt_explode", "in_acgt.vcf", "--seed 3 -O v --source 1 -explode 1 -d 2 -e 0.01 -GL 2"),
    ("in_acgt_rm1", "in_acgt.vcf", "--seed 3 -O v --source 1 -d 2 -e 0.01 -GL 2 --rm-invar-sites 1"),
    ("in_acgt_rm2", "in_acgt.vcf", "--seed 3 -O v --source 1 -d 2 -e 0.01 -GL 2 --rm-invar-sites 2"),
    ("in_acgt_rm3", "in_acgt.vcf", "--seed 3 -O v --source 1 -d 2 -e 0.01 -GL 2 --rm-invar-sites 3"),
    ("in_binary_plain", "in_binary.vcf", "--seed 5 -O v -d 2 -e 0.01 -GL 1"),
    ("in_binary_explode", "in_binary.vcf", "--seed 5 -O v -explode 1 -d 2 -e 0.01 -GL 1"),
    ("in_binary_rm1", "in_binary.vcf", "--seed 5 -O v -d 2 -e 0.01 -GL 1 --rm-invar-sites 1"),
    ("in_binary_rm2", "in_binary.vcf", "--seed 5 -O v -d 2 -e 0.01 -GL 1 --rm-invar-sites 2"),
    ("in_binary_rm3_explode", "in_binary.vcf", "--seed 5 -O v -explode 1 -d 2 -e 0.01 -GL 1 --rm-invar-sites 3"),
]


def gz_write(path, raw):
    with gzip.GzipFile(path, "wb", compresslevel=9, mtime=0) as g:
        g.write(raw)


def main():
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    tmp = tempfile.mkdtemp(prefix="vgl_inputs_")
    manifest = json.load(open(os.path.join(ROOT, "tests/golden/manifest.json")))
    # synthetic inputs of the extra golden cases, regenerated from their seeds (file names as in the manifest)
    synth_paths = {os.path.basename(f): f for _, f, _ in make_golden.extra_cases(tmp)}
    done = set()
    for cid, m in sorted(manifest.items()):
        name = m["input"]
        if name in done:
            continue
        src = synth_paths.get(name) or os.path.join(REF, "test/data", name)
        gz_write(os.path.join(OUT, name + ".gz"), open(src, "rb").read())
        done.add(name)
    print("stored %d golden input VCFs" % len(done))

    cases = {}
    texts = hand_written()
    for name, text in texts.items():
        open(os.path.join(tmp, name), "w").write(text)
        gz_write(os.path.join(OUT, name + ".gz"), text.encode())
    for cid, name, argline in CASES:
        argv = argline.split()
        pref = os.path.join(tmp, cid)
        dump = os.path.join(tmp, cid + ".vgld")
        make_golden.run(make_golden.BIN_DUMP, os.path.join(tmp, name), argv, pref, dump)
        sites = vgl_dump.read_dump(dump) if os.path.exists(dump) else []
        a = make_golden.vargs.parse_args(list(argv))
        cases[cid] = dict(input=name, argv=argv, source=a.source, explode=a.explode, rm_invar_sites=a.rm_invar_sites,
                          n_samples=int(sites[0].S) if sites else 0,
                          sites=[[int(d.pos), [int(x) for x in d.gts]] for d in sites],
                          pinned_by="instrumented reference binary run in the build container")
        print(cid, "%d sites reached simulate_record_values" % len(sites))
    json.dump(cases, open(os.path.join(OUT, "in_cases.json"), "w"), indent=0, sort_keys=True)
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
