#!/usr/bin/env python3
"""tools/make_bcf_inputs.py -- BCF versions of the input fixtures (container only; needs oracle/_ref).

tests/bcf_writer.py encodes every tests/golden/inputs/<name>.vcf.gz as uncompressed BCF; the UNMODIFIED reference is then run
on both files with the same arguments and must write the same VCF (so the encoding is one htslib accepts and reads as the same
genotypes).  Only then is the BCF stored as tests/golden/inputs/<name>.bcf.gz, with the GT dictionary id in bcf_inputs.json."""
import gzip
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bcf_writer as bw  # noqa: E402

BIN = os.path.join(ROOT, "oracle/_ref/vcfgl_ref")
INPUTS = os.path.join(ROOT, "tests/golden/inputs")
ARGS = {"default": "--seed 42 -O v -d 2 -e 0.01 -GL 1", "acgt": "--seed 42 -O v --source 1 -d 2 -e 0.01 -GL 1 -explode 1"}
ACGT = {"data4_acgt_biallelic.vcf", "data4_acgt_biallelic_a2g_c2t.vcf", "data5_acgt_multiallelic.vcf", "data6.vcf", "data7.vcf", "in_acgt.vcf"}


def main():
    tmp = tempfile.mkdtemp(prefix="vgl_bcfin_")
    manifest = {}
    for f in sorted(os.listdir(INPUTS)):
        if not f.endswith(".vcf.gz"):
            continue
        name = f[:-3]
        buf = gzip.open(os.path.join(INPUTS, f), "rb").read()
        bcf, first, offs, ids = bw.vcf_to_bcf(buf)
        pv, pb = os.path.join(tmp, name), os.path.join(tmp, name + ".bcf")
        open(pv, "wb").write(buf)
        open(pb, "wb").write(bcf)
        argv = ARGS["acgt" if name in ACGT else "default"].split()
        outs = []
        for p, tag in ((pv, "v"), (pb, "b")):
            r = subprocess.run([BIN, "-i", p, "-o", os.path.join(tmp, name + tag)] + argv, capture_output=True, text=True)
            if r.returncode != 0:
                raise SystemExit("%s (%s input): reference failed: %s" % (name, tag, r.stderr[-300:]))
            outs.append([l for l in open(os.path.join(tmp, name + tag + ".vcf")) if not l.startswith("##")])
        if outs[0] != outs[1] or len(outs[0]) < 2:
            raise SystemExit("%s: the reference's output differs between the VCF and the BCF input" % name)
        with gzip.GzipFile(os.path.join(INPUTS, name[:-4] + ".bcf.gz"), "wb", compresslevel=9, mtime=0) as g:
            g.write(bcf)
        manifest[name[:-4] + ".bcf"] = dict(vcf=name, gt_key=ids["GT"], first_record=first, n_records=len(offs) - 1,
                                            validated_by="reference output identical for VCF and BCF input (%s)" % " ".join(argv))
        print(name, "ok:", len(offs) - 1, "records, reference output identical (%d lines)" % len(outs[0]))
    json.dump(manifest, open(os.path.join(INPUTS, "bcf_inputs.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
