"""Live pin of oracle/bcf_oracle.py's record layout for INPUT records that carry FORMAT keys besides GT (the case
VGL_HOST_BCF does not serialise yet, DESIGN.md 4 "Limit of this mode"): random VCF texts with FORMAT GT / GT:DP / GT:DP:GQ
go through the unmodified reference with -O u and through the instrumented one; every record rebuilt by the oracle from the
capture plus the input's own blocks (in_fmt) must be the reference's bytes: an input DP is replaced in place by the simulated
FORMAT/DP, GQ stays in front of the simulated tags, and with -addFormatDP 0 the input's DP block is passed through.

Container only: skipped where oracle/_ref does not exist."""
import os
import random
import subprocess

import pytest

import bcf_util as bu
import vgl_dump
from test_vcfin_live_reference import random_vcf
from vcfgl_b200 import args as vargs

bo = bu.bo
BIN = os.path.join(bu.ROOT, "oracle", "_ref", "vcfgl_ref")
BIN_DUMP = os.path.join(bu.ROOT, "oracle", "_ref", "vcfgl_ref_dump")

pytestmark = pytest.mark.skipif(not os.path.exists(BIN_DUMP), reason="oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("block", range(3))
def test_records_with_input_format_blocks(block, tmp_path):
    rnd = random.Random(9950 + block)
    n_records = n_passed_through = n_in_place = 0
    for k in range(10):
        acgt = rnd.random() < 0.5
        S, buf = random_vcf(rnd, acgt)
        vcf = str(tmp_path / ("in%d.vcf" % k))
        open(vcf, "wb").write(buf)
        argv = ["--seed", str(rnd.randrange(1, 999)), "-O", "u", "--source", str(int(acgt)), "-d", rnd.choice(["1", "4", "40"]), "-e", "0.01",
                "-GL", str(rnd.choice([1, 2])), "-doUnobserved", str(rnd.randrange(0, 6)), "-addFormatDP", str(rnd.choice([0, 1, 1])),
                "-addPL", str(rnd.choice([0, 1])), "-addGL", "1", "-addFormatAD", str(rnd.choice([0, 1])), "-addInfoDP", str(rnd.choice([0, 1]))]
        a = vargs.parse_args(argv)
        dump = str(tmp_path / ("c%d.vgld" % k))
        for binary, env, out in ((BIN, dict(os.environ), "u%d" % k), (BIN_DUMP, dict(os.environ, VGL_DUMP_PATH=dump), "d%d" % k)):
            r = subprocess.run([binary, "-i", vcf, "-o", str(tmp_path / out)] + argv, capture_output=True, text=True, env=env)
            assert r.returncode == 0, (argv, r.stderr[-1500:])
        text, ids, recs = bo.read_bcf(str(tmp_path / ("u%d.bcf" % k)))
        kept = [d for d in vgl_dump.read_dump(dump) if d.ret == 0] if os.path.exists(dump) and os.path.getsize(dump) else []
        assert len(kept) == len(recs)
        # FORMAT keys of the input record at each position (no -explode here: one record per site)
        in_keys = {}
        names = [l.split("ID=")[1].split(",")[0] for l in buf.decode().splitlines() if l.startswith("##contig")]
        for line in buf.decode().splitlines():
            if line and not line.startswith("#"):
                f = line.split("\t")
                in_keys[(names.index(f[0]), int(f[1]) - 1)] = f[8].split(":")[1:]
        ftags, itags = bu.enabled_tags(a)
        sim_ids = {ids["FORMAT/" + t] for t in ftags}
        for d, rec in zip(kept, recs):
            r = bo.split_record(rec)
            n_in = r["n_info"] - len(itags)
            passthrough = r["filter_bytes"] + b"".join(b for _, b in r["infos"][:n_in])
            raw = {key: blk for key, n, t, blk in r["fmts"] if key not in sim_ids}     # blocks the reference copied from the input
            in_fmt = []
            for name in in_keys[(d.rid, d.pos)]:
                key = ids["FORMAT/" + name]
                in_fmt.append((key, raw.get(key)))
                n_in_place += key in sim_ids
                n_passed_through += key not in sim_ids
            alleles = bo.alleles_of_site(d.n_alleles, d.alleles2acgt, d.info_dp, a.do_unobserved, a.do_gvcf)
            fmt, info = bu.site_arrays(a, d)
            got = bo.encode_record(d.rid, d.pos, r["qual_bits"], r["id_bytes"], passthrough, n_in, alleles, d.S, ids, fmt, info, in_fmt=in_fmt)
            assert got == rec, (argv, d.pos, in_keys[(d.rid, d.pos)], [k_ for k_, *_ in r["fmts"]])
        n_records += len(recs)
    assert n_records > 50 and n_passed_through > 20 and n_in_place > 20, (n_records, n_passed_through, n_in_place)
