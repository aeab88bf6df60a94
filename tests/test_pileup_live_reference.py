"""Live pin of the -printPileup formatter (vcfgl_b200/pileup.py, host formatting of the simulator's draws; reference:
vcfgl.cpp:616-634) on the instrumented reference binary: the seeded random configurations of tests/fuzz_cases.py are run
with -printPileup 1, and the pileup text rebuilt from the captured draws must be the file the reference wrote, byte for byte.

Container only: skipped where oracle/_ref does not exist (the GPU box uses tests/golden/pileup/ instead)."""
import gzip
import os
import random
import subprocess

import pytest

import replay_util
import vgl_dump
from fuzz_cases import draw_case, reference_exited
from vcfgl_b200 import pileup

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN_DUMP = os.path.join(ROOT, "oracle", "_ref", "vcfgl_ref_dump")

pytestmark = pytest.mark.skipif(not os.path.exists(BIN_DUMP), reason="oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("block", range(2))
def test_pileup_equals_live_reference(block, tmp_path):
    rnd = random.Random(7700 + block)
    n_bytes = 0
    for k in range(12):
        ref_argv, a, vcf, _ = draw_case(rnd, str(tmp_path), k)
        dump = str(tmp_path / ("c%d.vgld" % k))
        r = subprocess.run([BIN_DUMP, "-i", vcf, "-o", str(tmp_path / ("c%d" % k))] + ref_argv + ["-printPileup", "1"],
                           capture_output=True, text=True, env=dict(os.environ, VGL_DUMP_PATH=dump))
        if r.returncode != 0 and reference_exited(r.stderr):
            continue
        assert r.returncode == 0, (ref_argv, r.stderr[-1500:])
        if not os.path.exists(dump) or os.path.getsize(dump) == 0:
            continue
        sites = vgl_dump.read_dump(dump)
        S = sites[0].S
        _, rp = replay_util.batch_from_dump(sites, a)
        # synth.write_vcf inputs: one contig "1", binary alleles -> REF is A at every site (vcfgl.cpp:103-127)
        got = pileup.format_pileup(a, ["1"] * len(sites), [d.pos for d in sites], [0] * len(sites), [d.ret for d in sites], rp, S)
        want = gzip.open(str(tmp_path / ("c%d.pileup.gz" % k)), "rb").read()
        assert got == want, (ref_argv, got[:300], want[:300])
        n_bytes += len(want)
    assert n_bytes > 10000
