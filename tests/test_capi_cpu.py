"""CPU-side checks of the drop-in boundary (no GPU needed): the C-ABI library loads, exports every
symbol include/vgl.h declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from vcfgl_b200 import args as vargs
from vcfgl_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "vgl.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vgl_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    names = header_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), "libvgl.so does not export %s" % n
    assert sorted(capi.EXPORTS) == names
    assert lib.vgl_abi_version() == capi.ABI_VERSION == 9


def test_struct_layouts_match_header():
    # sizes the C side was compiled with (natural alignment), mirrored by ctypes
    assert C.sizeof(capi.VglSiteOut) == 208 or C.sizeof(capi.VglSiteOut) == capi.SITE_DTYPE.itemsize
    assert capi.SITE_DTYPE.itemsize == C.sizeof(capi.VglSiteOut)
    assert C.sizeof(capi.VglParams) % 8 == 0


def test_ctypes_mirror_has_the_compiled_layout(tmp_path):
    # sizeof / offsetof as gcc lays out include/vgl.h == the ctypes mirror in vcfgl_b200/capi.py
    import subprocess
    src = tmp_path / "layout.c"
    src.write_text("""#include <stdio.h>
#include <stddef.h>
#include "vgl.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(vgl_in_site), offsetof(vgl_in_site, allele_acgt),
         sizeof(vgl_parse_out), offsetof(vgl_parse_out, sites), sizeof(vgl_gvcf_rec), sizeof(vgl_gvcf_out), sizeof(vgl_params), offsetof(vgl_params, host_output), offsetof(vgl_params, bcf_dict),
         offsetof(vgl_params, bcf_blob_bytes_per_site), sizeof(vgl_batch_out), offsetof(vgl_batch_out, bcf), offsetof(vgl_batch_out, bcf_bytes),
         sizeof(vgl_bcf_site_in), sizeof(vgl_site_out));
  return 0; }
""")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    P, B = capi.VglParams, capi.VglBatchOut
    want = [capi.IN_SITE_DTYPE.itemsize, capi.IN_SITE_DTYPE.fields["allele_acgt"][1], C.sizeof(capi.VglParseOut), capi.VglParseOut.sites.offset,
            capi.GVCF_REC_DTYPE.itemsize, C.sizeof(capi.VglGvcfOut), C.sizeof(P), P.host_output.offset, P.bcf_dict.offset, P.bcf_blob_bytes_per_site.offset, C.sizeof(B), B.bcf.offset,
            B.bcf_bytes.offset, C.sizeof(capi.VglBcfSiteIn), C.sizeof(capi.VglSiteOut)]
    assert got == want


def test_strerror():
    lib = capi.load()
    assert b"no CPU path" in lib.vgl_strerror(capi.VGL_ENODEV)
    assert lib.vgl_strerror(0) == b"ok"


def test_create_rejects_bad_params_like_the_reference_cli():
    lib = capi.load()
    a = vargs.parse_args("-d 10 -e 0.01 -GL 1".split())
    p = capi.params_from_args(a, 4, 16)
    h = C.c_void_p()
    p.gl_model, p.precise_gl = 1, 1        # io.cpp:935
    assert lib.vgl_create(C.byref(p), C.byref(h)) == capi.VGL_EINVAL
    p.precise_gl = 0
    p.error_rate = 1.0                      # io.cpp:868 [0, 1)
    assert lib.vgl_create(C.byref(p), C.byref(h)) == capi.VGL_EINVAL
    p.error_rate = 0.01
    p.abi_version = 99
    assert lib.vgl_create(C.byref(p), C.byref(h)) == capi.VGL_EINVAL


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    a = vargs.parse_args("-d 10 -e 0.01 -GL 1".split())
    with pytest.raises(capi.VglError) as e:
        capi.Context(capi.params_from_args(a, 4, 16))
    assert e.value.code == capi.VGL_ENODEV


def test_cli_mirror_validation():
    ok = vargs.parse_args("--seed 42 -d 1 -e 0.2 -GL 1 --adjust-qs 3 -addQS 1 -explode 1".split())
    assert ok.gl_model == 1 and ok.adjust_qs == 3 and ok.tag_mask & vargs.TAG_QS
    for bad in ("-e 0.01 -GL 1",                       # no depth
                "-d 1 -GL 1",                          # no error rate
                "-d 1 -e 0.01 -GL 1 --precise-gl 1",   # io.cpp:935
                "-d 1 -e 0.01 -eq 2",                  # needs beta variance
                "-d 1 -e 0 -eq 1 -bv 1e-5",            # needs error rate > 0
                "-d 1 -e 0.01 -bv 1e-5",               # beta variance without error-qs
                "-d 1 -e 0.01 --adjust-qs 2",          # needs -addQS
                "-d 501 -e 0.01",                      # depth range
                "-d 1 -e 0.01 --gvcf-dps 1,5",         # io.cpp:986-989: needs -doGVCF 1
                "-d 1 -e 0.01 --adjust-qs 4",          # io.cpp:891: needs -printPileup 1
                "-d 1 -e 0.01 --adjust-qs 8",          # io.cpp:894: needs -printQScores 1
                "-d 1 -e 0.01 --adjust-qs 16",         # io.cpp:897: needs -printGlError 1
                "-d 1 -e 0.01 -doGVCF 1"):             # gVCF requirements
        with pytest.raises(vargs.ArgError):
            vargs.parse_args(bad.split())
    assert vargs.parse_args("-d 1 -e 0.01 --adjust-qs 12 -printPileup 1 -printQScores 1".split()).adjust_qs == 12
    al, be = vargs.beta_shape(0.01, 1e-5)
    assert abs(al - 9.89) < 1e-9 and abs(be - 979.11) < 1e-9
