"""VGL_HOST_BGZF: the BCF record stream of a batch compressed on the device into BGZF blocks (csrc/bgzf.cu; the reference's
default container, -O b: htslib/bgzf.c, vcfgl.cpp:1791-1803).

The blocks use one dynamic Huffman code per context (RFC 1951 3.2.7), built by vgl_submit from the symbol statistics of
the context's first record stream (csrc/tables.cpp bgzf_build_code, pinned on zlib by tests/test_tables.py); a block the
code would expand beyond the block image (three quarters of its input) goes out as a stored block.

Oracle = zlib on the host: every block must be a well-formed BGZF gzip member (magic, "BC" extra field with the block size,
raw deflate data, CRC32 and length of the uncompressed bytes), and the inflated blocks, concatenated, must equal byte for
byte the uncompressed record stream VGL_HOST_BCF returns for the same submit (which tests/test_gpu_bcf.py pins to the
reference's -O u files).  A file made of a host-compressed header, the device's blocks and the EOF block must read back
through Python's gzip module as a multi-member gzip stream.
"""
import gzip
import io
import struct
import zlib

import numpy as np
import pytest

from vcfgl_b200 import args as vargs
from vcfgl_b200 import capi, synth

pytestmark = pytest.mark.gpu

EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")   # htslib/bgzf.c: the 28-byte empty block
IDS = dict(DP=1, GL=2, PL=3, GP=4, AD=5, ADF=6, ADR=7, QS=8, I16=9)

CASES = {
    # name: (argv, S, n_sites, batch, missing rate)
    "cfg2": ("--seed 42 -d 10 -e 0.01 -GL 1 -addGL 1 -addPL 1 -addFormatAD 1", 100, 700, 700, 0.0),
    "cfg2_batches": ("--seed 42 -d 10 -e 0.01 -GL 1 -addGL 1 -addPL 1 -addFormatAD 1", 100, 700, 256, 0.02),
    "gl2_alltags": ("--seed 5 -d 6 -e 0.02 -GL 2 -doUnobserved 1 -addGP 1 -addPL 1 -addInfoDP 1 -addFormatAD 1 -addInfoAD 1 "
                    "-addFormatADF 1 -addFormatADR 1", 37, 400, 400, 0.03),
    "trim": ("--seed 6 -d 2 -e 0.1 -GL 1 -doUnobserved 0 --rm-invar-sites 4 --rm-empty-sites 1 -addPL 1 -addFormatAD 1", 5, 3000, 3000, 0.0),
    "s1": ("--seed 10 -d 4 -e 0.05 -GL 1 -addPL 1 -addFormatAD 1", 1, 5000, 5000, 0.0),          # > 32 records per block: literals only
    "s1300": ("--seed 11 -d 8 -e 0.01 -GL 1 -addGL 1 -addPL 1 -addFormatAD 1", 1300, 40, 40, 0.0),   # a record spans several blocks
    "s2501_deep": ("--seed 14 -d 40 -e 0.02 -GL 1 -addGL 1 -addPL 1 -addFormatAD 1", 2501, 10, 10, 0.0),
    "aux_i16": ("--seed 7 -d 10 -e 0.001 -GL 1 -doUnobserved 1 -addPL 1 -addI16 1 -addQS 1", 100, 600, 600, 0.0),
}


def split_blocks(buf):
    """-> [(payload, crc, isize)] of a run of BGZF blocks; checks the fixed header fields"""
    out, o = [], 0
    while o < len(buf):
        assert buf[o:o + 4] == b"\x1f\x8b\x08\x04", (o, buf[o:o + 4])
        xlen = struct.unpack_from("<H", buf, o + 10)[0]
        assert xlen == 6 and buf[o + 12:o + 16] == b"BC\x02\x00"
        bsize = struct.unpack_from("<H", buf, o + 16)[0] + 1
        assert o + bsize <= len(buf)
        crc, isize = struct.unpack_from("<II", buf, o + bsize - 8)
        out.append((buf[o + 18:o + bsize - 8], crc, isize))
        o += bsize
    return out


def run(mode, a, S, gt, n_sites, batch):
    ctx = capi.Context(capi.params_from_args(a, S, max_batch_sites=batch, n_slots=1, host_output=mode, bcf_dict=IDS,
                                             bcf_blob_bytes_per_site=16))
    outs = []
    for s0 in range(0, n_sites, batch):
        m = min(batch, n_sites - s0)
        ctx.input_buffer(0)[:m] = gt[s0:s0 + m]
        sin, _ = ctx.bcf_input(0)
        sin[:m] = 0
        sin["pos"][:m] = np.arange(s0, s0 + m) * 3 + 1
        sin["qual_bits"][:m] = capi.F32_MISSING_BITS
        ctx.submit(0, s0, m)
        b = ctx.wait(0)
        assert b.status == 0
        if mode == capi.HOST_BGZF:
            assert b.bcf is None
            outs.append((bytes(b.bgzf), b.bcf_off.copy(), b.bcf_bytes, b.bgzf_blocks))
        else:
            outs.append((bytes(b.bcf), b.bcf_off.copy(), b.bcf_bytes, 0))
    ctx.close()
    return outs


@pytest.mark.parametrize("name", sorted(CASES))
def test_bgzf_blocks_inflate_to_the_record_stream(name):
    argv, S, n_sites, batch, missing = CASES[name]
    a = vargs.parse_args(argv.split())
    hap = synth.sfs_genotypes(n_sites, S, 4242, missing) if S > 1 else np.random.default_rng(1).integers(0, 2, (n_sites, 2)).astype(np.int8)
    gt = synth.pack_gt(hap)
    plain = run(capi.HOST_BCF, a, S, gt, n_sites, batch)
    packed = run(capi.HOST_BGZF, a, S, gt, n_sites, batch)
    tot_raw = tot_z = 0
    for (want, off_w, nb_w, _), (z, off_z, nb_z, n_blocks) in zip(plain, packed):
        assert nb_z == nb_w and np.array_equal(off_w, off_z)      # the offsets describe the uncompressed stream
        blocks = split_blocks(z)
        assert len(blocks) == n_blocks == (nb_w + 32767) // 32768
        got = bytearray()
        for payload, crc, isize in blocks:
            d = zlib.decompressobj(-15)
            raw = d.decompress(payload)
            assert d.eof and not d.unused_data, "deflate stream must end with the block"
            assert len(raw) == isize and zlib.crc32(raw) == crc
            assert isize == 32768 or payload is blocks[-1][0]
            got += raw
        assert bytes(got) == want, name
        tot_raw += len(want)
        tot_z += len(z)
    assert tot_raw > 0
    # a whole file: host-compressed header + the device's blocks + EOF block reads back as a multi-member gzip stream
    header = b"BCF\x02\x02" + struct.pack("<I", 4) + b"x\n\x00\x00"
    hz = zlib.compressobj(6, zlib.DEFLATED, -15)
    hd = hz.compress(header) + hz.flush()
    hblock = b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(hd) + 25) + hd + \
        struct.pack("<II", zlib.crc32(header), len(header))
    blob = hblock + b"".join(p[0] for p in packed) + EOF_BLOCK
    assert gzip.GzipFile(fileobj=io.BytesIO(blob)).read() == header + b"".join(p[0] for p in plain)
    if name == "cfg2":    # the simulated tags repeat: the stream must actually shrink
        assert tot_z < 0.5 * tot_raw, (tot_z, tot_raw)


def inflate_all(z):
    got = bytearray()
    kinds = set()
    for payload, crc, isize in split_blocks(z):
        d = zlib.decompressobj(-15)
        raw = d.decompress(payload)
        assert d.eof and not d.unused_data and len(raw) == isize and zlib.crc32(raw) == crc
        kinds.add((payload[0] >> 1) & 3)      # BTYPE of the block's only deflate block
        got += raw
    return bytes(got), kinds


def test_dynamic_code_beats_the_fixed_code(monkeypatch):
    argv, S, n_sites, batch, missing = CASES["cfg2"]
    a = vargs.parse_args(argv.split())
    gt = synth.pack_gt(synth.sfs_genotypes(n_sites, S, 4242, missing))
    plain = run(capi.HOST_BCF, a, S, gt, n_sites, batch)
    dyn = run(capi.HOST_BGZF, a, S, gt, n_sites, batch)
    monkeypatch.setenv("VGL_BGZF_FIXED", "1")
    fix = run(capi.HOST_BGZF, a, S, gt, n_sites, batch)
    raw_d, kinds_d = inflate_all(dyn[0][0])
    raw_f, kinds_f = inflate_all(fix[0][0])
    assert raw_d == raw_f == plain[0][0]
    assert kinds_d == {2} and kinds_f == {1}
    assert len(dyn[0][0]) < 0.9 * len(fix[0][0]), (len(dyn[0][0]), len(fix[0][0]))


def test_blocks_the_image_does_not_hold_are_stored():
    """the code is built from the first batch; a first batch of missing genotypes only (no reads: constant tags) gives a
    code under which the second batch's literals take up to 15 bits each"""
    argv, S, n_sites = "--seed 42 -d 10 -e 0.2 -GL 1 -addGL 1 -addPL 1 -addFormatAD 1", 1, 6000   # one sample: blocks without cells
    a = vargs.parse_args(argv.split())
    hap = np.random.default_rng(3).integers(0, 2, (2 * n_sites, 2)).astype(np.int8)
    hap[:n_sites] = -1
    gt = synth.pack_gt(hap)
    plain = run(capi.HOST_BCF, a, S, gt, 2 * n_sites, n_sites)
    packed = run(capi.HOST_BGZF, a, S, gt, 2 * n_sites, n_sites)
    kinds = set()
    for (want, _, _, _), (z, _, _, _) in zip(plain, packed):
        got, k = inflate_all(z)
        assert got == want
        kinds |= k
    assert kinds == {0, 2}, kinds       # BTYPE 00 = stored, 10 = dynamic Huffman
