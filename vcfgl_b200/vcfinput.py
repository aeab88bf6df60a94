"""Host side of the input path: VCF text -> sites of the simulation, mirroring the reference's driver loop.

The device does the per-record work (include/vgl.h "Input path": columns, allele map, genotypes, skip decision); what is
left for the host is what the reference's `main_simulate_record_values` (vcfgl.cpp:1469-1620) does AROUND records:

* the header: sample names from the `#CHROM` line, contig lengths from `##contig` (needed by -explode 1);
* the site sequence: every record that passes check_rec_alleles is a site; with `-explode 1` the positions between
  records (and from the last record to the end of the LAST contig of the file, vcfgl.cpp:1567-1611) become sites whose
  true genotypes are all 0|0 of a blank copy of the record being read when exploding first happened
  (`explode_rec`, vcfgl.cpp:1489-1503) -- so they are dropped as a whole by `--rm-invar-sites 1`;
* cutting that sequence into batches, `vgl_place_rows` + `vgl_submit(..., VGL_SUBMIT_GT_ON_DEVICE)`.

Nothing here touches genotypes: they stay on the device from text to tags.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

from . import capi

IN_STATUS_TEXT = {
    capi.IN_ENCOLS: "fewer than 10 tab-separated columns",
    capi.IN_EPOS: "position value is too large",
    capi.IN_ENALLELE: "multiallelic site with more alleles than supported (5; 2 with --source 0)",
    capi.IN_EALLELE: "allele is not a valid base (--source 1) / not 0 or 1 (--source 0)",
    capi.IN_ENOGT: "could not find GT tag",
    capi.IN_ENSAMPLES: "number of columns does not match the number of samples",
    capi.IN_EGTCHAR: "couldn't read GT data: value not a number or '.'",
    capi.IN_EPLOIDY: "a genotype is not diploid",
    capi.IN_EALLELEIDX: "GT allele index is not below the number of alleles",
    capi.IN_ESYMBOLIC: "GT refers to a symbolic allele",
}


class VcfInputError(RuntimeError):
    """what the reference reports with ERROR()/ASSERT() and exit(1) (shared.h:292-327)"""


@dataclass
class VcfHeader:
    samples: List[str]
    contigs: Dict[str, int]          # ##contig=<ID=..,length=..>
    body_offset: int                 # first byte after the #CHROM line
    text: bytes = b""


def read_header(buf: bytes) -> VcfHeader:
    """Header lines up to and including `#CHROM` (htslib/vcf.c vcf_hdr_read; only what the simulator needs)."""
    off = 0
    contigs: Dict[str, int] = {}
    while True:
        nl = buf.find(b"\n", off)
        if nl < 0:
            raise VcfInputError("no #CHROM line in the VCF header")
        line = buf[off:nl].rstrip(b"\r")
        if line.startswith(b"##contig=<"):
            body = line[len(b"##contig=<"):].rstrip(b">").decode()
            kv = dict(x.split("=", 1) for x in body.split(",") if "=" in x)
            if "ID" in kv:
                contigs[kv["ID"]] = int(kv.get("length", 0))
        elif line.startswith(b"#CHROM"):
            cols = line.decode().split("\t")
            if len(cols) < 10:
                raise VcfInputError("the VCF has no sample columns")
            return VcfHeader(cols[9:], contigs, nl + 1, buf[:nl + 1])
        elif not line.startswith(b"##"):
            raise VcfInputError("record before the #CHROM line")
        off = nl + 1


@dataclass
class SiteRun:
    """consecutive sites of one contig: site k is at 0-based position pos[k] and takes the genotypes of parsed record
    src[k], or the blank -explode record when src[k] < 0"""
    contig: str
    pos: np.ndarray
    src: np.ndarray
    slot: int = -1       # set by simulate_vcf_text: the slot the batch ran on (valid until the consumer asks for the next batch)


@dataclass
class SitePlanner:
    """the reference's site sequence (vcfgl.cpp:1469-1620) for records arriving chunk by chunk"""
    explode: int
    rm_invar: int
    contigs: Dict[str, int]
    max_run: int                                   # longest run handed out at once (<= batch capacity)
    contig: Optional[bytes] = None
    n_in_contig: int = 0                           # nSitesTotalInContig
    fill_acgt: int = -1                            # REF of explode_rec, once created
    n_skipped: int = 0                             # nSitesSkipped (input-side part)
    last_status: dict = field(default_factory=dict)

    def _explode_sites_kept(self) -> bool:
        return not (self.rm_invar & 1)             # blank record: allelesum == 0 -> -1 with --rm-invar-sites 1

    def feed(self, text: np.ndarray, sites: np.ndarray) -> Iterator[SiteRun]:
        """`sites`: vgl_in_site records of one parsed chunk, `text`: the chunk (for the CHROM column)"""
        n = len(sites)
        i = 0
        off = sites["line_off"].astype(np.int64)
        while i < n:
            # records i..j-1 share a contig: compare the line prefix with "<contig>\t"
            if self.contig is not None:
                key = np.frombuffer(self.contig + b"\t", np.uint8)
                idx = off[i:, None] + np.arange(len(key))[None, :]
                same = (text[np.minimum(idx, len(text) - 1)] == key[None, :]).all(axis=1)
                j = i + (int(np.argmin(same)) if not same.all() else n - i)
            else:
                j = i
            if j == i:                             # contig change (vcfgl.cpp:1484-1488)
                lo = int(off[i])
                tab = lo + int(np.argmax(text[lo:lo + 4096] == 9))
                self.contig = text[lo:tab].tobytes()
                self.n_in_contig = 0
                continue
            yield from self._run(sites[i:j], i)
            i = j

    def feed_rids(self, rids: np.ndarray, names: List[str], sites: np.ndarray) -> Iterator[SiteRun]:
        """the same for BCF records: the contig of a record is its rid (bcf1_t::rid), names[rid] its name"""
        n = len(sites)
        i = 0
        while i < n:
            name = names[int(rids[i])].encode()
            if self.contig != name:
                self.contig = name
                self.n_in_contig = 0
            j = i + 1
            while j < n and rids[j] == rids[i]:
                j += 1
            yield from self._run(sites[i:j], i)
            i = j

    def _run(self, recs: np.ndarray, first_index: int) -> Iterator[SiteRun]:
        contig = self.contig.decode()
        pos = recs["pos"].astype(np.int64)
        keep = recs["skip_code"] == 0
        self.n_skipped += int((~keep).sum())
        idx = np.arange(first_index, first_index + len(recs), dtype=np.int32)
        if not self.explode:
            p, s = pos[keep], idx[keep]
            for a in range(0, len(p), self.max_run):
                yield SiteRun(contig, p[a:a + self.max_run], s[a:a + self.max_run])
            self.n_in_contig += len(recs)
            return
        if pos[0] < self.n_in_contig or (len(pos) > 1 and (np.diff(pos) <= 0).any()):
            raise VcfInputError("-explode 1 needs positions in increasing order within a contig "
                                "(the reference never leaves its explode loop otherwise, vcfgl.cpp:1492-1496)")
        if self.fill_acgt < 0 and pos[0] != self.n_in_contig:       # explode_rec is a copy of THIS record (vcfgl.cpp:1498-1503)
            self.fill_acgt = int(recs["allele_acgt"][0][0])
        if self.fill_acgt < 0 and len(pos) > 1 and (np.diff(pos) > 1).any():
            k = int(np.argmax(np.diff(pos) > 1)) + 1
            self.fill_acgt = int(recs["allele_acgt"][k][0])
        yield from self._explode_range(contig, self.n_in_contig, int(pos[-1]) + 1, pos, idx, keep)
        self.n_in_contig = int(pos[-1]) + 1

    def _explode_range(self, contig, x0, x1, pos, idx, keep) -> Iterator[SiteRun]:
        blank_kept = self._explode_sites_kept()
        for a in range(x0, x1, self.max_run):
            b = min(a + self.max_run, x1)
            p = np.arange(a, b, dtype=np.int64)
            s = np.full(b - a, -1, np.int32)
            ok = np.ones(b - a, bool) if blank_kept else np.zeros(b - a, bool)
            if len(pos):
                lo, hi = np.searchsorted(pos, a), np.searchsorted(pos, b)
                s[pos[lo:hi] - a] = idx[lo:hi]
                ok[pos[lo:hi] - a] = keep[lo:hi]
            if not blank_kept:
                self.n_skipped += int((s < 0).sum())
            if ok.any():
                yield SiteRun(contig, p[ok], s[ok])

    def finish(self, last_record_acgt0: int = -1) -> Iterator[SiteRun]:
        """after the last record: -explode 1 runs to the end of the last contig (vcfgl.cpp:1567-1611)"""
        if not self.explode or self.contig is None:
            return
        contig = self.contig.decode()
        size = self.contigs.get(contig, 0)
        if self.n_in_contig > size:
            raise VcfInputError("positions beyond the contig length with -explode 1 (the reference does not terminate)")
        if self.n_in_contig == size:
            return
        if self.fill_acgt < 0:
            self.fill_acgt = last_record_acgt0
        e = np.zeros(0, np.int64)
        yield from self._explode_range(contig, self.n_in_contig, size, e, e.astype(np.int32), np.zeros(0, bool))
        self.n_in_contig = size


def raise_first_error(text: np.ndarray, res: capi.ParseResult):
    i = res.first_error_record
    s = res.sites[i]
    line = text[int(s["line_off"]):int(s["line_off"]) + min(int(s["line_len"]), 80)].tobytes().decode(errors="replace")
    raise VcfInputError("%s at position %d (record %d of the chunk: %r)"
                        % (IN_STATUS_TEXT.get(int(s["status"]), "status %d" % s["status"]), int(s["pos"]) + 1, i, line))


def simulate_vcf_text(ctx: capi.Context, parser: capi.Parser, body: bytes, *, gt_source: int, explode: int,
                      contigs: Dict[str, int], chunk_bytes: Optional[int] = None,
                      first_site_id: int = 0) -> Iterator[Tuple[SiteRun, capi.Batch]]:
    """Drive text -> parse -> place -> simulate for a whole VCF body.  Yields (sites of the batch, finished batch) in the
    reference's output order; slots alternate so that parsing / placing the next batch overlaps the previous batch's
    kernels and copies."""
    cap = int(parser.text.shape[0])
    assert parser.S == ctx.S
    chunk_bytes = min(chunk_bytes or cap, cap)
    rm_invar = int(ctx.params.rm_invar_sites) & 3
    planner = SitePlanner(explode, rm_invar, contigs, ctx.cap)
    n_slots = int(ctx.params.n_slots)
    pending: List[Tuple[int, SiteRun]] = []
    site_id = first_site_id
    slot = 0
    mv = np.frombuffer(body, np.uint8)
    off = 0
    last_acgt0 = -1

    def drain(keep: int):
        while len(pending) > keep:
            sl, run = pending.pop(0)
            run.slot = sl
            yield run, ctx.wait(sl)

    def submit(run: SiteRun):
        nonlocal slot, site_id
        yield from drain(n_slots - 1)
        fill = (planner.fill_acgt & 0xF) * 0x11 if planner.fill_acgt >= 0 else 0
        ctx.place_rows(slot, parser, len(run.src), row_map=run.src, fill_gt=fill)
        ctx.submit(slot, site_id, len(run.src), flags=capi.SUBMIT_GT_ON_DEVICE)
        pending.append((slot, run))
        site_id += len(run.src)
        slot = (slot + 1) % n_slots

    while off < len(mv):
        n = min(chunk_bytes, len(mv) - off)
        final = off + n == len(mv)
        res = parser.parse(mv[off:off + n], gt_source, capi.PARSE_FINAL if final else 0)
        if res.n_errors:
            raise_first_error(parser.text, res)
        if res.n_records == 0:
            raise VcfInputError("a record does not fit the parser's text buffer (%d bytes)" % cap)
        sites = res.sites.copy()
        last_acgt0 = int(sites["allele_acgt"][-1][0])
        for run in planner.feed(parser.text, sites):
            yield from submit(run)
        off += int(res.bytes_consumed)
        # the rows of this parse must be placed before the next parse overwrites them: vgl_parse_vcf orders itself after
        # the last vgl_place_rows on the device, nothing to do here
    for run in planner.finish(last_acgt0):
        if (run.src >= 0).any():
            raise AssertionError("tail sites cannot reference records")
        yield from submit(run)
    yield from drain(0)


def bcf_record_offsets(body) -> np.ndarray:
    """offsets [n + 1] of the records in the bytes after a BCF header: hop l_shared + l_indiv + 8 (htslib/vcf.c:1456-1500)"""
    import struct
    off, o, n = [0], 0, len(body)
    while o + 8 <= n:
        l_shared, l_indiv = struct.unpack_from("<II", body, o)
        if o + 8 + l_shared + l_indiv > n:
            break                                   # incomplete record: carry it over to the next chunk
        o += 8 + l_shared + l_indiv
        off.append(o)
    return np.array(off, np.uint32)


def simulate_bcf_records(ctx: capi.Context, parser: capi.Parser, body: bytes, *, gt_key: int, gt_source: int, explode: int,
                         contig_names: List[str], contig_lengths: Dict[str, int], first_site_id: int = 0,
                         max_records_per_chunk: Optional[int] = None) -> Iterator[Tuple[SiteRun, capi.Batch]]:
    """simulate_vcf_text for uncompressed BCF records (the bytes after the header, BGZF already inflated)."""
    assert parser.S == ctx.S
    rm_invar = int(ctx.params.rm_invar_sites) & 3
    planner = SitePlanner(explode, rm_invar, contig_lengths, ctx.cap)
    n_slots = int(ctx.params.n_slots)
    pending: List[Tuple[int, SiteRun]] = []
    site_id, slot = first_site_id, 0
    mv = np.frombuffer(body, np.uint8)
    off = bcf_record_offsets(body)
    per = max_records_per_chunk or ctx.cap
    last_acgt0 = -1

    def drain(keep: int):
        while len(pending) > keep:
            sl, run = pending.pop(0)
            run.slot = sl
            yield run, ctx.wait(sl)

    for r0 in range(0, len(off) - 1, per):
        r1 = min(r0 + per, len(off) - 1)
        lo, hi = int(off[r0]), int(off[r1])
        res = parser.parse_bcf(mv[lo:hi], off[r0:r1 + 1] - off[r0], gt_key, gt_source)
        if res.n_errors:
            s = res.sites[res.first_error_record]
            raise VcfInputError("%s at position %d (record %d)" % (IN_STATUS_TEXT.get(int(s["status"]), "status %d" % s["status"]),
                                                                    int(s["pos"]) + 1, r0 + res.first_error_record))
        sites = res.sites.copy()
        last_acgt0 = int(sites["allele_acgt"][-1][0])
        rids = mv[lo:hi].view(np.uint8)
        rid = np.array([int.from_bytes(rids[int(o) + 8:int(o) + 12].tobytes(), "little", signed=True) for o in sites["line_off"]], np.int32)
        for run in planner.feed_rids(rid, contig_names, sites):
            yield from drain(n_slots - 1)
            fill = (planner.fill_acgt & 0xF) * 0x11 if planner.fill_acgt >= 0 else 0
            ctx.place_rows(slot, parser, len(run.src), row_map=run.src, fill_gt=fill)
            ctx.submit(slot, site_id, len(run.src), flags=capi.SUBMIT_GT_ON_DEVICE)
            pending.append((slot, run))
            site_id += len(run.src)
            slot = (slot + 1) % n_slots
    for run in planner.finish(last_acgt0):
        yield from drain(n_slots - 1)
        fill = (planner.fill_acgt & 0xF) * 0x11 if planner.fill_acgt >= 0 else 0
        ctx.place_rows(slot, parser, len(run.src), row_map=run.src, fill_gt=fill)
        ctx.submit(slot, site_id, len(run.src), flags=capi.SUBMIT_GT_ON_DEVICE)
        pending.append((slot, run))
        site_id += len(run.src)
        slot = (slot + 1) % n_slots
    yield from drain(0)
