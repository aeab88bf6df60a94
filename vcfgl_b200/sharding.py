"""Site-range sharding over GPUs.

Every site is independent once the RNG is counter-based and keyed by the global site id, so the
path shards by contiguous site ranges, one per rank, with NO collective on the data path; the host
merges the shards' records in rank order (which is site order).  The reference itself cannot be
sharded reproducibly: its RNG streams are sequential over sites (vcfgl.cpp:214-219)."""
from __future__ import annotations

from typing import Iterator, List, Tuple


def shard_range(n_sites: int, world: int, rank: int) -> Tuple[int, int]:
    """contiguous [lo, hi) of rank; sizes differ by at most one site"""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n_sites, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def batches(lo: int, hi: int, batch_sites: int) -> Iterator[Tuple[int, int]]:
    """(first_site_id, n_sites) batches covering [lo, hi)"""
    s = lo
    while s < hi:
        n = min(batch_sites, hi - s)
        yield s, n
        s += n


def merge_order(ranges: List[Tuple[int, int]]) -> List[int]:
    """ranks in the order their records must be written; checks the ranges tile [0, n) exactly"""
    order = sorted(range(len(ranges)), key=lambda r: ranges[r][0])
    pos = 0
    for r in order:
        lo, hi = ranges[r]
        if lo != pos or hi < lo:
            raise ValueError("shards do not tile the site range: %r" % (ranges,))
        pos = hi
    return order


def shard_text(body, world: int, rank: int) -> Tuple[int, int]:
    """Byte range [lo, hi) of a VCF body (record lines) for rank: cut at the first line start at or after rank * len / world,
    so that the ranges tile the body and every record belongs to exactly one rank.  `body`: bytes-like with .find()."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    n = len(body)

    def line_start_at_or_after(p: int) -> int:
        if p <= 0:
            return 0
        if p >= n:
            return n
        nl = body.find(b"\n", p - 1)
        return n if nl < 0 else nl + 1
    return line_start_at_or_after(rank * n // world), line_start_at_or_after((rank + 1) * n // world)


def site_id_offsets(n_sites_per_rank: List[int]) -> List[int]:
    """first global site id of every rank = the sites of the ranks before it (what one all_gather of a single integer per rank
    gives; the only exchange of the sharded input path, and it is host metadata, not on the data path).  Without -explode the
    sites of a rank are the records it keeps; with -explode 1 the positions between two ranks' records belong to the later rank,
    which needs the earlier rank's last (contig, position) as its planner's start state -- pass it the same way."""
    out, acc = [], 0
    for k in n_sites_per_rank:
        out.append(acc)
        acc += int(k)
    return out
