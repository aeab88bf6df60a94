"""N>1 host logic on CPU: two gloo ranks agree on a tiling of the site range, on the merge order,
and on the max-over-ranks timing reduction bench.py uses."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vcfgl_b200 import sharding


def test_shard_range_tiles():
    for n in (0, 1, 7, 100, 1048576):
        for w in (1, 2, 3, 8):
            rs = [sharding.shard_range(n, w, r) for r in range(w)]
            assert sharding.merge_order(rs) == list(range(w))
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1
    assert list(sharding.batches(5, 12, 3)) == [(5, 3), (8, 3), (11, 1)]
    with pytest.raises(ValueError):
        sharding.merge_order([(0, 5), (6, 9)])


def _worker(rank, world, port, n_sites, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(n_sites, world, rank)
    got = [None] * world
    dist.all_gather_object(got, (lo, hi))
    order = sharding.merge_order(got)
    # per-rank "elapsed" -> max over ranks, whole-job throughput (bench.py)
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cells = torch.tensor([float((hi - lo) * 100)])
    dist.all_reduce(cells, op=dist.ReduceOp.SUM)
    q.put((rank, got, order, float(t.item()), float(cells.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, n_sites = 2, 1001
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_sites, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, got, order, tmax, cells in res:
        assert got == [(0, 501), (501, 1001)]
        assert order == [0, 1]
        assert tmax == 2.0
        assert cells == 1001 * 100


def _text_worker(rank, world, port, q):
    """sharded input path, host logic: byte ranges cut at line starts, global site ids from one all_gather"""
    import numpy as np
    import vcfin_oracle as vo
    from vcfgl_b200 import vcfinput
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    buf = vo.load_input("s40.in.vcf")
    hdr = vcfinput.read_header(buf)
    body = buf[hdr.body_offset:]
    lo, hi = sharding.shard_text(body, world, rank)
    sites, rows, used = vo.parse(body[lo:hi], len(hdr.samples), 0, rm_invar=1)       # (the device parser's CPU stand-in)
    n_kept = int(((sites["status"] == 0) & (sites["skip_code"] == 0)).sum())
    got = [None] * world
    dist.all_gather_object(got, (lo, hi, n_kept))
    first = sharding.site_id_offsets([g[2] for g in got])[rank]
    q.put((rank, got, first, sites["pos"].tolist(), rows.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_text_sharding():
    import numpy as np
    import vcfin_oracle as vo
    from vcfgl_b200 import vcfinput
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_text_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    buf = vo.load_input("s40.in.vcf")
    hdr = vcfinput.read_header(buf)
    body = buf[hdr.body_offset:]
    sites, rows, _ = vo.parse(body, len(hdr.samples), 0, rm_invar=1)
    ranges = [(g[0], g[1]) for g in res[0][1]]
    assert ranges[0][0] == 0 and ranges[-1][1] == len(body) and ranges[0][1] == ranges[1][0]
    assert all(body[lo - 1:lo] == b"\n" for lo, _ in ranges[1:])
    assert sum((r[3] for r in res), []) == sites["pos"].tolist()
    assert b"".join(r[4] for r in res) == rows.tobytes()
    kept = (sites["status"] == 0) & (sites["skip_code"] == 0)
    assert [r[2] for r in res] == [0, int(kept[:len(res[0][3])].sum())]


def test_shard_text_edges():
    body = b"a\nbb\nccc\n"
    for w in (1, 2, 3, 5, 11):
        rs = [sharding.shard_text(body, w, r) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == len(body)
        assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        assert all(lo == 0 or body[lo - 1:lo] == b"\n" for lo, _ in rs)
    assert sharding.shard_text(b"", 2, 1) == (0, 0)
    assert sharding.site_id_offsets([3, 0, 5]) == [0, 3, 3]
