"""Host-side logic of the data-parallel path on CPU: world_size-2 gloo processes exercise the bucketed gradient
all-reduce (s2vt_b200.dp.GradAllReducer), the bucket layout over a flat buffer, and the video sharding used by the
beam-search eval.  No CUDA kernels are involved (the reducer is device-agnostic)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import s2vt_b200
from s2vt_b200 import dp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _layout():
    m = s2vt_b200.S2VT(50, 24, 6, dim_hid=16, dim_embed=12)
    names = [n for n, _ in m.named_parameters()]
    sizes = [p.numel() for p in m.parameters()]
    offsets, n = [], 0
    for s in sizes:
        offsets.append(n)
        n += (s + 3) // 4 * 4
    return names, offsets, sizes, n


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    names, offsets, sizes, n = _layout()
    ranges = dp.bucket_ranges(names, offsets, sizes)
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(n, generator=g)
    mine = flat.clone()
    red = dp.GradAllReducer(flat, ranges, overlap=True)       # overlap silently off on CPU
    # backward announces buckets in production order; finish() must pick up whatever was not announced
    red.ready("out_linear")
    red.ready("word_rnn")
    red.finish()
    others = [torch.randn(n, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
    expect = sum(others) / world
    covered = torch.zeros(n, dtype=torch.bool)
    for a, b in ranges.values():
        covered[a:b] = True
    ok = torch.allclose(flat[covered], expect[covered], atol=1e-6) and torch.equal(flat[~covered], mine[~covered])
    # a second round reuses the reducer
    flat.copy_(mine)
    red.finish()
    ok = ok and torch.allclose(flat[covered], expect[covered], atol=1e-6)
    q.put((rank, bool(ok), red.bytes_reduced))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    names, offsets, sizes, n = _layout()
    assert all(2 * 4 * sum(sizes) <= b <= 2 * 4 * n for _, _, b in res)    # every parameter byte reduced once per round (+ alignment padding inside a bucket)


def test_bucket_ranges_cover_all_parameters_once():
    names, offsets, sizes, n = _layout()
    ranges = dp.bucket_ranges(names, offsets, sizes)
    assert list(ranges) == ["out_linear", "word_rnn", "embedding", "vid_rnn", "feat_linear"]
    spans = sorted(ranges.values())
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0                                        # disjoint
    total = sum(b - a for a, b in spans)
    assert sum(sizes) <= total <= n
    with pytest.raises(ValueError):                            # a bucket split by a foreign tensor is rejected
        bad = list(offsets)
        bad[names.index("out_linear.bias")] += 1000
        dp.bucket_ranges(names, bad, sizes)


def test_shard_range_1970_videos_over_8_ranks():
    parts = [dp.shard_range(1970, r, 8) for r in range(8)]
    assert parts[0] == (0, 247) and parts[-1][1] == 1970
    assert [b - a for a, b in parts] == [247, 247, 246, 246, 246, 246, 246, 246]
    assert all(parts[i][1] == parts[i + 1][0] for i in range(7))
    assert dp.shard_range(5, 7, 8) == (5, 5)                   # more ranks than items: empty shard
