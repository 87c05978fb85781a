"""Multi-rank host logic on CPU: world_size-2 gloo groups (the N > 1 path of bench.py / entry.get_distribution
without a GPU): shard ranges, the symbol-table all-reduce, and byte-identical per-rank bitstreams."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tf_image_compression_b200 import entry, parallel, range_coder


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, tmp, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the dataset: 7 "images" of symbols (seeded, same on every rank); each rank encodes its shard
        rs = np.random.RandomState(42)
        images = [(rs.rand(rs.randint(2000, 6000)) < 0.2).astype(np.uint8) for _ in range(7)]
        b, e = parallel.shard_range(len(images))
        local = np.zeros(2, np.int64)
        for im in images[b:e]:
            local += np.bincount(im, minlength=2)
        total = parallel.allreduce_counts(local)
        prob = parallel.distribution(total)
        cum = entry.cum_freq_table(prob, 4096)
        sizes = []
        for i in range(b, e):
            path = os.path.join(tmp, f"img{i}.encoded")
            enc = range_coder.RangeEncoder(path)
            enc.encode(images[i], cum)
            enc.close()
            sizes.append(enc.bytes_written)
        q.put((rank, (b, e), total.tolist(), cum, sizes))
    finally:
        dist.destroy_process_group()


def test_shard_ranges_cover_in_order():
    for n in (0, 1, 7, 64, 65):
        for w in (1, 2, 3, 8):
            r = [parallel.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_table_allreduce_and_bitstreams(tmp_path):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the single-process answer
    rs = np.random.RandomState(42)
    images = [(rs.rand(rs.randint(2000, 6000)) < 0.2).astype(np.uint8) for _ in range(7)]
    total = sum(np.bincount(im, minlength=2) for im in images)
    cum = entry.cum_freq_table(total / total.sum(), 4096)
    assert [r[1] for r in res] == [(0, 4), (4, 7)]
    for rank, _, tot, c, _ in res:
        assert tot == total.tolist() and c == cum  # every rank holds the dataset-wide table
    for i, im in enumerate(images):  # byte-identical to a one-rank run
        ref = tmp_path / f"ref{i}.encoded"
        enc = range_coder.RangeEncoder(str(ref))
        enc.encode(im, cum)
        enc.close()
        assert ref.read_bytes() == (tmp_path / f"img{i}.encoded").read_bytes()


def test_single_process_fallbacks():
    assert parallel.world() == (0, 1)
    assert parallel.allreduce_counts([3, 4]).tolist() == [3, 4]
    assert np.allclose(parallel.distribution([1, 3]), [0.25, 0.75])
