"""Multi-process host logic of the sharded modes on CPU: world_size 2, gloo.

The compute kernels need a GPU; what is covered here is everything around them that a 2..8-GPU run
depends on -- the batch / row-band partition, the single collective that gathers the results, and the
property that makes row-band sharding legal at all (the correlation path is independent per epipolar
row), checked with the oracle as the stand-in compute.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nndepth_b200.engine import Padder, gather_disparities, gather_row_bands, row_band, shard_range
from oracle import corr1d as oc


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    manager = mp.Manager()
    ret = manager.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def test_shard_range_partitions_exactly():
    for total in (1, 7, 8, 17, 136, 384):
        for world in (1, 2, 3, 4, 8):
            cuts = [shard_range(total, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            for (a0, a1), (b0, b1) in zip(cuts, cuts[1:]):
                assert a1 == b0
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_padder_matches_reference_arithmetic():
    # reference dataloaders/utils.py:9-15: pad = (((d // k) + 1) * k - d) % k
    for (h, w) in ((375, 1242), (384, 1248), (1080, 1920), (33, 65)):
        p = Padder((h, w), 32)
        ph, pw = (((h // 32) + 1) * 32 - h) % 32, (((w // 32) + 1) * 32 - w) % 32
        assert (p.top, p.bottom, p.left, p.right) == (0, ph, pw // 2, pw - pw // 2)
        x = torch.arange(h * w, dtype=torch.float32).view(1, 1, h, w)
        (xp,) = p.pad(x)
        assert xp.shape[-2] % 32 == 0 and xp.shape[-1] % 32 == 0
        assert torch.equal(p.unpad(xp), x)


def _batch_shard(rank, world):
    # 5 "pairs" over 2 ranks would be uneven: the engine's gather takes equal shards (8 pairs / N GPUs)
    torch.manual_seed(0)
    full = torch.randn(8, 1, 6, 10)
    b0, b1 = shard_range(8, rank, world)
    local = full[b0:b1] * 1.0                      # stand-in for this rank's forward
    out = gather_disparities(local, world)
    return bool(torch.equal(out, full))


def test_batch_shard_gather_world2():
    assert _run(_batch_shard) == [True, True]


def _row_band_corr(rank, world):
    rng = np.random.default_rng(5)
    B, C, H, W = 1, 16, 7, 24                      # 7 rows over 2 ranks: bands of 4 and 3
    f1 = rng.standard_normal((B, C, H, W), dtype=np.float32)
    f2 = rng.standard_normal((B, C, H, W), dtype=np.float32)
    coords = (np.broadcast_to(np.arange(W, dtype=np.float32), (B, 1, H, W)) - rng.uniform(0, 9, (B, 1, H, W))).astype(np.float32)
    full = oc.CorrBlock1D(f1, f2, 3, 3)(coords)
    b1, (h0, h1) = row_band(torch.from_numpy(f1), rank, world)
    b2, _ = row_band(torch.from_numpy(f2), rank, world)
    bc, _ = row_band(torch.from_numpy(coords), rank, world)
    local = oc.CorrBlock1D(b1.numpy(), b2.numpy(), 3, 3)(bc.numpy())
    gathered = gather_row_bands(torch.from_numpy(local), H, world)
    return bool(np.array_equal(gathered.numpy(), full)), (h0, h1)


def test_row_band_sharding_is_exact_world2():
    res = _run(_row_band_corr)
    assert [r[0] for r in res] == [True, True]
    assert [r[1] for r in res] == [(0, 4), (4, 7)]
