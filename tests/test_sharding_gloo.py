"""N>1 host path on CPU: two gloo ranks shard a frame sequence, run the per-frame work on their
block (here the CPU oracle stands in for the device, the test is about the sharding/gather logic),
all-gather the per-frame digests, and the result must equal the single-process run."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dr-slam_b200"))
sys.path.insert(0, ROOT)

TOTAL = 5   # odd on purpose: blocks of 3 and 2


def frame_digest(i):
    import drfe
    from oracle import oracle as orc
    gray, depth, K = drfe.synth_frame(320, 240, i % 3, 20260000 + i)
    kps, desc = orc.OrbOracle(300).extract(gray)
    return [i, len(kps), int(desc.astype(np.int64).sum()), int(kps["x"].sum())]


def worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard.frame_block(TOTAL, rank, world)
    mine = [frame_digest(i) for i in range(a, b)]
    pad = (TOTAL + world - 1) // world
    t = torch.full((pad, 4), -1, dtype=torch.int64)
    if mine:
        t[:len(mine)] = torch.tensor(mine, dtype=torch.int64)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    dist.barrier()
    if rank == 0:
        rows = torch.cat(parts)
        rows = rows[rows[:, 0] >= 0]
        np.save(out, rows.numpy())
    dist.destroy_process_group()


def test_frame_blocks_partition():
    import shard
    for total in (0, 1, 5, 256, 257):
        for world in (1, 2, 4, 8):
            blocks = [shard.frame_block(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert shard.weak_block(256, 3) == (768, 1024)
    assert shard.owner_of(3, 5, 2) == 1
    with pytest.raises(ValueError):
        shard.frame_block(4, 2, 2)


def test_two_rank_gloo_matches_single_process(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "gathered.npy")
    mp.spawn(worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    want = np.array([frame_digest(i) for i in range(TOTAL)], dtype=np.int64)
    assert np.array_equal(got[np.argsort(got[:, 0])], want)
