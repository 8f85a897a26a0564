"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes exercise the sharding and the descriptor
all-gather (padding / counts / per-peer blocks); results are checked against the oracle's matcher."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vulkansift_b200.dist import all_pairs_schedule, exchange_descriptor_blocks, shard_range
    import oracle
    # ragged blocks: rank 0 holds 37 descriptors, rank 1 holds 101
    n = [37, 101][rank]
    desc = np.random.default_rng(100 + rank).integers(0, 256, (n, 128), dtype=np.uint8)
    counts, blocks = exchange_descriptor_blocks(torch.from_numpy(desc))
    assert counts == [37, 101] and tuple(blocks.shape) == (2, 101, 128)
    assert np.array_equal(blocks[rank, :n].numpy(), desc)
    assert int(blocks[0, 37:].sum()) == 0  # padding rows are zero
    # single-collective form: fixed-capacity slots with the count in the last row
    counts1, blocks1 = exchange_descriptor_blocks(torch.from_numpy(desc), capacity=128)
    assert counts1 == [37, 101] and tuple(blocks1.shape) == (2, 129, 128)
    for j in range(2):
        assert torch.equal(blocks1[j, :counts1[j]], blocks[j, :counts[j]])
        assert int(blocks1[j, counts1[j]:128].sum()) == 0
    res = {}
    for j in all_pairs_schedule(rank, world):
        peer = blocks[j, :counts[j]].numpy()
        res[j] = oracle.match_descriptors(desc, peer, 1)
    np.save(os.path.join(out_dir, "m%d.npy" % rank), res[1 - rank])
    np.save(os.path.join(out_dir, "d%d.npy" % rank), desc)
    # image sharding covers every item exactly once
    b, e = shard_range(7, rank, world)
    t = torch.zeros(7, dtype=torch.int64)
    t[b:e] = 1
    dist.all_reduce(t)
    assert t.tolist() == [1] * 7
    dist.barrier()
    dist.destroy_process_group()


def test_descriptor_exchange_world2_gloo(tmp_path, oracle_mod):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    d0, d1 = np.load(tmp_path / "d0.npy"), np.load(tmp_path / "d1.npy")
    m0, m1 = np.load(tmp_path / "m0.npy"), np.load(tmp_path / "m1.npy")
    assert m0.tobytes() == oracle_mod.match_descriptors(d0, d1, 1).tobytes()
    assert m1.tobytes() == oracle_mod.match_descriptors(d1, d0, 1).tobytes()


def test_shard_range_is_a_partition():
    from vulkansift_b200.dist import all_pairs_schedule, shard_range
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                b, e = shard_range(n, r, world)
                assert 0 <= b <= e <= n
                cover += list(range(b, e))
            assert cover == list(range(n))
            sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert all_pairs_schedule(2, 4) == [3, 0, 1]
    assert all_pairs_schedule(0, 1) == []
