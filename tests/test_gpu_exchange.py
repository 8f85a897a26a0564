"""The library's descriptor exchange over peer memory (csrc/exchange.cu, vksiftx_exchange*): three processes (gloo only carries
the 64-byte IPC handles and the barriers) share ONE GPU here -- the protocol (push into the peers' slots, flag, wait, zero the
padding, ONE search against all received blocks in place, double buffering over several rounds with ragged and degenerate block
sizes, so that slots hold stale rows of earlier, larger blocks) is the same as between GPUs over NVLink; bench.py --gpus N checks
the real multi-GPU case against the oracle on every rank."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from vulkansift_b200 import api as a
    a.load()
    a.lib.vksift_setLogLevel(a.VKSIFT_LOG_WARNING)
    return a

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROUNDS = 5
WORLD = 3
COUNTS = [[700, 900, 300], [1024, 2, 513], [1, 333, 640], [513, 640, 129], [40, 300, 170]]  # per round, per rank (a block of one row is not matched against)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _desc(rnd, rank):
    sys.path.insert(0, ROOT)
    from vulkansift_b200.synth import random_descriptors
    n = COUNTS[rnd][rank]
    if rnd == 4:
        # distances beyond d^2 = 2^22 with float-sqrt collisions (test_large_distances_follow_the_float_comparison): rank 0 against the
        # others goes through the float-order rescan of the grouped search
        rng = np.random.default_rng(40 + rank)
        d = np.zeros((n, 128), np.uint8)
        if rank > 0:
            d[:, :73] = 255
            d[:, 73] = 68
        d[:, 127] = rng.integers(0, 2, n)
        return d
    return random_descriptors(n, 1000 + 10 * rnd + rank)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vulkansift_b200 import api
    from vulkansift_b200.dist import PeerExchange
    api.load()
    api.lib.vksift_setLogLevel(api.VKSIFT_LOG_WARNING)
    with api.Instance(max_nb_sift_per_buffer=1024, input_image_max_size=1 << 20, gpu_device_index=0) as inst:
        px = PeerExchange(inst, 1024)
        for rnd in range(ROUNDS):
            d = _desc(rnd, rank)
            f = np.zeros(len(d), api.FEATURE_DTYPE)
            f["descriptor"] = d
            inst.upload_features(f, 0)
            counts, res = px.match_all_peers(0)
            assert counts == COUNTS[rnd], (counts, COUNTS[rnd])
            for peer in range(world):
                if peer == rank:
                    continue
                if res[peer] is None:
                    assert counts[peer] < 2
                    np.save(os.path.join(out_dir, "m%d_%d_%d.npy" % (rnd, rank, peer)), np.zeros(0, api.MATCH_DTYPE))
                else:
                    np.save(os.path.join(out_dir, "m%d_%d_%d.npy" % (rnd, rank, peer)), res[peer])
        # the exchange alone: counts and the peer's rows, in place
        counts, ptr, stride = px.allgather(0)
        assert counts == COUNTS[ROUNDS - 1] and stride == 1024 * 128 and ptr % 128 == 0
        px.close()
    dist.barrier()
    dist.destroy_process_group()


def test_peer_memory_exchange_three_processes_match_oracle(tmp_path, oracle_mod):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(WORLD, _free_port(), str(tmp_path)), nprocs=WORLD, join=True)
    for rnd in range(ROUNDS):
        d = [_desc(rnd, r) for r in range(WORLD)]
        for rank in range(WORLD):
            for peer in range(WORLD):
                if peer == rank:
                    continue
                got = np.load(tmp_path / ("m%d_%d_%d.npy" % (rnd, rank, peer)))
                if len(d[peer]) < 2:
                    assert len(got) == 0
                    continue
                exp = oracle_mod.match_descriptors(d[rank], d[peer])
                assert got.dtype == exp.dtype and got.tobytes() == exp.tobytes(), "round %d rank %d peer %d" % (rnd, rank, peer)


def test_exchange_reports_a_missing_peer_instead_of_hanging(api):
    """world_size 2 with nobody on the other side: the wait gives up after two seconds and the error callback fires."""
    with api.Instance(max_nb_sift_per_buffer=256, input_image_max_size=1 << 20) as inst:
        h = inst.exchange_create(0, 2, 256)
        # "connect" to a second region in the same process is not possible (a process cannot open its own handle), so the
        # unconnected exchange must refuse to run rather than push to a null pointer
        f = np.zeros(10, api.FEATURE_DTYPE)
        inst.upload_features(f, 0)
        with pytest.raises(api.VksiftError):
            inst.exchange_allgather(0)
        assert len(h) == 64
        inst.exchange_destroy()
