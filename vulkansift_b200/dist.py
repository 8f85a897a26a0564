"""Multi-GPU host logic: image sharding and the descriptor exchange of cross-image matching.

One process per GPU (torchrun).  Detection shards by image with no data-path collective
(SURVEY 8e); all-pairs matching has exactly one exchange step: every rank contributes its
descriptor block and receives everyone else's.  Two implementations:

* `PeerExchange` (the product path on a GPU node): the library's own exchange over NVLink peer memory
  (csrc/exchange.cu: every rank pushes its block into a slot on every peer with one kernel, flags in peer
  memory, blocks matched in place).  torch.distributed only carries the 64-byte IPC handles once, at setup.
* `gather_instance_descriptors` + `match_against_peers_batched`: one NCCL all-gather plus a count read-back
  (round 1; kept as the baseline `bench.py` times next to the peer-memory path).  These functions only move
  tensors, so the same code runs on CPU tensors with the gloo backend in the tests.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced [begin, end) of n_items for `rank` (first n_items % world ranks get one more)."""
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def exchange_descriptor_blocks(local_desc, group=None, capacity=None):
    """All-gather of per-rank descriptor blocks of different lengths.

    local_desc: (n_local, 128) uint8 tensor on the backend's device.
    Returns (counts, blocks): counts[j] = rows contributed by rank j; blocks is (world, rows, 128),
    rank j's descriptors are blocks[j, :counts[j]] and the padding rows are zero.
    capacity=None: two collectives, 8-byte counts, then blocks padded to the largest count (0.4 MB/rank at ~3k features).
    capacity=c (every rank holds at most c rows and passes the same c): ONE collective of (c+1)-row slots whose last row
    carries the count, so the exchange costs one launch and one host read instead of two of each; the step is latency
    bound (0.4 MB per rank against 900 GB/s per direction of NVLink), so the extra padding rows are free.
    """
    world = dist.get_world_size(group)
    dev = local_desc.device
    assert local_desc.dtype == torch.uint8 and local_desc.dim() == 2 and local_desc.shape[1] == 128
    if capacity is not None:
        n = local_desc.shape[0]
        assert n <= capacity, "descriptor block of %d rows exceeds the exchange capacity %d" % (n, capacity)
        send = torch.zeros((capacity + 1, 128), dtype=torch.uint8, device=dev)
        send[:n] = local_desc
        send[capacity, :8] = torch.tensor([n], dtype=torch.int64).view(torch.uint8).to(dev)
        blocks = torch.empty((world, capacity + 1, 128), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(blocks.view(world * (capacity + 1), 128), send, group=group)
        counts = [int(c) for c in blocks[:, capacity, :8].contiguous().view(torch.int64).flatten().tolist()]
        return counts, blocks
    n_local = torch.tensor([local_desc.shape[0]], dtype=torch.int64, device=dev)
    counts_t = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts_t, n_local, group=group)
    counts = [int(c) for c in counts_t.tolist()]
    max_n = max(max(counts), 1)
    send = torch.zeros((max_n, 128), dtype=torch.uint8, device=dev)
    send[:local_desc.shape[0]] = local_desc
    blocks = torch.empty((world, max_n, 128), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(blocks.view(world * max_n, 128), send, group=group)
    return counts, blocks


def all_pairs_schedule(rank, world):
    """Peers whose block this rank matches its own features against: every j != rank, starting after
    the own rank so that the ranks do not all read the same block at the same time."""
    return [(rank + d) % world for d in range(1, world)]


class _ExchangeBuffers:
    """Persistent send slot / receive buffer of the single-collective exchange (allocated once per (device, world, capacity))."""

    def __init__(self, dev, world, capacity):
        self.capacity = capacity
        self.send = torch.zeros((capacity + 1, 128), dtype=torch.uint8, device=dev)
        self.recv = torch.empty((world, capacity + 1, 128), dtype=torch.uint8, device=dev)
        self.count_host = torch.zeros(1, dtype=torch.int64).pin_memory()
        self.counts_host = torch.zeros((world, 8), dtype=torch.uint8).pin_memory()  # the in-band counts of all ranks, read back with one strided copy


_exchange_buffers = {}


def gather_instance_descriptors(inst, buffer_id, group=None, capacity=None):
    """CUDA glue: all-gather the descriptors of one feature buffer of a vulkansift_b200.api.Instance.

    capacity=None: the local block is copied device-to-device into a torch tensor by vksiftx_copyDescriptorsToDevice and
    exchanged with two collectives (counts, then blocks padded to the largest count).
    capacity=c: vksiftx_copyDescriptorsToDevice writes the block (zero padded) straight into a persistent (c+1)-row send
    slot, the count goes into the slot's last row, and ONE all-gather moves everything; the host then reads the world's
    counts back.  Nothing else touches the host.  The returned blocks alias the persistent receive buffer: they are valid
    until the next gather with the same capacity.
    """
    dev = torch.device("cuda", inst.device_index)
    if capacity is None:
        n = inst.features_number(buffer_id)
        local = torch.empty((max(n, 1), 128), dtype=torch.uint8, device=dev)
        inst.copy_descriptors_to_device(buffer_id, local.data_ptr(), max(n, 1))
        return exchange_descriptor_blocks(local[:n], group)
    world = dist.get_world_size(group)
    key = (inst.device_index, world, capacity)
    xb = _exchange_buffers.get(key)
    if xb is None:
        xb = _exchange_buffers[key] = _ExchangeBuffers(dev, world, capacity)
    n = inst.copy_descriptors_to_device(buffer_id, xb.send.data_ptr(), capacity)  # rows [n, capacity) are zero filled
    xb.count_host[0] = n
    xb.send[capacity].view(torch.int64)[:1].copy_(xb.count_host, non_blocking=True)
    dist.all_gather_into_tensor(xb.recv.view(world * (capacity + 1), 128), xb.send, group=group)
    xb.counts_host.copy_(xb.recv[:, capacity, :8], non_blocking=True)  # one strided device-to-host copy, no staging kernel
    torch.cuda.current_stream(dev).synchronize()
    counts = [int(c) for c in xb.counts_host.view(torch.int64).flatten().tolist()]
    return counts, xb.recv


def match_against_peers_batched(inst, buffer_a, counts, blocks, rank, out=None):
    """Match the local features (buffer_a) against EVERY peer block of the gathered tensor with one library call
    (vksiftx_matchFeaturesAgainstBlocks: the searches are enqueued back to back, no host round trip in between) and ONE
    download.  Returns {peer: matches or None (peer holds fewer than 2 descriptors)}."""
    assert blocks.data_ptr() % 128 == 0 and (blocks.stride(0) * blocks.element_size()) % 128 == 0
    n_a = inst.features_number(buffer_a)
    inst.match_against_blocks(buffer_a, blocks.data_ptr(), counts, blocks.stride(0) * blocks.element_size(), skip_block=rank)
    res = inst.download_matches_blocks(n_a, out=out)
    return {j: (res[j] if (j != rank and counts[j] >= 2) else None) for j in range(len(counts)) if j != rank}


def match_against_peers(inst, buffer_a, scratch_buffer, counts, blocks, rank, world, download=True):
    """Match the local features (buffer_a) against every peer block, read in place from the gathered tensor
    (vksiftx_matchFeaturesAgainstDevice; a block whose address is not 128-byte aligned goes through `scratch_buffer`).
    Returns {peer: matches or None}."""
    out = {}
    for j in all_pairs_schedule(rank, world):
        if counts[j] < 2:
            out[j] = None
            continue
        ptr = blocks[j].data_ptr()
        if ptr % 128 == 0:
            inst.match_against_device(buffer_a, ptr, counts[j])
        else:
            inst.upload_descriptors_device(ptr, counts[j], scratch_buffer)
            inst.match(buffer_a, scratch_buffer)
        out[j] = inst.download_matches() if download else None
    return out


class PeerExchange:
    """All-pairs exchange step through the library's NVLink peer-memory all-gather (vksiftx_exchange*).

    Setup (once): every rank allocates its receive region, the 64-byte CUDA IPC handles are all-gathered with
    torch.distributed (any backend: the handles travel as a CPU or CUDA uint8 tensor), every rank maps its peers.
    Step: `match_all_peers(buffer)` = ONE library call (push + wait + searches against the seven received blocks in place)
    and ONE download.  Collective: every rank must call it the same number of times."""

    def __init__(self, inst, slot_rows, group=None):
        self.inst, self.group = inst, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        backend = dist.get_backend(group)
        dev = torch.device("cuda", inst.device_index) if backend == "nccl" else torch.device("cpu")
        # every step below is followed by an agreement on its outcome: a rank that cannot allocate or map (no peer access, IPC not
        # permitted in this container) must not leave the others waiting in the next collective
        err, handle = None, bytes(64)
        try:
            handle = inst.exchange_create(self.rank, self.world, slot_rows)
        except Exception as e:  # VksiftError from the library's error callback
            err = e
        mine = torch.frombuffer(bytearray(handle), dtype=torch.uint8).to(dev)
        every = torch.empty((self.world, 64), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(every.view(-1), mine, group=group)
        if err is None:
            try:
                inst.exchange_connect(every.cpu().numpy().tobytes())
            except Exception as e:
                err = e
        ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)  # also the barrier: every region is mapped everywhere before the first push
        if int(ok.item()) == 0:
            try:
                inst.exchange_destroy()
            except Exception:
                pass
            raise RuntimeError("descriptor exchange over peer memory is not available on this node: %s" % (err or "a peer could not map the regions"))

    def allgather(self, buffer_id):
        """(counts, device pointer of block 0, stride in bytes): the peers' descriptor blocks, in place in this rank's region."""
        return self.inst.exchange_allgather(buffer_id)

    def match_all_peers(self, buffer_id, out=None):
        """{peer: matches or None (peer holds fewer than 2 descriptors)} of the local features against every peer's."""
        n_a = self.inst.features_number(buffer_id)
        counts = self.inst.exchange_match_all_peers(buffer_id)
        res = self.inst.download_matches_blocks(n_a, out=out)
        return counts, {j: (res[j] if counts[j] >= 2 else None) for j in range(self.world) if j != self.rank}

    def close(self):
        dist.barrier(self.group)  # nobody pushes into a region that is about to be unmapped
        self.inst.exchange_destroy()
