"""N > 1 host logic on the CPU: two gloo ranks shard a batch, run the START jobs through the
oracle test double, and the gathered result must equal the single-process result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hdl_deflate_b200 import sharding, workload, compress_bound
    from oracle import hdlz_oracle
    first, last = sharding.shard_range(n_total, rank, world)
    stride = compress_bound(2048)
    out = torch.zeros((last - first, stride), dtype=torch.uint8)
    lens = torch.zeros(last - first, dtype=torch.int32)
    for i in range(first, last):
        s = hdlz_oracle.compress(workload.block(i, 2048))[1]
        out[i - first, :len(s)] = torch.frombuffer(bytearray(s), dtype=torch.uint8)
        lens[i - first] = len(s)
    # pack the shard the way hdlz_pack_batch does: stream starts rounded up to 4, in block order
    padded = (lens.to(torch.int64) + 3) & ~3
    loff = torch.cumsum(padded, 0) - padded
    packed = torch.zeros(int(padded.sum()) + 16, dtype=torch.uint8)
    for i in range(last - first):
        packed[int(loff[i]):int(loff[i]) + int(lens[i])] = out[i, :int(lens[i])]
    # the input side: rank 0 owns all blocks, every rank receives its shard
    blocks = torch.zeros((n_total, 2048), dtype=torch.uint8)
    if rank == 0:
        for i in range(n_total):
            blocks[i] = torch.frombuffer(bytearray(workload.block(i, 2048)), dtype=torch.uint8)
    mine = sharding.broadcast_blocks(blocks, src=0)
    assert mine.shape[0] == last - first and mine[0].numpy().tobytes() == workload.block(first, 2048)
    mine2 = sharding.scatter_blocks(blocks if rank == 0 else None, n_total, 2048, src=0)
    assert torch.equal(mine2, mine)
    all_len = sharding.gather_lengths(lens, n_total)
    off, total = sharding.packed_offsets(all_len)
    gathered, off2, all_len2 = sharding.gather_streams(packed, lens, n_total, dst=0)
    assert torch.equal(all_len, all_len2) and torch.equal(off, off2)
    if rank == 0:
        q.put((all_len.numpy(), off.numpy(), total, gathered.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partition():
    sys.path.insert(0, ROOT)
    from hdl_deflate_b200 import sharding
    for n in (0, 1, 7, 8, 9, 1 << 20, (1 << 23) + 5):
        for world in (1, 2, 3, 4, 8):
            edges = [sharding.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gather_equals_single_process():
    sys.path.insert(0, ROOT)
    from hdl_deflate_b200 import workload, compress_bound
    from oracle import hdlz_oracle
    n_total, world = 37, 2          # odd: ragged shards
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    all_len, off, total, gathered = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    want = [hdlz_oracle.compress(workload.block(i, 2048))[1] for i in range(n_total)]
    assert list(all_len) == [len(s) for s in want]
    r4 = [(len(s) + 3) & ~3 for s in want]
    assert total == sum(r4) == gathered.shape[0]          # packed: only stream bytes (starts 4-byte aligned)
    assert list(off) == list(np.cumsum([0] + r4[:-1]))
    for i, s in enumerate(want):
        assert gathered[int(off[i]):int(off[i]) + len(s)].tobytes() == s
    assert total < 0.7 * n_total * compress_bound(2048)   # far less than the fixed slots
