"""Multi-GPU layout of the block workload (BASELINE configs[4]).

Blocks are independent zlib streams (own header and Adler-32, 32-byte window that never
reaches before position 0; deflate.py:753-757, 788-814, 989), so the path shards with no
exchange step: rank r owns the contiguous range shard_range(n, r, world).  The only
collectives sit outside the kernels: an all-gather of the per-block stream lengths, from which
every rank derives the packed offsets, and (optionally) a gather of the streams themselves.
Works on any torch.distributed backend (nccl on the GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_total, rank, world):
    """[first, last) of the blocks rank `rank` owns: contiguous, sizes differ by at most one."""
    base, rem = divmod(int(n_total), int(world))
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def gather_lengths(local_len, n_total, group=None):
    """All-gather the per-block lengths (int32 tensor of this rank's shard) -> int32 [n_total]
    in global block order, on the same device."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    assert local_len.numel() == counts[rank]
    pad = max(counts)
    buf = torch.zeros(pad, dtype=local_len.dtype, device=local_len.device)
    buf[:counts[rank]] = local_len
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    return torch.cat([o[:c] for o, c in zip(outs, counts)])


def packed_offsets(all_len):
    """Exclusive prefix sum (int64) of the stream lengths rounded up to 4 — the layout hdlz_pack_batch
    writes, so every stream starts 4-byte aligned and stays on the lane-per-stream inflate route:
    where block i starts in the packed output (the `d_in_off` array hdlz_decompress_batch accepts)
    and the total size."""
    ln = (all_len.to(torch.int64) + 3) & ~3
    off = torch.cumsum(ln, 0) - ln
    return off, int(ln.sum())


def shard_byte_ranges(off, all_len, n_total, world):
    """[lo, hi) of every rank's shard in the packed buffer described by `off` (packed_offsets)."""
    ln = (all_len.to(torch.int64) + 3) & ~3
    out = []
    for r in range(world):
        first, last = shard_range(n_total, r, world)
        lo = int(off[first]) if first < n_total else int(ln.sum())
        hi = int(off[last - 1] + ln[last - 1]) if last > first else lo
        out.append((lo, hi))
    return out


def gather_streams(local_packed, local_len, n_total, dst=0, group=None):
    """Gather the PACKED streams of every rank on `dst` (BASELINE configs[4] "NCCL gather"): two-phase —
    all-gather of the lengths, from which every rank derives the global packed offsets, then each rank
    sends exactly its packed bytes (what hdlz_pack_batch wrote: starts rounded up to 4) into its range
    of the destination buffer.  Only real stream bytes cross the links.
    -> (uint8 [total], int64 offsets [n_total], int32 lengths [n_total]) on dst,
       (None, offsets, lengths) elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    all_len = gather_lengths(local_len, n_total, group)
    off, total = packed_offsets(all_len)
    ranges = shard_byte_ranges(off, all_len, n_total, world)
    lo, hi = ranges[rank]
    mine = local_packed.view(torch.uint8).reshape(-1)[:hi - lo]
    if rank == dst:
        buf = torch.empty(max(total, 1), dtype=torch.uint8, device=local_packed.device)
        buf[lo:hi] = mine
        ops = [dist.P2POp(dist.irecv, buf[a:b], r, group) for r, (a, b) in enumerate(ranges) if r != dst and b > a]
    else:
        buf = None
        ops = [dist.P2POp(dist.isend, mine.contiguous(), dst, group)] if hi > lo else []
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return (buf[:total] if buf is not None else None), off, all_len


def broadcast_blocks(blocks, src=0, group=None):
    """Rank `src` owns the input blocks (uint8 [n_total, stride] on its device); every rank receives its
    contiguous shard.  The other half of configs[4]'s "NCCL used only to broadcast inputs and gather
    outputs": a scatter expressed as one broadcast of the whole batch (NVSwitch delivers it to all peers
    at once) followed by a local slice.  `blocks` must be an allocated tensor of the full shape on
    every rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dist.broadcast(blocks, src=src, group=group)
    first, last = shard_range(blocks.shape[0], rank, world)
    return blocks[first:last]


def scatter_blocks(all_blocks, n_total, stride, src=0, group=None, out=None):
    """Rank `src` owns all input blocks (uint8 [n_total, stride] on its device; None elsewhere); every rank
    receives exactly its contiguous shard (point-to-point over NVLink: each byte crosses one link once).
    -> uint8 [shard, stride] on every rank (`out` if given)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    first, last = shard_range(n_total, rank, world)
    if rank == src:
        mine = all_blocks[first:last]
        ops = []
        for r in range(world):
            a, b = shard_range(n_total, r, world)
            if r != src and b > a:
                ops.append(dist.P2POp(dist.isend, all_blocks[a:b].contiguous(), r, group))
        if out is not None:
            out.copy_(mine)
            mine = out
    else:
        device = out.device if out is not None else torch.device("cpu")
        mine = out if out is not None else torch.empty((last - first, stride), dtype=torch.uint8, device=device)
        ops = [dist.P2POp(dist.irecv, mine, src, group)] if last > first else []
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return mine
