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
    """Exclusive prefix sum (int64) of the stream lengths: where block i starts in the packed
    output (the `d_in_off` array hdlz_decompress_batch accepts) and the total size."""
    ln = all_len.to(torch.int64)
    off = torch.cumsum(ln, 0) - ln
    return off, int(ln.sum())


def gather_streams(local_out, local_len, out_stride, n_total, dst=0, group=None):
    """Gather the fixed-stride stream slots of every rank on `dst` (two-phase: lengths first).
    -> (uint8 [n_total, out_stride], int32 [n_total]) on dst, (None, lengths) elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    all_len = gather_lengths(local_len, n_total, group)
    counts = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    pad = max(counts)
    buf = torch.zeros((pad, out_stride), dtype=torch.uint8, device=local_out.device)
    buf[:counts[rank]] = local_out.view(-1, out_stride)[:counts[rank]]
    if rank == dst:
        outs = [torch.empty_like(buf) for _ in range(world)]
        dist.gather(buf, outs, dst=dst, group=group)
        return torch.cat([o[:c] for o, c in zip(outs, counts)]), all_len
    dist.gather(buf, None, dst=dst, group=group)
    return None, all_len
