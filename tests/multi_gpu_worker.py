"""Worker of tests/test_gpu_multi.py (launched by torch.distributed.run, one rank per GPU): every rank
compresses its shard of a seeded batch on its own GPU, packs the streams, and the NCCL gather of the packed
bytes on rank 0 must equal what one GPU produces for the whole batch."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hdl_deflate_b200 as hz  # noqa: E402
from hdl_deflate_b200 import sharding  # noqa: E402


def compress_packed(eng, dev, first, count, block=2048):
    stream = torch.cuda.current_stream().cuda_stream
    stride = hz.compress_bound(block)
    d_in = torch.empty((count, block), dtype=torch.uint8, device=dev)
    eng.generate_blocks(d_in, block, block, count, first_block=first, stream=stream)
    d_out = torch.empty(count * stride, dtype=torch.uint8, device=dev)
    d_len = torch.zeros(count, dtype=torch.int32, device=dev)
    d_st = torch.zeros(count, dtype=torch.int32, device=dev)
    eng.compress_batch(d_in, block, None, block, d_out, stride, d_len, d_st, count, stream=stream)
    d_packed = torch.empty(count * stride, dtype=torch.uint8, device=dev)
    d_off = torch.zeros(count, dtype=torch.int64, device=dev)
    d_tot = torch.zeros(1, dtype=torch.int64, device=dev)
    eng.pack_batch(d_out, stride, d_len, d_packed, d_off, d_tot, count, stream=stream)
    torch.cuda.synchronize()
    assert int(d_st.abs().sum()) == 0
    return d_in, d_packed[:int(d_tot.item())], d_off, d_len


def main():
    n_total = int(sys.argv[1])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    eng = hz.Engine(local)
    first, last = sharding.shard_range(n_total, rank, world)
    # inputs: rank 0 owns the whole batch and scatters the shards
    d_all = None
    if rank == 0:
        d_all = torch.empty((n_total, 2048), dtype=torch.uint8, device=dev)
        eng.generate_blocks(d_all, 2048, 2048, n_total, first_block=0, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    recv = torch.empty((last - first, 2048), dtype=torch.uint8, device=dev)
    sharding.scatter_blocks(d_all, n_total, 2048, src=0, out=recv)
    d_in, packed, off, lens = compress_packed(eng, dev, first, last - first)
    assert torch.equal(recv, d_in), "scattered shard differs from the locally generated one"
    g_buf, g_off, g_len = sharding.gather_streams(packed, lens, n_total, dst=0)
    if rank == 0:
        _, want, woff, wlen = compress_packed(eng, dev, 0, n_total)
        assert torch.equal(g_len, wlen) and torch.equal(g_off, woff)
        assert g_buf.numel() == want.numel() and torch.equal(g_buf, want), "gathered bytes differ from the single-GPU result"
        # and they inflate back to the blocks
        d_back = torch.empty(n_total * 2048, dtype=torch.uint8, device=dev)
        d_bl = torch.zeros(n_total, dtype=torch.int32, device=dev)
        d_bs = torch.zeros(n_total, dtype=torch.int32, device=dev)
        eng.decompress_batch(g_buf, g_off, 0, g_len, d_back, 2048, 2048, d_bl, d_bs, n_total, flags=hz.F_VERIFY_ADLER,
                             stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert int(d_bs.abs().sum()) == 0 and torch.equal(d_back.view(n_total, 2048), d_all)
        print("MULTI_GPU_OK world=%d blocks=%d packed_bytes=%d" % (world, n_total, g_buf.numel()), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
