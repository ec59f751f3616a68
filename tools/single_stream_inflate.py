import os, sys, time, zlib
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import hdl_deflate_b200 as hz
from hdl_deflate_b200 import workload
eng = hz.Engine(0)
dev = torch.device("cuda:0")
data = b"".join(workload.blocks(100, 2048, 2048))   # 4 MiB
s = torch.cuda.current_stream().cuda_stream
for name, z in (("own fixed-block stream", eng.compress(data)), ("zlib level 6", zlib.compress(data, 6)), ("zlib Z_FIXED", (lambda c: c.compress(data) + c.flush())(zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_FIXED)))):
    n = len(z)
    d_in = torch.frombuffer(bytearray(z) + bytearray(64), dtype=torch.uint8).to(dev)
    d_len = torch.tensor([n], dtype=torch.int32, device=dev)
    d_out = torch.zeros(len(data) + 64, dtype=torch.uint8, device=dev)
    d_olen = torch.zeros(1, dtype=torch.int32, device=dev); d_st = torch.zeros(1, dtype=torch.int32, device=dev)
    cap = len(data)
    eng.decompress_batch(d_in, None, (n + 15) & ~15, d_len, d_out, (cap + 15) & ~15, cap, d_olen, d_st, 1, stream=s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        eng.decompress_batch(d_in, None, (n + 15) & ~15, d_len, d_out, (cap + 15) & ~15, cap, d_olen, d_st, 1, stream=s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    assert int(d_st.item()) == 0 and bytes(d_out[:cap].cpu().numpy()) == data
    print("%-24s %8d -> %8d bytes  %8.3f ms = %7.1f MB/s of output" % (name, n, cap, ms, cap / ms / 1e3), flush=True)
