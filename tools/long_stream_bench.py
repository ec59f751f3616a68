"""Developer tool: ONE long stream through the compressor — spread over the whole grid (k_compress<.., kLong>)
against the one-warp kernel (HDLZ_NO_LONG=1): kernel time with device-resident buffers (CUDA events), and the
host call hdlz_compress_stream with pageable buffers beside it."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hdl_deflate_b200 as hz  # noqa: E402
from hdl_deflate_b200 import workload  # noqa: E402

eng = hz.Engine(0)
dev = torch.device("cuda:0")
data = b"".join(workload.blocks(100, 8192, 2048))[:(1 << 24) - 4096]
s = torch.cuda.current_stream().cuda_stream
for n in (1 << 16, 1 << 20, 1 << 22, len(data)):
    d = data[:n]
    d_in = torch.frombuffer(bytearray(d) + bytearray(64), dtype=torch.uint8).to(dev)
    cap = hz.compress_bound(n)
    d_out = torch.zeros(cap, dtype=torch.uint8, device=dev)
    d_len = torch.zeros(1, dtype=torch.int32, device=dev)
    d_st = torch.zeros(1, dtype=torch.int32, device=dev)
    stride = (n + 15) & ~15
    for mode in ("grid", "warp"):
        if mode == "warp":
            os.environ["HDLZ_NO_LONG"] = "1"
        else:
            os.environ.pop("HDLZ_NO_LONG", None)
        reps = 5 if mode == "grid" or n <= (1 << 20) else 2
        eng.compress_batch(d_in, stride, None, n, d_out, cap, d_len, d_st, 1, stream=s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.compress_batch(d_in, stride, None, n, d_out, cap, d_len, d_st, 1, stream=s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        eng.compress(d)                      # (grows the engine's pinned staging buffers)
        t0 = time.perf_counter()
        z = eng.compress(d)
        host_ms = (time.perf_counter() - t0) * 1e3
        assert int(d_st.item()) == 0 and int(d_len.item()) == len(z)
        assert bytes(d_out[:len(z)].cpu().numpy()) == z
        print("%9d bytes  %-5s kernel %8.3f ms = %8.2f GB/s   Engine.compress %8.2f ms   out %d"
              % (n, mode, ms, n / ms / 1e6, host_ms, len(z)), flush=True)
os.environ.pop("HDLZ_NO_LONG", None)
