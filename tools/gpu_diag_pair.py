"""Developer tool: run stream B right after stream A on the same persistent warp (device batch API)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import __graft_entry__; __graft_entry__.build()
import hdl_deflate_b200 as hz
from hdl_deflate_b200 import workload
from oracle import hdlz_oracle as O

L = 2048; S = hz.compress_bound(L)
eng = hz.Engine(0); dev = torch.device("cuda:0")
NW = 148 * 8 * 4
B = int(os.environ.get("B", 880395))
def run(first_blocks, label):
    n = NW + 1
    arr = np.zeros((n, L), dtype=np.uint8)
    for i in range(n): arr[i] = np.frombuffer(workload.block(1000 + i, L), dtype=np.uint8)
    arr[0] = first_blocks
    arr[NW] = np.frombuffer(workload.block(B, L), dtype=np.uint8)
    d_in = torch.from_numpy(arr).to(dev).contiguous()
    d_out = torch.zeros(n * S, dtype=torch.uint8, device=dev)
    d_len = torch.zeros(n, dtype=torch.int32, device=dev); d_st = torch.zeros(n, dtype=torch.int32, device=dev)
    eng.compress_batch(d_in, L, None, L, d_out, S, d_len, d_st, n, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out = d_out.view(n, S).cpu().numpy(); ln = d_len.cpu().numpy()
    bad = []
    for i in range(n):
        w = O.compress(arr[i].tobytes())[1]
        if out[i, :ln[i]].tobytes() != w: bad.append(i)
    print(label, "bad streams:", bad[:10])
run(np.frombuffer(workload.block(B - NW, L), dtype=np.uint8), "pred=real predecessor")
run(np.zeros(L, dtype=np.uint8), "pred=zeros")
run(np.frombuffer(workload.block(5, L), dtype=np.uint8), "pred=other block")
# alone through the stream API
print("alone:", eng.compress(workload.block(B, L)) == O.compress(workload.block(B, L))[1])
