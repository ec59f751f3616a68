"""Developer tool: compress 2^20 blocks with ONE device-batch launch and compare every stream with the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import __graft_entry__; __graft_entry__.build()
import hdl_deflate_b200 as hz
from oracle import hdlz_oracle as O
from gpu_diag import explain

n, L = int(os.environ.get("N", 1 << 20)), 2048
S = hz.compress_bound(L)
eng = hz.Engine(0)
dev = torch.device("cuda:0")
d_in = torch.empty(n * L, dtype=torch.uint8, device=dev)
d_out = torch.zeros(n * S, dtype=torch.uint8, device=dev)
d_len = torch.zeros(n, dtype=torch.int32, device=dev); d_st = torch.zeros(n, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
eng.generate_blocks(d_in, L, L, n, stream=s)
for rep in range(3):
    eng.compress_batch(d_in, L, None, L, d_out, S, d_len, d_st, n, stream=s)
torch.cuda.synchronize()
h_in = d_in.view(n, L).cpu().numpy(); h_out = d_out.view(n, S).cpu().numpy(); h_len = d_len.cpu().numpy().astype(np.uint32)
want = np.zeros((n, S), dtype=np.uint8)
wlen, st = O.batch(O.KIND_PORT_COMPRESS, h_in, np.arange(n, dtype=np.uint64) * L, np.full(n, L, np.uint32), want,
                   np.arange(n, dtype=np.uint64) * S, S, os.cpu_count())
badlen = np.nonzero(h_len != wlen)[0]
mask = np.arange(S)[None, :] < wlen[:, None]
badbytes = np.nonzero(((h_out != want) & mask).any(axis=1))[0]
print("status nonzero", int((d_st != 0).sum()), "len mismatches", len(badlen), "byte mismatches", len(badbytes))
print("first bad:", badbytes[:30], "mod 4:", (badbytes[:30] % 4), "warp slot:", (badbytes[:30] // 4) % (148 * 8))
for i in badbytes[:3]:
    print("block", i)
    explain(h_in[i].tobytes(), h_out[i, :h_len[i]].tobytes(), want[i, :wlen[i]].tobytes(), print)
for i in badbytes[:3]:
    g = h_out[i, :h_len[i]]; w = want[i, :wlen[i]]
    d = np.nonzero(g != w)[0]
    print("diff byte idx", d, "got", [hex(x) for x in g[d]], "want", [hex(x) for x in w[d]], "xor", [hex(a ^ b) for a, b in zip(g[d], w[d])])
    lo = (d[0] // 4) * 4 - 8
    print("got words ", [g[k:k+4].tobytes()[::-1].hex() for k in range(lo, lo + 32, 4)])
    print("want words", [w[k:k+4].tobytes()[::-1].hex() for k in range(lo, lo + 32, 4)])
    np.save(os.path.join(ROOT, "gpurun_out", "bad_block_%d.npy" % i), h_in[i])
