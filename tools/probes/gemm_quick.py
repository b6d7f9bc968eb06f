import sys, os, torch
sys.path.insert(0, "/root/repo")
import cpmusic
from cpmusic import ops
def timed(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3
dev = torch.device("cuda:0")
T = 131072
for name, N, K in [("qkv", 1536, 512), ("out", 512, 512), ("ff2", 512, 2048), ("longK", 512, 8192)]:
    x = torch.randn(T, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16(); bias = torch.randn(N, device=dev)
    t = timed(lambda: ops.gemm_nt(x, w, bias))
    ref = torch.addmm(bias.bfloat16(), x[:4096], w.t()).float()
    got = ops.gemm_nt(x, w, bias)[:4096].float()
    print(name, round(t, 1), "us", round(2.0 * T * N * K / t / 1e6, 1), "TF", "maxerr", float((got - ref).abs().max()))
