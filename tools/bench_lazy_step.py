"""Recurrent state kernel variants at the rollout shape (256 sequences x 8 heads x 12 layers, L2 flushed): one-CTA-per-tile (default),
deferred write-back by pending count, read-only / write-back halves, plain device copies for scale, persistent bulk-copy-staged kernel.
    python tools/bench_lazy_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cpmusic
from cpmusic import ops
dev = torch.device("cuda:0")
N, H, layers = 256, 8, 12
S = torch.zeros(layers, N, H, 64, 64, device=dev); Z = torch.zeros(layers, N, H, 64, device=dev)
ring = torch.zeros(layers, N, H, 8, 128, device=dev)
qkv = torch.randn(N, 1536, device=dev).bfloat16()
q, k, v = (qkv[:, j * 512:(j + 1) * 512].unflatten(-1, (H, 64)) for j in range(3))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
step = torch.zeros(1, dtype=torch.int32, device=dev)
def timeit(fn):
    for _ in range(2): fn()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = 0
    for _ in range(5):
        flush.zero_(); a.record(); g.replay(); b.record(); torch.cuda.synchronize(); t += a.elapsed_time(b)
    return t / 5 / layers * 1e3
print("eager", round(timeit(lambda: [ops.linattn_step(q, k, v, S[l], Z[l]) for l in range(layers)]), 2), "us/launch")
for p in range(8):
    step.fill_(p)
    print("lazy p =", p, round(timeit(lambda: [ops.linattn_step_lazy(q, k, v, S[l], Z[l], ring[l], step) for l in range(layers)]), 2), "us/launch")
kvp = torch.zeros(layers, N, H, 128, device=dev)
print("step_out (S read only)", round(timeit(lambda: [ops.linattn_step_out(q, k, v, S[l], Z[l], kvp[l]) for l in range(layers)]), 2), "us/launch")
print("state_update (S read+write, no output)", round(timeit(lambda: [ops.linattn_state_update(S[l], kvp[l]) for l in range(layers)]), 2), "us/launch")
Sc = S.clone()
print("torch copy S->Sc (read+write 67 MB)", round(timeit(lambda: [Sc[l].copy_(S[l]) for l in range(layers)]), 2), "us/launch")
print("torch sum(S) (read 33.5 MB)", round(timeit(lambda: [S[l].sum() for l in range(layers)]), 2), "us/launch")
for c in (1, 2, 3):
    print("persistent TMA step,", c, "CTAs/SM", round(timeit(lambda: [ops.linattn_step(q, k, v, S[l], Z[l], tma_ctas=c) for l in range(layers)]), 2), "us/launch")
