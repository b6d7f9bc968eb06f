"""Per-kernel GPU time of one chunk-parallel forward + backward call (torch profiler, CUDA activities, L2 flushed before each call).
    python tools/profile_linattn_kernels.py [N L H]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cpmusic
from torch.profiler import profile, ProfilerActivity
N, L, H = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (128, 1024, 8)
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
qkv = torch.randn(N, L, 3 * H * 64, generator=g).to(dev).bfloat16()
q, k, v = (qkv[..., i * H * 64:(i + 1) * H * 64].unflatten(-1, (H, 64)) for i in range(3))
go = torch.randn(N, L, H, 64, generator=g).to(dev).bfloat16()
gqkv = torch.empty_like(qkv)
gq, gk, gv = (gqkv[..., i * H * 64:(i + 1) * H * 64].unflatten(-1, (H, 64)) for i in range(3))
saved = cpmusic.ops.linattn_saved(N, L, H, dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def once():
    flush.zero_()
    out, den = cpmusic.ops.linattn_fwd_raw(q, k, v, impl=0, saved=saved)
    flush.zero_()
    cpmusic.ops.linattn_bwd_raw(q, k, v, out, den, go, gq, gk, gv, impl=0, saved=saved)
for _ in range(3): once()
torch.cuda.synchronize()
REP = 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(REP): once()
    torch.cuda.synchronize()
import re
NAME = lambda k: (re.search(r"(cp_\w+|linattn\w+)", k) or [k, k])[1][:40]
tot = 0.0
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
    if "cp_" in e.key or "linattn" in e.key:
        print(f"{NAME(e.key):40s} {e.device_time_total / REP:9.1f} us per call  ({e.count // REP} launch)")
        tot += e.device_time_total / REP
print(f"{'sum':40s} {tot:9.1f} us  -> {N * L * H * 1408 / tot / 1e3:.0f} GB/s algorithmic")
