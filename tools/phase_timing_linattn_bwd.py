"""Phase timestamps of the chunk-parallel backward kernel cp_bwd_main (clock64 of thread 0 at the phase boundaries, per CTA).
    python tools/phase_timing_linattn_bwd.py [N L H]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cpmusic

N, L, H = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (128, 1024, 8)
dev = torch.device("cuda:0")
lib = cpmusic._lib.load()
g = torch.Generator().manual_seed(0)
qkv = torch.randn(N, L, 3 * H * 64, generator=g).to(dev).bfloat16()
q, k, v = (qkv[..., i * H * 64:(i + 1) * H * 64].unflatten(-1, (H, 64)) for i in range(3))
go = torch.randn(N, L, H, 64, generator=g).to(dev).bfloat16()
gqkv = torch.empty_like(qkv)
gq, gk, gv = (gqkv[..., i * H * 64:(i + 1) * H * 64].unflatten(-1, (H, 64)) for i in range(3))
saved = cpmusic.ops.linattn_saved(N, L, H, dev)
out, den = cpmusic.ops.linattn_fwd_raw(q, k, v, impl=3, saved=saved)
for _ in range(3):
    cpmusic.ops.linattn_bwd_raw(q, k, v, out, den, go, gq, gk, gv, impl=3, saved=saved)
NCTA = 296
buf = torch.zeros(NCTA * 128, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev); flush.zero_()
lib.cpm_debug_linattn_timing(buf.data_ptr())
cpmusic.ops.linattn_bwd_raw(q, k, v, out, den, go, gq, gk, gv, impl=3, saved=saved)
torch.cuda.synchronize()
lib.cpm_debug_linattn_timing(None)
b = buf.view(NCTA, 128).cpu()
names = ["wait Q,V,Sp + phi(Q)", "wait K,go,Rs + G', phi(K) + sync", "round 1 (X)", "convert X + sync", "round 2 (dQf, dKf, PT)",
         "convert PT + sync", "dq, dk rows out", "round 3 (dv) rest", "dv rows out + sync"]
P = len(names)
for cta in (0, 1, 147, 295):
    row = b[cta]
    n = int((row != 0).sum())
    d = (row[1:n] - row[:n - 1]).tolist()
    print(f"CTA {cta}: stamps {n}")
    for t in range(min(3, (n - 1) // P)):
        print("   tile", t, d[P * t:P * t + P])
tot = torch.zeros(P)
cnt = 0
for cta in range(NCTA):
    row = b[cta]; n = int((row != 0).sum())
    d = (row[1:n] - row[:n - 1]).float()
    for t in range(1, (n - 1) // P):                 # skip each CTA's first tile (cold start)
        tot += d[P * t:P * t + P]; cnt += 1
mean = tot / max(cnt, 1)
print("mean cycles per phase over", cnt, "tiles (first tile of every CTA left out):")
for nm, x in zip(names, mean):
    print(f"   {nm:38s} {float(x):8.0f}")
print("   sum", float(mean.sum()))
