"""Phase timestamps of the chunk-parallel forward kernel (clock64 at phase boundaries, per CTA).
    python tools/phase_timing_linattn.py [N L H]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cpmusic

N, L, H = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (16, 1024, 8)
dev = torch.device("cuda:0")
lib = cpmusic._lib.load()
qkv = torch.randn(N, L, 3 * H * 64, device=dev).bfloat16()
q, k, v = (qkv[..., i * H * 64:(i + 1) * H * 64].unflatten(-1, (H, 64)) for i in range(3))
for _ in range(3):
    cpmusic.ops.linattn_fwd_raw(q, k, v, impl=3)
buf = torch.zeros(444 * 64, dtype=torch.int64, device=dev)
lib.cpm_debug_linattn_timing(buf.data_ptr())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev); flush.zero_()
cpmusic.ops.linattn_fwd_raw(q, k, v, impl=3)
torch.cuda.synchronize()
lib.cpm_debug_linattn_timing(None)
b = buf.view(444, 64).cpu()
names = ["load wait", "phi+sync", "mma1", "convert+sync", "mma2", "epilogue+sync"]
for cta in (0, 1, 147, 148, 300, 443):
    row = b[cta]
    n = int((row != 0).sum())
    d = (row[1:n] - row[:n - 1]).tolist()
    print(f"CTA {cta}: stamps {n}; per-phase cycles:")
    for t in range((n - 1) // 6):
        print("   tile", t, {nm: d[6 * t + i] for i, nm in enumerate(names)})
tot = torch.zeros(6)
cnt = 0
for cta in range(444):
    row = b[cta]; n = int((row != 0).sum())
    d = (row[1:n] - row[:n - 1]).float()
    for t in range((n - 1) // 6):
        tot += d[6 * t:6 * t + 6]; cnt += 1
print("mean cycles per phase over", cnt, "tiles:", {nm: round(float(x), 0) for nm, x in zip(names, tot / cnt)})
print("kernel span (cycles, per CTA first->last stamp): min/mean/max",
      [float(f((b.max(1).values - b[:, 0]).float())) for f in (torch.min, torch.mean, torch.max)])
