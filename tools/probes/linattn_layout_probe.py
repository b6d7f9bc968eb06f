"""Does the token-major fused layout (a tile = 128 rows of 128 B, 3 KB apart) cost HBM efficiency?  Same work, two layouts:
(a) 128 x 1024 x 8 heads as column slices of one (N, L, 1536) buffer (what the model runs), (b) 1024 x 1024 x 1 head with q, k, v,
go, out and the gradients each a contiguous (N, L, 64) tensor - every tile is 16 KB of consecutive bytes.
    python tools/probes/linattn_layout_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, cpmusic
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(iters):
        flush.zero_(); a.record(); g.replay(); b.record(); torch.cuda.synchronize(); tot += a.elapsed_time(b)
    return tot / iters

for name, (N, L, H, fused) in {"fused 128x1024x8": (128, 1024, 8, True), "contiguous 1024x1024x1": (1024, 1024, 1, False)}.items():
    gen = torch.Generator().manual_seed(0)
    if fused:
        qkv = torch.randn(N, L, 3 * H * 64, generator=gen).to(dev).bfloat16()
        q, k, v = (qkv[..., i * H * 64:(i + 1) * H * 64].unflatten(-1, (H, 64)) for i in range(3))
        gqkv = torch.empty_like(qkv)
        gq, gk, gv = (gqkv[..., i * H * 64:(i + 1) * H * 64].unflatten(-1, (H, 64)) for i in range(3))
    else:
        q, k, v = (torch.randn(N, L, H, 64, generator=gen).to(dev).bfloat16() for _ in range(3))
        gq, gk, gv = (torch.empty_like(q) for _ in range(3))
    go = torch.randn(N, L, H, 64, generator=gen).to(dev).bfloat16()
    saved = cpmusic.ops.linattn_saved(N, L, H, dev)
    out, den = cpmusic.ops.linattn_fwd_raw(q, k, v, impl=0, saved=saved)
    f = timeit(lambda: cpmusic.ops.linattn_fwd_raw(q, k, v, impl=0, saved=saved))
    b = timeit(lambda: cpmusic.ops.linattn_bwd_raw(q, k, v, out, den, go, gq, gk, gv, impl=0, saved=saved))
    print(json.dumps({"layout": name, "impl": cpmusic.ops.linattn_last_impl(), "fwd_ms": round(f, 4), "bwd_ms": round(b, 4)}))
