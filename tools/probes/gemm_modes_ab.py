"""A/B of the two cpm_gemm_nt schedules (1 = 256 x 256 tiles, two accumulator stages; 3 = 256 x 512 tiles, early release) at the update
shapes: modes interleaved, 7 rounds of 10 launches each, median per mode (single measurements differ by up to 10 % with the clocks)."""
import os, sys, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cpmusic
from cpmusic import ops

def timed(fn, iters=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3

dev = torch.device("cuda:0")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
shapes = [("qkv fwd", 1536, 512, True), ("qkv dgrad", 512, 1536, False), ("out fwd", 512, 512, True), ("ff1 fwd", 2048, 512, True), ("ff1 dgrad", 512, 2048, False),
          ("ff2 fwd", 512, 2048, True), ("ff2 dgrad", 2048, 512, False), ("in fwd", 512, 1216, True), ("in dgrad", 1216, 512, False), ("heads fwd", 344, 512, True), ("heads dgrad", 512, 344, False)]
for name, N, K, has_bias in shapes:
    x = torch.randn(T, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16(); bias = torch.randn(N, device=dev) if has_bias else None
    res = {1: [], 3: [], 0: []}
    for r in range(7):
        for mode in (1, 3, 0):
            ops.gemm_set_mode(mode)
            res[mode].append(timed(lambda: ops.gemm_nt(x, w, bias)))
    ops.gemm_set_mode(0)
    m = {k: statistics.median(v) for k, v in res.items()}
    print(f"{name:12s} N {N:5d} K {K:5d}  stream {m[1]:7.1f}  wide {m[3]:7.1f}  auto {m[0]:7.1f} us   ({2.0 * T * N * K / min(m[1], m[3]) / 1e6:6.0f} TF best)")
