"""clock64 phase stamps of cpm_tc_linear at the rollout shapes.  python tools/phase_timing_tc_linear.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cpmusic
from cpmusic import ops

dev = torch.device("cuda:0")
lib = cpmusic._lib.load()
names = ["setup", "producer issued", "first tile landed", "mma all issued", "(epi reached wait)", "accumulator ready", "epilogue done"]
for (M, N, K, bn, epi) in [(256, 1536, 512, 64, ops.TL_BIAS), (256, 512, 512, 32, ops.TL_RES), (256, 2048, 512, 64, ops.TL_GELU), (256, 512, 2048, 32, ops.TL_RES)]:
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    b = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev).bfloat16()
    bn = 64
    sk = ops.tc_linear_split(M, N, K, bn)
    kw = dict(epilogue=epi, residual=res if epi == ops.TL_RES else None, block_n=bn, split_k=sk)
    for _ in range(3):
        ops.tc_linear(a, w, b, **kw)
    ncta = (N // bn) * ((M + 127) // 128) * sk
    buf = torch.zeros(ncta * 8, dtype=torch.int64, device=dev)
    lib.cpm_debug_tc_linear_timing(buf.data_ptr())
    ops.tc_linear(a, w, b, **kw)
    torch.cuda.synchronize()
    lib.cpm_debug_tc_linear_timing(None)
    t = buf.view(ncta, 8).double()
    rel = (t - t[:, :1])
    print(f"M{M} N{N} K{K} bn{bn} split_k{sk} epi{epi}: {ncta} CTAs; mean cycles since CTA start:")
    for i, nm in enumerate(["setup done", "accumulator parked in smem", "block sync passed", "cluster sync passed", "row stats done",
                            "accumulator ready", "finalise done"]):
        print(f"    {nm:40s} {rel[:, i + 1].mean().item():9.0f}   (max {rel[:, i + 1].max().item():.0f})")
    import time
    torch.cuda.synchronize(); a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            ops.tc_linear(a, w, b, **kw)
    g.replay(); torch.cuda.synchronize()
    a0.record(); g.replay(); a1.record(); torch.cuda.synchronize()
    print(f"    back-to-back in a graph: {a0.elapsed_time(a1) * 1e3 / 20:.2f} us per launch")
