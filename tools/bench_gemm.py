"""Times the tcgen05 GEMM family (cpm_gemm_nt / cpm_gemm_tn) at the update-phase shapes of bench.py (T = 131072 tokens) next
to the library GEMM torch dispatches to, CUDA events around back-to-back launches (operands >> L2 at these sizes).

    python tools/bench_gemm.py [--tokens 131072] [--iters 20]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpmusic  # noqa: E402
from cpmusic import ops  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3          # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=131072)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    T = args.tokens
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    peak = peaks.get("bf16_tflops", 1651.8)
    rows = []
    for name, N, K in [("qkv", 1536, 512), ("out", 512, 512), ("ff1", 2048, 512), ("ff2", 512, 2048), ("in", 512, 1216), ("heads", 344, 512)]:
        x = torch.randn(T, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
        wt = w.t().contiguous()
        bias = torch.randn(N, device=dev)
        bias_bf = bias.bfloat16()
        dy = torch.randn(T, N, device=dev).bfloat16()
        dw = torch.zeros(N, K, device=dev)
        fl = 2.0 * T * N * K
        t_own = timed(lambda: ops.gemm_nt(x, w, bias), args.iters)
        t_lib = timed(lambda: torch.addmm(bias_bf, x, w.t()), args.iters)
        t_dg = timed(lambda: ops.gemm_nt(dy, wt), args.iters)
        t_dg_lib = timed(lambda: dy @ w, args.iters)
        t_wg = timed(lambda: (ops.gemm_tn_acc(dy, x, dw), ops.colsum(dy)), args.iters)
        t_wg_lib = timed(lambda: (torch.mm(dy.t(), x, out_dtype=torch.float32), dy.sum(0, dtype=torch.float32)), args.iters)
        row = dict(layer=name, N=N, K=K, fwd_us=round(t_own, 1), fwd_lib_us=round(t_lib, 1), fwd_tf=round(fl / t_own / 1e6, 1),
                   dgrad_us=round(t_dg, 1), dgrad_lib_us=round(t_dg_lib, 1), dgrad_tf=round(fl / t_dg / 1e6, 1),
                   wgrad_us=round(t_wg, 1), wgrad_lib_us=round(t_wg_lib, 1), wgrad_tf=round(fl / t_wg / 1e6, 1), frac_of_burst_peak=round(fl / t_own / 1e6 / peak, 3))
        for mode, tag in ((1, "stream"), (3, "wide")):
            ops.gemm_set_mode(mode)
            row[f"fwd_{tag}_us"] = round(timed(lambda: ops.gemm_nt(x, w, bias), args.iters), 1)
            row[f"dgrad_{tag}_us"] = round(timed(lambda: ops.gemm_nt(dy, wt), args.iters), 1)
        ops.gemm_set_mode(0)
        if name == "ff1":
            t_g = timed(lambda: ops.gemm_nt(x, w, bias, epilogue=ops.GEMM_GELU, p_drop=0.1, seed=1, rng_offset=0), args.iters)
            t_g0 = timed(lambda: ops.gemm_nt(x, w, bias, epilogue=ops.GEMM_GELU, p_drop=0.0), args.iters)
            t_g_lib = timed(lambda: ops.gelu_dropout(torch.addmm(bias_bf, x, w.t()), 0.1), args.iters)
            row.update(fwd_gelu_us=round(t_g, 1), fwd_gelu_nodrop_us=round(t_g0, 1), fwd_gelu_unfused_us=round(t_g_lib, 1))
        if name == "ff2":
            h = torch.randn(T, K, device=dev).bfloat16()
            dyo = torch.randn(T, N, device=dev).bfloat16()
            t_g = timed(lambda: ops.gemm_nt(dyo, wt, epilogue=ops.GEMM_DGELU, aux=h, p_drop=0.1, seed=1, rng_offset=0), args.iters)
            row.update(dgrad_dgelu_us=round(t_g, 1))
        print(json.dumps(row), flush=True)
        rows.append(row)
        del x, w, wt, dy, dw
    tot_own = sum(r["fwd_us"] + r["dgrad_us"] + r["wgrad_us"] for r in rows[:4])
    tot_lib = sum(r["fwd_lib_us"] + r["dgrad_lib_us"] + r["wgrad_lib_us"] for r in rows[:4])
    print(json.dumps(dict(per_layer_fwd_bwd_us_own=round(tot_own, 1), per_layer_fwd_bwd_us_lib=round(tot_lib, 1))))


if __name__ == "__main__":
    main()
