"""Micro-benchmark of the linear-attention kernels (CUDA events, L2 flushed between iterations).
    python tools/bench_linattn.py [--impl 1 2] [--bwd]
Prints one JSON line per (shape, impl, direction) with achieved algorithmic GB/s vs the measured HBM peak."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cpmusic

SHAPES = {"cfg2 32x512x8": (32, 512, 8), "ppo-update 16x1024x8": (16, 1024, 8), "bench minibatch 64x1024x8": (64, 1024, 8), "bench minibatch 128x1024x8": (128, 1024, 8),
          "cfg5 1x8192x16": (1, 8192, 16), "cfg1 4x512x8": (4, 512, 8)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", type=int, nargs="+", default=[1, 2])
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--shape", default=None, help="substring filter on the shape name")
    ap.add_argument("--dirs", nargs="+", default=["fwd", "bwd"])
    ap.add_argument("--hot", type=int, default=0, help="run this many 8192^3 bf16 matmuls right before every timed call: the chip is then at "
                    "its power cap with lowered SM clocks, as inside the update phase of bench.py")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ha, hb = (torch.randn(8192, 8192, device=dev).bfloat16() for _ in range(2)) if args.hot else (None, None)
    for name, (N, L, H) in SHAPES.items():
        if args.shape and args.shape not in name:
            continue
        g = torch.Generator().manual_seed(0)
        qkv = torch.randn(N, L, 3 * H * 64, generator=g).to(dev).bfloat16()
        q, k, v = (qkv[..., i * H * 64:(i + 1) * H * 64].unflatten(-1, (H, 64)) for i in range(3))
        go = torch.randn(N, L, H, 64, generator=g).to(dev).bfloat16()
        gqkv = torch.empty_like(qkv)
        gq, gk, gv = (gqkv[..., i * H * 64:(i + 1) * H * 64].unflatten(-1, (H, 64)) for i in range(3))
        for impl in args.impl:
            for direction in args.dirs:
                saved = cpmusic.ops.linattn_saved(N, L, H, dev) if impl in (0, 3) else None
                out, den = cpmusic.ops.linattn_fwd_raw(q, k, v, impl=impl, saved=saved)
                def run_eager():
                    if direction == "fwd":
                        cpmusic.ops.linattn_fwd_raw(q, k, v, impl=impl, saved=saved)
                    else:
                        cpmusic.ops.linattn_bwd_raw(q, k, v, out, den, go, gq, gk, gv, impl=impl, saved=saved)
                try:
                    for _ in range(3):
                        run_eager()
                except Exception as e:
                    print(json.dumps({"shape": name, "impl": impl, "dir": direction, "error": str(e)[:100]}))
                    continue
                used = cpmusic.ops.linattn_last_impl()
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()          # GPU time only: no Python / tensor-map-encode gaps between launches
                with torch.cuda.graph(graph):
                    run_eager()
                run = graph.replay
                tot = 0.0
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                for _ in range(args.iters):
                    for _ in range(args.hot):
                        ha @ hb
                    flush.zero_()
                    a.record(); run(); b.record()
                    torch.cuda.synchronize()
                    tot += a.elapsed_time(b)
                ms = tot / args.iters
                per = 512 if direction == "fwd" else 896
                byt = N * L * H * per
                print(json.dumps({"shape": name, "impl": used, "dir": direction, "ms": round(ms, 4), "alg_GBps": round(byt / ms / 1e6, 1),
                                  "frac_of_measured_hbm": round(byt / ms / 1e6 / peak, 4), "alg_MB": round(byt / 1e6, 1)}))


if __name__ == "__main__":
    main()
