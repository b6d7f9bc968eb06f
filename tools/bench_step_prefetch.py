"""What an L2 prefetch of the recurrent state can buy: 12 x [prefetch S_i ; ~20 us of unrelated work ; step kernel on S_i]
against 12 x [work ; step] (CUDA graphs, L2 flushed before each replay).  The difference per layer is the most the
rollout engine's prefetch_state modes can save.   python tools/bench_step_prefetch.py [--songs 256]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cpmusic
from cpmusic import ops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--songs", type=int, default=256)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--sleep", type=int, default=40000, help="cycles of unrelated work between prefetch and step")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    N, H, nl = args.songs, 8, 12
    S = torch.zeros(nl, N, H, 64, 64, device=dev)
    Z = torch.zeros(nl, N, H, 64, device=dev)
    qkv = torch.randn(N, 1536, device=dev).bfloat16()
    q, k, v = (qkv[:, j * 512:(j + 1) * 512].unflatten(-1, (H, 64)) for j in range(3))
    sink = torch.zeros(nl, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def chain(kind):
        for i in range(nl):
            if kind == "prefetch":
                ops.l2_prefetch(S[i])
            if kind == "touch":                         # a real read of the tile set (default caching): upper bound for any prefetch
                sink[i] = S[i].view(-1, 4096).sum()
            if kind != "step-only":
                torch.cuda._sleep(args.sleep)
            if kind == "sleep-only":
                continue
            if kind == "fused-next":
                ops.linattn_step(q, k, v, S[i], Z[i], prefetch=S[(i + 1) % nl], prefetch_when=2)
            else:
                ops.linattn_step(q, k, v, S[i], Z[i])

    res = {}
    with torch.cuda.stream(torch.cuda.Stream()):
        for kind in ("step-only", "sleep-only", "plain", "prefetch", "fused-next", "touch"):
            chain(kind)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                chain(kind)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tot = 0.0
            for _ in range(args.iters):
                flush.zero_()
                a.record(); g.replay(); b.record()
                torch.cuda.synchronize()
                tot += a.elapsed_time(b)
            res[kind] = round(tot / args.iters * 1e3 / nl, 2)
    res["step_us_plain"] = round(res["plain"] - res["sleep-only"], 2)
    res["step_us_prefetched"] = round(res["prefetch"] - res["sleep-only"], 2)
    res["step_us_after_touch(+touch kernel)"] = round(res["touch"] - res["sleep-only"], 2)
    res["step_us_fused_next"] = round(res["fused-next"] - res["sleep-only"], 2)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
