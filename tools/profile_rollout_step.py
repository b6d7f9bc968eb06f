"""Runs a few eager (non-graph) rollout token steps of the full-size model — a compact target for
`ncu --metrics gpu__time_duration.sum` (per-kernel durations of one token step).
    python tools/profile_rollout_step.py [--mode mega|fused|unfused] [--steps 4] [--graph]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cpmusic

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="mega", choices=["mega", "fused", "unfused", "tc", "fold"])
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--graph", action="store_true")
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--pdl", action="store_true")
ap.add_argument("--lazy", action="store_true")
args = ap.parse_args()
VOCAB = [56, 135, 18, 87, 18, 25]
torch.manual_seed(0)
dev = torch.device("cuda:0")
m = cpmusic.LinearTransformer(VOCAB).to(dev).eval()
eng = cpmusic.RolloutEngine(m, args.batch, max(args.steps, 64), greedy=False, use_graph=args.graph, mode=args.mode, pdl=args.pdl, lazy_state=args.lazy)
init = torch.stack([torch.randint(0, n, (args.batch,)) for n in VOCAB], -1).to(dev)
eng.generate(init, args.steps)
torch.cuda.synchronize()
t0 = time.perf_counter()
torch.cuda.nvtx.range_push("steps")
eng.generate(init, args.steps)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("us/step (incl. host overhead when eager):", (time.perf_counter() - t0) / args.steps * 1e6, "launches/step", eng.launches_per_step)
