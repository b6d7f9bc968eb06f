"""Per-token time of the rollout step, persistent kernel vs kernel chain (256 songs, full-size model):
    python tools/bench_rollout.py [songs] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cpmusic
VOCAB = [56, 135, 18, 87, 18, 25]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = cpmusic.LinearTransformer(VOCAB, dropout=0.1).to(dev)
init = torch.stack([torch.randint(0, n, (N,)) for n in VOCAB], -1).to(dev)
for mode in (sys.argv[3:] or ["persistent", "chain"]):
    eng = cpmusic.RolloutEngine(m, N, T, greedy=False, seed=1, mode=mode)
    eng.generate(init, n_steps=8)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        a.record()
        eng.generate(init)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f"{mode:10s} songs={N} steps={T}: {best * 1e3 / T:7.1f} us per token step ({eng.launches_per_step} "
          f"{'stages' if eng.mode == 'persistent' else 'kernels'} per step), mode ran: {eng.mode}", flush=True)
