import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, cpmusic
VOCAB = [56, 135, 18, 87, 18, 25]
dev = torch.device("cuda:0")
for N, cfg in ((5, dict(d_model=128, n_layer=2, n_head=2, d_inner=256)), (256, dict())):
    torch.manual_seed(5)
    m = cpmusic.LinearTransformer(VOCAB, dropout=0.0, **cfg).to(dev).eval()
    g = torch.Generator().manual_seed(12)
    init = torch.stack([torch.randint(0, n, (N,), generator=g) for n in VOCAB], -1).to(dev)
    for T in (1, 2, 3, 6):
        p = cpmusic.RolloutEngine(m, N, 8, greedy=True, mode="persistent")
        c = cpmusic.RolloutEngine(m, N, 8, greedy=True, mode="chain")
        a, b = p.generate(init, n_steps=T), c.generate(init, n_steps=T)
        dS = (p.S - c.S).abs().amax(dim=(1, 2, 3, 4)).tolist()
        dZ = (p.Z - c.Z).abs().amax(dim=(1, 2, 3)).tolist()
        tok = (a["tokens"] != b["tokens"]).float().mean(dim=(0, 2)).tolist()
        print(f"N={N} T={T}: token mismatch per step {[round(x, 3) for x in tok]} | max|dS| per layer {[f'{x:.1e}' for x in dS[:4]]} | max|dZ| {[f'{x:.1e}' for x in dZ[:4]]}", flush=True)
