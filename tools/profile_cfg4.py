"""In-situ kernel breakdown (torch profiler) of one cfg4 DQN / AIRL update (1024 windows x 50 tokens).  python tools/profile_cfg4.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, cpmusic
from torch.profiler import ProfilerActivity, profile
import bench_configs as bc
dev = torch.device("cuda:0")
VOCAB = bc.VOCAB
q = cpmusic.LinearTransformer(VOCAB, dropout=0.1).to(dev).train()
tgt = cpmusic.LinearTransformer(VOCAB, dropout=0.1).to(dev).eval()
head = cpmusic.rl.RewardHead(VOCAB, d_model=512).to(dev)
opt = torch.optim.Adam(q.parameters(), lr=1e-4, fused=True)
s, s2, mask = bc.batch(1024, 50, dev, seed=1)
act = torch.stack([torch.randint(0, n, (1024, 25), device=dev) for n in VOCAB], -1)
done = torch.zeros(1024, device=dev)
hidden = torch.randn(1024, 50, 512, device=dev).bfloat16()
def step():
    reward = head(hidden)
    td = cpmusic.rl.dqn_td_loss(q, tgt, s, s2, act, reward, done, compat=False)
    ce = sum(q.train_step(s, s2, mask)) / 6
    opt.zero_grad(set_to_none=True)
    (0.3 * td + 0.7 * ce).backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(r[2] for r in rows)
print(f"cfg4 update: {tot / 1e3:.2f} ms of kernels, {sum(r[1] for r in rows)} launches")
for k, n, t in sorted(rows, key=lambda r: -r[2])[:24]:
    print(f"{100 * t / tot:5.1f} %  n={n:5d}  avg {t / n:8.2f} us   {k[:100]}")
