"""Kernel-level breakdown (torch profiler, CUDA activities) of one cfg2 pretraining step: 32 x 512 tokens, bf16, train_step + clip 3 + Adam.
    python tools/profile_pretrain_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cpmusic
VOCAB = [56, 135, 18, 87, 18, 25]
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = cpmusic.TransformerModel(VOCAB, dropout=0.1).to(dev).train()
opt = torch.optim.Adam(m.parameters(), lr=1e-4, fused=True)
g = torch.Generator().manual_seed(1)
x = torch.stack([torch.randint(0, n, (32, 512), generator=g) for n in VOCAB], -1).to(dev); y = x.roll(-1, 1); mask = torch.ones(32, 512, device=dev)
def step():
    losses = m.train_step(x, y, mask)
    opt.zero_grad(set_to_none=True)
    (sum(losses) / 6).backward()
    torch.nn.utils.clip_grad_norm_(m.parameters(), 3.0, foreach=True)
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=26, max_name_column_width=70))
