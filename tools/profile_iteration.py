"""In-situ kernel breakdown (torch profiler / CUPTI, no replay, warm caches) of the bench iteration's two phases:
    python tools/profile_iteration.py [rollout_steps]
Prints one table for `rollout_steps` token steps of the captured rollout graph and one for the whole update phase
(value pass + GAE + actor update + critic update at the bench shapes)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
it = bench.PPOIteration(0, 1, dev)
for _ in range(2):
    it.step(it.init_dev)
it.flush()
torch.cuda.synchronize()


def table(prof, title, div):
    print(f"==== {title}")
    rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0]
    tot = sum(r[2] for r in rows)
    rows.sort(key=lambda r: -r[2])
    print(f"total device time {tot / 1e3:.2f} ms over {sum(r[1] for r in rows)} kernels; per unit ({div}): {tot / div:.1f} us")
    for k, n, t in rows[:32]:
        print(f"{100 * t / tot:5.1f} %  n={n:6d}  avg {t / n:8.2f} us   {k[:110]}")


eng = it.engine
eng.generate(it.init_dev, n_steps=8)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    eng.generate(it.init_dev, n_steps=steps)
    torch.cuda.synchronize()
table(prof, f"rollout: {steps} token steps of the captured graph ({eng.launches_per_step} cpmusic kernels per step)", steps)

# the update phase alone: feed step() a finished rollout by patching generate
roll = eng.generate(it.init_dev)
torch.cuda.synchronize()
eng_generate = eng.generate
eng.generate = lambda *_a, **_k: roll
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    it.step(it.init_dev)
    it.flush()
    torch.cuda.synchronize()
eng.generate = eng_generate
table(prof, "update phase (value pass, GAE, actor + critic update, 2 minibatches of 128 x 1024 each)", 1)
