"""Where does the overlapped critic update go?  CUDA events on both streams of bench.py's iteration: the rollout's duration with and
without the critic update queued under it, and the critic update's own duration under the rollout / alone.
    python tools/probes/overlap_timeline.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
import cpmusic
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
cpmusic._lib.load()
it = bench.PPOIteration(0, 1, dev)
for _ in range(3):
    it.step(it.init_dev)
it.flush()
torch.cuda.synchronize()
E = lambda: torch.cuda.Event(enable_timing=True)
# (a) rollout alone
a0, a1 = E(), E()
a0.record(); it.engine.generate(it.init_dev); a1.record(); torch.cuda.synchronize()
print("rollout alone: %.1f ms" % a0.elapsed_time(a1))
# (b) a pending critic update exists after a step; time the next rollout with it queued underneath
for rep in range(2):
    it.step(it.init_dev)                       # leaves a pending critic update
    torch.cuda.synchronize()
    r0, r1, c0, c1 = E(), E(), E(), E()
    r0.record()
    roll = it.engine.generate(it.init_dev)
    r1.record()
    it.cstream.wait_event(r0)
    with torch.cuda.stream(it.cstream):
        c0.record()
    it._launch_pending()
    with torch.cuda.stream(it.cstream):
        c1.record()
    torch.cuda.synchronize()
    print("overlapped: rollout %.1f ms, critic update %.1f ms (starts %.1f ms after the rollout, ends %.1f ms after its start)"
          % (r0.elapsed_time(r1), c0.elapsed_time(c1), r0.elapsed_time(c0), r0.elapsed_time(c1)))
# (c) critic update alone
it.step(it.init_dev)
torch.cuda.synchronize()
c0, c1 = E(), E()
with torch.cuda.stream(it.cstream):
    c0.record()
it._launch_pending()
with torch.cuda.stream(it.cstream):
    c1.record()
torch.cuda.synchronize()
print("critic update alone: %.1f ms" % c0.elapsed_time(c1))
