"""Per-stage timeline of the persistent rollout-step kernel (cpm_debug_rollout_timing): for one token step, the time every
stage takes from "previous barrier passed" to "work done" (max / median over CTAs) and the barrier wait after it.
    python tools/phase_timing_rollout.py [songs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cpmusic
VOCAB = [56, 135, 18, 87, 18, 25]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = cpmusic.LinearTransformer(VOCAB, dropout=0.1).to(dev)
init = torch.stack([torch.randint(0, n, (N,)) for n in VOCAB], -1).to(dev)
eng = cpmusic.RolloutEngine(m, N, 64, greedy=False, seed=1, mode="persistent")
eng.generate(init, n_steps=8)
lib = cpmusic._lib.load()
n_ph, G = eng._plan["phases"], torch.cuda.get_device_properties(0).multi_processor_count
buf = torch.zeros(G, 4, n_ph, 8, dtype=torch.int64, device=dev)
lib.cpm_debug_rollout_timing(buf.data_ptr())
eng.generate(init, n_steps=8)
torch.cuda.synchronize()
lib.cpm_debug_rollout_timing(None)
t = buf.cpu().double()
st = 2                                                   # a warm step
done, passed = t[:, st, :, 0], t[:, st, :, 1]
prev = torch.cat([t[:, st - 1, -1:, 1], passed[:, :-1]], 1)           # barrier before the stage
work = (done - prev) / 1e3
wait = (passed - done) / 1e3
staged, accr = t[:, st, :, 2], t[:, st, :, 3]
has = t[:, st, :, 3] > 0                                          # CTAs that owned a tile of the stage
t_stage = torch.where(has, (staged - prev) / 1e3, torch.zeros_like(work))
t_mma = torch.where(has, (accr - staged) / 1e3, torch.zeros_like(work))
t_epi = torch.where(has, (done - accr) / 1e3, torch.zeros_like(work))
landed, lnd = t[:, st, :, 4], t[:, st, :, 5]
hl = landed > 0
t_land = torch.where(hl, (landed - prev) / 1e3, torch.zeros_like(work))
t_ln = torch.where(hl, (lnd - landed) / 1e3, torch.zeros_like(work))
t_fence = torch.where(hl, (staged - lnd) / 1e3, torch.zeros_like(work))
ta, ti = t[:, st, :, 6], t[:, st, :, 7]
hm = ta > 0
for i in (1, 3, 4, 5):
    n_has = max(int(hm[:, i].sum()), 1)
    f = lambda x: torch.where(hm[:, i], x[:, i], torch.zeros_like(x[:, i])).sum().item() / n_has / 1e3
    print(f"stage {i}: UMMA thread: activation barrier seen {f(ta - prev):.2f} us after the grid barrier ({f(ta - staged):.2f} after thread 0 staged), "
          f"all UMMAs issued {f(ti - ta):.2f} us later, accumulator seen by the workers {f(accr - ti):.2f} us after that")
for i in (1, 4):
    n_has = max(int(hl[:, i].sum()), 1)
    print(f"stage {i}: LN staging of thread 0: loads landed after {t_land[:, i].sum() / n_has:.2f} us, LayerNorm + stores {t_ln[:, i].sum() / n_has:.2f} us, "
          f"fence + arrive {t_fence[:, i].sum() / n_has:.2f} us")
ck = t[:, st, :, 2]
print(f"SM clock during the step (clock64 / globaltimer between the first state stage and the sampler, CTA 0): "
      f"{(ck[0, -1] - ck[0, 2]) / (done[0, -1] - done[0, 2]) * 1e3:.0f} MHz")
names = ["in"] + [f"L{l}.{n}" for l in range((n_ph - 3) // 5) for n in ("qkv", "state", "out", "ff1", "ff2")] + ["heads", "sample"]
print(f"step {st}: {(passed[:, -1].max() - prev[:, 0].min()) / 1e3:.1f} us total")
agg = {}
for i, n in enumerate(names):
    k = n.split(".")[-1]
    a = agg.setdefault(k, [0.0, 0.0, 0.0, 0, 0.0, 0.0, 0.0])
    a[0] += work[:, i].max().item(); a[1] += work[:, i].median().item(); a[2] += wait[:, i].min().item(); a[3] += 1
    n_has = max(int(has[:, i].sum()), 1)
    a[4] += t_stage[:, i].sum().item() / n_has; a[5] += t_mma[:, i].sum().item() / n_has; a[6] += t_epi[:, i].sum().item() / n_has
    if i < 7 or i >= n_ph - 2:
        print(f"{n:10s} work max {work[:, i].max():7.2f} med {work[:, i].median():7.2f} us | barrier wait min {wait[:, i].min():6.2f} med {wait[:, i].median():6.2f}")
print("--- per stage kind, averaged over its occurrences")
for k, (mx, md, wt, c, ts, tm, te) in agg.items():
    print(f"{k:8s} x{c:3d}: work max {mx / c:7.2f} med {md / c:7.2f} us, barrier (min wait) {wt / c:6.2f} us -> {mx + wt:8.1f} us per step"
          f" | mean over owning CTAs: stage A {ts / c:5.2f}, wait UMMA {tm / c:5.2f}, epilogue {te / c:5.2f}")
