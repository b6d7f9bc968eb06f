"""Timeline of the token-step GEMM kernels inside the rollout chain (cpm_debug_small_timing): for one token step, per launch of
cpm_gemm_nt_small[_ln], when block (0,0) entered, finished its set-up, got past griddepcontrol.wait, saw its first activation
block, had its accumulator, and had stored its tile - and the gap to the next GEMM kernel.
    python tools/phase_timing_chain.py [songs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cpmusic
VOCAB = [56, 135, 18, 87, 18, 25]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = cpmusic.LinearTransformer(VOCAB, dropout=0.1).to(dev)
init = torch.stack([torch.randint(0, n, (N,)) for n in VOCAB], -1).to(dev)
lib = cpmusic._lib.load()
cap = 4096
log = torch.zeros(1 + cap * 8, dtype=torch.int64, device=dev)
lib.cpm_debug_small_timing(log.data_ptr(), cap)          # baked into the graph captured below
eng = cpmusic.RolloutEngine(m, N, 64, greedy=False, seed=1, mode="chain")
eng.generate(init, n_steps=4)
torch.cuda.synchronize()
log.zero_()
eng.generate(init, n_steps=6)
torch.cuda.synchronize()
lib.cpm_debug_small_timing(None, 0)
n = int(log[0])
t = log[1:1 + n * 8].view(n, 8).cpu().double()
per = 51                                                  # GEMM launches per token step (in, 12 x 4, heads, + layer-0 plain QKV counted)
per = n // 6
step = t[-per:]                                           # the last token step
t0 = step[0, 0]
print(f"{n} GEMM launches logged, {per} per token step; last step spans {(step[-1, 5] - t0) / 1e3:.1f} us from the first GEMM's entry")
print(" idx    N     K | entry  setup  dep-wait  A-landed  acc-ready  stored | (us since this kernel's entry)   gap from previous GEMM's store to this dep-wait return")
tot = {"entry->dep": 0.0, "dep->A": 0.0, "A->acc": 0.0, "acc->stored": 0.0, "gap": 0.0}
for i in range(per):
    e, s_, d, a, c, st = (step[i, j] for j in range(6))
    gap = (d - step[i - 1, 5]) / 1e3 if i else float("nan")
    if i < 14 or i >= per - 3:
        print(f"{i:4d} {int(step[i, 6]):5d} {int(step[i, 7]):5d} | {0:5.2f} {(s_ - e) / 1e3:6.2f} {(d - e) / 1e3:8.2f} {((a - e) / 1e3 if a > 0 else float("nan")):9.2f} {(c - e) / 1e3:10.2f} {(st - e) / 1e3:7.2f} | gap {gap:6.2f}")
    if a > 0:                                            # K > 512 launches stream through the ring and do not stamp "A landed"
        tot["dep->A"] += (a - d) / 1e3; tot["A->acc"] += (c - a) / 1e3
    else:
        tot["A->acc"] += (c - d) / 1e3
    tot["acc->stored"] += (st - c) / 1e3
    if i: tot["gap"] += gap
print("sums over the step (us):", {k: round(float(v), 1) for k, v in tot.items() if k != "entry->dep"})
