"""Two half-batch token-step chains side by side: does the GPU run two 128-song rollouts concurrently faster than one 256-song
rollout?  Each half has its own RolloutEngine (own recurrent state, own CUDA graph, global sequence ids so the tokens are those of
the single engine) and its own stream - plain streams, or two complementary SM partitions (green contexts).
    python tools/probes/dual_rollout_probe.py [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, cpmusic
VOCAB = [56, 135, 18, 87, 18, 25]
T = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = cpmusic.LinearTransformer(VOCAB, dropout=0.1).to(dev)
init = torch.stack([torch.randint(0, n, (256,)) for n in VOCAB], -1).to(dev)

def timed(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best * 1e3 / T

one = cpmusic.RolloutEngine(m, 256, T, greedy=False, seed=1)
one.generate(init, n_steps=4)
ref = one.generate(init)["tokens"]
print(f"one engine, 256 songs: {timed(lambda: one.generate(init)):7.1f} us per token step", flush=True)

def dual(streams, label):
    K = len(streams)
    W = 256 // K
    halves = [cpmusic.RolloutEngine(m, W, T, greedy=False, seed=1, seq_base=W * i) for i in range(K)]
    for i, (e, s) in enumerate(zip(halves, streams)):
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            e.generate(init[W * i:W * i + W], n_steps=4)          # captures the graph on this stream
    torch.cuda.synchronize()
    def run():
        cur = torch.cuda.current_stream()
        for i, (e, s) in enumerate(zip(halves, streams)):
            e.model.eval()
            e.reset(init[W * i:W * i + W])
            e.model.refresh_packs()
            if e.fold:
                e._fold_refresh()
            s.wait_stream(cur)
        for _ in range(T):
            for e, s in zip(halves, streams):
                with torch.cuda.stream(s):
                    e.graph.replay()
        for s in streams:
            cur.wait_stream(s)
    us = timed(run)
    toks = torch.cat([torch.cat([init[W * i:W * i + W][None], e.hist_tok[:T]], 0).permute(1, 0, 2) for i, e in enumerate(halves)], 0)
    same = bool((toks == ref).all())
    print(f"{K} engines of {W} songs, {label}: {us:7.1f} us per token step (both halves); tokens equal the single engine's: {same}", flush=True)

dual([torch.cuda.Stream(priority=-1), torch.cuda.Stream(priority=-1)], "two plain streams")
if not os.environ.get("QUICK"):
    dual([torch.cuda.Stream(priority=-1) for _ in range(4)], "four plain streams")
for W1 in (() if os.environ.get("QUICK") else (128, 64, 32)):
    e1 = cpmusic.RolloutEngine(m, W1, T, greedy=False, seed=1)
    e1.generate(init[:W1], n_steps=4)
    print(f"one engine, {W1} songs alone: {timed(lambda: e1.generate(init[:W1])):7.1f} us per token step", flush=True)
    del e1
try:
    from cuda.bindings import driver as drv
    def ck(r):
        if r[0] != drv.CUresult.CUDA_SUCCESS: raise RuntimeError(str(r[0]))
        return r[1:] if len(r) > 2 else r[1]
    cudev = ck(drv.cuDeviceGet(0))
    res = ck(drv.cuDeviceGetDevResource(cudev, drv.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
    groups, nb, rem = ck(drv.cuDevSmResourceSplitByCount(1, res, 0, 72))
    streams, keep = [], []
    for r in (groups[0], rem):
        desc = ck(drv.cuDevResourceGenerateDesc([r], 1))
        g = ck(drv.cuGreenCtxCreate(desc, cudev, drv.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
        h = ck(drv.cuGreenCtxStreamCreate(g, drv.CUstream_flags.CU_STREAM_NON_BLOCKING, -1))
        keep.append((g, h)); streams.append(torch.cuda.ExternalStream(int(h)))
    print("partitions:", groups[0].sm.smCount, rem.sm.smCount, flush=True)
    dual(streams, "two SM partitions")
except Exception as e:
    print("green contexts:", e)
