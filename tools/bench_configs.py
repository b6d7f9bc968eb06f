"""Throughput of the other BASELINE.json configurations on one GPU (the headline metric, configs[2], is bench.py):
  cfg2  teacher-forced pretraining step, bf16, batch 32 x seq 512, 12 layers / d 512 / 8 heads (train_step + Adam + clip 3)
  cfg4  DQN / AIRL update: Q and target nets on 1024 windows of 50 tokens + fused TD loss + CE term + Adam; reward head on 1024 windows
  cfg5  long-sequence stress: 24 layers, d_model 1024 (16 heads), seq 8192, batch 1: train_step forward + backward
CUDA events, 3 warm-up + `--iters` timed iterations.   python tools/bench_configs.py [--iters 10] [--only cfg2]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cpmusic

VOCAB = [56, 135, 18, 87, 18, 25]


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def batch(N, L, dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.stack([torch.randint(0, n, (N, L), generator=g) for n in VOCAB], -1).to(dev)
    mask = torch.ones(N, L, device=dev)
    return x, x.roll(-1, 1), mask


CEIL_CFG2 = 5.94e6          # tokens/s: 236.1 MFLOP per token fwd+bwd at the sustained 1402 TFLOP/s


def run_cfg2(dev, iters=10, graphed=True, eager=True):
    out = []
    if eager:
        m = cpmusic.TransformerModel(VOCAB, dropout=0.1).to(dev).train()
        opt = torch.optim.Adam(m.parameters(), lr=1e-4, fused=True)
        x, y, mask = batch(32, 512, dev)

        def step():
            losses = m.train_step(x, y, mask)
            opt.zero_grad(set_to_none=True)
            (sum(losses) / 6).backward()
            torch.nn.utils.clip_grad_norm_(m.parameters(), 3.0, foreach=True)
            opt.step()
        ms = timed(step, iters)
        tok = 32 * 512
        out.append({"config": "cfg2 pretrain step 32x512 bf16 (agent_pretrain.py:557-565)", "ms_per_step": round(ms, 3),
                    "tokens_per_s": round(tok / ms * 1e3), "model_tflops": round(tok * 236.1e6 / ms / 1e9, 1),
                    "frac_of_ceiling": round(tok / ms * 1e3 / CEIL_CFG2, 3)})
        del m, opt
    if graphed:
        # the same step captured once as a CUDA graph (cpmusic.GraphedTrainStep): no host launch cost per kernel
        m = cpmusic.TransformerModel(VOCAB, dropout=0.1).to(dev).train()
        opt = torch.optim.Adam(m.parameters(), lr=1e-4, fused=True, capturable=True)
        gs = cpmusic.GraphedTrainStep(m, opt, batch_size=32, seq_len=512, max_grad_norm=3.0)
        x, y, mask = batch(32, 512, dev)
        ms = timed(lambda: gs(x, y, mask), iters)
        tok = 32 * 512
        out.append({"config": "cfg2 pretrain step 32x512 bf16, whole-step CUDA graph", "ms_per_step": round(ms, 3),
                    "tokens_per_s": round(tok / ms * 1e3), "model_tflops": round(tok * 236.1e6 / ms / 1e9, 1),
                    "frac_of_ceiling": round(tok / ms * 1e3 / CEIL_CFG2, 3)})
        del m, opt, gs
    return out


def run_cfg4(dev, iters=10):
    q = cpmusic.LinearTransformer(VOCAB, dropout=0.1).to(dev).train()
    tgt = cpmusic.LinearTransformer(VOCAB, dropout=0.1).to(dev).eval()
    head = cpmusic.rl.RewardHead(VOCAB, d_model=512).to(dev)
    opt = torch.optim.Adam(q.parameters(), lr=1e-4, fused=True)
    s, s2, mask = batch(1024, 50, dev, seed=1)
    act = torch.stack([torch.randint(0, n, (1024, 25), device=dev) for n in VOCAB], -1)
    done = torch.zeros(1024, device=dev)
    hidden = torch.randn(1024, 50, 512, device=dev).bfloat16()          # stands in for the Longformer body's last hidden state

    def step():
        reward = head(hidden)                                            # AIRL reward read-out (fused)
        td = cpmusic.rl.dqn_td_loss(q, tgt, s, s2, act, reward, done, compat=False)
        ce = sum(q.train_step(s, s2, mask)) / 6
        opt.zero_grad(set_to_none=True)
        (0.3 * td + 0.7 * ce).backward()                                 # IRL_dqn_train.py:335
        opt.step()
    ms = timed(step, iters)
    # work: Q net fwd+bwd on s (x2: the TD term and the CE term each run the net) + target net fwd on s2
    flop = 1024 * 50 * 78.69e6 * (3 + 3 + 1)
    return [{"config": "cfg4 DQN/AIRL update, replay batch 1024 x 50 tokens (IRL_dqn_train.py:285-345)", "ms_per_update": round(ms, 3),
             "sequences_per_s": round(1024 / ms * 1e3), "tokens_per_s": round(1024 * 50 / ms * 1e3),
             "frac_of_ceiling": round(flop / (ms * 1e-3) / 1402.2e12, 3)}]


def run_cfg5(dev, iters=5, heads_list=(16, 8)):
    out = []
    for heads in heads_list:                    # BASELINE.json does not say which: 16 x 64 or 8 x 128
        m = cpmusic.TransformerModel(VOCAB, d_model=1024, n_layer=24, n_head=heads, d_inner=4096, dropout=0.1).to(dev).train()
        x, y, mask = batch(1, 8192, dev, seed=2)

        def step():
            losses = m.train_step(x, y, mask)
            m.zero_grad(set_to_none=True)
            (sum(losses) / 6).backward()
        ms = timed(step, iters)
        params = sum(p.numel() for p in m.parameters())
        flop = 8192 * 6.0 * (params - sum(t.numel() for t in m._tables()))        # 6 FLOP per weight and token, fwd + bwd
        out.append({"config": f"cfg5 long sequence: 24 layers, d 1024, {heads} heads x {1024 // heads}, seq 8192, batch 1, fwd+bwd",
                    "ms_per_step": round(ms, 3), "tokens_per_s": round(8192 / ms * 1e3), "params_M": round(params / 1e6, 1),
                    "frac_of_ceiling": round(flop / (ms * 1e-3) / 1402.2e12, 3)})
        del m
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    rows = []
    if not args.only or args.only == "cfg2":
        rows += run_cfg2(dev, args.iters)
    if not args.only or args.only == "cfg4":
        rows += run_cfg4(dev, args.iters)
    if not args.only or args.only == "cfg5":
        rows += run_cfg5(dev, max(args.iters // 2, 2))
    for r in rows:
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
