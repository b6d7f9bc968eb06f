"""Per-kernel roofline table: every cpmusic kernel on the PPO update / rollout path timed alone at the
bench shapes (CUDA graph replay = GPU time only, L2 flushed before every replay, CUDA events) against the
HBM roofline with its ALGORITHMIC bytes (SURVEY §8d / DESIGN §4).

    python tools/bench_kernels.py [--tokens 65536] [--iters 10] [--only substr]

Prints one JSON line per kernel and a markdown table at the end (copied into profiles/)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cpmusic
from cpmusic import ops

VOCAB = [56, 135, 18, 87, 18, 25]
EMB = [128, 256, 64, 512, 128, 128]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=65536, help="tokens per update minibatch (64 x 1024)")
    ap.add_argument("--songs", type=int, default=256)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    T, d, dff, B = args.tokens, 512, 2048, args.songs
    bf = torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *s, dt=bf: torch.randn(*s, device=dev, generator=g).to(dt)
    rows = []

    def timeit(name, fn, alg_bytes, note=""):
        if args.only and args.only not in name:
            return
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        for _ in range(args.iters):
            flush.zero_()
            a.record(); graph.replay(); b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        ms = tot / args.iters
        gbs = alg_bytes / ms / 1e6
        r = {"kernel": name, "us": round(ms * 1e3, 1), "alg_MB": round(alg_bytes / 1e6, 1), "alg_GBps": round(gbs), "frac_hbm": round(gbs / peak, 3), "note": note}
        rows.append(r)
        print(json.dumps(r), flush=True)

    def fwd_bwd(name, make_out, inputs, bytes_f, bytes_b, note=""):
        """times forward alone and backward alone (backward = grad of a captured forward's output)."""
        timeit(name + " fwd", lambda: make_out(), bytes_f, note)
        if args.only and args.only not in name + " bwd":
            return
        out = make_out()
        go = torch.randn_like(out)
        timeit(name + " bwd", lambda: torch.autograd.grad(out, inputs, go, retain_graph=True), bytes_b, note)

    # ---- embedding gather (+ scatter-add backward)
    idx = torch.stack([torch.randint(0, n, (T,), device=dev, generator=g) for n in VOCAB], -1)
    tables = [torch.randn(n, e, device=dev, generator=g).requires_grad_() for n, e in zip(VOCAB, EMB)]
    fwd_bwd("embed", lambda: ops.cp_embed(idx, tables, bf), tables, T * (6 * 8 + 1216 * 2), T * (6 * 8 + 1216 * 2))
    # ---- positional encoding + dropout
    x = rnd(T // 1024, 1024, d).requires_grad_()
    pe = torch.randn(1, 20000, d, device=dev, generator=g)
    fwd_bwd("add_pe+dropout", lambda: ops.add_pe(x, pe, 1024, 0, None, 0.1), [x], T * d * 4, T * d * 4)
    # ---- residual + dropout + LayerNorm
    xr, res = rnd(T, d).requires_grad_(), rnd(T, d).requires_grad_()
    gam, bet = torch.ones(d, device=dev, requires_grad=True), torch.zeros(d, device=dev, requires_grad=True)
    fwd_bwd("ln_residual(+dropout)", lambda: ops.ln_residual(xr, res, gam, bet, 1e-5, 0.1), [xr, res, gam, bet],
            T * d * 2 * 4, T * d * 2 * 4, "fwd: x,res in; y,s out.  bwd: gy,s in; gs,gres out")
    rb = torch.zeros(d, device=dev, requires_grad=True)
    fwd_bwd("ln_residual(+bias+dropout)", lambda: ops.ln_residual(xr, res, gam, bet, 1e-5, 0.1, res_bias=rb), [xr, res, gam, bet, rb],
            T * d * 2 * 4, T * d * 2 * 4, "residual-branch bias folded in; bwd also returns its gradient")
    # ---- GELU + dropout
    h = rnd(T, dff).requires_grad_()
    fwd_bwd("gelu+dropout", lambda: ops.gelu_dropout(h, 0.1), [h], T * dff * 4, T * dff * 6)
    hb = torch.zeros(dff, device=dev, requires_grad=True)
    fwd_bwd("gelu+bias+dropout", lambda: ops.gelu_dropout(h, 0.1, bias=hb), [h, hb], T * dff * 4, T * dff * 6, "bwd also returns the bias gradient")
    gyb = rnd(T, dff)
    timeit("torch column sum (T x 2048 bf16 -> fp32)", lambda: gyb.sum(0, dtype=torch.float32), T * dff * 2, "what cpm_colsum replaces")
    timeit("colsum (T x 2048 bf16 -> fp32)", lambda: ops.colsum(gyb), T * dff * 2, "bias gradient of linear1")
    gy5 = rnd(T, 1536)[:, :512]
    timeit("colsum (T x 512 slice of 1536, bf16)", lambda: ops.colsum(gy5), T * 512 * 2, "strided view")
    # ---- heads: log-prob / entropy, masked CE
    seg = ops.seg_offsets(VOCAB)
    lg = rnd(T, 344).requires_grad_()
    tok = torch.stack([torch.randint(0, n, (T,), device=dev, generator=g) for n in VOCAB], -1)
    fwd_bwd("heads_logp+entropy", lambda: torch.cat(ops.heads_logp(lg, tok, seg, True), -1), [lg],
            T * (344 * 2 + 48 + 48), T * (344 * 4 + 48 + 48))
    mask = torch.ones(T, device=dev)
    fwd_bwd("masked_ce", lambda: ops.masked_ce(lg, tok, mask, seg), [lg], T * (344 * 2 + 48 + 4), T * (344 * 4 + 48 + 4 + 24))
    # ---- sampling (rollout: one row per song)
    lgs = rnd(B, 344)
    temp, topp = [1.2, 1.0, 1.2, 1.0, 2.0, 5.0], [0.9, 0.99, None, 0.9, 0.9, None]
    timeit("heads_sample (nucleus)", lambda: ops.heads_sample(lgs, seg, temp, topp, greedy=False, seed=1, want_logp=True), B * (344 * 2 + 72),
           "latency bound at 256 rows")
    # ---- recurrent step
    S, Z = torch.zeros(B, 8, 64, 64, device=dev), torch.zeros(B, 8, 64, device=dev)
    qkv = rnd(B, 1536)
    q, k, v = (qkv[:, j * 512:(j + 1) * 512].unflatten(-1, (8, 64)) for j in range(3))
    timeit("linattn_step", lambda: ops.linattn_step(q, k, v, S, Z), B * 270336 // 1)
    # ---- chunked linear attention
    N = T // 1024
    qkv2 = rnd(N, 1024, 1536).requires_grad_()
    fwd_bwd("linattn (tcgen05-cp)", lambda: ops.causal_linear_attention_fused(qkv2, 8), [qkv2], T * 8 * 512, T * 8 * 896)
    # ---- RL maths
    rew, val, done = torch.rand(B, 1024, device=dev), torch.randn(B, 1024, device=dev), torch.zeros(B, 1024, device=dev)
    last = torch.zeros(B, device=dev)
    timeit("returns_scan (GAE)", lambda: ops.returns_scan(rew, 0.99, "gae", val, done, last, 0.95), B * 1024 * 20, "latency bound (1 MB)")
    nl, ol, ad, en = (torch.randn(T, 6, device=dev, generator=g) * 0.1 for _ in range(4))
    nl.requires_grad_()
    timeit("ppo_loss fwd+bwd", lambda: ops.ppo_loss_standard(nl, ol, ad, None, None, en), T * 6 * 4 * 6)
    ql, nx = rnd(1024, 50, 344).requires_grad_(), rnd(1024, 50, 344)
    act = torch.stack([torch.randint(0, n, (1024, 25), device=dev, generator=g) for n in VOCAB], -1)
    rw, dn = torch.rand(1024, device=dev), torch.zeros(1024, device=dev)
    timeit("dqn_td fwd+bwd (cfg4)", lambda: ops.dqn_td_loss(ql, nx, act, rw, dn, seg, 25, 0.95, False), 1024 * 50 * 344 * 2 * 3,
           "reads Q, Q', writes dQ")
    print("\n| kernel | us | algorithmic MB | GB/s | frac of measured HBM (%.0f GB/s) | note |\n|---|---|---|---|---|---|" % peak)
    for r in rows:
        print(f"| {r['kernel']} | {r['us']} | {r['alg_MB']} | {r['alg_GBps']} | {r['frac_hbm']} | {r['note']} |")


if __name__ == "__main__":
    with torch.cuda.stream(torch.cuda.Stream()):     # autograd backward replays on the forward's stream: keep off the legacy stream
        main()
