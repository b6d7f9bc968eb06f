"""Small invocations of the hand-rolled mbarrier / TMEM / TMA kernels for compute-sanitizer (memcheck, racecheck, synccheck):
chunk-parallel linear attention fwd + bwd (per-chunk-state and streaming dispatch), recurrent step, 2-CTA GEMMs (nt with each
epilogue, tn), the token-step GEMM with and without the LayerNorm fold, a few rollout steps through the chain and through the
persistent kernel.   compute-sanitizer --tool memcheck python tools/sanitize_kernels.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cpmusic
from cpmusic import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
which = sys.argv[1:] or ["attn", "step", "gemm", "small", "rollout"]
if "attn" in which:
    for (N, L, H) in ((2, 256, 4), (13, 256, 8)):            # 8 chains: per-chunk states + scan; 104 chains: streaming kernels
        q, k, v, go = (torch.randn(N, L, H, 64, device=dev).bfloat16().requires_grad_() for _ in range(4))
        out = ops.causal_linear_attention(q, k, v)
        out.backward(go.detach())
        torch.cuda.synchronize()
        print("attn", (N, L, H), ops.linattn_last_impl(), float(out.float().abs().mean()), flush=True)
    for (N, L, H, E) in ((3, 50, 4, 64), (13, 200, 8, 64), (2, 256, 2, 128), (2, 200, 2, 128)):    # ragged lengths; 128-wide heads
        q, k, v, go = (torch.randn(N, L, H, E, device=dev).bfloat16().requires_grad_() for _ in range(4))
        out = ops.causal_linear_attention(q, k, v)
        out.backward(go.detach())
        torch.cuda.synchronize()
        print("attn", (N, L, H, E), ops.linattn_last_impl(), float(out.float().abs().mean()), flush=True)
if "step" in which:
    S, Z = torch.zeros(5, 4, 64, 64, device=dev), torch.zeros(5, 4, 64, device=dev)
    qkv = torch.randn(5, 3 * 256, device=dev).bfloat16()
    q, k, v = (qkv[:, j * 256:(j + 1) * 256].unflatten(-1, (4, 64)) for j in range(3))
    for _ in range(3):
        o = ops.linattn_step(q, k, v, S, Z)
    torch.cuda.synchronize()
    print("step", float(o.float().abs().mean()), flush=True)
if "gemm" in which:
    a = torch.randn(1300, 512, device=dev).bfloat16()
    w = torch.randn(1024, 512, device=dev).bfloat16() * 0.05
    b = torch.randn(1024, device=dev)
    for mode in (1, 3):
        ops.gemm_set_mode(mode)
        d = ops.gemm_nt(a, w, b)
        h, g = ops.gemm_nt(a, w, b, epilogue=ops.GEMM_GELU, p_drop=0.1, seed=3, rng_offset=16)
        dh = ops.gemm_nt(a, w, None, epilogue=ops.GEMM_DGELU, aux=h, p_drop=0.1, seed=3, rng_offset=16)
    ops.gemm_set_mode(0)
    gw = torch.zeros(1024, 512, device=dev)
    ops.gemm_tn_acc(d, a, gw)
    torch.cuda.synchronize()
    print("gemm", float(d.float().abs().mean()), float(gw.abs().mean()), flush=True)
if "small" in which:
    a = torch.randn(70, 512, device=dev).bfloat16()
    w = torch.randn(344, 512, device=dev).bfloat16() * 0.05
    b = torch.randn(344, device=dev)
    d = ops.gemm_nt_small(a, w, b, gelu=True)
    st = torch.empty(70, 2, device=dev)
    f = ops.gemm_nt_small_ln(a, w, b, fold_c1=w.float().sum(1), stats_out=st)
    r = ops.gemm_nt_small_ln(a, w, b, resid=d, r_stats=st, r_gamma=torch.ones(344, device=dev), r_beta=torch.zeros(344, device=dev))
    torch.cuda.synchronize()
    print("small", float(d.float().abs().mean()), float(f.float().abs().mean()), float(r.float().abs().mean()), flush=True)
if "rollout" in which:
    VOCAB = [56, 135, 18, 87, 18, 25]
    m = cpmusic.LinearTransformer(VOCAB, dropout=0.0, d_model=128, n_layer=2, n_head=2, d_inner=256).to(dev).eval()
    init = torch.stack([torch.randint(0, n, (7,)) for n in VOCAB], -1).to(dev)
    for mode in ("chain", "persistent"):
        eng = cpmusic.RolloutEngine(m, 7, 6, greedy=False, seed=1, mode=mode, use_graph=(mode == "persistent"))
        t = eng.generate(init)["tokens"]
        torch.cuda.synchronize()
        print("rollout", mode, eng.mode, t.shape, flush=True)
print("sanitize_kernels: done")
