"""The persistent rollout-step kernel (csrc/rollout_step.cu, cpm_rollout_run) against the kernel chain it replaces, the
teacher-forced parallel model and - through them - the oracle.  Reference loop: testing-no-type-cp.py:157-167."""
import pytest
import torch

pytestmark = pytest.mark.gpu

VOCAB = [56, 135, 18, 87, 18, 25]


def _init(N, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.stack([torch.randint(0, n, (N,), generator=g) for n in VOCAB], -1)


def _model(cpm, cuda, seed=3, **cfg):
    torch.manual_seed(seed)
    return cpm.LinearTransformer(VOCAB, dropout=0.0, **cfg).to(cuda).eval()


def _pair(cpm, m, N, T, **kw):
    p = cpm.RolloutEngine(m, N, T, mode="persistent", **kw)
    c = cpm.RolloutEngine(m, N, T, mode="chain", fold_ln=False, **kw)       # like with like: the chain with its LayerNorm kernels
    assert p.mode == "persistent" and c.mode == "chain"
    return p, c


@pytest.fixture
def like_with_like(cpm):
    """Persistent kernel vs chain: the chain without its K split for linear2 (a different fp32 summation order), so that the two
    executions differ only in what the persistent kernel's tests are about."""
    lib = cpm._lib.load()
    lib.cpm_gemm_small_set_split(0)
    yield
    lib.cpm_gemm_small_set_split(1)


def _agree(a, b):
    return (a == b).float().mean().item()


@pytest.mark.parametrize("N,cfg", [(5, dict(d_model=128, n_layer=2, n_head=2, d_inner=256)),
                                   (37, dict(d_model=256, n_layer=3, n_head=4, d_inner=512)),
                                   (256, dict()), (300, dict())])
def test_one_step_logits_and_state_equal_the_chain(cuda, cpm, like_with_like, N, cfg):
    """One token step: logits, sampled tokens, recorded log-probs and the recurrent state (S, Z of every layer) of the
    persistent kernel against the kernel chain.  Every rounding sits where the chain has it; what may differ is the fp32
    accumulation inside the tensor core (weights vs songs on the UMMA M axis): about one bf16 ulp in one of 1e5 values."""
    m = _model(cpm, cuda, **cfg)
    init = _init(N, 11).to(cuda)
    p, c = _pair(cpm, m, N, 4, greedy=False, seed=77)
    a, b = p.generate(init, n_steps=1), c.generate(init, n_steps=1)
    assert p.mode == "persistent"                                   # did not fall back
    assert torch.allclose(p.S, c.S, rtol=2e-2, atol=2e-2) and (p.S != c.S).float().mean() < 1e-3
    assert torch.allclose(p.Z, c.Z, rtol=2e-2, atol=2e-2)
    c.reset(init)
    with torch.no_grad():
        m.refresh_packs()
        lc = c._logits()
    got = p._plan["scratch"]["logits"][:, :m.seg[-1]].float()
    want = lc[:, :m.seg[-1]].float()
    assert (got - want).abs().max() < 3e-2 and (got != want).float().mean() < 2e-2
    assert _agree(a["tokens"], b["tokens"]) > 0.995 and (a["logp"] - b["logp"]).abs().max() < 5e-2


@pytest.mark.parametrize("greedy", [True, False])
def test_many_steps_track_the_chain_full_size(cuda, cpm, like_with_like, greedy):
    """48 tokens of 256 songs on the 12-layer / d512 model.  Against the kernel chain: the first tokens agree and the
    agreement decays slowly (a one-ulp difference flips a near-tie now and then, and a song that took another token is a
    different song from there on).  Within the kernel: a second run after reset reproduces the first bit for bit, and one
    launch of 48 steps equals 48 launches of one step through the C-ABI."""
    m = _model(cpm, cuda, seed=5)
    N, T = 256, 48
    init = _init(N, 12).to(cuda)
    p, c = _pair(cpm, m, N, T, greedy=greedy, seed=9)
    a, b = p.generate(init), c.generate(init)
    assert _agree(a["tokens"][:, :3], b["tokens"][:, :3]) > 0.995
    assert _agree(a["tokens"][:, :9], b["tokens"][:, :9]) > 0.95
    S1, lp1 = p.S.clone(), a["logp"].clone()
    again = p.generate(init)
    assert torch.equal(again["tokens"], a["tokens"]) and torch.equal(again["logp"], lp1) and torch.equal(p.S, S1)
    p.reset(init)
    p._plan["barrier"].zero_()
    for _ in range(T):                                              # step by step through the C-ABI
        cpm._lib.check(cpm._lib.load().cpm_rollout_run(p._plan["handle"], 1, torch.cuda.current_stream().cuda_stream))
    assert torch.equal(p.hist_tok[:T].permute(1, 0, 2), a["tokens"][:, 1:]) and torch.equal(p.S, S1)
    assert int(p.step_dev.item()) == T


def test_persistent_greedy_equals_teacher_forced_argmax(cuda, cpm):
    """The greedy tokens of the persistent kernel are the arg-max of the PARALLEL (chunked tcgen05 attention) model run
    teacher-forced over them, wherever the top-2 margin exceeds the bf16 noise: an independent check of positions, state
    recurrence, LayerNorm placement and heads."""
    m = _model(cpm, cuda, seed=8)
    mp = cpm.TransformerModel(VOCAB, dropout=0.0).to(cuda).eval()
    mp.load_state_dict(m.state_dict())
    N, T = 64, 128
    init = _init(N, 4).to(cuda)
    out = cpm.RolloutEngine(m, N, T, greedy=True, mode="persistent").generate(init)
    with torch.no_grad():
        lc = mp.logits_concat(mp.hidden(out["tokens"][:, :-1])).float()
    ok = tot = 0
    for a in range(6):
        seg = lc[..., mp.seg[a]:mp.seg[a + 1]]
        top2 = seg.topk(2, -1).values
        sure = (top2[..., 0] - top2[..., 1]) > 0.08
        ok += ((seg.argmax(-1) == out["tokens"][:, 1:, a]) & sure).sum().item()
        tot += sure.sum().item()
    assert tot > 0.5 * N * T * 6 and ok == tot, (ok, tot)
    lp, _ = cpm.ops.heads_logp(lc.bfloat16(), out["tokens"][:, 1:], mp.seg, False)
    assert (out["logp"] - lp).abs().max() < 8e-2


def test_persistent_rollout_sees_optimizer_updates_and_seed_changes(cuda, cpm, like_with_like):
    m = _model(cpm, cuda, seed=2, d_model=128, n_layer=2, n_head=2, d_inner=256)
    N, T = 9, 12
    init = _init(N, 6).to(cuda)
    p, c = _pair(cpm, m, N, T, greedy=False, seed=1)
    before = p.generate(init)["tokens"].clone()
    assert _agree(before[:, :4], c.generate(init)["tokens"][:, :4]) > 0.97
    with torch.no_grad():
        for q in m.parameters():
            q.add_(0.05 * torch.randn_like(q))
    after = p.generate(init)["tokens"].clone()                      # the plan reads the refreshed packs by address
    assert _agree(after[:, :4], c.generate(init)["tokens"][:, :4]) > 0.97 and not torch.equal(after, before)
    other = p.generate(init, seed=2)["tokens"]
    assert _agree(other[:, :4], c.generate(init, seed=2)["tokens"][:, :4]) > 0.97 and not torch.equal(other, after)
    # sharding invariance: songs 4.. on their own engine draw the same tokens (same kernel, same Philox streams)
    part = cpm.RolloutEngine(m, N - 4, T, greedy=False, seed=2, seq_base=4, mode="persistent").generate(init[4:])["tokens"]
    assert torch.equal(part, other[4:])


def test_unsupported_shapes_fall_back_to_the_chain_and_bad_tokens_raise(cuda, cpm):
    m7 = cpm.LinearTransformer([56, 135, 18, 4, 87, 18, 25], dropout=0.0, d_model=128, n_layer=1, n_head=2, d_inner=256).to(cuda).eval()
    init7 = torch.zeros(3, 7, dtype=torch.int64, device=cuda)
    eng = cpm.RolloutEngine(m7, 3, 4, greedy=True, mode="persistent")
    assert eng.generate(init7)["tokens"].shape == (3, 5, 7) and eng.mode == "persistent"      # 7 attributes are fine
    wide = cpm.LinearTransformer(VOCAB, dropout=0.0, d_model=256, n_layer=1, n_head=2, d_inner=256).to(cuda).eval()     # 128-wide heads
    eng = cpm.RolloutEngine(wide, 3, 4, greedy=True, mode="persistent")
    assert eng.generate(init7[:, :6])["tokens"].shape == (3, 5, 6) and eng.mode == "chain"        # fell back
    m = _model(cpm, cuda, d_model=128, n_layer=1, n_head=2, d_inner=256)
    bad = _init(4, 1).to(cuda)
    bad[2, 1] = 999
    cpm.ops.IndexGuard.reset(cuda)
    cpm.RolloutEngine(m, 4, 2, greedy=True, mode="persistent").generate(bad)
    with pytest.raises(IndexError):
        cpm.ops.IndexGuard.check(cuda)
    cpm.ops.IndexGuard.check(cuda)                                   # the flag is cleared once reported


# ---------------------------------------------------------------- the default chain: LayerNorm folded into the Linear kernels
@pytest.mark.parametrize("N,cfg", [(5, dict(d_model=128, n_layer=2, n_head=2, d_inner=256)), (256, dict()), (300, dict())])
def test_folded_chain_one_step_vs_layernorm_kernels(cuda, cpm, N, cfg):
    """cpm_gemm_nt_small_ln: the step with LayerNorm folded into the Linear kernels (raw rows times gamma o W, statistics from
    the activation tile, residual LayerNorm rebuilt in the epilogue) against the same chain with its LayerNorm kernels: logits
    within bf16 noise, the recurrent state likewise, nearly all greedy tokens equal."""
    m = _model(cpm, cuda, seed=13, **cfg)
    init = _init(N, 3).to(cuda)
    f = cpm.RolloutEngine(m, N, 4, greedy=True, mode="chain", fold_ln=True)
    c = cpm.RolloutEngine(m, N, 4, greedy=True, mode="chain", fold_ln=False)
    assert f.fold and not c.fold
    for e in (f, c):
        e.reset(init)
    with torch.no_grad():
        m.refresh_packs()
        f._fold_refresh()
        lf, lc = f._logits().float()[:, :m.seg[-1]], c._logits().float()[:, :m.seg[-1]]
    assert (lf - lc).abs().max() < 0.12 and (lf - lc).abs().mean() < 0.012, ((lf - lc).abs().max(), (lf - lc).abs().mean())
    rel = lambda x, y: ((x - y).norm() / y.norm()).item()
    assert rel(f.S, c.S) < 2e-2 and rel(f.Z, c.Z) < 1e-2, (rel(f.S, c.S), rel(f.Z, c.Z))
    a, b = f.generate(init, n_steps=1), c.generate(init, n_steps=1)
    assert _agree(a["tokens"], b["tokens"]) > 0.9                  # random-weight logits are nearly flat: near-ties flip
    assert f.launches_per_step == c.launches_per_step - 2 * len(m.transformer_encoder.layers)      # two LayerNorm launches per layer gone


def test_folded_chain_greedy_equals_teacher_forced_argmax(cuda, cpm):
    """The default rollout (LayerNorm-folded chain, CUDA graph) against the PARALLEL model run teacher-forced over the tokens it
    generated: every greedy choice with a clear top-2 margin is the parallel model's arg-max; recorded log-probs agree."""
    m = _model(cpm, cuda, seed=8)
    mp = cpm.TransformerModel(VOCAB, dropout=0.0).to(cuda).eval()
    mp.load_state_dict(m.state_dict())
    N, T = 64, 128
    init = _init(N, 4).to(cuda)
    eng = cpm.RolloutEngine(m, N, T, greedy=True)
    assert eng.mode == "chain" and eng.fold
    out = eng.generate(init)
    with torch.no_grad():
        lc = mp.logits_concat(mp.hidden(out["tokens"][:, :-1])).float()
    ok = tot = 0
    for a in range(6):
        seg = lc[..., mp.seg[a]:mp.seg[a + 1]]
        top2 = seg.topk(2, -1).values
        sure = (top2[..., 0] - top2[..., 1]) > 0.12
        ok += ((seg.argmax(-1) == out["tokens"][:, 1:, a]) & sure).sum().item()
        tot += sure.sum().item()
    assert tot > 0.4 * N * T * 6 and ok >= tot - 2, (ok, tot)
    lp, _ = cpm.ops.heads_logp(lc.bfloat16(), out["tokens"][:, 1:], mp.seg, False)
    assert (out["logp"] - lp).abs().max() < 0.15 and (out["logp"] - lp).abs().mean() < 0.02
