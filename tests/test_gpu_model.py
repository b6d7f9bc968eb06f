"""GPU parity tests, module level: the reference model surface (train_step / forward_hidden /
forward_output / recurrent inference / RL read-outs) on cpmusic kernels vs the oracle restatement
and the committed golden vectors.  fp32 compute mode gives the tight comparison; bf16 mode is the
production dtype with its tolerance stated."""
import numpy as np
import pytest
import torch

from oracle import model_oracle as mo, rl_oracle as rl, sampling_oracle as so

pytestmark = pytest.mark.gpu
VOCAB = [56, 135, 18, 87, 18, 25]
SMALL = dict(d_model=128, n_layer=2, n_head=2, d_inner=256, dropout=0.0)


def _cmp(a, b, atol, rtol=0.0, what=""):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).double().cpu()
    err = (a - b).abs()
    assert bool((err <= atol + rtol * b.abs()).all()), f"{what}: max err {err.max().item():.3e}, ref scale {b.abs().max().item():.3e}"


def _load_small(cpm, g, cuda, cls=None, dtype=torch.float32, is_training=True):
    cls = cls or cpm.LinearTransformer
    m = cls(VOCAB, is_training, compute_dtype=dtype, **SMALL)
    sd = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}
    res = m.load_state_dict(sd, strict=False)
    assert res.missing_keys == ["pos_emb.pe"] and not res.unexpected_keys
    return m.to(cuda)


def test_train_step_fp32_golden(cuda, cpm, golden):
    g = golden("model_small")
    m = _load_small(cpm, g, cuda).train()            # dropout=0.0 in this config
    x, y, mask = (torch.from_numpy(g[k]).to(cuda) for k in ("x", "y", "mask"))
    h = m.forward_hidden(x)
    _cmp(h, g["h"], 2e-4, 1e-4, "hidden")
    logits = torch.cat(m.forward_output(h, y), -1)
    _cmp(logits, g["logits"], 3e-4, 1e-4, "logits")
    losses = m.train_step(x, y, mask)
    assert isinstance(losses, tuple) and len(losses) == 6
    _cmp(torch.stack(losses), g["losses"], 1e-4, 1e-5, "losses")
    (sum(losses) / 6).backward()
    params = dict(m.named_parameters())
    for k in g.files:
        if k.startswith("grad::"):
            ref = g[k]
            _cmp(params[k[6:]].grad, ref, 2e-5 + 2e-3 * np.abs(ref).max(), 2e-3, k)
    assert m.project_concat_type.weight.grad is None      # allocated, never used


def test_train_step_bf16(cuda, cpm, golden):
    """Production dtype: bf16 activations/GEMMs, fp32 masters.  Loss within 2e-2 of the fp64 oracle,
    hidden states within 6e-2 absolute (post-LayerNorm O(1) values after 2 layers of bf16 rounding)."""
    g = golden("model_small")
    m = _load_small(cpm, g, cuda, dtype=torch.bfloat16).train()
    x, y, mask = (torch.from_numpy(g[k]).to(cuda) for k in ("x", "y", "mask"))
    _cmp(m.forward_hidden(x), g["h"], 6e-2, 2e-2, "hidden bf16")
    losses = torch.stack(m.train_step(x, y, mask))
    _cmp(losses, g["losses"], 2e-2, 1e-2, "losses bf16")
    (losses.sum() / 6).backward()
    ref = g["grad::in_linear.weight"]
    got = m.in_linear.weight.grad.double().cpu().numpy()
    cos = (got * ref).sum() / (np.linalg.norm(got) * np.linalg.norm(ref))
    assert cos > 0.99, cos


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-3), (torch.bfloat16, 6e-2)])
def test_pretraining_loss_curve_matches_oracle(cuda, cpm, golden, dtype, tol):
    """Reference-matching loss curves (north_star): 10 optimizer steps of the pretraining loop (agent_pretrain.py:557-565:
    train_step, mean of the six losses, clip_grad_norm_ 3, Adam) on the same batches from the same initial weights —
    cpmusic on the GPU with torch's FUSED Adam (which updates parameters without bumping their version counters, so this
    also checks that every step's compute copies see the previous update) against the oracle on the CPU with plain Adam.
    fp32 compute follows the oracle to 2e-3 over the whole curve; bf16 compute to 6e-2, and both must descend."""
    g = golden("model_small")
    m = _load_small(cpm, g, cuda, dtype=dtype).train()
    o = mo.OracleCPModel(VOCAB, **SMALL).train()
    o.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}, strict=False)
    om = torch.optim.Adam(m.parameters(), lr=1e-3, fused=True)
    oo = torch.optim.Adam(o.parameters(), lr=1e-3)
    gen = torch.Generator().manual_seed(31)
    batches = []
    for _ in range(2):                                   # two batches, revisited: something to fit, so the curve descends
        x = torch.stack([torch.randint(0, n, (3, 96), generator=gen) for n in VOCAB], -1)
        batches.append((x, x.roll(-1, 1), (torch.rand(3, 96, generator=gen) > 0.2).float()))
    curve_m, curve_o = [], []
    for step in range(10):
        x, y, mask = batches[step % 2]
        lo = sum(o.train_step(x, y, mask)) / 6
        oo.zero_grad()
        lo.backward()
        torch.nn.utils.clip_grad_norm_(o.parameters(), 3.0)
        oo.step()
        lm = sum(m.train_step(x.to(cuda), y.to(cuda), mask.to(cuda))) / 6
        om.zero_grad()
        lm.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 3.0)
        om.step()
        curve_m.append(lm.item())
        curve_o.append(lo.item())
    cm, co = torch.tensor(curve_m), torch.tensor(curve_o)
    _cmp(cm, co, tol, 0.0, f"loss curve {curve_m} vs {curve_o}")
    assert co[-2:].mean() < co[:2].mean() - 0.1 and cm[-2:].mean() < cm[:2].mean() - 0.1


def test_module_vs_live_oracle_random_weights(cuda, cpm):
    """Fresh random weights (module default init) shared through state_dict — the drop-in contract:
    an oracle/reference checkpoint loads strictly and produces the same numbers."""
    torch.manual_seed(7)
    o = mo.OracleCPModel(VOCAB, d_model=192, n_layer=3, n_head=3, d_inner=384, dropout=0.0).eval()
    m = cpm.TransformerModel(VOCAB, d_model=192, n_layer=3, n_head=3, d_inner=384, dropout=0.0, compute_dtype=torch.float32)
    m.load_state_dict(o.state_dict())
    m = m.to(cuda)
    gen = torch.Generator().manual_seed(8)
    x = torch.stack([torch.randint(0, n, (2, 129), generator=gen) for n in VOCAB], -1)
    mask = torch.ones(2, 129)
    mask[1, 100:] = 0
    with torch.no_grad():
        ref = torch.stack(o.train_step(x, x.roll(-1, 1), mask))
        got = torch.stack(m.train_step(x.to(cuda), x.roll(-1, 1).to(cuda), mask.to(cuda)))
    _cmp(got, ref, 2e-4, 1e-4, "losses")
    # forward(x, target) == forward_output(forward_hidden(x), target)  (dqn_policy/model.py:252-255)
    with torch.no_grad():
        a = m(x.to(cuda), None)
        b = o(x, None)
    for ya, yb in zip(a, b):
        _cmp(ya, yb, 5e-4, 1e-4, "forward logits")
    # compute_loss with the reference's (N, n_i, L) layout
    with torch.no_grad():
        cl = m.compute_loss(a[3].permute(0, 2, 1), x[..., 3].roll(-1, 1).to(cuda), mask.to(cuda))
    _cmp(cl, ref[3], 2e-4, 1e-4, "compute_loss")


def test_seven_attribute_layout_vs_live_oracle(cuda, cpm):
    """The data files' seven-attribute compound word (`type` kept at column 3): 7 embeddings (Σ = 1248), 7 heads.
    Losses and gradients against the oracle, greedy recurrent decoding bit-exact against the oracle's argmax."""
    vocab7 = [56, 135, 18, 4, 87, 18, 25]
    torch.manual_seed(21)
    cfg = dict(d_model=128, n_layer=2, n_head=2, d_inner=256, dropout=0.0)
    o = mo.OracleCPModel(vocab7, **cfg).eval()
    m = cpm.TransformerModel(vocab7, compute_dtype=torch.float32, **cfg)
    assert m.attrs[3] == "type" and len(m.state_dict()) == len(o.state_dict())
    m.load_state_dict(o.state_dict())
    m = m.to(cuda)
    gen = torch.Generator().manual_seed(22)
    x = torch.stack([torch.randint(0, n, (3, 70), generator=gen) for n in vocab7], -1)
    mask = torch.ones(3, 70)
    mask[2, 40:] = 0
    ref = torch.stack(o.train_step(x, x.roll(-1, 1), mask))
    got = torch.stack(m.train_step(x.to(cuda), x.roll(-1, 1).to(cuda), mask.to(cuda)))
    assert got.shape == (7,)
    _cmp(got, ref, 2e-4, 1e-4, "losses")
    (ref.sum() / 7).backward()
    (got.sum() / 7).backward()
    po = dict(o.named_parameters())
    for name, prm in m.named_parameters():
        if po[name].grad is not None:
            r = po[name].grad
            _cmp(prm.grad, r, 2e-5 + 2e-3 * r.abs().max().item(), 2e-3, f"grad {name}")
    # greedy decoding, recurrent, true positions: tokens bit-exact against the oracle's parallel argmax on the decoded prefix
    mr = cpm.TransformerModel(vocab7, is_training=False, compute_dtype=torch.float32, **cfg)
    mr.load_state_dict(o.state_dict())
    mr = mr.to(cuda).eval()
    toks = cpm.RolloutEngine(mr, 2, 12, greedy=True).generate(x[:2, 0].to(cuda))["tokens"].cpu()      # (2, 13, 7)
    with torch.no_grad():
        logits = o.forward_output(o.forward_hidden(toks[:, :-1]))
    for a, lg in enumerate(logits):
        top2 = lg.topk(2, -1).values
        sure = (top2[..., 0] - top2[..., 1]) > 1e-3                      # skip numerical near-ties
        assert torch.equal(lg.argmax(-1)[sure], toks[:, 1:, a][sure]), f"attribute {a}"


def test_module_with_128_wide_heads_vs_live_oracle(cuda, cpm):
    """BASELINE cfg5's other reading (d_model = heads x 128): the same module surface with query/value dimensions 128 —
    teacher-forced losses and gradients against the oracle, and the recurrent path (generic step kernel, state
    (N,H,128,128)) against the parallel one, all in fp32 compute."""
    torch.manual_seed(17)
    cfg = dict(d_model=256, n_layer=2, n_head=2, d_inner=256, dropout=0.0)
    o = mo.OracleCPModel(VOCAB, **cfg).eval()
    m = cpm.TransformerModel(VOCAB, compute_dtype=torch.float32, **cfg)
    m.load_state_dict(o.state_dict())
    m = m.to(cuda)
    gen = torch.Generator().manual_seed(18)
    x = torch.stack([torch.randint(0, n, (2, 140), generator=gen) for n in VOCAB], -1)
    mask = torch.ones(2, 140)
    mask[0, 90:] = 0
    ref = torch.stack(o.train_step(x, x.roll(-1, 1), mask))
    got = torch.stack(m.train_step(x.to(cuda), x.roll(-1, 1).to(cuda), mask.to(cuda)))
    _cmp(got, ref, 2e-4, 1e-4, "losses")
    (ref.sum() / 6).backward()
    (got.sum() / 6).backward()
    po = dict(o.named_parameters())
    for name, prm in m.named_parameters():
        if po[name].grad is not None:
            r = po[name].grad
            _cmp(prm.grad, r, 2e-5 + 2e-3 * r.abs().max().item(), 2e-3, f"grad {name}")
    mr = cpm.TransformerModel(VOCAB, is_training=False, compute_dtype=torch.float32, **cfg)
    mr.load_state_dict(o.state_dict())
    mr = mr.to(cuda).eval()
    with torch.no_grad():
        par = m.eval().forward_hidden(x[:1, :20].to(cuda))
        mem, hs = None, []
        for t in range(20):
            h, mem = mr.forward_hidden(x[:1, t:t + 1].to(cuda), mem, is_training=False, pos_offset=t)
            hs.append(h)
    assert mem[0][0].shape == (1, 2, 128, 128)
    _cmp(torch.stack(hs, 1), par, 3e-4, 1e-4, "recurrent vs parallel at 128-wide heads")


def test_recurrent_forward_hidden(cuda, cpm, golden):
    """Recurrent mode: default reproduces the reference's position-0 quirk; pos_offset
    gives the true position and then matches the parallel path."""
    g = golden("model_small")
    m = _load_small(cpm, g, cuda, is_training=False).eval()
    x = torch.from_numpy(g["x"]).to(cuda)
    for key, true_pos in (("h_rec_pos0", False), ("h_rec_true", True)):
        mem, hs = None, []
        for t in range(12):
            inp = x[:1, t:t + 1]                                   # (1,1,6) like testing-no-type-cp.py:151
            if true_pos:
                h, mem = m.forward_hidden(inp, mem, is_training=False, pos_offset=t)
            else:
                h, mem = m.forward_hidden(inp, mem, is_training=False)
            assert h.shape == (1, 128) and len(mem) == 2 and mem[0][0].shape == (1, 2, 64, 64)
            hs.append(h)
        _cmp(torch.cat(hs, 0), g[key], 2e-4, 1e-4, key)
    with pytest.raises(RuntimeError):
        m.forward_hidden(x)                                        # parallel call on a recurrent model
    out = m.forward_output_sampling(hs[-1], seed=5)
    assert out.shape == (6,) and all(0 <= int(out[a]) < VOCAB[a] for a in range(6))


def test_rollout_engine_graph_equals_eager_equals_teacher_forced(cuda, cpm, golden):
    """Greedy generation: (i) CUDA-graph replay == eager stepping, bit-exact tokens; (ii) feeding the
    generated sequence to the parallel (teacher-forced) model reproduces every greedy choice."""
    g = golden("model_small")
    mr = _load_small(cpm, g, cuda, is_training=False).eval()
    mp = _load_small(cpm, g, cuda, is_training=True).eval()
    N, T = 5, 40
    init = torch.stack([torch.randint(0, n, (N,), generator=torch.Generator().manual_seed(9)) for n in VOCAB], -1).to(cuda)
    eng_g = cpm.RolloutEngine(mr, N, T, greedy=True, true_positions=True, use_graph=True)
    eng_e = cpm.RolloutEngine(mr, N, T, greedy=True, true_positions=True, use_graph=False)
    a, b = eng_g.generate(init), eng_e.generate(init)
    assert torch.equal(a["tokens"], b["tokens"]) and a["tokens"].shape == (N, T + 1, 6)
    assert torch.equal(eng_g.generate(init)["tokens"], a["tokens"])        # replay after reset is deterministic
    with torch.no_grad():
        lc = mp.logits_concat(mp.hidden(a["tokens"][:, :-1]))
        tf_tok, tf_lp, _ = cpm.ops.heads_sample(lc.reshape(N * T, -1), mp.seg, greedy=True, want_logp=True)
    match = (tf_tok.view(N, T, 6) == a["tokens"][:, 1:]).float().mean().item()
    assert match > 0.995, match           # fp32: only near-ties between two logits may differ
    _cmp(a["logp"], tf_lp.view(N, T, 6), 2e-3, 1e-3, "recorded log-probs")
    # sampled rollouts: identical tokens for the same seed regardless of how sequences are sharded
    eng_s = cpm.RolloutEngine(mr, N, T, greedy=False, seed=123, seq_base=0)
    full = eng_s.generate(init)["tokens"]
    part = cpm.RolloutEngine(mr, 2, T, greedy=False, seed=123, seq_base=3).generate(init[3:])["tokens"]
    assert torch.equal(full[3:], part)
    other = cpm.RolloutEngine(mr, N, T, greedy=False, seed=124).generate(init)["tokens"]
    assert not torch.equal(full, other)


def test_ppo_and_dqn_readouts(cuda, cpm, golden):
    g = golden("model_small")
    m = _load_small(cpm, g, cuda, cls=cpm.LinearTransformer).eval()
    o = mo.OracleCPModel(VOCAB, **SMALL).eval()
    o.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}, strict=False)
    gen = torch.Generator().manual_seed(10)
    x = torch.stack([torch.randint(0, n, (4, 50), generator=gen) for n in VOCAB], -1)
    with torch.no_grad():
        ol = o.forward_output(o.forward_hidden(x[:1]))
        act_ref, lp_ref = rl.ppo_choose_action_compat(ol)
        act, lp = cpm.rl.ppo_choose_action(m, x[:1].to(cuda))
        assert torch.equal(act.cpu(), act_ref)
        _cmp(lp, lp_ref, 2e-3, 1e-3, "choose_action logp")
        assert torch.equal(cpm.rl.dqn_choose_action(m, x[:1].to(cuda)).cpu(), rl.dqn_choose_action_compat(ol))
        ob = o.forward_output(o.forward_hidden(x))
        a_ref, l_ref = rl.ppo_select_update_compat(ob)
        a_got, l_got = cpm.rl.ppo_select_update(m, x.to(cuda))
        assert torch.equal(a_got.cpu(), a_ref)
        _cmp(l_got, l_ref, 2e-3, 1e-3, "select_update logp")
        a_all, l_all = cpm.rl.ppo_select_update(m, x.to(cuda), compat=False)
        ra, rlp = rl.action_logp_all(ob)
        assert torch.equal(a_all.cpu(), ra)
        _cmp(l_all, rlp, 2e-3, 1e-3, "all logp")


def test_actor_critic_value_paths(cuda, cpm):
    torch.manual_seed(11)
    vocab = [49, 19, 19, 89, 67, 25]                         # PPO dictionary sizes (prepare_data.py:243-295)
    oc = mo.OracleCritic(vocab, **SMALL).eval()
    c = cpm.Critic_Transformer(vocab, compute_dtype=torch.float32, **SMALL)
    c.load_state_dict(oc.state_dict())
    c = c.to(cuda).eval()
    x = torch.stack([torch.randint(0, n, (3, 50)) for n in vocab], -1)
    with torch.no_grad():
        _cmp(c.value_produce(x.to(cuda)), oc.value_produce(x), 2e-4, 1e-4, "value_produce")
    oa = mo.OracleCPModel(vocab, variant="actor", **SMALL).eval()
    a = cpm.Actor_Transformer(vocab, compute_dtype=torch.float32, **SMALL)
    a.load_state_dict(oa.state_dict())
    a = a.to(cuda).eval()
    with torch.no_grad():
        h = a.forward_hidden(x[:1].to(cuda))
        _cmp(a.value_funtion(h.squeeze(0)), oa.value_funtion(oa.forward_hidden(x[:1]).squeeze(0)), 2e-4, 1e-4, "value_funtion")
        assert len(a.forward_output(h)) == 6                  # PPO arity: forward_output(h)


def test_fast_transformers_shim_runs_reference_style_model(cuda, cpm):
    """An ft-style caller (fp32 in/out, TriangularCausalMask, `memory=` keyword) on the shim."""
    from oracle import ft_oracle
    torch.manual_seed(12)
    kw = dict(n_layers=2, n_heads=2, query_dimensions=64, value_dimensions=64, feed_forward_dimensions=256,
              activation="gelu", dropout=0.0, attention_type="causal-linear")
    ref = ft_oracle.TransformerEncoderBuilder.from_kwargs(**kw).get().eval()
    enc = cpm.TransformerEncoderBuilder.from_kwargs(compute_dtype=torch.float32, **kw).get()
    enc.load_state_dict(ref.state_dict())
    enc = enc.to(cuda).eval()
    x = torch.randn(2, 33, 128)
    with torch.no_grad():
        _cmp(enc(x.to(cuda), cpm.TriangularCausalMask(33, device=cuda)), ref(x, ft_oracle.TriangularCausalMask(33)), 2e-4, 1e-4, "encoder")
    rref = ft_oracle.RecurrentEncoderBuilder.from_kwargs(**kw).get().eval()
    rref.load_state_dict(ref.state_dict())
    renc = cpm.RecurrentEncoderBuilder.from_kwargs(compute_dtype=torch.float32, **kw).get()
    renc.load_state_dict(ref.state_dict())
    renc = renc.to(cuda).eval()
    mem_a, mem_b = None, None
    with torch.no_grad():
        for t in range(6):
            ya, mem_a = renc(x[:, t].to(cuda), memory=mem_a)
            yb, mem_b = rref(x[:, t], memory=mem_b)
            _cmp(ya, yb, 2e-4, 1e-4, f"recurrent step {t}")
    _cmp(mem_a[1][0], mem_b[1][0], 1e-4, 1e-4, "Si")


def test_full_size_model_bf16_smoke_properties(cuda, cpm):
    """Reference-size model (12 layers, d 512, 8 heads; 38,982,227 parameters), BASELINE cfg1 batch
    (4 x 512): bf16 losses vs the fp32 kernel path on identical weights, finite gradients, loss
    close to ln(vocab) at init, and causality of the hidden states."""
    torch.manual_seed(13)
    m = cpm.LinearTransformer(VOCAB, dropout=0.0).to(cuda).train()
    assert sum(p.numel() for p in m.parameters()) == 38982227
    gen = torch.Generator().manual_seed(14)
    x = torch.stack([torch.randint(0, n, (4, 512), generator=gen) for n in VOCAB], -1).to(cuda)
    mask = torch.ones(4, 512, device=cuda)
    l16 = torch.stack(m.train_step(x, x.roll(-1, 1), mask))
    (l16.sum() / 6).backward()
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)
    m.zero_grad()
    m.set_compute_dtype(torch.float32)
    with torch.no_grad():
        l32 = torch.stack(m.train_step(x, x.roll(-1, 1), mask))
        h_a = m.hidden(x)
        x2 = x.clone()
        x2[:, 300:] = 0
        h_b = m.hidden(x2)
    _cmp(l16, l32, 3e-2, 1e-2, "bf16 vs fp32 losses")
    assert torch.equal(h_a[:, :300], h_b[:, :300])                        # causal: bit-identical prefix
    for a, n in enumerate(VOCAB):
        assert abs(l32[a].item() - np.log(n)) < 0.6


def test_rollout_256_songs_full_size(cuda, cpm):
    """BASELINE cfg3 rollout shape at full model size (12 layers, d 512; 256 songs): CUDA-graph replay equals eager
    stepping bit for bit, the recorded log-probs match a teacher-forced parallel forward over the generated tokens,
    and sharding the songs over two engines (as two GPUs would) reproduces the same tokens."""
    torch.manual_seed(21)
    m = cpm.LinearTransformer(VOCAB, dropout=0.0).to(cuda).eval()
    N, T = 256, 12
    init = torch.stack([torch.randint(0, n, (N,), generator=torch.Generator().manual_seed(22)) for n in VOCAB], -1).to(cuda)
    a = cpm.RolloutEngine(m, N, T, greedy=False, seed=5, use_graph=True).generate(init)
    b = cpm.RolloutEngine(m, N, T, greedy=False, seed=5, use_graph=False).generate(init)
    assert torch.equal(a["tokens"], b["tokens"]) and torch.equal(a["logp"], b["logp"])
    halves = [cpm.RolloutEngine(m, 128, T, greedy=False, seed=5, seq_base=128 * r).generate(init[128 * r:128 * (r + 1)])["tokens"] for r in (0, 1)]
    assert torch.equal(torch.cat(halves, 0), a["tokens"])
    with torch.no_grad():
        lc = m.logits_concat(m.hidden(a["tokens"][:, :-1]))
        lp, _ = cpm.ops.heads_logp(lc, a["tokens"][:, 1:], m.seg, False)
    _cmp(a["logp"], lp, 6e-2, 2e-2, "rollout log-probs vs teacher-forced forward (bf16)")


def test_rollout_graph_sees_optimizer_updates(cuda, cpm, golden):
    """The captured rollout graph reads packed bf16 weights by address: after an optimizer step the
    packs are refreshed in place, so graph replay must equal eager stepping on the NEW weights."""
    g = golden("model_small")
    m = _load_small(cpm, g, cuda, dtype=torch.bfloat16)
    N, T = 3, 16
    init = torch.stack([torch.randint(0, n, (N,), generator=torch.Generator().manual_seed(15)) for n in VOCAB], -1).to(cuda)
    eng = cpm.RolloutEngine(m, N, T, greedy=True, use_graph=True)
    before = eng.generate(init)["tokens"].clone()
    opt = torch.optim.SGD(m.parameters(), lr=0.5)
    x = torch.from_numpy(g["x"]).to(cuda)
    m.train()
    loss = sum(m.train_step(x, x.roll(-1, 1), torch.ones(2, 70, device=cuda)))
    loss.backward()
    opt.step()
    after_graph = eng.generate(init)["tokens"].clone()
    after_eager = cpm.RolloutEngine(m, N, T, greedy=True, use_graph=False).generate(init)["tokens"]
    assert torch.equal(after_graph, after_eager)
    assert not torch.equal(before, after_graph)          # lr 0.5 really changed the policy


def test_graphed_train_step_equals_eager_and_keeps_dropout_fresh(cuda, cpm, golden):
    """The whole pretraining step as one CUDA graph: (i) with dropout off, three replays reproduce three eager steps
    (losses within fp32 round-off of each other, same parameters afterwards to 1e-5); (ii) with dropout on, replays of the
    same batch give DIFFERENT losses (the device-side RNG base advances inside the graph) and training still descends."""
    g = golden("model_small")
    cfgs = dict(SMALL)
    x = torch.from_numpy(g["x"]).to(cuda)[:2, :128]
    if x.shape[1] < 128:
        x = x.repeat(1, 128 // x.shape[1] + 1, 1)[:, :128]
    y, mask = x.roll(-1, 1), torch.ones(x.shape[:2], device=cuda)
    ma = _load_small(cpm, g, cuda, dtype=torch.float32).train()
    mb = _load_small(cpm, g, cuda, dtype=torch.float32).train()
    oa = torch.optim.Adam(ma.parameters(), lr=1e-3, fused=True, capturable=True)
    ob = torch.optim.Adam(mb.parameters(), lr=1e-3, fused=True, capturable=True)

    def eager(xx, yy, mm):
        le = torch.stack(mb.train_step(xx, yy, mm))
        ob.zero_grad(set_to_none=False)
        (le.sum() / 6).backward()
        torch.nn.utils.clip_grad_norm_(mb.parameters(), 3.0, foreach=True)
        ob.step()
        return le.detach()

    step = cpm.GraphedTrainStep(ma, oa, batch_size=2, seq_len=128, max_grad_norm=3.0, warmup=2)
    for _ in range(2):                                         # the constructor's two warm-up steps ran on its zero-filled buffers
        eager(torch.zeros_like(x), torch.zeros_like(y), mask)
    for _ in range(3):
        lg = step(x, y, mask).clone()
        _cmp(lg, eager(x, y, mask), 2e-4, 1e-4, "graphed vs eager losses")
    for (k, a), (_, b) in zip(ma.state_dict().items(), mb.state_dict().items()):
        _cmp(a, b, 2e-4, 1e-3, f"parameters after 3 steps: {k}")
    # dropout on: masks must differ between replays of the same batch
    md = cpm.LinearTransformer(VOCAB, compute_dtype=torch.bfloat16, **dict(cfgs, dropout=0.3)).to(cuda).train()
    od = torch.optim.Adam(md.parameters(), lr=0.0, fused=True, capturable=True)      # lr 0: only the masks change between replays
    sd = cpm.GraphedTrainStep(md, od, batch_size=2, seq_len=128)
    l1, l2, l3 = (sd(x, y, mask).clone() for _ in range(3))
    assert not torch.equal(l1, l2) and not torch.equal(l2, l3)
    assert sd.rng_counters_per_step > 0


def test_greedy_rollout_tokens_equal_oracle_recurrent_decode(cuda, cpm, golden):
    """north_star: sampled token indices bit-exact under greedy decoding.  The RolloutEngine's greedy tokens (fp32 compute,
    CUDA-graph step, true positions) against the ORACLE's own recurrent greedy decode (ft RecurrentLinearAttention restated,
    oracle/ft_oracle.py + model_oracle.py; argmax per attribute, the t -> 0 limit of testing-no-type-cp.py's sampling loop) on
    the CPU, token by token.  The oracle is stepped on the tokens the kernels produced, so every one of the N x T x 6
    sub-tokens is checked against the oracle's argmax for the SAME history.  A difference is tolerated only where the oracle's
    own top-2 logit gap for that attribute is below 2e-3 (fp32 accumulation order can flip such a near-tie; a random-init model
    has many); everywhere else the indices must be bit-exact, and at least 97 % of all sub-tokens must agree outright."""
    g = golden("model_small")
    mr = _load_small(cpm, g, cuda, is_training=False).eval()
    o = mo.OracleCPModel(VOCAB, is_training=False, **SMALL).eval()
    o.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}, strict=False)
    N, T = 6, 48
    init = torch.stack([torch.randint(0, n, (N,), generator=torch.Generator().manual_seed(3)) for n in VOCAB], -1)
    got = cpm.RolloutEngine(mr, N, T, greedy=True, true_positions=True, use_graph=True).generate(init.to(cuda))["tokens"].cpu()
    assert torch.equal(got[:, 0], init)
    agree = total = 0
    for n in range(N):
        mem = None
        with torch.no_grad():
            for t in range(T):
                z = o.pos_emb(o.embed(got[n:n + 1, t][:, None, :]), t).squeeze(1)
                h, mem = o.transformer_encoder(z, memory=mem)
                for a, lg in enumerate(o.forward_output(h)):
                    top2 = lg[0].topk(2)
                    total += 1
                    if int(got[n, t + 1, a]) == int(top2.indices[0]):
                        agree += 1
                    else:
                        gap = float(top2.values[0] - top2.values[1])
                        assert gap < 2e-3 and int(got[n, t + 1, a]) == int(top2.indices[1]), \
                            f"sequence {n} step {t} attribute {a}: kernel {int(got[n, t + 1, a])} vs oracle {int(top2.indices[0])} (top-2 gap {gap:.2e})"
    assert agree >= 0.97 * total, f"{agree} of {total} greedy sub-tokens bit-exact"


def test_train_step_cfg2_shape_bf16_vs_oracle(cuda, cpm):
    """BASELINE cfg2 (pretraining, bf16, 32 x 512, full-size 12-layer model): the six losses and the gradients of the production
    path - own tcgen05 GEMMs with fused GELU epilogues, chunk-parallel attention with the streaming state kernels, fused
    LayerNorm / CE kernels - against the fp32 ORACLE (model_oracle.py with the C causal-product clone) on the same weights and
    batch.  Tolerances: losses 3e-2 absolute (bf16 activations through 12 layers); gradient direction cosine >= 0.98 for the first
    and last layers' weights and the embedding tables (bf16 backward through 12 layers)."""
    from oracle.causal_product_c import causal_dot_product_c
    torch.manual_seed(5)
    m = cpm.LinearTransformer(VOCAB, dropout=0.0).to(cuda).train()
    o = mo.OracleCPModel(VOCAB, is_training=True, dropout=0.0).train()
    o.load_state_dict({k: v.detach().cpu() for k, v in m.state_dict().items()}, strict=True)
    o.transformer_encoder.product = causal_dot_product_c
    gen = torch.Generator().manual_seed(6)
    N, L = 32, 512
    x = torch.stack([torch.randint(0, n, (N, L), generator=gen) for n in VOCAB], -1)
    y = x.roll(-1, 1)
    lens = torch.randint(L // 2, L + 1, (N,), generator=gen)
    mask = (torch.arange(L)[None, :] < lens[:, None]).float()
    losses = torch.stack(m.train_step(x.to(cuda), y.to(cuda), mask.to(cuda)))
    (losses.sum() / 6).backward()
    ref = torch.stack(o.train_step(x, y, mask))
    (ref.sum() / 6).backward()
    _cmp(losses, ref, 3e-2, 1e-2, "cfg2 losses bf16 vs fp32 oracle")
    op = dict(o.named_parameters())
    for name in ("in_linear.weight", "transformer_encoder.layers.0.attention.query_projection.weight", "transformer_encoder.layers.0.linear1.weight",
                 "transformer_encoder.layers.11.linear2.weight", "transformer_encoder.layers.11.attention.out_projection.bias",
                 "transformer_encoder.layers.5.norm1.weight", "proj_pitch.weight", "word_emb_pitch.lut.weight"):
        a, b = dict(m.named_parameters())[name].grad.double().cpu().flatten(), op[name].grad.double().flatten()
        cos = float((a * b).sum() / (a.norm() * b.norm()))
        assert cos > 0.98, f"{name}: gradient cosine {cos:.4f}"
        assert 0.9 < float(a.norm() / b.norm()) < 1.1, f"{name}: gradient norm ratio {float(a.norm() / b.norm()):.3f}"


@pytest.mark.parametrize("dtype,atol", [(torch.float32, 3e-4), (torch.bfloat16, 6e-2)])
def test_encoder_length_mask_vs_oracle(cuda, cpm, dtype, atol):
    """ft's key-padding path: ``encoder(x, TriangularCausalMask, LengthMask(lengths))`` multiplies the feature-mapped keys by
    the length matrix in every layer.  Outputs at every position (valid and padded) and the input gradient
    against the oracle; the padded keys really are excluded (the output changes when the mask is dropped)."""
    from oracle import ft_oracle
    torch.manual_seed(31)
    kw = dict(n_layers=2, n_heads=2, query_dimensions=64, value_dimensions=64, feed_forward_dimensions=256,
              activation="gelu", dropout=0.0, attention_type="causal-linear")
    ref = ft_oracle.TransformerEncoderBuilder.from_kwargs(**kw).get().eval()
    enc = cpm.TransformerEncoderBuilder.from_kwargs(compute_dtype=dtype, **kw).get()
    enc.load_state_dict(ref.state_dict())
    enc = enc.to(cuda).eval()
    N, L = 3, 256
    lengths = torch.tensor([256, 130, 17])
    x = torch.randn(N, L, 128)
    xr = x.clone().requires_grad_()
    xg = x.clone().to(cuda).requires_grad_()
    yr = ref(xr, ft_oracle.TriangularCausalMask(L), ft_oracle.LengthMask(lengths, L))
    yg = enc(xg, cpm.TriangularCausalMask(L, device=cuda), cpm.LengthMask(lengths.to(cuda), L))
    _cmp(yg, yr, atol, atol, "encoder output under a length mask")
    w = torch.randn(N, L, 128)
    (yr * w).sum().backward()
    (yg * w.to(cuda)).sum().backward()
    _cmp(xg.grad, xr.grad, 20 * atol, 5e-2, "input gradient under a length mask")
    with torch.no_grad():
        y_nomask = enc(x.to(cuda), cpm.TriangularCausalMask(L, device=cuda))
    assert (y_nomask[1, 200] - yg[1, 200]).abs().max() > 1e-2           # position 200 of song 1 saw padded keys without the mask
    assert (y_nomask[0] - yg[0]).abs().max() < atol                     # song 0 has no padding
    with pytest.raises(ValueError):
        enc(x.to(cuda), cpm.TriangularCausalMask(L, device=cuda), cpm.LengthMask(lengths.to(cuda), L + 1))


def test_sm_partition_stream_runs_the_same_kernels_on_fewer_sms(cuda, cpm):
    """graphs.sm_partition_stream: a torch stream confined to a subset of the SMs (CUDA green context).  The library's kernels -
    a persistent 2-CTA GEMM sized for the whole device, the chunk-parallel attention pair, a row kernel - must give bit-identical
    results there (their grids queue in waves) and the GEMM must take visibly longer on a third of the SMs."""
    try:
        side, got = cpm.graphs.sm_partition_stream(48, cuda)
    except RuntimeError as e:
        pytest.skip(f"green contexts unavailable: {e}")
    assert 8 <= got <= 64
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(16384, 512, generator=gen).to(cuda).bfloat16()
    w = (torch.randn(1536, 512, generator=gen) / 16).to(cuda).bfloat16()
    b = torch.randn(1536, generator=gen).to(cuda)

    def work():
        qkv = cpm.ops.gemm_nt(x, w, b).view(16, 1024, 1536)
        att = cpm.ops.causal_linear_attention_fused(qkv, 8)
        return qkv, att, cpm.ops.colsum(att.view(-1, 512))

    def timed(stream):
        with torch.cuda.stream(stream):
            res = work()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                cpm.ops.gemm_nt(x, w, b)
            e1.record()
        torch.cuda.synchronize()
        return res, e0.elapsed_time(e1)

    ref, t_full = timed(torch.cuda.current_stream())
    side.wait_stream(torch.cuda.current_stream())
    out, t_part = timed(side)
    for name, a, r in zip(("gemm", "attention", "colsum"), out, ref):
        assert torch.equal(a, r), f"{name} differs on the SM partition"
    assert t_part > 1.5 * t_full, f"the partition stream is not confined: {t_part:.3f} ms vs {t_full:.3f} ms on the whole device"


def test_pack_cache_one_launch_refresh_equals_per_layer_copies(cuda, cpm, golden):
    """encoder.PackCache refreshes every bf16 packing of a model with ONE cpm_pack_weights launch after an optimizer step (row-major
    copy, transposed copy, bf16 and fp32 bias per master; the zero rows that pad the 339 head rows to 344 stay zero).  Bit-identical
    to the per-layer torch copies, for the encoder layers, the input projection and the concatenated heads; and a training step
    after the refresh sees the new weights."""
    torch.manual_seed(3)
    VOCAB = [56, 135, 18, 87, 18, 25]
    m = cpm.TransformerModel(VOCAB, d_model=256, n_layer=2, n_head=4, d_inner=512, dropout=0.0).to(cuda).train()
    gen = torch.Generator().manual_seed(1)
    x = torch.stack([torch.randint(0, n, (2, 128), generator=gen) for n in VOCAB], -1).to(cuda)
    mask = torch.ones(2, 128, device=cuda)
    opt = torch.optim.Adam(m.parameters(), lr=1e-2, fused=True)
    caches = [c for c in (getattr(mod, "_cache", None) for mod in m.modules()) if isinstance(c, cpm.encoder.PackCache)]
    assert caches
    l0 = sum(m.train_step(x, x.roll(-1, 1), mask))
    l0.backward()
    opt.step()                                             # every packing is stale now
    lib_calls = cpm._lib.COUNTS["cpm_pack_weights"]
    l1 = sum(m.train_step(x, x.roll(-1, 1), mask))         # first get() of each cache refreshes all of its packings at once
    assert cpm._lib.COUNTS["cpm_pack_weights"] - lib_calls == len([c for c in caches if c._table is not None]) >= 1
    assert float(l1.detach()) < float(l0.detach())
    n_checked = 0
    for c in caches:
        for key, (stamp, (wc, bc, rows, masters, wt, b32)) in c._store.items():
            n = len(rows)
            ref_w = torch.zeros_like(wc)
            ref_b = torch.zeros_like(b32)
            r0 = 0
            for w, b in zip(masters[:n], masters[n:]):
                ref_w[r0:r0 + w.shape[0]] = w.detach().to(torch.bfloat16)
                ref_b[r0:r0 + w.shape[0]] = b.detach()
                r0 += w.shape[0]
            assert torch.equal(wc, ref_w), f"{key}: row-major packing"
            assert torch.equal(wt, ref_w.t()), f"{key}: transposed packing"
            assert torch.equal(b32, ref_b) and torch.equal(bc, ref_b.to(torch.bfloat16)), f"{key}: bias"
            n_checked += 1
    assert n_checked >= 2 * 4 + 2
