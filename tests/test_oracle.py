"""CPU: the oracle against its committed golden vectors and against itself (three equivalent forms
of the causal product, recurrent == parallel, C clone == PyTorch), Philox known-answer vectors,
RL formula identities.  The encoder internals are unpinned w.r.t. the real fast_transformers (absent); everything around them
is held to the executed reference in tests/test_ref_golden.py."""
import numpy as np
import pytest
import torch

from oracle import ft_oracle as ft, model_oracle as mo, rl_oracle as rl, sampling_oracle as so
from oracle.causal_product_c import causal_dot_product_c

VOCAB = [56, 135, 18, 87, 18, 25]


def test_causal_product_three_forms_agree():
    torch.manual_seed(0)
    Q, K, V = (torch.rand(2, 3, 50, 16, dtype=torch.float64) for _ in range(3))
    a = ft.causal_dot_product_quadratic(Q, K, V)
    b = ft.causal_dot_product_scan(Q, K, V)
    c = causal_dot_product_c(Q.float(), K.float(), V.float())
    assert torch.allclose(a, b, atol=1e-12)
    assert torch.allclose(a.float(), c, atol=1e-4, rtol=1e-5)


def test_causal_product_backward_forms_agree():
    torch.manual_seed(1)
    Q, K, V, G = (torch.randn(1, 2, 40, 16, dtype=torch.float64, requires_grad=True) for _ in range(4))
    ft.causal_dot_product_quadratic(Q, K, V).backward(G)
    gq, gk, gv = ft.causal_dot_product_backward_scan(Q.detach(), K.detach(), V.detach(), G.detach())
    for x, y in ((gq, Q.grad), (gk, K.grad), (gv, V.grad)):
        assert torch.allclose(x, y, atol=1e-11)
    Qf, Kf, Vf = (t.detach().float().requires_grad_() for t in (Q, K, V))
    causal_dot_product_c(Qf, Kf, Vf).backward(G.detach().float())
    for x, y in ((Qf.grad, Q.grad), (Kf.grad, K.grad), (Vf.grad, V.grad)):
        assert torch.allclose(x.double(), y, atol=1e-3, rtol=1e-4)


def test_linattn_golden(golden):
    g = golden("linattn")
    q, k, v, go = (torch.from_numpy(g[n]).double().requires_grad_() for n in ("q", "k", "v", "go"))
    out = ft.causal_linear_attention(q, k, v, product=ft.causal_dot_product_scan)
    out.backward(go.detach())
    assert np.allclose(out.detach().numpy(), g["out"], atol=2e-6)
    for name, t in (("gq", q), ("gk", k), ("gv", v)):
        assert np.allclose(t.grad.numpy(), g[name], atol=1e-5, rtol=1e-5)


def test_recurrent_equals_parallel_and_golden(golden):
    g = golden("recurrent")
    q, k, v = (torch.from_numpy(g[n]).double() for n in ("q", "k", "v"))      # (T,N,H,E)
    par = ft.causal_linear_attention(q.permute(1, 0, 2, 3), k.permute(1, 0, 2, 3), v.permute(1, 0, 2, 3))
    assert np.allclose(par.permute(1, 0, 2, 3).numpy(), g["out"], atol=1e-6)


def test_fla_naive_cross_check():
    """Independent third-party statement of the same recurrence (flash-linear-attention's pure
    PyTorch naive kernel; eps 1e-10 instead of 1e-6)."""
    naive = pytest.importorskip("fla.ops.linear_attn.naive")
    fn = getattr(naive, "naive_chunk_linear_attn", None)
    if fn is None:
        pytest.skip("fla naive_chunk_linear_attn not available")
    torch.manual_seed(2)
    q, k, v = (torch.randn(1, 128, 2, 64, dtype=torch.float64) for _ in range(3))
    mine = ft.causal_linear_attention(q, k, v)
    try:
        theirs = fn(ft.feature_map(q), ft.feature_map(k), v, scale=1.0, normalize=True)
    except Exception as e:  # signature drift between fla versions
        pytest.skip(f"fla naive signature: {e}")
    if theirs.shape != mine.shape:
        theirs = theirs.transpose(1, 2)
    assert torch.allclose(mine, theirs, atol=1e-5, rtol=1e-4)


def test_model_golden_and_state_dict_contract(golden):
    g = golden("model_small")
    m = mo.OracleCPModel(VOCAB, d_model=128, n_layer=2, n_head=2, d_inner=256, dropout=0.0).double().eval()
    sd = {k[4:]: torch.from_numpy(g[k]).double() for k in g.files if k.startswith("sd::")}
    missing = m.load_state_dict(sd, strict=False)
    assert missing.missing_keys == ["pos_emb.pe"] and not missing.unexpected_keys
    x, y, mask = torch.from_numpy(g["x"]), torch.from_numpy(g["y"]), torch.from_numpy(g["mask"]).double()
    h = m.forward_hidden(x)
    assert np.allclose(h.detach().numpy(), g["h"], atol=1e-5)
    losses = torch.stack(m.train_step(x, y, mask))
    assert np.allclose(losses.detach().numpy(), g["losses"], atol=1e-6)
    full = mo.OracleCPModel(VOCAB)
    assert len(full.state_dict()) == 217
    assert sum(p.numel() for p in full.parameters()) == 38982227
    assert full.state_dict()["pos_emb.pe"].shape == (1, 20000, 512)


def test_recurrent_pos0_quirk_differs_from_parallel(golden):
    g = golden("model_small")
    assert np.abs(g["h_rec_true"] - g["h"][0, :12]).max() < 1e-4      # true positions == parallel
    assert np.abs(g["h_rec_pos0"] - g["h"][0, :12]).max() > 1e-2      # reference quirk


def test_philox_known_answers():
    assert so.philox4x32_10((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert so.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert so.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_choice_from_uniform_is_numpy_choice():
    """np.random.choice(a, p=p) == choice_from_uniform(a, p, u) for the u it draws."""
    rng = np.random.RandomState(3)
    for _ in range(200):
        n = rng.randint(2, 40)
        p = rng.rand(n)
        p /= p.sum()
        cand = rng.permutation(n)
        state = rng.get_state()
        want = rng.choice(cand, size=1, p=p)[0]
        rng.set_state(state)
        u = rng.random_sample()
        assert so.choice_from_uniform(cand, p, u) == want


def test_sampling_golden_and_literal_path(golden):
    g = golden("sampling")
    logits, seg = g["logits"], g["seg"]
    for r in range(logits.shape[0]):
        for a, name in enumerate(so.ATTRS):
            lg = logits[r, seg[a]:seg[a + 1]]
            t, p = so.SAMPLING_CFG[name]
            u = so.philox_uniform(int(g["seed"]), int(g["seq_base"]) + r, int(g["step"]), a)
            assert abs(u - g["u"][r, a]) == 0
            assert so.sampling_from_uniform(lg, u, p=p, t=t) == g["sampled"][r, a]
            assert so.greedy(lg) == g["greedy"][r, a]
    # nucleus keeps the prefix up to and including the first index whose cumsum exceeds p
    probs = np.array([0.5, 0.3, 0.15, 0.05], dtype=np.float32)
    cand, cp = so.nucleus_candidates(probs.copy(), 0.7)
    assert list(cand) == [0, 1] and abs(cp.sum() - 1) < 1e-6
    out = so.forward_output_sampling({n: logits[0, seg[a]:seg[a + 1]] for a, n in enumerate(so.ATTRS)}, np.random.RandomState(0))
    assert out.shape == (6,)


def test_returns_compat_closed_form(golden):
    g = golden("rl")
    r = torch.from_numpy(g["rewards"]).double().reshape(-1)
    T = len(r)
    closed = torch.stack([sum(0.99 ** (T - 1 - t - j) * r[j] for j in range(T - t)) for t in range(T)])
    assert np.allclose(closed.numpy(), g["ret_raw"].reshape(-1), atol=1e-5)
    togo = torch.stack([sum(0.99 ** (j - t) * r[j] for j in range(t, T)) for t in range(T)])
    assert not np.allclose(togo.numpy(), g["ret_raw"].reshape(-1), atol=1e-3)        # NOT reward-to-go
    ret = rl.calculate_returns_compat(torch.from_numpy(g["rewards"]).double(), 0.99)
    assert np.allclose(ret.numpy(), g["ret"], atol=1e-5)
    adv = rl.calculate_advantages_compat(ret, torch.from_numpy(g["values"]).double())
    assert np.allclose(adv.numpy(), g["adv"], atol=1e-5)


def test_rl_golden(golden):
    g = golden("rl")
    seg = g["seg"]
    split = lambda t: [t[..., seg[i]:seg[i + 1]] for i in range(6)]
    ql, nq = torch.from_numpy(g["ql"]).double(), torch.from_numpy(g["nq"]).double()
    act, rw, dn = torch.from_numpy(g["action"]), torch.from_numpy(g["rw"]).double(), torch.from_numpy(g["dn"]).double()
    assert abs(rl.dqn_td_loss_compat(split(ql), split(nq), act, rw, dn).item() - float(g["td_compat"])) < 1e-5
    assert abs(rl.dqn_td_loss_standard(split(ql), split(nq), act, rw, dn).item() - float(g["td_standard"])) < 1e-5
    # the gather quirk: batch element 0, sequence position = batch index
    i = 3
    q = split(ql)[i].gather(2, act[:, :, i].unsqueeze(0)).squeeze(0)
    assert torch.equal(q, torch.stack([split(ql)[i][0, j, act[j, :, i]] for j in range(act.shape[0])]))
    win = torch.from_numpy(g["win"]).double()
    a_ppo, lp = rl.ppo_choose_action_compat(split(win))
    assert np.array_equal(a_ppo.numpy(), g["act_ppo"]) and np.allclose(lp.numpy(), g["lp_ppo"], atol=1e-5)
    assert np.array_equal(rl.dqn_choose_action_compat(split(win)).numpy(), g["act_dqn"])
    gadv, gret = rl.gae_standard(*(torch.from_numpy(g[n]).double() for n in ("r2", "v2", "d2", "lv")), 0.99, 0.95)
    assert np.allclose(gadv.numpy(), g["gae_adv"], atol=1e-5) and np.allclose(gret.numpy(), g["gae_ret"], atol=1e-5)
