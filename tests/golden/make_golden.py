"""Regenerates the golden fixtures in tests/golden/ from the oracle (fixed seeds).

    PYTHONPATH=/root/repo python tests/golden/make_golden.py

The reference ships no golden vectors and its arithmetic dependency (fast_transformers 0.4.0) is
absent, so these vectors pin the ORACLE ("parity unpinned"); they exist so that the GPU
parity tests and the oracle's own regression tests compare against committed numbers rather than
only against a live re-run.  Everything is computed in float64 where the oracle allows and stored
as float32.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ft_oracle as ft, model_oracle as mo, rl_oracle as rl, sampling_oracle as so  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
VOCAB = [56, 135, 18, 87, 18, 25]


def f32(t):
    return t.detach().to(torch.float32).numpy()


def linattn():
    g = torch.Generator().manual_seed(101)
    N, L, H, E = 2, 200, 2, 64
    q, k, v, go = (torch.randn(N, L, H, E, generator=g, dtype=torch.float64) for _ in range(4))
    q.requires_grad_(), k.requires_grad_(), v.requires_grad_()
    out = ft.causal_linear_attention(q, k, v)
    out.backward(go)
    den = torch.einsum("nlhi,nlhi->nlh", ft.feature_map(q), ft.feature_map(k).cumsum(1)) + ft.EPS
    np.savez_compressed(os.path.join(OUT, "linattn.npz"), q=f32(q), k=f32(k), v=f32(v), go=f32(go), out=f32(out),
                        den=f32(den), gq=f32(q.grad), gk=f32(k.grad), gv=f32(v.grad))


def recurrent():
    g = torch.Generator().manual_seed(102)
    T, N, H, E = 5, 3, 2, 64
    q, k, v = (torch.randn(T, N, H, E, generator=g, dtype=torch.float64) for _ in range(3))
    st, outs = None, []
    for t in range(T):
        o, st = ft.recurrent_linear_attention(q[t], k[t], v[t], st)
        outs.append(o)
    np.savez_compressed(os.path.join(OUT, "recurrent.npz"), q=f32(q), k=f32(k), v=f32(v), out=f32(torch.stack(outs)),
                        S=f32(st[0]), Z=f32(st[1]))


def model_small():
    torch.manual_seed(103)
    cfg = dict(d_model=128, n_layer=2, n_head=2, d_inner=256, dropout=0.0)
    m = mo.OracleCPModel(VOCAB, **cfg).double().eval()
    g = torch.Generator().manual_seed(104)
    N, L = 2, 70
    x = torch.stack([torch.randint(0, n, (N, L), generator=g) for n in VOCAB], -1)
    y = x.roll(-1, 1)
    mask = torch.zeros(N, L, dtype=torch.float64)
    mask[0, :50] = 1
    mask[1, :61] = 1
    h = m.forward_hidden(x)
    logits = m.forward_output(h)
    losses = m.train_step(x, y, mask)
    (sum(losses) / 6).backward()
    grad_keys = ["in_linear.weight", "in_linear.bias", "transformer_encoder.layers.0.attention.query_projection.weight",
                 "transformer_encoder.layers.1.attention.value_projection.bias", "transformer_encoder.layers.0.linear1.weight",
                 "transformer_encoder.layers.1.norm2.weight", "transformer_encoder.norm.bias", "word_emb_pitch.lut.weight",
                 "word_emb_barbeat.lut.weight", "proj_chord.weight", "proj_velocity.bias"]
    params = dict(m.named_parameters())
    blob = {"x": x.numpy(), "y": y.numpy(), "mask": f32(mask), "h": f32(h), "logits": f32(torch.cat(logits, -1)),
            "losses": f32(torch.stack(losses))}
    for k_, v_ in m.state_dict().items():
        if k_ != "pos_emb.pe":
            blob["sd::" + k_] = f32(v_)
    for k_ in grad_keys:
        blob["grad::" + k_] = f32(params[k_].grad)
    # recurrent hidden states with the reference quirk (position 0 every step) and with true positions
    mr = mo.OracleCPModel(VOCAB, is_training=False, **cfg).double().eval()
    mr.load_state_dict(m.state_dict())
    with torch.no_grad():
        for name, true_pos in (("h_rec_pos0", False), ("h_rec_true", True)):
            mem, hs = None, []
            for t in range(12):
                hh, mem = mr.forward_hidden(x[:1, t:t + 1], mem, is_training=False, pos_offset=t if true_pos else 0)
                hs.append(hh)
            blob[name] = f32(torch.cat(hs, 0))
    np.savez_compressed(os.path.join(OUT, "model_small.npz"), **blob)


def sampling():
    rng = np.random.RandomState(105)
    seg = np.concatenate([[0], np.cumsum(VOCAB)])
    rows = 24
    logits = (rng.randn(rows, seg[-1]) * 2.0).astype(np.float32)
    seed = 20261018
    greedy = np.zeros((rows, 6), np.int64)
    sampled = np.zeros((rows, 6), np.int64)
    us = np.zeros((rows, 6), np.float64)
    margin = np.zeros((rows, 6), np.float64)       # distance of u from the nearest CDF edge
    step = 7
    for r in range(rows):
        for a, name in enumerate(so.ATTRS):
            lg = logits[r, seg[a]:seg[a + 1]]
            greedy[r, a] = so.greedy(lg)
            t, p = so.SAMPLING_CFG[name]
            u = so.philox_uniform(seed, 1000 + r, step, a)
            us[r, a] = u
            sampled[r, a] = so.sampling_from_uniform(lg, u, p=p, t=t)
            probs = so.softmax_with_temperature(lg, t)
            if p is not None:
                cand, cp = so.nucleus_candidates(probs, p)
            else:
                probs = probs / sum(probs)
                cand = np.argsort(probs)[::-1]
                cp = probs[cand]
            cdf = np.cumsum(cp.astype(np.float64))
            cdf /= cdf[-1]
            margin[r, a] = np.min(np.abs(cdf - u))
    np.savez_compressed(os.path.join(OUT, "sampling.npz"), logits=logits, seg=seg, seed=seed, seq_base=1000, step=step,
                        greedy=greedy, sampled=sampled, u=us, margin=margin)


def rl_vectors():
    g = torch.Generator().manual_seed(106)
    T, A = 30, 25
    rewards = torch.rand(T, 1, generator=g, dtype=torch.float64)
    values = torch.randn(T, 1, generator=g, dtype=torch.float64) * 0.3
    ret = rl.calculate_returns_compat(rewards, 0.99)
    ret_raw = rl.calculate_returns_compat(rewards, 0.99, normalize=False)
    adv = rl.calculate_advantages_compat(ret, values)
    new_logp = (-torch.rand(A, 6, generator=g, dtype=torch.float64) * 2).requires_grad_()
    old_long = (-torch.rand(T, A, 6, generator=g, dtype=torch.float64) * 2.5).long().double()
    ploss = rl.ppo_policy_loss_compat(new_logp, old_long, adv)
    ploss.backward()
    # standard forms
    B, TT = 4, 77
    r2 = torch.rand(B, TT, generator=g, dtype=torch.float64)
    v2 = torch.randn(B, TT, generator=g, dtype=torch.float64)
    d2 = (torch.rand(B, TT, generator=g) < 0.05).double()
    lv = torch.randn(B, generator=g, dtype=torch.float64)
    gadv, gret = rl.gae_standard(r2, v2, d2, lv, 0.99, 0.95)
    togo = rl.rewards_to_go_standard(r2, d2, 0.99)
    n = 500
    nl = (-torch.rand(n, generator=g, dtype=torch.float64)).requires_grad_()
    ol = -torch.rand(n, generator=g, dtype=torch.float64)
    ad = torch.randn(n, generator=g, dtype=torch.float64)
    en = torch.rand(n, generator=g, dtype=torch.float64).requires_grad_()
    va = torch.randn(60, generator=g, dtype=torch.float64).requires_grad_()
    rt = torch.randn(60, generator=g, dtype=torch.float64)
    sl, spl, svl, sel = rl.ppo_loss_standard(nl, ol, ad, va, rt, en)
    sl.backward()
    # DQN
    Bq, L = 6, 50
    seg = np.concatenate([[0], np.cumsum(VOCAB)])
    ql = torch.randn(Bq, L, int(seg[-1]), generator=g, dtype=torch.float64).requires_grad_()
    nq = torch.randn(Bq, L, int(seg[-1]), generator=g, dtype=torch.float64)
    action = torch.stack([torch.randint(0, nv, (Bq, A), generator=g) for nv in VOCAB], -1)
    rw = torch.rand(Bq, 1, generator=g, dtype=torch.float64)
    dn = (torch.rand(Bq, 1, generator=g) < 0.3).double()
    split = lambda t: [t[..., seg[i]:seg[i + 1]] for i in range(6)]
    tdc = rl.dqn_td_loss_compat(split(ql), split(nq), action, rw, dn)
    tdc.backward()
    gq_c = ql.grad.clone()
    ql.grad = None
    tds = rl.dqn_td_loss_standard(split(ql), split(nq), action, rw, dn)
    tds.backward()
    # action read-outs on one window
    win = torch.randn(1, L, int(seg[-1]), generator=g, dtype=torch.float64)
    act_ppo, lp_ppo = rl.ppo_choose_action_compat(split(win))
    act_dqn = rl.dqn_choose_action_compat(split(win))
    np.savez_compressed(
        os.path.join(OUT, "rl.npz"), rewards=f32(rewards), values=f32(values), ret=f32(ret), ret_raw=f32(ret_raw), adv=f32(adv),
        new_logp=f32(new_logp), old_long=f32(old_long), ploss=f32(ploss), dnew=f32(new_logp.grad),
        r2=f32(r2), v2=f32(v2), d2=f32(d2), lv=f32(lv), gae_adv=f32(gadv), gae_ret=f32(gret), togo=f32(togo),
        nl=f32(nl), ol=f32(ol), ad=f32(ad), en=f32(en), va=f32(va), rt=f32(rt), std_losses=f32(torch.stack([sl, spl, svl, sel])),
        d_nl=f32(nl.grad), d_en=f32(en.grad), d_va=f32(va.grad),
        seg=seg, ql=f32(ql), nq=f32(nq), action=action.numpy(), rw=f32(rw), dn=f32(dn), td_compat=f32(tdc), gq_compat=f32(gq_c),
        td_standard=f32(tds), gq_standard=f32(ql.grad), win=f32(win), act_ppo=act_ppo.numpy(), lp_ppo=f32(lp_ppo),
        act_dqn=act_dqn.numpy())


if __name__ == "__main__":
    linattn()
    recurrent()
    model_small()
    sampling()
    rl_vectors()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
