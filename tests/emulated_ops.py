"""TEST INFRASTRUCTURE: stand-ins for the CUDA entry points of ``cpmusic.ops`` written with plain torch / the oracle, so
that the product's HOST LOGIC (model.py, encoder.py, rl.py, data.py, midi.py: packing caches, reshapes, segment offsets,
position handling, read-out index arithmetic, buffer plumbing) can be executed on a machine without a GPU and held to the
same executed-reference vectors as the CUDA path.  Nothing here is reachable from the product: the stand-ins are patched
onto the module only inside ``emulated()`` by tests, and they say nothing about the kernels themselves (those are tested
on the GPU)."""
import contextlib

import torch
import torch.nn.functional as F

from oracle import ft_oracle as ft, rl_oracle as rl, sampling_oracle as so


def _segments(x, seg):
    return [x[..., seg[i]:seg[i + 1]] for i in range(len(seg) - 1)]


def cp_embed(idx, tables, dtype=torch.bfloat16):
    parts = [F.embedding(idx[..., a], t) * (t.shape[1] ** 0.5) for a, t in enumerate(tables)]
    return torch.cat(parts, -1).to(dtype)


def add_pe(x, pe, L, pos_offset=0, pos_dev=None, p_drop=0.0):
    assert p_drop == 0.0, "emulation covers the deterministic path"
    d = x.shape[-1]
    rows = x.numel() // d
    base = int(pos_dev[0]) if pos_dev is not None else pos_offset       # the kernel reads the device-side step counter
    pos = ((torch.arange(rows) % L) + base).clamp(max=pe.reshape(-1, d).shape[0] - 1)
    return (x.reshape(rows, d).float() + pe.reshape(-1, d)[pos]).to(x.dtype).view(x.shape)


def ln_residual(x, res, gamma, beta, eps=1e-5, p_drop=0.0, res_bias=None):
    assert p_drop == 0.0
    y = x.float()
    if res is not None:
        y = y + res.float() + (res_bias if res_bias is not None else 0.0)
    return F.layer_norm(y, (y.shape[-1],), gamma, beta, eps).to(x.dtype)


def gelu_dropout(x, p_drop=0.0, bias=None):
    assert p_drop == 0.0
    return F.gelu(x.float() + (bias if bias is not None else 0.0)).to(x.dtype)


def colsum(x):
    return x.sum(0, dtype=torch.float32)


def causal_linear_attention_fused(qkv, n_heads, eps=1e-6, impl=0, key_mask=None):
    N, L, W = qkv.shape
    E = W // (3 * n_heads)
    q, k, v = (qkv[..., j * n_heads * E:(j + 1) * n_heads * E].reshape(N, L, n_heads, E).float() for j in range(3))
    klm = None if key_mask is None else key_mask.float()
    return ft.causal_linear_attention(q, k, v, key_lengths_mask=klm, eps=eps).reshape(N, L, n_heads * E).to(qkv.dtype)


def linattn_step(q, k, v, S, Z, eps=1e-6, **_kw):
    if S.shape[0] != q.shape[0]:
        raise ValueError("The batch size changed during iteration")
    out, (S2, Z2) = ft.recurrent_linear_attention(q.float(), k.float(), v.float(), [S, Z], eps=eps)
    S.copy_(S2)                                                         # the kernel updates the caller's buffers in place
    Z.copy_(Z2)
    return out.to(q.dtype)


def heads_sample(logits, seg, temperature=None, top_p=None, greedy=True, seed=0, seq_base=0, step=0, step_dev=None,
                 want_logp=False, want_entropy=False, tokens_out=None, logp_out=None):
    if step_dev is not None:
        step = int(step_dev[0])
    want_logp = want_logp or logp_out is not None
    A = len(seg) - 1
    l2 = logits.reshape(-1, logits.shape[-1]).float()
    rows = l2.shape[0]
    t = temperature if temperature is not None else [1.0] * A
    p = top_p if top_p is not None else [None] * A
    tok = torch.empty(rows, A, dtype=torch.int64)
    for a, lg in enumerate(_segments(l2, seg)):
        if greedy:
            tok[:, a] = lg.argmax(-1)
        else:
            for r in range(rows):
                u = so.philox_uniform(seed, seq_base + r, step, a)
                tok[r, a] = so.sampling_from_uniform(lg[r].detach().numpy(), u, p=p[a] or None, t=t[a])
    logp = ent = None
    if want_logp or want_entropy:
        ls = [torch.log_softmax(lg, -1) for lg in _segments(l2, seg)]
        if want_logp:
            logp = torch.stack([ls[a].gather(-1, tok[:, a:a + 1])[:, 0] for a in range(A)], -1)
        if want_entropy:
            ent = torch.stack([-(x.exp() * x).sum(-1) for x in ls], -1)
    if tokens_out is not None:
        tokens_out.copy_(tok)
        tok = tokens_out
    if logp_out is not None and logp is not None:
        logp_out.copy_(logp)
        logp = logp_out
    return tok, logp, ent


def heads_logp(logits, tokens, seg, want_entropy=False):
    A = len(seg) - 1
    l2 = logits.reshape(-1, logits.shape[-1]).float()
    tk = tokens.reshape(-1, A)
    ls = [torch.log_softmax(lg, -1) for lg in _segments(l2, seg)]
    logp = torch.stack([ls[a].gather(-1, tk[:, a:a + 1])[:, 0] for a in range(A)], -1).view(tokens.shape)
    ent = torch.stack([-(x.exp() * x).sum(-1) for x in ls], -1).view(tokens.shape) if want_entropy else None
    return logp, ent


def masked_ce(logits, targets, mask, seg, group=None):
    assert group is None
    A = len(seg) - 1
    l2 = logits.reshape(-1, logits.shape[-1]).float()
    tg = targets.reshape(-1, A)
    m = mask.reshape(-1).float()
    ce = torch.stack([F.cross_entropy(lg, tg[:, a], reduction="none") for a, lg in enumerate(_segments(l2, seg))], -1)
    return (ce * m[:, None]).sum(0) / m.sum()


def returns_scan(rewards, gamma, mode="compat", values=None, dones=None, last_value=None, lam=0.95):
    r = rewards.float()
    if mode == "compat":
        return torch.stack([rl.calculate_returns_compat(row, gamma, normalize=False)[:, 0] for row in r])
    if mode == "togo":
        return rl.rewards_to_go_standard(r, dones.float(), gamma)
    return rl.gae_standard(r, values.float(), dones.float(), last_value.float(), gamma, lam)


def zscore(x, sub=None, unbiased=True, eps=0.0, group=None):
    assert group is None
    y = x.float() - (sub.float() if sub is not None else 0.0)
    return (y - y.mean()) / (y.std(unbiased=unbiased) + eps)


def ppo_loss_compat(new_logp, old_logp, adv, clip=0.2):
    T = old_logp.shape[0]
    return rl.ppo_policy_loss_compat(new_logp.reshape(1, -1), old_logp.reshape(T, 1, -1), adv.reshape(T, 1), clip)


def dqn_td_loss(q_logits, next_logits, action, reward, done, seg, n_actions=25, gamma=0.95, compat=True):
    fn = rl.dqn_td_loss_compat if compat else rl.dqn_td_loss_standard
    q6, n6 = _segments(q_logits.float(), seg), _segments(next_logits.detach().float(), seg)
    B = q_logits.shape[0]
    loss = fn(q6, n6, action, reward.reshape(B, 1).float(), done.reshape(B, 1).float(), gamma, n_actions)
    return loss, None


def rollout_advance(tokens, history_tok, vals, history_f, step_dev, max_steps):
    step = int(step_dev[0])
    if step < max_steps:
        if history_tok is not None:
            history_tok[step].copy_(tokens.view_as(history_tok[step]))
        if history_f is not None:
            history_f[step].copy_(vals.view_as(history_f[step]))
    step_dev += 1


def reward_head(h, u, c, want_scores=False):
    scores = torch.sigmoid(h.float().mean(1) @ u.t() + c)
    return (scores.mean(-1), scores) if want_scores else scores.mean(-1)


def rowdot(h, u, c=None):
    out = h.float() @ u.float()
    return out if c is None else out + c


EMULATED = dict(rowdot=rowdot, cp_embed=cp_embed, add_pe=add_pe, ln_residual=ln_residual, gelu_dropout=gelu_dropout, colsum=colsum,
                causal_linear_attention_fused=causal_linear_attention_fused, linattn_step=linattn_step, heads_sample=heads_sample,
                heads_logp=heads_logp, masked_ce=masked_ce, returns_scan=returns_scan, zscore=zscore, ppo_loss_compat=ppo_loss_compat,
                dqn_td_loss=dqn_td_loss, reward_head=reward_head, rollout_advance=rollout_advance)


@contextlib.contextmanager
def emulated(cpm):
    """Patches the stand-ins onto ``cpm.ops`` for the duration of the block.  Refuses to do so where a GPU exists: there the
    real kernels are what gets tested, and nothing may stand in for them."""
    if torch.cuda.is_available():
        raise RuntimeError("emulated_ops is for GPU-less hosts only; on a GPU box run the -m gpu tests against the real kernels")
    saved = {k: getattr(cpm.ops, k) for k in EMULATED}
    try:
        for k, fn in EMULATED.items():
            setattr(cpm.ops, k, fn)
        yield
    finally:
        for k, fn in saved.items():
            setattr(cpm.ops, k, fn)


__all__ = ["emulated", "EMULATED"]
