"""PPO / DQN arithmetic of the reference on the cpmusic kernels.

``*_compat`` reproduce what the reference computes, quirks included;
the others are the standard forms BASELINE.json's north_star names.  Anchors:
ppo_policy/ppo_train.py:251-402 and dqn_policy/IRL_dqn_train.py:240-336.
All tensors stay on the device; nothing here calls ``.item()``/``.cpu()``.
"""
from __future__ import annotations

import torch

from . import ops

N_ACTIONS = 25


# ------------------------------------------------------------------------------------------ PPO
def _tail_rows(lc, n_actions):
    """(B,L,W) concatenated logits -> (B,n_actions,W) with row r = position -(r+1)."""
    return lc[:, -n_actions:, :].flip(1).contiguous()


def ppo_choose_action(model, state_x, n_actions=N_ACTIONS, compat=True):
    """``PPO.choose_action`` (ppo_train.py:251-290): parallel forward over the window, greedy tokens
    at positions -1..-n_actions, log of their softmax probability.  state_x (1,L,A).
    compat: tempo/chord log-probs are read at the vocabulary index of the token argmax'd at
    position +(r+1) (ppo_train.py:273-274).  -> action (n_actions,A) int64, logp (n_actions,A)."""
    lc = model.logits_concat(model.hidden(state_x))                   # (1,L,W)
    L = lc.shape[1]
    tok_all, _, _ = ops.heads_sample(lc[0], model.seg, greedy=True)   # (L,A)
    rows = _tail_rows(lc, n_actions)[0]                               # (n_actions,W)
    action = tok_all[L - n_actions:].flip(0).contiguous()
    sel = action.clone()
    if compat:
        sel[:, :2] = tok_all[1:n_actions + 1, :2]
    logp, _ = ops.heads_logp(rows, sel, model.seg)
    return action, logp


def ppo_select_update(model, states, n_actions=N_ACTIONS, compat=True, want_entropy=False):
    """``PPO.select_udpate`` (ppo_train.py:293-346) without the Python double loop.
    states (B,L,A).  compat: returns only the LAST batch element's (action, logp) like the
    reference; otherwise all (B,n_actions,A)."""
    lc = model.logits_concat(model.hidden(states))
    rows = _tail_rows(lc, n_actions)                                   # (B,n_actions,W)
    if compat:
        rows = rows[-1:]
    flat = rows.reshape(-1, rows.shape[-1])
    action, _, _ = ops.heads_sample(flat, model.seg, greedy=True)
    logp, ent = ops.heads_logp(flat, action, model.seg, want_entropy)
    shape = (n_actions, len(model.seg) - 1) if compat else (rows.shape[0], n_actions, len(model.seg) - 1)
    out = (action.view(shape), logp.view(shape))
    return out + (ent.view(shape),) if want_entropy else out


def calculate_returns_compat(rewards, gamma, normalize=True, group=None):
    """``calculate_returns`` (ppo_train.py:348-357): rewards (T,1)|(T,) -> (T,1)."""
    ret = ops.returns_scan(rewards.reshape(1, -1), gamma, "compat").reshape(-1, 1)
    return ops.zscore(ret, unbiased=True, eps=0.0, group=group) if normalize else ret


def calculate_advantages_compat(returns, values, normalize=True, group=None):
    """``calculate_advantages`` (ppo_train.py:359-363)."""
    if normalize:
        return ops.zscore(returns, sub=values.to(returns.device), unbiased=True, eps=0.0, group=group)
    return returns - values


def gae(rewards, values, dones, last_value, gamma=0.99, lam=0.95, normalize=True, group=None):
    """GAE(λ): rewards/values/dones (B,T), last_value (B,) -> (adv, ret); adv z-scored over the
    global batch when normalize (moments all-reduced over ``group``)."""
    adv, ret = ops.returns_scan(rewards, gamma, "gae", values=values, dones=dones, last_value=last_value, lam=lam)
    if normalize:
        adv = ops.zscore(adv, unbiased=False, eps=1e-8, group=group)
    return adv, ret


def ppo_policy_loss_compat(new_logp, old_logp_long, advantages, clip=0.2):
    """ppo_train.py:388-396: new_logp (A,C'), old (T,A,C') int64-truncated, advantages (T,1)."""
    return ops.ppo_loss_compat(new_logp.reshape(-1), old_logp_long.reshape(old_logp_long.shape[0], -1).float(),
                               advantages.reshape(-1), clip)


def value_loss_compat(returns, value_pred):
    return torch.nn.functional.mse_loss(returns, value_pred).sum()


# ------------------------------------------------------------------------------------------ DQN
def dqn_choose_action(model, x, n_actions=N_ACTIONS, compat=True):
    """``DQN.choose_action`` (IRL_dqn_train.py:240-264): greedy tokens at positions
    [0,-1,...,-(n_actions-1)] (compat, because ``-0 == 0``) or [-1..-n_actions].  x (1,L,A)."""
    lc = model.logits_concat(model.hidden(x))
    tok_all, _, _ = ops.heads_sample(lc[0], model.seg, greedy=True)    # (L,A)
    L = tok_all.shape[0]
    if compat:
        pos = torch.tensor([0] + [L - i for i in range(1, n_actions)], device=tok_all.device)
    else:
        pos = torch.tensor([L - i for i in range(1, n_actions + 1)], device=tok_all.device)
    return tok_all[pos]


def dqn_td_loss(eval_net, target_net, state, next_state, action, reward, done, gamma=0.95, n_actions=N_ACTIONS,
                compat=True):
    """TD block of ``DQN.update`` (IRL_dqn_train.py:285-330) as one fused kernel over the two nets'
    concatenated logits.  Returns the mean-of-attributes MSE (differentiable in eval_net)."""
    q = eval_net.logits_concat(eval_net.hidden(state))
    with torch.no_grad():
        nq = target_net.logits_concat(target_net.hidden(next_state))
    loss, _ = ops.dqn_td_loss(q, nq, action, reward, done, eval_net.seg, n_actions, gamma, compat)
    return loss


# ------------------------------------------------------------------------------------------ reward head
class RewardHead(torch.nn.Module):
    """The read-out of the PPO reward model ``LongFormer.token_forward`` (ppo_policy/model.py:459-494) — six ``proj_*``
    ``Linear(d, n_a)``, six ``eval_*`` ``Linear(n_a, 1)``, mean over the sequence, sigmoid, average — with the reference's
    parameter names, evaluated by one fused kernel on the body's last hidden state.  The Longformer body itself is out
    of scope (HF ``transformers``); pass its ``last_hidden_state``."""
    ATTRS = ("tempo", "chord", "barbeat", "pitch", "duration", "velocity")

    def __init__(self, n_token, d_model=512):
        super().__init__()
        if len(n_token) != len(self.ATTRS):
            raise ValueError("RewardHead follows the reference's 6-attribute layout")
        for a, n in zip(self.ATTRS, n_token):
            setattr(self, f"proj_{a}", torch.nn.Linear(d_model, int(n)))
            setattr(self, f"eval_{a}", torch.nn.Linear(int(n), 1))

    def collapsed(self):
        """(u (6,d), c (6,)): eval_a(proj_a(h)) == h . u_a + c_a."""
        us, cs = [], []
        for a in self.ATTRS:
            proj, ev = getattr(self, f"proj_{a}"), getattr(self, f"eval_{a}")
            us.append(ev.weight.float() @ proj.weight.float())                      # (1,d)
            cs.append(ev.weight.float() @ proj.bias.float() + ev.bias.float())      # (1,)
        return torch.cat(us, 0), torch.cat(cs, 0)

    @torch.no_grad()
    def forward(self, hidden, want_scores=False):
        u, c = self.collapsed()
        return ops.reward_head(hidden, u, c, want_scores)


class DiscriminatorHead(torch.nn.Module):
    """The read-out of the DQN-side AIRL discriminator (dqn_policy/AIRL_model.py:91-98,117-120): the Longformer's
    ``last_hidden_state`` averaged over ALL positions (the reference ignores the attention mask here) followed by
    ``score_classifier`` = Linear(d,128) -> BatchNorm1d(128) -> Tanh -> Linear(128,64) -> Tanh -> Linear(64,1) -> Sigmoid, with
    the reference's parameter names (``score_classifier.{0,1,3,5}.*``) so its checkpoints load.  (N, L, d) -> (N, 1) in
    (0, 1).  Differentiable (the reference trains it with BCE, AIRL.py:61-90); BatchNorm follows ``train()`` / ``eval()``.
    A few thousand rows of a 128-wide MLP: plain library ops on whatever device ``hidden`` lives on - like
    ``value_funtion``, not worth a kernel.  The Longformer body is out of scope (HF ``transformers``)."""

    def __init__(self, d_model=512):
        super().__init__()
        nn = torch.nn
        self.score_classifier = nn.Sequential(nn.Linear(d_model, 128), nn.BatchNorm1d(128), nn.Tanh(), nn.Linear(128, 64), nn.Tanh(),
                                              nn.Linear(64, 1), nn.Sigmoid())

    def forward(self, hidden, masks=None):
        return self.score_classifier(hidden.to(self.score_classifier[0].weight.dtype).mean(dim=1))      # bf16 body output -> fp32 masters
