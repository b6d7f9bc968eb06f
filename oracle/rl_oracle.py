"""Oracle restatement of the reference's RL arithmetic (PPO + DQN), plain PyTorch.

TEST INFRASTRUCTURE ONLY — see ``oracle/__init__.py``.  PINNED against the executed
reference (tests/golden/make_ref_golden.py -> tests/test_ref_golden.py).

``*_compat`` functions are transcriptions of what the reference *actually*
computes, quirks included; ``*_standard`` functions are the
textbook forms BASELINE.json's north_star names (GAE(λ), ratio·A clipped
surrogate, entropy bonus, detached per-position TD target).

Reference anchors:
* ``PPO.choose_action``        ppo_policy/ppo_train.py:251-290
* ``PPO.select_udpate``        ppo_policy/ppo_train.py:293-346
* ``calculate_returns``        ppo_policy/ppo_train.py:348-357
* ``calculate_advantages``     ppo_policy/ppo_train.py:359-363
* ``update_policy`` losses     ppo_policy/ppo_train.py:388-402
* ``DQN.choose_action``        dqn_policy/IRL_dqn_train.py:240-264
* ``DQN.update`` TD block      dqn_policy/IRL_dqn_train.py:285-336
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

N_ACTIONS = 25


# ------------------------------------------------------------------------- PPO
def ppo_choose_action_compat(logits6, n_actions=N_ACTIONS):
    """logits6: 6 tensors (1, L, n_i).  Action row r (0-based) is the argmax token
    at position -(r+1); its log-prob is read at position -(r+1) but, for tempo and
    chord, at the vocabulary index of the token argmax'd at position +(r+1)
    (ppo_train.py:273-274 quirk).  → action (n_actions,6) int64, logp (n_actions,6)."""
    probs = [torch.softmax(y, dim=-1) for y in logits6]
    tok = [p.argmax(dim=-1) for p in probs]
    act, lp = [], []
    for idx in range(1, n_actions + 1):
        act.append(torch.stack([t[0, -idx] for t in tok]))
        sel = [probs[a][0, -idx, tok[a][0, idx if a < 2 else -idx]] for a in range(6)]
        lp.append(torch.log(torch.stack(sel)))
    return torch.stack(act), torch.stack(lp)


def ppo_select_update_compat(logits6, n_actions=N_ACTIONS):
    """Batched twin used in the update; returns the LAST batch element's
    (action, logp) only (ppo_train.py:346), all indices at position -(r+1)."""
    probs = [torch.softmax(y, dim=-1) for y in logits6]
    tok = [p.argmax(dim=-1) for p in probs]
    b = logits6[0].shape[0] - 1
    act, lp = [], []
    for idx in range(1, n_actions + 1):
        act.append(torch.stack([t[b, -idx] for t in tok]))
        lp.append(torch.log(torch.stack([probs[a][b, -idx, tok[a][b, -idx]] for a in range(6)])))
    return torch.stack(act), torch.stack(lp)


def action_logp_all(logits6, n_actions=N_ACTIONS):
    """Clean batched read-out: for every batch element, argmax token and its
    log-softmax at positions -1..-n_actions → (B,n_actions,6) each."""
    acts, lps = [], []
    for y in logits6:
        tail = y[:, -n_actions:, :].flip(1)                 # row r ↔ position -(r+1)
        ls = torch.log_softmax(tail, dim=-1)
        a = tail.argmax(dim=-1)
        acts.append(a)
        lps.append(ls.gather(-1, a[..., None])[..., 0])
    return torch.stack(acts, -1), torch.stack(lps, -1)


def calculate_returns_compat(rewards, gamma, normalize=True):
    """rewards (T,1) or (T,).  R_i = r_i + γ R_{i-1} accumulated FORWARD, each new R
    inserted at the front: returns[t] = Σ_{j<=T-1-t} γ^{T-1-t-j} r_j.  z-scored with
    the unbiased std and no epsilon."""
    r = rewards.reshape(-1)
    out, R = [], torch.zeros((), dtype=r.dtype)
    for x in r:
        R = x + R * gamma
        out.insert(0, R)
    ret = torch.stack(out)[:, None]
    if normalize:
        ret = (ret - ret.mean()) / ret.std()
    return ret


def calculate_advantages_compat(returns, values, normalize=True):
    adv = returns - values
    if normalize:
        adv = (adv - adv.mean()) / adv.std()
    return adv


def ppo_policy_loss_compat(new_logp, old_logp_long, advantages, clip=0.2):
    """new_logp (A,6); old_logp_long (T,A,6) — the int64-truncated stored
    log-probs (ppo_train.py:135); advantages (T,1).
    ``-mean(min(0.2·A, clamp(ratio)·A))`` — first arm is NOT ratio·A (ppo_train.py:391)."""
    ratio = (new_logp - old_logp_long).exp()
    arm1 = (0.2 * advantages).unsqueeze(2)
    arm2 = torch.clamp(ratio, 1.0 - clip, 1.0 + clip) * advantages.unsqueeze(2)
    return -torch.min(arm1, arm2).mean()


def value_loss_compat(returns, value_pred):
    return F.mse_loss(returns, value_pred).sum()


def rewards_to_go_standard(rewards, dones, gamma):
    """Textbook discounted reward-to-go: G_t = r_t + γ(1-d_t) G_{t+1}.  (B,T)."""
    G = torch.zeros_like(rewards)
    run = torch.zeros_like(rewards[:, 0])
    for t in range(rewards.shape[1] - 1, -1, -1):
        run = rewards[:, t] + gamma * (1.0 - dones[:, t]) * run
        G[:, t] = run
    return G


def gae_standard(rewards, values, dones, last_value, gamma, lam):
    """GAE(λ) (Schulman et al. 2016). rewards/values/dones (B,T), last_value (B,).
    δ_t = r_t + γ(1-d_t)V_{t+1} - V_t ; A_t = δ_t + γλ(1-d_t)A_{t+1} ; ret = A + V."""
    B, T = rewards.shape
    adv = torch.zeros_like(rewards)
    run = torch.zeros_like(last_value)
    nxt = last_value
    for t in range(T - 1, -1, -1):
        nd = 1.0 - dones[:, t]
        delta = rewards[:, t] + gamma * nd * nxt - values[:, t]
        run = delta + gamma * lam * nd * run
        adv[:, t] = run
        nxt = values[:, t]
    return adv, adv + values


def ppo_loss_standard(new_logp, old_logp, adv, value, ret, entropy, clip=0.2, vf_coef=0.5, ent_coef=0.01):
    """Clipped surrogate + value MSE − entropy bonus, all means over elements.
    new_logp/old_logp/adv/entropy share a shape; value/ret share a shape."""
    ratio = (new_logp - old_logp).exp()
    surr = torch.min(ratio * adv, torch.clamp(ratio, 1 - clip, 1 + clip) * adv)
    pl = -surr.mean()
    vl = F.mse_loss(value, ret)
    el = entropy.mean()
    return pl + vf_coef * vl - ent_coef * el, pl, vl, el


# ------------------------------------------------------------------------- DQN
def dqn_choose_action_compat(logits6, n_actions=N_ACTIONS):
    """Positions [0,-1,...,-(n_actions-1)] because ``-0 == 0``
    (IRL_dqn_train.py:257-258). logits6: (1,L,n_i) → (n_actions,6) int64."""
    tok = [y.argmax(dim=-1) for y in logits6]
    return torch.stack([torch.stack([t[0, -idx] for t in tok]) for idx in range(n_actions)])


def dqn_td_loss_compat(q_logits6, next_logits6, action, reward, done, gamma=0.95, n_actions=N_ACTIONS):
    """q_logits6/next_logits6: 6×(B,L,n_i); action (B,A,6) int64; reward, done (B,1).
    Q(s,a) is gathered with a (1,B,A) index, i.e. reads ``y[0, j, action[j,k,i]]``
    (batch element 0, sequence position = batch index; needs B <= L).  Target =
    r + γ(1-done)·topk_A(max_vocab Q'(s')), not detached.  Mean of 6 MSEs."""
    total = 0
    for i in range(6):
        q = q_logits6[i].gather(2, action[:, :, i].unsqueeze(0)).squeeze(0)
        nxt = next_logits6[i].max(2)[0].topk(n_actions, dim=1)[0]
        tgt = reward + gamma * (1 - done) * nxt
        total = total + F.mse_loss(q, tgt)
    return total / 6


def dqn_td_loss_standard(q_logits6, next_logits6, action, reward, done, gamma=0.95, n_actions=N_ACTIONS):
    """Clean variant: Q(s,a)[b,k] = y[b, pos_k, action[b,k,i]] with pos_k = -(k+1)
    (the position action k is read from in the clean read-out); target uses the
    same positions of the target net's max over vocabulary, detached."""
    total = 0
    for i in range(6):
        tail = q_logits6[i][:, -n_actions:, :].flip(1)
        q = tail.gather(2, action[:, :, i:i + 1])[..., 0]
        nxt = next_logits6[i][:, -n_actions:, :].flip(1).max(2)[0].detach()
        tgt = reward + gamma * (1 - done) * nxt
        total = total + F.mse_loss(q, tgt)
    return total / 6
