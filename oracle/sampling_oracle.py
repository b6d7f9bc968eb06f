"""Oracle restatement of the reference's host-side numpy samplers.

TEST INFRASTRUCTURE ONLY — see ``oracle/__init__.py``.  PINNED against the executed
reference (tests/golden/make_ref_golden.py -> tests/test_ref_golden.py).

Follows ``dqn_policy/model.py:19-55`` (identical twins at
``dqn_policy/agent_pretrain.py:136-172`` and ``ppo_policy/model.py``):
``softmax_with_temperature`` → ``nucleus`` (if p given) | ``weighted_sampling``.
The per-attribute (t, p) settings and the call/return orders are those of
``forward_output_sampling`` (``dqn_policy/model.py:282-297``).

The reference draws with the global ``np.random`` stream, which a device sampler
cannot reproduce bit for bit.  ``np.random.choice(a, p=p)`` is *defined* as
``a[searchsorted(cumsum(p)/cumsum(p)[-1], u, side='right')]`` with
``u ~ U[0,1)``; ``choice_from_uniform`` restates exactly that, so a device
sampler fed the same ``u`` must return the same index.  The uniforms come from
the counter-based Philox4x32-10 generator restated below (Salmon et al. 2011,
Random123 known-answer vectors are checked in the tests).
"""
from __future__ import annotations

import numpy as np

ATTRS = ("tempo", "chord", "barbeat", "pitch", "duration", "velocity")
# (temperature, nucleus p) per attribute, dqn_policy/model.py:282-287
SAMPLING_CFG = {
    "tempo": (1.2, 0.9), "barbeat": (1.2, None), "chord": (1.0, 0.99),
    "pitch": (1.0, 0.9), "duration": (2.0, 0.9), "velocity": (5.0, None),
}
# the reference *calls* the samplers in this order (matters for the np.random stream)
CALL_ORDER = ("tempo", "barbeat", "chord", "pitch", "duration", "velocity")


# ------------------------------------------------------------------ literal path
def softmax_with_temperature(logits, temperature):
    e = np.exp(logits / temperature)
    return e / np.sum(e)


def weighted_sampling(probs, rng=np.random):
    probs = probs / sum(probs)
    order = np.argsort(probs)[::-1]
    return rng.choice(order, size=1, p=probs[order])[0]


def nucleus_candidates(probs, p):
    """Returns (candidate indices, renormalised candidate probs): the sorted prefix
    up to AND INCLUDING the first index whose cumulative mass exceeds ``p``
    (dqn_policy/model.py:31-43)."""
    probs = probs / (sum(probs) + 1e-5)
    order = np.argsort(probs)[::-1]
    csum = np.cumsum(probs[order])
    over = csum > p
    last = (np.where(over)[0][0] + 1) if over.sum() > 0 else len(order)
    cand = order[:last]
    cp = np.asarray([probs[i] for i in cand])
    cp = cp / sum(cp)
    return cand, cp


def nucleus(probs, p, rng=np.random):
    cand, cp = nucleus_candidates(probs, p)
    return rng.choice(cand, size=1, p=cp)[0]


def sampling(logit, p=None, t=1.0, rng=np.random):
    """dqn_policy/model.py:48-55 with an injectable RandomState."""
    logit = np.asarray(logit, dtype=np.float32).squeeze()
    probs = softmax_with_temperature(logit, t)
    return nucleus(probs, p, rng) if p is not None else weighted_sampling(probs, rng)


def forward_output_sampling(logits6, rng=np.random):
    """Restates the sampling half of ``forward_output_sampling``: ``logits6`` maps
    attribute → 1-D logits; returns int array in ATTRS order, drawing in CALL_ORDER."""
    words = {}
    for a in CALL_ORDER:
        t, p = SAMPLING_CFG[a]
        words[a] = sampling(logits6[a], p=p, t=t, rng=rng)
    return np.array([words[a] for a in ATTRS])


# ------------------------------------------------------------ explicit-uniform path
def choice_from_uniform(candidates, probs, u):
    """What ``np.random.choice(candidates, p=probs)`` returns when its internal
    uniform draw equals ``u``."""
    cdf = np.cumsum(np.asarray(probs, dtype=np.float64))
    cdf /= cdf[-1]
    return candidates[int(np.searchsorted(cdf, u, side="right"))]


def sampling_from_uniform(logit, u, p=None, t=1.0):
    logit = np.asarray(logit, dtype=np.float32).squeeze()
    probs = softmax_with_temperature(logit, t)
    if p is not None:
        cand, cp = nucleus_candidates(probs, p)
    else:
        probs = probs / sum(probs)
        cand = np.argsort(probs)[::-1]
        cp = probs[cand]
    return int(choice_from_uniform(cand, cp, u))


def greedy(logit):
    """argmax — the decoding the reference RL loops use (ppo_train.py:266-267,
    IRL_dqn_train.py:249-250); first maximal index like torch/numpy."""
    return int(np.argmax(np.asarray(logit)))


# ------------------------------------------------------------------------- Philox
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """Philox4x32-10 block function. counter: 4 uint32, key: 2 uint32 → 4 uint32."""
    c = [np.uint64(int(x) & 0xFFFFFFFF) for x in counter]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c[0]
        p1 = _M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return [int(x) for x in c]


def philox_uniform(seed: int, seq_id: int, step: int, attr: int) -> float:
    """The uniform the device sampler uses for (sequence, step, attribute):
    counter = (seq_id, step, attr, 0), key = (seed lo, seed hi); u = top 24 bits / 2^24."""
    r = philox4x32_10((seq_id, step, attr, 0), (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    return (r[0] >> 8) * (1.0 / 16777216.0)
