"""Oracle restatement of pytorch-fast-transformers 0.4.0 (causal-linear encoder).

TEST INFRASTRUCTURE ONLY — see ``oracle/__init__.py``.  PARITY UNPINNED: the
real package is absent (reference ``requirements.txt:54``); this file restates
its published semantics as recorded in SURVEY.md Appendix A.1/A.2 and anchors on
the reference's call sites:

* ``TransformerEncoderBuilder.from_kwargs(n_layers=12, n_heads=8,
  query_dimensions=64, value_dimensions=64, feed_forward_dimensions=2048,
  activation='gelu', dropout=0.1, attention_type="causal-linear").get()``
  — reference ``dqn_policy/model.py:128-137``, ``dqn_policy/agent_pretrain.py:244-253``,
  ``ppo_policy/model.py:129-138,313-321``.
* ``RecurrentEncoderBuilder`` with the same kwargs — ``dqn_policy/model.py:141-150``.
* ``TriangularCausalMask(L, device=)`` — ``dqn_policy/model.py:231``.

Everything is plain PyTorch and dtype-generic, so the same code yields the fp32
"reference CPU path" and the fp64 ground truth used to set tolerances.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

EPS = 1e-6  # CausalLinearAttention / RecurrentLinearAttention eps (App. A.1/A.2)


# --------------------------------------------------------------------------- #
# feature map + causal product, three equivalent forms
# --------------------------------------------------------------------------- #
def feature_map(x: torch.Tensor) -> torch.Tensor:
    """elu(x)+1, the default ft feature map for linear attention (App. A.1)."""
    return F.elu(x) + 1.0


def causal_dot_product_quadratic(Q, K, V):
    """``out[n,h,l] = sum_{j<=l} (Q[n,h,l].K[n,h,j]) V[n,h,j]`` via a masked L×L
    score matrix.  Q,K: (N,H,L,E)  V: (N,H,L,M).  (ft ``causal_product`` fwd,
    SURVEY §2.3 K1.)"""
    L = Q.shape[2]
    scores = torch.matmul(Q, K.transpose(-1, -2))
    mask = torch.ones(L, L, dtype=torch.bool, device=Q.device).tril()
    scores = scores.masked_fill(~mask, 0.0)
    return torch.matmul(scores, V)


def causal_dot_product_scan(Q, K, V):
    """Same product as a literal left-to-right KV-state scan (what ft's
    ``causal_product_cpu`` does per (n,h)): ``kv += k⊗v ; out = qᵀ kv``."""
    N, H, L, E = Q.shape
    M = V.shape[-1]
    kv = torch.zeros(N, H, E, M, dtype=Q.dtype, device=Q.device)
    out = torch.empty(N, H, L, M, dtype=Q.dtype, device=Q.device)
    for l in range(L):
        kv = kv + K[:, :, l, :, None] * V[:, :, l, None, :]
        out[:, :, l] = torch.einsum("nhe,nhem->nhm", Q[:, :, l], kv)
    return out


def causal_dot_product_backward_scan(Q, K, V, G):
    """ft ``causal_product`` backward: forward scan for gQ,
    reverse scan for gK and gV."""
    N, H, L, E = Q.shape
    M = V.shape[-1]
    gQ = torch.zeros_like(Q)
    gK = torch.zeros_like(K)
    gV = torch.zeros_like(V)
    kv = torch.zeros(N, H, E, M, dtype=Q.dtype, device=Q.device)
    for l in range(L):
        kv = kv + K[:, :, l, :, None] * V[:, :, l, None, :]
        gQ[:, :, l] = torch.einsum("nhem,nhm->nhe", kv, G[:, :, l])
    r = torch.zeros(N, H, E, M, dtype=Q.dtype, device=Q.device)
    for l in range(L - 1, -1, -1):
        r = r + Q[:, :, l, :, None] * G[:, :, l, None, :]
        gK[:, :, l] = torch.einsum("nhem,nhm->nhe", r, V[:, :, l])
        gV[:, :, l] = torch.einsum("nhem,nhe->nhm", r, K[:, :, l])
    return gQ, gK, gV


def causal_linear_attention(q, k, v, key_lengths_mask=None, eps: float = EPS,
                            product=causal_dot_product_quadratic):
    """ft ``CausalLinearAttention.forward`` (App. A.1).  q,k: (N,L,H,E) raw
    projections (feature map applied here), v: (N,L,H,M) → (N,L,H,M)."""
    Q = feature_map(q)
    K = feature_map(k)
    if key_lengths_mask is not None:          # K * k_len.float_matrix[:, :, None, None]
        K = K * key_lengths_mask[:, :, None, None].to(K.dtype)
    Z = 1.0 / (torch.einsum("nlhi,nlhi->nlh", Q, K.cumsum(1)) + eps)
    out = product(Q.permute(0, 2, 1, 3).contiguous(),
                  K.permute(0, 2, 1, 3).contiguous(),
                  v.permute(0, 2, 1, 3).contiguous()).permute(0, 2, 1, 3)
    return out * Z[..., None]


def recurrent_linear_attention(q, k, v, state=None, eps: float = EPS):
    """ft ``RecurrentLinearAttention.forward`` (App. A.2). q,k: (N,H,E) v: (N,H,M);
    state = [Si (N,H,E,M), Zi (N,H,E)] updated and returned."""
    Q = feature_map(q)
    K = feature_map(k)
    N, H, E = Q.shape
    M = v.shape[-1]
    if state is None:
        Si = torch.zeros(N, H, E, M, dtype=Q.dtype, device=Q.device)
        Zi = torch.zeros(N, H, E, dtype=Q.dtype, device=Q.device)
    else:
        Si, Zi = state
    if len(Si) != N:
        raise ValueError("The batch size changed during iteration")
    Zi = Zi + K
    Si = Si + torch.einsum("nhd,nhm->nhdm", K, v)
    Z = 1.0 / (torch.einsum("nhd,nhd->nh", Q, Zi) + eps)
    out = torch.einsum("nhd,nhdm,nh->nhm", Q, Si, Z)
    return out, [Si, Zi]


# --------------------------------------------------------------------------- #
# masks (only the flag the reference uses, plus ft's LengthMask for the key-padding path the encoder signature carries)
class LengthMask:
    """``fast_transformers.masking.LengthMask`` restated: ``float_matrix[n, l] = 1 if l < lengths[n] else 0``."""

    def __init__(self, lengths, max_len=None, device=None):
        self.lengths = torch.as_tensor(lengths).long()
        self.max_len = int(max_len) if max_len is not None else int(self.lengths.max())
        self.lower_triangular = False

    @property
    def bool_matrix(self):
        return torch.arange(self.max_len)[None, :] < self.lengths[:, None]

    @property
    def float_matrix(self):
        return self.bool_matrix.float()



# --------------------------------------------------------------------------- #
class TriangularCausalMask:
    """``fast_transformers.masking.TriangularCausalMask`` restated: carries the
    ``lower_triangular`` flag the causal attention checks (App. A.1)."""

    def __init__(self, N, device="cpu"):
        self.N = N
        self.device = device
        self.lower_triangular = True


class FullMask:
    def __init__(self, N, device="cpu"):
        self.N = N
        self.device = device
        self.lower_triangular = False


# --------------------------------------------------------------------------- #
# encoder modules with ft's parameter names
# --------------------------------------------------------------------------- #
class AttentionLayer(nn.Module):
    def __init__(self, d_model, n_heads, d_keys=None, d_values=None):
        super().__init__()
        d_keys = d_keys or d_model // n_heads
        d_values = d_values or d_model // n_heads
        self.n_heads = n_heads
        self.query_projection = nn.Linear(d_model, d_keys * n_heads)
        self.key_projection = nn.Linear(d_model, d_keys * n_heads)
        self.value_projection = nn.Linear(d_model, d_values * n_heads)
        self.out_projection = nn.Linear(d_values * n_heads, d_model)

    def forward(self, x, attn_mask, product=causal_dot_product_quadratic, key_lengths_mask=None):
        if not getattr(attn_mask, "lower_triangular", False):
            raise RuntimeError("CausalLinearAttention only supports full lower triangular masks")
        N, L, _ = x.shape
        H = self.n_heads
        q = self.query_projection(x).view(N, L, H, -1)
        k = self.key_projection(x).view(N, L, H, -1)
        v = self.value_projection(x).view(N, L, H, -1)
        a = causal_linear_attention(q, k, v, key_lengths_mask=key_lengths_mask, product=product).reshape(N, L, -1)
        return self.out_projection(a)

    def step(self, x, state):
        N = x.shape[0]
        H = self.n_heads
        q = self.query_projection(x).view(N, H, -1)
        k = self.key_projection(x).view(N, H, -1)
        v = self.value_projection(x).view(N, H, -1)
        a, state = recurrent_linear_attention(q, k, v, state)
        return self.out_projection(a.reshape(N, -1)), state


class TransformerEncoderLayer(nn.Module):
    """POST-norm layer, exact-erf GELU, dropout after attention / activation /
    linear2 (App. A.1)."""

    def __init__(self, d_model, n_heads, d_ff, dropout=0.1):
        super().__init__()
        self.attention = AttentionLayer(d_model, n_heads)
        self.linear1 = nn.Linear(d_model, d_ff)
        self.linear2 = nn.Linear(d_ff, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)

    def _tail(self, x, a):
        x = x + self.dropout(a)
        y = x = self.norm1(x)
        y = self.dropout(F.gelu(self.linear1(y)))
        y = self.dropout(self.linear2(y))
        return self.norm2(x + y)

    def forward(self, x, attn_mask, product=causal_dot_product_quadratic, key_lengths_mask=None):
        return self._tail(x, self.attention(x, attn_mask, product, key_lengths_mask))

    def step(self, x, state):
        a, state = self.attention.step(x, state)
        return self._tail(x, a), state


class TransformerEncoder(nn.Module):
    """``TransformerEncoderBuilder(...).get()`` product: N post-norm layers then a
    final LayerNorm (``final_normalization=True``)."""

    def __init__(self, n_layers=12, n_heads=8, query_dimensions=64, value_dimensions=64,
                 feed_forward_dimensions=2048, dropout=0.1):
        super().__init__()
        d_model = value_dimensions * n_heads
        self.layers = nn.ModuleList([
            TransformerEncoderLayer(d_model, n_heads, feed_forward_dimensions, dropout)
            for _ in range(n_layers)])
        self.norm = nn.LayerNorm(d_model)
        self.product = causal_dot_product_quadratic

    def forward(self, x, attn_mask=None, length_mask=None):
        attn_mask = attn_mask or FullMask(x.shape[1])
        klm = None                                   # ft: K = K * k_len.float_matrix[:, :, None, None] in every layer (App. A.1)
        if length_mask is not None:
            klm = length_mask.float_matrix if hasattr(length_mask, "float_matrix") else torch.as_tensor(length_mask).float()
        for layer in self.layers:
            x = layer(x, attn_mask, self.product, klm)
        return self.norm(x)


class RecurrentTransformerEncoder(TransformerEncoder):
    """``RecurrentEncoderBuilder(...).get()`` product (App. A.2): same parameters,
    one-token step, ``memory`` is the deprecated alias of ``state``."""

    def forward(self, x, state=None, memory=None):
        state = state if state is not None else memory
        if state is None:
            state = [None] * len(self.layers)
        for i, layer in enumerate(self.layers):
            x, s = layer.step(x, state[i])
            state[i] = s
        return self.norm(x), state


class _Builder:
    _cls = TransformerEncoder

    def __init__(self, **kw):
        self.kw = kw

    @classmethod
    def from_kwargs(cls, **kw):
        return cls(**kw)

    def get(self):
        kw = dict(self.kw)
        if kw.pop("attention_type", "causal-linear") != "causal-linear":
            raise ValueError("oracle restates only attention_type='causal-linear'")
        if kw.pop("activation", "gelu") != "gelu":
            raise ValueError("oracle restates only activation='gelu'")
        return self._cls(**kw)


class TransformerEncoderBuilder(_Builder):
    _cls = TransformerEncoder


class RecurrentEncoderBuilder(_Builder):
    _cls = RecurrentTransformerEncoder


def sinusoidal_pe(max_len: int, d_model: int) -> torch.Tensor:
    """The ``pe`` buffer of the reference ``PositionalEncoding``
    (``dqn_policy/model.py:82-88``): interleaved sin/cos, shape (1,max_len,d)."""
    pos = torch.arange(max_len, dtype=torch.float32)[:, None]
    freq = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * (-math.log(10000.0) / d_model))
    pe = torch.empty(max_len, d_model)
    pe[:, 0::2] = torch.sin(pos * freq)
    pe[:, 1::2] = torch.cos(pos * freq)
    return pe[None]
