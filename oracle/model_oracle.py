"""Oracle restatement of the reference CP agent models (CPU, plain PyTorch).

TEST INFRASTRUCTURE ONLY — see ``oracle/__init__.py``.  PINNED against the executed reference for everything in this file
(tests/test_ref_golden.py); the encoder it calls (ft_oracle.py) stays unpinned.

Follows, without copying, the reference modules:
* ``Embeddings`` / ``PositionalEncoding`` — ``dqn_policy/agent_pretrain.py:185-210``
  (= ``dqn_policy/model.py:67-92``).
* ``TransformerModel`` — ``dqn_policy/agent_pretrain.py:213-375``;
  ``LinearTransformer`` — ``dqn_policy/model.py:97-255``;
  ``Actor_Transformer`` — ``ppo_policy/model.py:98-280``;
  ``Critic_Transformer`` — ``ppo_policy/model.py:285-394``.
The state_dict key set equals SURVEY App. A.3.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ft_oracle as ft

ATTRS = ("tempo", "chord", "barbeat", "pitch", "duration", "velocity")
EMB_SIZES = (128, 256, 64, 512, 128, 128)
# seven-attribute layout of the data files (`type` at column 3, agent_pretrain.py:525-526 drops it before the model): the same
# restatement with one more independent embedding / head pair — the reference has no type-conditioned head
# (project_concat_type is allocated and never used, dqn_policy/model.py:153), so neither does this.
ATTRS7 = ("tempo", "chord", "barbeat", "type", "pitch", "duration", "velocity")
EMB_SIZES7 = (128, 256, 64, 32, 512, 128, 128)


class Embeddings(nn.Module):
    """``lut[idx] * sqrt(emb)`` (agent_pretrain.py:185-192)."""

    def __init__(self, n_token, d_emb):
        super().__init__()
        self.lut = nn.Embedding(n_token, d_emb)
        self.d_model = d_emb

    def forward(self, idx):
        return self.lut(idx) * math.sqrt(self.d_model)


class PositionalEncoding(nn.Module):
    """``x + pe[:, :L]`` then dropout; ``pe`` is a saved buffer (1,20000,d)
    (agent_pretrain.py:195-210).  With a one-token input this always adds
    position 0 — the reference's recurrent quirk."""

    def __init__(self, d_model, dropout=0.1, max_len=20000):
        super().__init__()
        self.dropout = nn.Dropout(dropout)
        self.register_buffer("pe", ft.sinusoidal_pe(max_len, d_model))

    def forward(self, x, pos_offset: int = 0):
        L = x.size(1)
        return self.dropout(x + self.pe[:, pos_offset:pos_offset + L].to(x.dtype))


class OracleCPModel(nn.Module):
    """TransformerModel / LinearTransformer / Actor_Transformer restated.

    ``variant``: 'dqn' (has the unused ``project_concat_type``; forward_output(h, y))
    or 'actor' (has ``value_funtion``; forward_output(h))."""

    def __init__(self, n_token, is_training=True, variant="dqn", d_model=512, n_layer=12,
                 n_head=8, d_inner=2048, dropout=0.1):
        super().__init__()
        self.n_token = list(n_token)
        self.attrs, self.emb_sizes = (ATTRS7, EMB_SIZES7) if len(self.n_token) == 7 else (ATTRS, EMB_SIZES)
        self.d_model, self.n_layer, self.n_head = d_model, n_layer, n_head
        self.variant = variant
        self.loss_func = nn.CrossEntropyLoss(reduction="none")
        for name, n, e in zip(self.attrs, self.n_token, self.emb_sizes):
            setattr(self, f"word_emb_{name}", Embeddings(n, e))
        self.pos_emb = PositionalEncoding(d_model, dropout)
        self.in_linear = nn.Linear(sum(self.emb_sizes), d_model)
        builder = ft.TransformerEncoderBuilder if is_training else ft.RecurrentEncoderBuilder
        self.transformer_encoder = builder.from_kwargs(
            n_layers=n_layer, n_heads=n_head, query_dimensions=d_model // n_head,
            value_dimensions=d_model // n_head, feed_forward_dimensions=d_inner,
            activation="gelu", dropout=dropout, attention_type="causal-linear").get()
        if variant == "dqn":
            self.project_concat_type = nn.Linear(d_model, d_model)   # allocated, unused
        else:
            self.value_funtion = nn.Sequential(nn.Linear(d_model, 128), nn.ReLU(), nn.Linear(128, 1))
        for name, n in zip(self.attrs, self.n_token):
            setattr(self, f"proj_{name}", nn.Linear(d_model, n))

    # -- pieces ------------------------------------------------------------- #
    def embed(self, x):
        embs = [getattr(self, f"word_emb_{a}")(x[..., i]) for i, a in enumerate(self.attrs)]
        return self.in_linear(torch.cat(embs, dim=-1))

    def forward_hidden(self, x, memory=None, is_training=True, pos_offset: int = 0):
        z = self.pos_emb(self.embed(x), pos_offset)
        if is_training:
            mask = ft.TriangularCausalMask(z.size(1), device=x.device)
            return self.transformer_encoder(z, mask)
        z = z.squeeze(0)
        return self.transformer_encoder(z, memory=memory)

    def forward_output(self, h, y=None):
        return tuple(getattr(self, f"proj_{a}")(h) for a in self.attrs)

    def forward(self, x, target=None):
        return self.forward_output(self.forward_hidden(x), target)

    def compute_loss(self, predict, target, loss_mask):
        loss = self.loss_func(predict, target) * loss_mask
        return torch.sum(loss) / torch.sum(loss_mask)

    def train_step(self, x, target, loss_mask):
        logits = self.forward_output(self.forward_hidden(x), target)
        return tuple(self.compute_loss(lg.permute(0, 2, 1), target[..., i], loss_mask)
                     for i, lg in enumerate(logits))


class OracleCritic(nn.Module):
    """Critic_Transformer.value_produce restated (ppo_policy/model.py:285-394):
    encoder → 6 heads → 6 Linear(n_i,1) → mean over sequence → average of 6."""

    def __init__(self, n_token, d_model=512, n_layer=12, n_head=8, d_inner=2048, dropout=0.1):
        super().__init__()
        self.n_token = list(n_token)
        for name, n, e in zip(ATTRS, self.n_token, EMB_SIZES):
            setattr(self, f"word_emb_{name}", Embeddings(n, e))
        self.pos_emb = PositionalEncoding(d_model, dropout)
        self.in_linear = nn.Linear(sum(EMB_SIZES), d_model)
        self.transformer_encoder = ft.TransformerEncoderBuilder.from_kwargs(
            n_layers=n_layer, n_heads=n_head, query_dimensions=d_model // n_head,
            value_dimensions=d_model // n_head, feed_forward_dimensions=d_inner,
            activation="gelu", dropout=dropout, attention_type="causal-linear").get()
        for name, n in zip(ATTRS, self.n_token):
            setattr(self, f"proj_{name}", nn.Linear(d_model, n))
        for name, n in zip(ATTRS, self.n_token):
            setattr(self, f"{name}_value", nn.Linear(n, 1))

    def value_produce(self, x):
        embs = [getattr(self, f"word_emb_{a}")(x[..., i]) for i, a in enumerate(ATTRS)]
        z = self.pos_emb(self.in_linear(torch.cat(embs, dim=-1)))
        h = self.transformer_encoder(z, ft.TriangularCausalMask(z.size(1)))
        total = 0
        for a in ATTRS:
            y = getattr(self, f"proj_{a}")(h)
            total = total + getattr(self, f"{a}_value")(y).mean(dim=1)
        return total / len(ATTRS)
