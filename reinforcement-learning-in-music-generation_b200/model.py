"""CP linear-transformer agent: the reference model surface on cpmusic kernels.

One class, ``CPLinearTransformer``, with the aliases the reference scripts import:

* ``TransformerModel``   — dqn_policy/agent_pretrain.py:213 (and cp-pretrain.py:213)
* ``LinearTransformer``  — dqn_policy/model.py:97
* ``Actor_Transformer``  — ppo_policy/model.py:98 (adds ``value_funtion``, drops ``project_concat_type``)
* ``Critic_Transformer`` — ppo_policy/model.py:285 (``value_produce``)

Same constructor ``(n_token, is_training=True)``, same method names and argument meaning
(``train_step``, ``forward_hidden``, ``forward_output`` in both arities, ``forward``,
``forward_output_sampling``, ``compute_loss``), same attribute / ``state_dict`` names
(217 keys for the 6-head DQN model).  ``inference`` is the batched,
device-resident replacement of ``inference_from_scratch`` (testing-no-type-cp.py:126).

``reference_compat`` selects the reference's quirks where they change numbers:
recurrent positional encoding always at position 0.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .encoder import (PackCache, RecurrentEncoderBuilder, TransformerEncoderBuilder, TriangularCausalMask,
                      cached_linear)

ATTRS6 = ("tempo", "chord", "barbeat", "pitch", "duration", "velocity")
ATTRS7 = ("tempo", "chord", "barbeat", "type", "pitch", "duration", "velocity")   # upstream CP layout
EMB6 = (128, 256, 64, 512, 128, 128)
EMB7 = (128, 256, 64, 32, 512, 128, 128)
# per-attribute (temperature, nucleus p) of forward_output_sampling (dqn_policy/model.py:282-287)
SAMPLING_CFG = {"tempo": (1.2, 0.9), "barbeat": (1.2, None), "chord": (1.0, 0.99), "pitch": (1.0, 0.9),
                "duration": (2.0, 0.9), "velocity": (5.0, None), "type": (1.0, 0.9)}


def sinusoidal_pe(max_len: int, d_model: int) -> torch.Tensor:
    pos = torch.arange(max_len, dtype=torch.float32)[:, None]
    freq = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * (-math.log(10000.0) / d_model))
    pe = torch.empty(max_len, d_model)
    pe[:, 0::2] = torch.sin(pos * freq)
    pe[:, 1::2] = torch.cos(pos * freq)
    return pe[None]


class Embeddings(nn.Module):
    """Holds ``lut.weight``; a lone call gathers through the fused kernel too."""

    def __init__(self, n_token, d_emb):
        super().__init__()
        self.lut = nn.Embedding(n_token, d_emb)
        self.d_model = d_emb

    def forward(self, x):
        return ops.cp_embed(x[..., None], [self.lut.weight], torch.float32)


class PositionalEncoding(nn.Module):
    """Keeps the reference's saved ``pe`` buffer (1,20000,d) and dropout rate."""

    def __init__(self, d_model, dropout=0.1, max_len=20000):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        self.register_buffer("pe", sinusoidal_pe(max_len, d_model))

    def forward(self, x, pos_offset: int = 0):
        p = self.dropout.p if self.training else 0.0
        return ops.add_pe(x, self.pe, x.size(1), pos_offset, None, p)


class CPLinearTransformer(nn.Module):
    variant = "dqn"

    def __init__(self, n_token: Sequence[int], is_training: bool = True, *, d_model: int = 512, n_layer: int = 12,
                 n_head: int = 8, d_inner: int = 2048, dropout: float = 0.1, compute_dtype=torch.bfloat16,
                 reference_compat: bool = True, return_fp32: bool = True, verbose: bool = False):
        super().__init__()
        self.n_token = [int(n) for n in n_token]        # taken verbatim, never inferred
        if len(self.n_token) == 6:
            self.attrs, self.emb_sizes = ATTRS6, list(EMB6)
        elif len(self.n_token) == 7:
            self.attrs, self.emb_sizes = ATTRS7, list(EMB7)
        else:
            raise ValueError("n_token must list 6 (reference) or 7 (upstream CP with 'type') vocabulary sizes")
        self.d_model, self.n_layer, self.n_head = d_model, n_layer, n_head
        self.d_head, self.d_inner, self.dropout = d_model // n_head, d_inner, dropout
        self.compute_dtype = compute_dtype
        self.reference_compat = reference_compat
        self.return_fp32 = return_fp32
        self.recurrent = not is_training
        self.loss_func = nn.CrossEntropyLoss(reduction="none")       # attribute kept for parity; unused
        if verbose:
            print("Token_class >>>>>:", self.n_token)
        for a, n, e in zip(self.attrs, self.n_token, self.emb_sizes):
            setattr(self, f"word_emb_{a}", Embeddings(n, e))
        self.pos_emb = PositionalEncoding(d_model, dropout)
        self.in_linear = nn.Linear(int(np.sum(self.emb_sizes)), d_model)
        builder = RecurrentEncoderBuilder if self.recurrent else TransformerEncoderBuilder
        self.transformer_encoder = builder.from_kwargs(
            n_layers=n_layer, n_heads=n_head, query_dimensions=d_model // n_head, value_dimensions=d_model // n_head,
            feed_forward_dimensions=d_inner, activation="gelu", dropout=dropout, attention_type="causal-linear",
            compute_dtype=compute_dtype).get()
        self._build_extras()
        for a, n in zip(self.attrs, self.n_token):
            setattr(self, f"proj_{a}", nn.Linear(d_model, n))
        self.seg = ops.seg_offsets(self.n_token)
        self.logits_width = -(-self.seg[-1] // 8) * 8                # row stride of the concatenated logits
        self._cache = PackCache()
        self._sample_step = 0

    def _build_extras(self):
        # allocated but never used in forward (dqn_policy/model.py:153); kept for state_dict parity
        self.project_concat_type = nn.Linear(self.d_model, self.d_model)

    # ------------------------------------------------------------------ internals (compute dtype)
    def set_compute_dtype(self, dtype):
        self.compute_dtype = dtype
        self.transformer_encoder.compute_dtype = dtype
        self._cache.clear()
        self.transformer_encoder._cache.clear()
        return self

    def _apply(self, fn, *a, **k):
        self._cache.clear()
        return super()._apply(fn, *a, **k)

    def refresh_packs(self):
        """Bring every cached bf16 weight packing up to date in place (needed before replaying a CUDA
        graph captured over them, e.g. the rollout step after an optimizer step)."""
        self._cache.refresh_all()
        self.transformer_encoder._cache.refresh_all()

    def invalidate_packs(self):
        """Force every cached packing to be rebuilt (in place) at its next use — see PackCache.invalidate."""
        self._cache.invalidate()
        self.transformer_encoder._cache.invalidate()

    def _tables(self):
        return [getattr(self, f"word_emb_{a}").lut.weight for a in self.attrs]

    def _heads(self):
        return [getattr(self, f"proj_{a}") for a in self.attrs]

    def _embed(self, x, pos_offset=0, pos_dev=None):
        """x (N,L,A) int64 -> (N,L,d) compute dtype: gather+scale+concat, in_linear, +PE, dropout."""
        if x.dtype != torch.int64:
            x = x.long()
        e = ops.cp_embed(x, self._tables(), self.compute_dtype)
        ops.IndexGuard.poll(x.device)       # out-of-range ids raise IndexError (one call late, without stalling the stream)
        z = cached_linear(self._cache, "in", [self.in_linear], e, self.compute_dtype)
        p = self.pos_emb.dropout.p if self.training else 0.0
        return ops.add_pe(z, self.pos_emb.pe, x.shape[-2] if x.dim() >= 2 else 1, pos_offset, pos_dev, p)

    def hidden(self, x, pos_offset=0):
        """Teacher-forced trunk in compute dtype: (N,L,A) -> (N,L,d)."""
        return self.transformer_encoder.forward_fused(self._embed(x, pos_offset))

    def logits_concat(self, h):
        """All heads as one GEMM: (…,d) -> (…, logits_width); attribute a = columns seg[a]:seg[a+1]."""
        return cached_linear(self._cache, "heads", self._heads(), h.to(self.compute_dtype), self.compute_dtype, pad_rows_to=8)

    def _split(self, lc):
        outs = tuple(lc[..., self.seg[i]:self.seg[i + 1]] for i in range(len(self.attrs)))
        return tuple(o.float() for o in outs) if self.return_fp32 else outs

    # ------------------------------------------------------------------ reference surface
    def forward_hidden(self, x, memory=None, is_training=True, pos_offset: Optional[int] = None):
        """Training: x (N,L,A) -> h (N,L,d).  Recurrent (``is_training=False``): x (1,1,A) [or (N,1,A)]
        -> (h (N,d), memory); the reference adds PE position 0 at every step — that is the
        default under ``reference_compat``; pass ``pos_offset`` for the true position."""
        if is_training:
            if self.recurrent:
                raise RuntimeError("model was built with is_training=False (recurrent encoder)")
            h = self.hidden(x, 0 if pos_offset is None else pos_offset)
            return h.float() if self.return_fp32 else h
        if not self.recurrent:
            raise RuntimeError("model was built with is_training=True (parallel encoder)")
        if x.dim() == 2:
            x = x[:, None, :]
        if pos_offset is None:
            pos_offset = 0 if self.reference_compat else self._infer_pos(memory)
        with torch.no_grad():
            z = self._embed(x, pos_offset).reshape(x.shape[0], self.d_model)
            h, memory = self.transformer_encoder.step_fused(z, memory)
        return (h.float() if self.return_fp32 else h), memory

    @staticmethod
    def _infer_pos(memory):
        return 0 if memory is None else int(getattr(memory, "pos", 0))

    def forward_output(self, h, y=None):
        """6 (or 7) logits tensors in attribute order; ``y`` is accepted and ignored exactly like the
        reference (dqn_policy/model.py:241-249; the PPO variant takes only ``h``)."""
        return self._split(self.logits_concat(h))

    def forward(self, x, target=None):
        return self.forward_output(self.forward_hidden(x), target)

    def compute_loss(self, predict, target, loss_mask):
        """Reference signature: predict (N, n_i, L) logits of ONE attribute, target (N,L), mask (N,L)."""
        lg = predict.permute(0, 2, 1).contiguous()
        if lg.dtype not in (torch.float32, torch.bfloat16):
            lg = lg.float()
        return ops.masked_ce(lg, target[..., None], loss_mask, [0, lg.shape[-1]])[0]

    def train_step(self, x, target, loss_mask, group=None):
        """Teacher-forced step: returns the per-attribute masked-mean CE losses (tuple of scalars,
        attribute order), one fused kernel for all heads.  ``group``: data-parallel process group
        whose ranks share the loss denominator (global sum of the mask)."""
        lc = self.logits_concat(self.hidden(x))
        if target.dtype != torch.int64:
            target = target.long()
        losses = ops.masked_ce(lc, target, loss_mask, self.seg, group)
        return tuple(losses[i] for i in range(len(self.attrs)))

    # ------------------------------------------------------------------ decoding
    def sampling_config(self, temperature=None, top_p=None):
        t = [SAMPLING_CFG[a][0] for a in self.attrs] if temperature is None else list(temperature)
        p = [SAMPLING_CFG[a][1] for a in self.attrs] if top_p is None else list(top_p)
        return t, p

    def decode(self, h, greedy=True, seed=0, seq_base=0, step=0, temperature=None, top_p=None, want_logp=False):
        """h (rows,d) -> tokens (rows,A) int64 [, logp (rows,A)] on device, no host sync."""
        lc = self.logits_concat(h.reshape(-1, h.shape[-1]))
        t, p = self.sampling_config(temperature, top_p)
        tok, lp, _ = ops.heads_sample(lc, self.seg, t, p, greedy=greedy, seed=seed, seq_base=seq_base, step=step,
                                      want_logp=want_logp)
        return (tok, lp) if want_logp else tok

    def forward_output_sampling(self, h, seed: Optional[int] = None):
        """Reference API (dqn_policy/model.py:259-298): h (1,d) -> np.array of A ints sampled with the
        reference's per-attribute temperature / nucleus settings — on the device (Philox) instead of
        host numpy; one D2H copy of A integers instead of six logits transfers."""
        seed = np.random.randint(0, 2 ** 31 - 1) if seed is None else seed
        tok = self.decode(h, greedy=False, seed=seed, step=self._sample_step)
        self._sample_step += 1
        return tok[0].cpu().numpy()

    def inference(self, init_tokens, n_steps, greedy=False, seed=0, true_positions=None, **kw):
        """Batched recurrent generation (replacement of ``inference_from_scratch``)."""
        from .rollout import RolloutEngine
        eng = RolloutEngine(self, batch=init_tokens.shape[0], max_steps=n_steps, greedy=greedy,
                            true_positions=(not self.reference_compat) if true_positions is None else true_positions, **kw)
        return eng.generate(init_tokens, n_steps, seed=seed)


class TransformerModel(CPLinearTransformer):
    pass


class LinearTransformer(CPLinearTransformer):
    pass


class Actor_Transformer(CPLinearTransformer):
    variant = "actor"

    def _build_extras(self):
        self.value_funtion = nn.Sequential(nn.Linear(self.d_model, 128), nn.ReLU(), nn.Linear(128, 1))

    def forward_output(self, h, y=None):
        return super().forward_output(h)


class Critic_Transformer(CPLinearTransformer):
    """Same trunk and heads + ``{attr}_value = Linear(n_i, 1)``; ``value_produce`` = mean over the
    sequence of the per-position value, averaged over attributes (ppo_policy/model.py:345-394)."""
    variant = "critic"

    def __init__(self, n_token, **kw):
        super().__init__(n_token, True, **kw)

    def _build_extras(self):
        for a, n in zip(self.attrs, self.n_token):
            setattr(self, f"{a}_value", nn.Linear(n, 1))

    def value_collapse(self):
        """The read-out is linear in the hidden state: mean_a(value_a(proj_a(h))) = h . u + c with
        u = mean_a(W_a^T w_a) (d,) and c = mean_a(w_a . b_a + beta_a).  Built with autograd through these tiny products, so
        the gradients reach proj_* and *_value exactly as through the six logits tensors - which are never formed."""
        u, c = 0.0, 0.0
        for a in self.attrs:
            proj, val = getattr(self, f"proj_{a}"), getattr(self, f"{a}_value")
            u = u + (val.weight[0][:, None] * proj.weight).sum(0)       # (n_a,) x (n_a, d): element-wise, no library GEMV
            c = c + (val.weight[0] * proj.bias).sum() + val.bias[0]
        return u / len(self.attrs), c / len(self.attrs)

    def value_per_position(self, x):
        """(N,L,A) -> (N,L) fp32: mean_a( Linear_a(logits_a) ), evaluated as one row dot of the hidden state (no logits GEMM)."""
        u, c = self.value_collapse()
        return ops.rowdot(self.hidden(x), u, c)

    def value_produce(self, x):
        return self.value_per_position(x).mean(dim=1, keepdim=True)
