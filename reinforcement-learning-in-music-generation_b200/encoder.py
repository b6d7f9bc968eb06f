"""fast_transformers-shaped causal-linear encoder running on the cpmusic kernels.

Mirrors the plugin interface the reference builds through
``TransformerEncoderBuilder.from_kwargs(...).get()`` / ``RecurrentEncoderBuilder``
(reference dqn_policy/model.py:128-150, agent_pretrain.py:244-266,
ppo_policy/model.py:129-151,313-321): same constructor kwargs, same submodule and
parameter names (``layers.{i}.attention.{query,key,value,out}_projection``, ``linear1``,
``linear2``, ``norm1``, ``norm2``, final ``norm`` — SURVEY App. A.3), same call signatures
(``forward(x, attn_mask, length_mask)`` and ``forward(x, state=None, memory=None) -> (y, state)``)
and the same error behaviour (RuntimeError for a non-causal mask, ValueError when the batch size
changes between recurrent steps).

Per layer (post-norm, exact-erf GELU):  fused QKV GEMM → chunked causal linear attention kernel
→ out-proj GEMM → fused residual+dropout+LayerNorm → FFN1 GEMM → fused GELU+dropout → FFN2 GEMM →
fused residual+dropout+LayerNorm.  fp32 master parameters, bf16 (or fp32) compute copies cached
and re-packed when a parameter changes.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from . import ops


import os as _os
# Bias folding (A/B switch for tools/, default off): "all" runs out-projection, linear1 and linear2 bias-less, adds their
# biases inside the LayerNorm / GELU kernels and takes the bias gradients as by-products of those kernels' backward (no
# column-sum passes over the gradient tensors); "gelu": linear1 only.  Measured on B200 (update phase of bench.py):
# off 293-296 ms, gelu 296 ms, all 289-293 ms - the saved reductions (84 us for the (T, 2048) gradient, 21 us for the 512-wide
# ones) are paid back by the heavier kernels (GELU backward 183 -> 309 us with the fused column sums), so the plain path stays.
FUSED_BIAS_GRADS = {"0": False, "1": "all"}.get(_os.environ.get("CPM_FUSED_BIAS_GRADS", "0"), _os.environ.get("CPM_FUSED_BIAS_GRADS", "0"))


# GELU (+ dropout) inside the GEMM epilogues (CPM_FUSED_GELU=1; default off).  Measured on B200 at T = 131072 tokens
# (profiles/r02_gemm_microbench_v2_modes.jsonl): linear1 with the fused epilogue 865 us against 462 us for the same GEMM with a
# bias epilogue plus 279 us for the stand-alone GELU kernel; linear2's data gradient with the fused GELU backward 894 us against
# 468 + 350.  ~40 instructions per element (erf + Philox) issued by the 8 epilogue warps of a CTA cannot keep up with what
# 64 resident warps per SM do in the stand-alone kernels; the epilogue, not the tensor pipe, then sets the tile time.
FUSED_GELU_EPILOGUE = _os.environ.get("CPM_FUSED_GELU", "0") == "1"


class TriangularCausalMask:
    """``fast_transformers.masking.TriangularCausalMask(N, device=)``: only the
    ``lower_triangular`` flag is consulted by causal-linear attention."""

    def __init__(self, N, device="cpu"):
        self.N = N
        self.device = device
        self.lower_triangular = True


class FullMask:
    def __init__(self, N=None, M=None, device="cpu"):
        self.N = N
        self.device = device
        self.lower_triangular = False


class LengthMask:
    """``fast_transformers.masking.LengthMask``: sequences of ``lengths[n]`` valid tokens padded to ``max_len``."""

    def __init__(self, lengths, max_len=None, device=None):
        lengths = torch.as_tensor(lengths, device=device).long()
        self.lengths = lengths
        self.max_len = int(max_len) if max_len is not None else int(lengths.max())
        self.lower_triangular = False

    @property
    def bool_matrix(self):
        return torch.arange(self.max_len, device=self.lengths.device)[None, :] < self.lengths[:, None]

    @property
    def float_matrix(self):
        return self.bool_matrix.float()


# torch's fused optimizers (``Adam(fused=True)``, with or without ``capturable``) update parameters WITHOUT bumping
# ``Tensor._version`` (measured: version unchanged across ``step()``), so the version counters alone cannot tell a stale
# packing.  Every optimizer step therefore advances an epoch kept PER PARAMETER (only the parameters the stepping optimizer
# owns: a frozen target network, or the critic while the actor steps, keeps its packings), and the epochs are part of the stamp.
# Parameters are keyed by their storage address, which is what a packing was filled from.
_PARAM_EPOCH = {}


def _optimizer_stepped(optimizer, *_args, **_kwargs):
    for group in getattr(optimizer, "param_groups", ()):
        for p in group.get("params", ()):
            key = p.data_ptr()
            _PARAM_EPOCH[key] = _PARAM_EPOCH.get(key, 0) + 1


from torch.optim.optimizer import register_optimizer_step_post_hook as _register_post_hook  # noqa: E402
_register_post_hook(_optimizer_stepped)


class PackCache:
    """Compute-dtype packings of fp32 master parameters (row-concatenated, optionally zero-padded).
    A packing is refreshed IN PLACE when a master changes (in-place write seen through ``_version``, or any optimizer
    step since), so its device address is stable — CUDA graphs captured over these buffers (the rollout step) stay
    valid across optimizer steps; call ``refresh_all()`` before replaying such a graph."""

    def __init__(self):
        self._store = {}
        self._table, self._table_key = None, None          # ops.PackTable over every bf16 packing of this cache (one-launch refresh)

    @staticmethod
    def _stamp(params, dtype):
        return tuple((p.data_ptr(), p._version, _PARAM_EPOCH.get(p.data_ptr(), 0)) for p in params) + (dtype,)

    @staticmethod
    def _fill(wc, bc, ws, bs, wt=None, b32=None):
        r0 = 0
        for w, b in zip(ws, bs):
            wc[r0:r0 + w.shape[0]].copy_(w)
            bc[r0:r0 + w.shape[0]].copy_(b)
            if b32 is not None:
                b32[r0:r0 + w.shape[0]].copy_(b)
            r0 += w.shape[0]
        if wt is not None:
            wt.copy_(wc.t())

    # ---- one-launch refresh: after an optimizer step every packing of a model is stale at once
    def _batched_entries(self):
        out = []
        for key, (stamp, packed) in self._store.items():
            wc, bc, rows, masters, wt, b32 = packed
            n = len(rows)
            if wc.dtype != torch.bfloat16 or not wc.is_cuda or wt is None or b32 is None:
                return None
            r0 = 0
            for w, b in zip(masters[:n], masters[n:]):
                if w.dtype != torch.float32 or not w.is_contiguous() or b is None or b.dtype != torch.float32 or not b.is_contiguous():
                    return None
                out.append((w, b, wc, wt, bc, b32, r0))
                r0 += w.shape[0]
        return out

    def _refresh_batched(self) -> bool:
        """Refreshes EVERY packing of this cache from its masters with one kernel launch (cpm_pack_weights) and re-stamps them.
        False when the cache holds something the kernel does not cover (fp32 parity packings, CPU tensors) or the item table
        would have to be rebuilt during a CUDA-graph capture."""
        if not self._store:
            return False
        key = tuple((k, p[1][0].data_ptr(), p[1][4].data_ptr() if p[1][4] is not None else 0, tuple(m.data_ptr() for m in p[1][3]))
                    for k, p in self._store.items())
        if key != self._table_key:
            if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
                return False
            entries = self._batched_entries()
            if entries is None:
                self._table, self._table_key = None, key
                return False
            self._table, self._table_key = ops.PackTable(entries, entries[0][2].device), key
        if self._table is None:
            return False
        with torch.no_grad():
            self._table.launch()
        for k, (stamp, packed) in list(self._store.items()):
            self._store[k] = (self._stamp(packed[3], stamp[-1]), packed)
        return True

    def get(self, key, linears, dtype, pad_rows_to: int = 1):
        ws = [l.weight for l in linears]
        bs = [l.bias for l in linears]
        stamp = self._stamp(ws + bs, dtype)
        hit = self._store.get(key)
        if hit is not None and hit[0] == stamp:
            return hit[1]
        if (hit is not None and dtype == torch.bfloat16 and hit[1][0].dtype == dtype and len(hit[1][3]) == len(ws + bs)
                and all(m is p for m, p in zip(hit[1][3], ws + bs)) and self._refresh_batched()):
            return self._store[key][1]
        with torch.no_grad():
            rows = [int(w.shape[0]) for w in ws]
            padded = -(-sum(rows) // pad_rows_to) * pad_rows_to
            if (hit is not None and hit[1][0].dtype == dtype and hit[1][0].device == ws[0].device
                    and hit[1][0].shape == (padded, ws[0].shape[1])):
                wc, bc, wt, b32 = hit[1][0], hit[1][1], hit[1][4], hit[1][5]       # same buffers, new values
            else:
                wc = torch.zeros(padded, ws[0].shape[1], dtype=dtype, device=ws[0].device)
                bc = torch.zeros(padded, dtype=dtype, device=ws[0].device)
                # operands of the own GEMMs: the transposed copy feeds the data-gradient GEMM (dX = dY . W as an "NT" product
                # against W^T), the fp32 bias its epilogue
                wt = torch.zeros(ws[0].shape[1], padded, dtype=dtype, device=ws[0].device) if dtype == torch.bfloat16 else None
                b32 = torch.zeros(padded, dtype=torch.float32, device=ws[0].device) if dtype == torch.bfloat16 else None
            self._fill(wc, bc, ws, bs, wt, b32)
        packed = (wc, bc, tuple(rows), tuple(ws + bs), wt, b32)
        self._store[key] = (stamp, packed)
        return packed

    def get_gemm_pack(self, key, linears, dtype, pad_rows_to: int = 1):
        """(wc (N,K), wt (K,N), b32 (N,), rows), masters  -  the operand set of ops.tc_linear_packed / ops.tc_ffn."""
        wc, _bc, rows, masters, wt, b32 = self.get(key, linears, dtype, pad_rows_to)
        return (wc, wt, b32, rows), masters

    def refresh_all(self):
        """Re-sync every existing packing with its masters (in place)."""
        if any(self._stamp(packed[3], stamp[-1]) != stamp for stamp, packed in self._store.values()) and self._refresh_batched():
            return
        with torch.no_grad():
            for key, (stamp, packed) in list(self._store.items()):
                wc, bc, rows, masters, wt, b32 = packed
                dtype = stamp[-1]
                now = self._stamp(masters, dtype)
                if now != stamp:
                    n = len(rows)
                    self._fill(wc, bc, masters[:n], masters[n:], wt, b32)
                    self._store[key] = (now, packed)

    def invalidate(self):
        """Marks every packing stale WITHOUT dropping its buffers: the next get() refills it in place.  Used right before
        a CUDA-graph capture of a training step so that the refill copies are part of the graph."""
        for key, (stamp, packed) in list(self._store.items()):
            self._store[key] = (("stale", stamp[-1]), packed)

    def clear(self):
        self._store.clear()
        self._table, self._table_key = None, None


def cached_linear(cache: PackCache, key, linears, x, dtype, pad_rows_to=1, use_bias=True):
    """use_bias=False: the GEMM runs bias-less; the caller adds the (fp32 master) bias inside the next fused kernel."""
    wc, bc, rows, masters, wt, b32 = cache.get(key, linears, dtype, pad_rows_to)
    if ops.use_own_gemm(x) and wt is not None and wc.shape[0] % 8 == 0:
        return ops.tc_linear_packed(x, (wc, wt, b32, rows), masters if use_bias else masters[:len(rows)], use_bias)
    if not use_bias:
        return ops.packed_linear(x, wc, None, rows, masters[:len(rows)])
    return ops.packed_linear(x, wc, bc, rows, masters)


class AttentionLayer(nn.Module):
    def __init__(self, d_model, n_heads, d_keys=None, d_values=None):
        super().__init__()
        d_keys = d_keys or d_model // n_heads
        d_values = d_values or d_model // n_heads
        if d_keys != d_values or d_keys not in (64, 128):
            raise ValueError("cpmusic kernels implement query_dimensions = value_dimensions = 64 (the reference's) or 128")
        self.n_heads = n_heads
        self.query_projection = nn.Linear(d_model, d_keys * n_heads)
        self.key_projection = nn.Linear(d_model, d_keys * n_heads)
        self.value_projection = nn.Linear(d_model, d_values * n_heads)
        self.out_projection = nn.Linear(d_values * n_heads, d_model)


class TransformerEncoderLayer(nn.Module):
    def __init__(self, d_model, n_heads, d_ff, dropout=0.1, d_head=None):
        super().__init__()
        self.attention = AttentionLayer(d_model, n_heads, d_head, d_head)
        self.linear1 = nn.Linear(d_model, d_ff)
        self.linear2 = nn.Linear(d_ff, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)


class TransformerEncoder(nn.Module):
    """Parallel (teacher-forced) causal-linear encoder."""

    def __init__(self, n_layers=12, n_heads=8, query_dimensions=64, value_dimensions=64,
                 feed_forward_dimensions=2048, dropout=0.1, compute_dtype=torch.bfloat16):
        super().__init__()
        if query_dimensions != value_dimensions or query_dimensions not in (64, 128):
            raise ValueError("cpmusic kernels implement query_dimensions = value_dimensions = 64 (the reference's) or 128")
        d_model = value_dimensions * n_heads
        self.d_model, self.n_heads, self.d_head = d_model, n_heads, query_dimensions
        self.layers = nn.ModuleList([
            TransformerEncoderLayer(d_model, n_heads, feed_forward_dimensions, dropout, query_dimensions) for _ in range(n_layers)])
        self.norm = nn.LayerNorm(d_model)
        self.compute_dtype = compute_dtype
        # 0 auto | 1 simt | 2 tcgen05, one CTA per (batch, head, segment) | 3 tcgen05 chunk-parallel (cpm_linattn_fwd `impl`);
        # CPM_LINATTN_IMPL overrides the default for whole-run A/B measurements (explicit 2 / 3 need bf16 and L % 128 == 0)
        self.attn_impl = int(_os.environ.get("CPM_LINATTN_IMPL", "0"))
        self._cache = PackCache()

    # ---- fused path used by the CP model (stays in compute dtype) -------------------------
    def _layer(self, i, layer, x, key_mask=None):
        p = layer.dropout.p if self.training else 0.0
        dt, c, at = self.compute_dtype, self._cache, layer.attention
        qkv = cached_linear(c, ("qkv", i), [at.query_projection, at.key_projection, at.value_projection], x, dt)
        a = ops.causal_linear_attention_fused(qkv, self.n_heads, ops.EPS_ATTN, self.attn_impl, key_mask)
        # out-projection, linear1 and linear2 run bias-less: their biases are added inside the LayerNorm / GELU kernels,
        # whose backward kernels return the bias gradients as by-products (no reduction pass over the gradient tensors)
        if not FUSED_BIAS_GRADS:
            o = cached_linear(c, ("out", i), [at.out_projection], a, dt)
            x = ops.ln_residual(x, o, layer.norm1.weight, layer.norm1.bias, layer.norm1.eps, p)
            if FUSED_GELU_EPILOGUE and ops.use_own_gemm(x) and dt == torch.bfloat16:
                # linear1 + GELU + dropout + linear2 with the activation in linear1's epilogue and its backward in the epilogue of
                # linear2's data gradient (opt-in: measured slower than the separate GELU kernels, see FUSED_GELU_EPILOGUE)
                pack1, _ = c.get_gemm_pack(("ff1", i), [layer.linear1], dt)
                pack2, _ = c.get_gemm_pack(("ff2", i), [layer.linear2], dt)
                f = ops.tc_ffn(x, pack1, pack2, p, layer.linear1, layer.linear2)
                return ops.ln_residual(x, f, layer.norm2.weight, layer.norm2.bias, layer.norm2.eps, p)
            h = cached_linear(c, ("ff1", i), [layer.linear1], x, dt)
            g = ops.gelu_dropout(h, p)
            f = cached_linear(c, ("ff2", i), [layer.linear2], g, dt)
            return ops.ln_residual(x, f, layer.norm2.weight, layer.norm2.bias, layer.norm2.eps, p)
        if FUSED_BIAS_GRADS == "all":     # also the two 512-wide biases through the LayerNorm kernels (measured: a wash)
            o = cached_linear(c, ("out", i), [at.out_projection], a, dt, use_bias=False)
            x = ops.ln_residual(x, o, layer.norm1.weight, layer.norm1.bias, layer.norm1.eps, p, res_bias=at.out_projection.bias)
        else:
            o = cached_linear(c, ("out", i), [at.out_projection], a, dt)
            x = ops.ln_residual(x, o, layer.norm1.weight, layer.norm1.bias, layer.norm1.eps, p)
        h = cached_linear(c, ("ff1", i), [layer.linear1], x, dt, use_bias=False)
        g = ops.gelu_dropout(h, p, bias=layer.linear1.bias)
        if FUSED_BIAS_GRADS == "all":
            f = cached_linear(c, ("ff2", i), [layer.linear2], g, dt, use_bias=False)
            return ops.ln_residual(x, f, layer.norm2.weight, layer.norm2.bias, layer.norm2.eps, p, res_bias=layer.linear2.bias)
        f = cached_linear(c, ("ff2", i), [layer.linear2], g, dt)
        return ops.ln_residual(x, f, layer.norm2.weight, layer.norm2.bias, layer.norm2.eps, p)

    def forward_fused(self, x, key_mask=None):
        """x (N,L,d) in compute dtype -> (N,L,d) in compute dtype.  key_mask (N,L): ft's key-padding (length) mask or None."""
        for i, layer in enumerate(self.layers):
            x = self._layer(i, layer, x, key_mask)
        return ops.ln_residual(x, None, self.norm.weight, self.norm.bias, self.norm.eps, 0.0)

    # ---- recurrent (one token per call) path; shares every parameter with the parallel path ------
    def new_state(self, N, device):
        H, E = self.n_heads, self.d_head
        return [[torch.zeros(N, H, E, E, dtype=torch.float32, device=device),
                 torch.zeros(N, H, E, dtype=torch.float32, device=device)] for _ in self.layers]

    def _step_layer(self, i, layer, x, st):
        p = layer.dropout.p if self.training else 0.0
        dt, c, at, H = self.compute_dtype, self._cache, layer.attention, self.n_heads
        qkv = cached_linear(c, ("qkv", i), [at.query_projection, at.key_projection, at.value_projection], x, dt)
        N = qkv.shape[0]
        E = self.d_head
        q, k, v = (qkv[:, j * H * E:(j + 1) * H * E].unflatten(-1, (H, E)) for j in range(3))
        a = ops.linattn_step(q, k, v, st[0], st[1]).view(N, H * E)
        o = cached_linear(c, ("out", i), [at.out_projection], a, dt)
        x = ops.ln_residual(x, o, layer.norm1.weight, layer.norm1.bias, layer.norm1.eps, p)
        if p == 0.0 and ops.use_own_gemm(x) and x.shape[0] < ops.SMALL_GEMM_ROWS and not torch.is_grad_enabled():
            # generation: GELU in linear1's epilogue (one launch less per layer; bit-identical to GEMM + gelu kernel)
            pack1, _ = c.get_gemm_pack(("ff1", i), [layer.linear1], dt)
            g = ops.gemm_nt_small(x, pack1[0], pack1[2], gelu=True)
        else:
            h = cached_linear(c, ("ff1", i), [layer.linear1], x, dt)
            g = ops.gelu_dropout(h, p)
        f = cached_linear(c, ("ff2", i), [layer.linear2], g, dt)
        return ops.ln_residual(x, f, layer.norm2.weight, layer.norm2.bias, layer.norm2.eps, p)

    def step_fused(self, x, state):
        """x (N,d) compute dtype; state as above (allocated when None)."""
        if state is None:
            state = self.new_state(x.shape[0], x.device)
        for i, layer in enumerate(self.layers):
            if state[i] is None:
                state[i] = self.new_state(x.shape[0], x.device)[0]
            x = self._step_layer(i, layer, x, state[i])
        return ops.ln_residual(x, None, self.norm.weight, self.norm.bias, self.norm.eps, 0.0), state

    # ---- ft signature ------------------------------------------------------------------------
    def forward(self, x, attn_mask=None, length_mask=None):
        if attn_mask is None or not getattr(attn_mask, "lower_triangular", False):
            raise RuntimeError("CausalLinearAttention only supports full lower triangular masks")
        key_mask = None
        if length_mask is not None:                # K = K * k_len.float_matrix; the reference never passes one
            key_mask = length_mask.bool_matrix if hasattr(length_mask, "bool_matrix") else torch.as_tensor(length_mask).bool()
            if key_mask.shape != x.shape[:2]:
                raise ValueError(f"length_mask covers {tuple(key_mask.shape)}, the input is {tuple(x.shape[:2])}")
        return self.forward_fused(x.to(self.compute_dtype), key_mask).to(x.dtype)

    def _apply(self, fn, *a, **k):
        self._cache.clear()
        return super()._apply(fn, *a, **k)


class RecurrentTransformerEncoder(TransformerEncoder):
    """One-token-per-call encoder.  ``state`` is a list (one entry per layer) of
    ``[Si (N,H,64,64) fp32, Zi (N,H,64) fp32]`` updated in place; ``memory`` is ft's deprecated
    alias and the keyword the reference uses (dqn_policy/model.py:237)."""
    def forward(self, x, state=None, memory=None):
        state = state if state is not None else memory
        with torch.no_grad():       # the recurrent kernels update state in place (ft does so under no_grad)
            y, state = self.step_fused(x.to(self.compute_dtype), state)
        return y.to(x.dtype), state


class _Builder:
    _cls = TransformerEncoder
    _known = ("n_layers", "n_heads", "query_dimensions", "value_dimensions", "feed_forward_dimensions", "dropout")

    def __init__(self, **kw):
        self.kw = kw

    @classmethod
    def from_kwargs(cls, **kw):
        return cls(**kw)

    def get(self):
        kw = dict(self.kw)
        attention_type = kw.pop("attention_type", "full")
        if attention_type != "causal-linear":
            raise ValueError(f"cpmusic implements attention_type='causal-linear' only (got {attention_type!r})")
        activation = kw.pop("activation", "relu")
        if activation != "gelu":
            raise ValueError(f"cpmusic implements activation='gelu' only (got {activation!r})")
        unknown = set(kw) - set(self._known) - {"compute_dtype"}
        if unknown:
            raise ValueError(f"unsupported builder arguments: {sorted(unknown)}")
        return self._cls(**kw)


class TransformerEncoderBuilder(_Builder):
    _cls = TransformerEncoder


class RecurrentEncoderBuilder(_Builder):
    _cls = RecurrentTransformerEncoder


def install_fast_transformers_shim():
    """Register ``fast_transformers.builders`` / ``fast_transformers.masking`` in sys.modules so the
    unmodified reference model files import this encoder (``from fast_transformers.builders import
    TransformerEncoderBuilder`` — dqn_policy/model.py:9-11)."""
    import sys
    import types
    pkg = types.ModuleType("fast_transformers")
    builders = types.ModuleType("fast_transformers.builders")
    masking = types.ModuleType("fast_transformers.masking")
    builders.TransformerEncoderBuilder = TransformerEncoderBuilder
    builders.RecurrentEncoderBuilder = RecurrentEncoderBuilder
    masking.TriangularCausalMask = TriangularCausalMask
    masking.FullMask = FullMask
    masking.LengthMask = LengthMask
    pkg.builders, pkg.masking = builders, masking
    pkg.__path__ = []
    sys.modules["fast_transformers"] = pkg
    sys.modules["fast_transformers.builders"] = builders
    sys.modules["fast_transformers.masking"] = masking
    return pkg
