"""Batched, device-resident recurrent generation (the rollout half of the PPO loop).

Replaces the reference's batch-1 host-driven loop (testing-no-type-cp.py:126-179: per token six
``.cpu()`` syncs, numpy sampling, one H2D copy) with: per-sequence recurrent state resident in HBM,
sampling on the device (Philox stream keyed by (seed, sequence id, step, attribute) so results do
not depend on how sequences are sharded over GPUs), a device-side step counter, and the whole
one-token step (embedding → 12 layers → heads → sample → bookkeeping) captured once in a CUDA graph
and replayed with no host synchronisation.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib, ops


class RolloutEngine:
    def __init__(self, model, batch: int, max_steps: int, greedy: bool = False, true_positions: bool = True,
                 temperature=None, top_p=None, seed: int = 0, seq_base: int = 0, use_graph: bool = True,
                 fused: Optional[bool] = None):
        self.model, self.N, self.max_steps = model, batch, max_steps
        self.greedy, self.true_positions = greedy, true_positions
        self.temperature, self.top_p = model.sampling_config(temperature, top_p)
        self.seed, self.seq_base, self.use_graph = seed, seq_base, use_graph
        dev = next(model.parameters()).device
        enc = model.transformer_encoder
        A, H, nl = len(model.attrs), enc.n_heads, len(enc.layers)
        self.S = torch.zeros(nl, batch, H, 64, 64, dtype=torch.float32, device=dev)
        self.Z = torch.zeros(nl, batch, H, 64, dtype=torch.float32, device=dev)
        self.state = [[self.S[i], self.Z[i]] for i in range(nl)]
        self.cur = torch.zeros(batch, A, dtype=torch.int64, device=dev)
        self.logp = torch.zeros(batch, A, dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.hist_tok = torch.zeros(max_steps, batch, A, dtype=torch.int64, device=dev)
        self.hist_logp = torch.zeros(max_steps, batch, A, dtype=torch.float32, device=dev)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_step = None
        self.fused = self.fused_supported() if fused is None else bool(fused)
        if self.fused and not self.fused_supported():
            raise ValueError("fused rollout step needs bf16 compute, <= 32 sequences and 64-aligned widths <= 2048")

    # ---- logits of the next token for every sequence, given self.cur and the recurrent state -------
    def fused_supported(self) -> bool:
        """The fused skinny-GEMM step needs bf16, <= 32 sequences (activation tile in shared memory
        up to K = d_inner = 2048) and 64-aligned widths."""
        m = self.model
        widths = (m.d_model, m.d_inner, int(sum(m.emb_sizes)))
        return (m.compute_dtype == torch.bfloat16 and self.N <= 32 and all(w % 64 == 0 and w <= 2048 for w in widths))

    def _logits_unfused(self):
        m = self.model
        z = m._embed(self.cur[:, None, :], 0, self.step_dev if self.true_positions else None)
        h, _ = m.transformer_encoder.step_fused(z.view(self.N, m.d_model), self.state)
        return m.logits_concat(h)

    def _logits_fused(self):
        """Same arithmetic with one launch per Linear layer: LayerNorms are folded into the consumer
        GEMM's prologue, bias / GELU / residual / positional encoding into its epilogue
        (5 launches per layer instead of 8-9)."""
        m, enc, N = self.model, self.model.transformer_encoder, self.N
        dt, H = torch.bfloat16, enc.n_heads
        emb = ops.cp_embed(self.cur[:, None, :], m._tables(), dt).view(N, -1)
        w, b, _, _ = m._cache.get("in", [m.in_linear], dt)
        x = ops.skinny_linear(emb, w, b, epilogue=ops.EPI_PE, pe=m.pos_emb.pe,
                              pos_dev=self.step_dev if self.true_positions else None)
        s_prev, prev = None, None
        for i, layer in enumerate(enc.layers):
            at = layer.attention
            wq, bq, _, _ = enc._cache.get(("qkv", i), [at.query_projection, at.key_projection, at.value_projection], dt)
            if i == 0:
                xin, qkv = x, ops.skinny_linear(x, wq, bq)
            else:
                xin = torch.empty(N, m.d_model, dtype=dt, device=x.device)
                qkv = ops.skinny_linear(s_prev, wq, bq, ln=(prev.norm2.weight, prev.norm2.bias, prev.norm2.eps), xout=xin)
            q, k, v = (qkv[:, j * H * 64:(j + 1) * H * 64].unflatten(-1, (H, 64)) for j in range(3))
            a = ops.linattn_step(q, k, v, self.state[i][0], self.state[i][1]).view(N, H * 64)
            wo, bo, _, _ = enc._cache.get(("out", i), [at.out_projection], dt)
            s1 = ops.skinny_linear(a, wo, bo, epilogue=ops.EPI_RESIDUAL, residual=xin)
            w1, b1, _, _ = enc._cache.get(("ff1", i), [layer.linear1], dt)
            x1 = torch.empty(N, m.d_model, dtype=dt, device=x.device)
            hmid = ops.skinny_linear(s1, w1, b1, ln=(layer.norm1.weight, layer.norm1.bias, layer.norm1.eps), xout=x1,
                                     epilogue=ops.EPI_GELU)
            w2, b2, _, _ = enc._cache.get(("ff2", i), [layer.linear2], dt)
            s_prev = ops.skinny_linear(hmid, w2, b2, epilogue=ops.EPI_RESIDUAL, residual=x1)
            prev = layer
        xl = ops.ln_residual(s_prev, None, prev.norm2.weight, prev.norm2.bias, prev.norm2.eps, 0.0)
        wh, bh, _, _ = m._cache.get("heads", m._heads(), dt, 8)
        return ops.skinny_linear(xl, wh, bh, ln=(enc.norm.weight, enc.norm.bias, enc.norm.eps))

    # one token for every sequence: reads self.cur, overwrites self.cur with the sampled token
    def _step(self):
        m = self.model
        lc = self._logits_fused() if self.fused else self._logits_unfused()
        ops.heads_sample(lc, m.seg, self.temperature, self.top_p, greedy=self.greedy, seed=self.seed,
                         seq_base=self.seq_base, step_dev=self.step_dev, tokens_out=self.cur, logp_out=self.logp)
        ops.rollout_advance(self.cur, self.hist_tok, self.logp, self.hist_logp, self.step_dev, self.max_steps)

    def reset(self, init_tokens):
        self.S.zero_()
        self.Z.zero_()
        self.step_dev.zero_()
        self.cur.copy_(init_tokens.to(self.cur.device, torch.int64))

    def _capture(self):
        was_training = self.model.training
        self.model.eval()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):               # warm-up: cuBLAS handles / workspaces, weight packs
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        before = _lib.kernel_launches()
        with torch.no_grad(), torch.cuda.graph(g):
            self._step()
        self.launches_per_step = _lib.kernel_launches() - before      # cpmusic kernels captured per token step
        self.graph = g
        self.model.train(was_training)

    @torch.no_grad()
    def generate(self, init_tokens, n_steps: Optional[int] = None, seed: Optional[int] = None):
        """init_tokens (N,A) int64 -> dict(tokens (N, n_steps+1, A) incl. the initial token,
        logp (N, n_steps, A) log-prob of each sampled sub-token under the T=1 policy)."""
        n_steps = self.max_steps if n_steps is None else n_steps
        if n_steps > self.max_steps:
            raise ValueError(f"n_steps {n_steps} > max_steps {self.max_steps}")
        if seed is not None and seed != self.seed:
            self.seed, self.graph = seed, None
        was_training = self.model.training
        self.model.eval()
        if self.use_graph and self.graph is None:
            self._capture()
        self.reset(init_tokens)
        self.model.refresh_packs()          # the graph reads the packed weights by address
        if self.use_graph:
            for _ in range(n_steps):
                self.graph.replay()
            _lib.EXTRA_LAUNCHES[0] += n_steps * self.launches_per_step
        else:
            for _ in range(n_steps):
                self._step()
        self.model.train(was_training)
        toks = torch.cat([init_tokens.to(self.cur.device, torch.int64)[None], self.hist_tok[:n_steps]], 0)
        return {"tokens": toks.permute(1, 0, 2).contiguous(), "logp": self.hist_logp[:n_steps].permute(1, 0, 2).contiguous()}
