"""Batched, device-resident recurrent generation (the rollout half of the PPO loop).

Replaces the reference's batch-1 host-driven loop (testing-no-type-cp.py:126-179: per token six
``.cpu()`` syncs, numpy sampling, one H2D copy) with: per-sequence recurrent state resident in HBM,
sampling on the device (Philox stream keyed by (seed, sequence id, step, attribute) so results do
not depend on how sequences are sharded over GPUs), a device-side step counter, and the whole
one-token step (embedding -> 12 layers -> heads -> sample -> bookkeeping) captured once in a CUDA graph
and replayed with no host synchronisation.  Every kernel of the step is one of this library's own
(small-M tcgen05 GEMMs with the GELU in linear1's epilogue, the recurrent state kernel, fused
residual + LayerNorm, embedding gather, sampler) and is launched with programmatic dependent launch,
so each kernel's set-up - and each GEMM's weight fetch - overlaps its predecessor's tail.

(Round 1 also carried a cooperative megakernel, skinny-GEMM, LayerNorm-fold, deferred / split state
write-back, L2-prefetch and grouped-graph variants of the step; all measured slower than this one and
were removed - the write-ups are profiles/r01_summary.md sections F, K, M.)
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import _lib, ops


class RolloutEngine:
    def __init__(self, model, batch: int, max_steps: int, greedy: bool = False, true_positions: bool = True,
                 temperature=None, top_p=None, seed: int = 0, seq_base: int = 0, use_graph: bool = True,
                 chain_pdl: Optional[bool] = None):
        self.model, self.N, self.max_steps = model, batch, max_steps
        self.greedy, self.true_positions = greedy, true_positions
        self.temperature, self.top_p = model.sampling_config(temperature, top_p)
        self.seed, self.seq_base, self.use_graph = seed, seq_base, use_graph
        dev = next(model.parameters()).device
        enc = model.transformer_encoder
        A, H, nl = len(model.attrs), enc.n_heads, len(enc.layers)
        E = enc.d_head
        self.S = torch.zeros(nl, batch, H, E, E, dtype=torch.float32, device=dev)
        self.Z = torch.zeros(nl, batch, H, E, dtype=torch.float32, device=dev)
        self.state = [[self.S[i], self.Z[i]] for i in range(nl)]
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.cur = torch.zeros(batch, A, dtype=torch.int64, device=dev)
        self.logp = torch.zeros(batch, A, dtype=torch.float32, device=dev)
        self.hist_tok = torch.zeros(max_steps, batch, A, dtype=torch.int64, device=dev)
        self.hist_logp = torch.zeros(max_steps, batch, A, dtype=torch.float32, device=dev)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_step = None
        if chain_pdl is None:
            chain_pdl = os.environ.get("CPM_CHAIN_PDL", "1") == "1"
        # programmatic dependent launch needs every kernel of the step to be a chain kernel (csrc/cpm_common.cuh): true when
        # the Linear layers run on the own GEMMs (bf16); library GEMMs in between would simply not overlap
        self.chain_pdl = bool(chain_pdl) and model.compute_dtype == torch.bfloat16 and ops.GEMM_IMPL == "own"

    def _logits(self):
        """Logits of the next token for every sequence, given self.cur and the recurrent state."""
        m = self.model
        z = m._embed(self.cur[:, None, :], 0, self.step_dev if self.true_positions else None)
        h, _ = m.transformer_encoder.step_fused(z.view(self.N, m.d_model), self.state)
        return m.logits_concat(h)

    def _step(self):
        """One token for every sequence."""
        if self.chain_pdl:
            ops.set_chain_pdl(True)
        try:
            m = self.model
            lc = self._logits()
            ops.heads_sample(lc, m.seg, self.temperature, self.top_p, greedy=self.greedy, seed=self.seed,
                             seq_base=self.seq_base, step_dev=self.step_dev, tokens_out=self.cur, logp_out=self.logp)
            ops.rollout_advance(self.cur, self.hist_tok, self.logp, self.hist_logp, self.step_dev, self.max_steps)
        finally:
            if self.chain_pdl:
                ops.set_chain_pdl(False)

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
        # captured on a high-priority stream: the kernel nodes inherit it, so the token step's small dependent kernels are
        # scheduled ahead of bulk work queued on other (default-priority) streams
        with torch.no_grad(), ops.graph_capture(g, stream=torch.cuda.Stream(priority=-1)):
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


