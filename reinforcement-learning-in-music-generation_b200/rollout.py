"""Batched, device-resident recurrent generation (the rollout half of the PPO loop).

Replaces the reference's batch-1 host-driven loop (testing-no-type-cp.py:126-179: per token six
``.cpu()`` syncs, numpy sampling, one H2D copy) with: per-sequence recurrent state resident in HBM,
sampling on the device (Philox stream keyed by (seed, sequence id, step, attribute) so results do
not depend on how sequences are sharded over GPUs), a device-side step counter, and the whole
one-token step (embedding -> 12 layers -> heads -> sample -> bookkeeping) run with no host synchronisation.

Two executions of the step, same arithmetic:
  * ``mode="chain"`` (default): the step as 67 kernels of this library (small-M tcgen05 GEMMs with LayerNorm folded in and
    the GELU in linear1's epilogue, the recurrent state kernel, embedding gather, sampler) launched with programmatic
    dependent launch - each kernel's set-up and each GEMM's weight fetch overlap its predecessor's tail - and captured once in
    a CUDA graph.  442 us per token step at 256 songs on B200.
  * ``mode="persistent"`` (opt-in, bf16 models with 64-wide heads; also CPM_ROLLOUT_MODE=persistent): ONE cooperative kernel
    for the whole rollout (csrc/rollout_step.cu, cpm_rollout_run): one CTA per SM, the 63 dependent stages of a token separated
    by a device-wide barrier, every CTA streaming the weight tiles of its own output tiles ahead of the barriers through a TMA
    ring, tcgen05 tiles with the weight rows on the UMMA M axis, LayerNorm applied while the activation tile is staged.
    700-760 us per token step: measured slower, kept as the one non-default mode with its per-stage timeline (see its header).

(Round 1 also carried a first cooperative megakernel, skinny-GEMM, LayerNorm-fold, deferred / split state write-back,
L2-prefetch and grouped-graph variants of the step; all measured slower than the chain and were removed - the write-ups
are profiles/r01_summary.md sections F, K, M.)
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

from . import _lib, ops
from ._lib import check


class RolloutEngine:
    def __init__(self, model, batch: int, max_steps: int, greedy: bool = False, true_positions: bool = True,
                 temperature=None, top_p=None, seed: int = 0, seq_base: int = 0, use_graph: bool = True,
                 chain_pdl: Optional[bool] = None, mode: Optional[str] = None, fold_ln: Optional[bool] = None):
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
        mode = mode or os.environ.get("CPM_ROLLOUT_MODE", "chain")
        if mode not in ("persistent", "chain"):
            raise ValueError("mode must be 'persistent' or 'chain'")
        if mode == "persistent" and (model.compute_dtype != torch.bfloat16 or not use_graph):
            mode = "chain"                    # fp32 parity mode / eager debugging: the kernel chain
        self.mode = mode
        self._plan = None                     # (handle, plan buffer, scratch, seed) of the persistent kernel
        # LayerNorm folded into the chain's Linear kernels (cpm_gemm_nt_small_ln): 25 LayerNorm launches per token become 1.
        # Needs the own bf16 GEMMs, 64-wide-or-not heads alike, and d_model <= 512 (the statistics come from the K <= 512 tile).
        fold = os.environ.get("CPM_ROLLOUT_FOLD", "1") == "1" if fold_ln is None else bool(fold_ln)
        self.fold = (fold and model.compute_dtype == torch.bfloat16 and ops.GEMM_IMPL == "own" and model.d_model <= 512
                     and model.d_model % 8 == 0 and batch < ops.SMALL_GEMM_ROWS)
        self._folded = None

    # ------------------------------------------------------------------ persistent kernel (csrc/rollout_step.cu)
    def _build_plan(self):
        m, enc, dev = self.model, self.model.transformer_encoder, self.cur.device
        lib = _lib.load()
        dt = torch.bfloat16
        B, d, A = self.N, m.d_model, len(m.attrs)
        cfg = _lib.RolloutConfig()
        keep = []                             # tensors the plan points at

        def ptr(t):
            keep.append(t)
            return t.data_ptr()

        cfg.batch, cfg.d_model, cfg.n_heads, cfg.d_ff = B, d, enc.n_heads, enc.layers[0].linear1.out_features
        cfg.n_layers, cfg.n_attr = len(enc.layers), A
        for a, t in enumerate(m._tables()):
            cfg.n_tokens[a], cfg.emb[a], cfg.tables[a] = t.shape[0], t.shape[1], ptr(t)
        wc, _bc, _rows, _ms, _wt, b32 = m._cache.get("in", [m.in_linear], dt)
        cfg.w_in, cfg.b_in = ptr(wc), ptr(b32)
        pe = m.pos_emb.pe[0]
        cfg.pe, cfg.pe_len, cfg.true_positions = ptr(pe), pe.shape[0], int(self.true_positions)
        for i, layer in enumerate(enc.layers):
            at, L = layer.attention, cfg.layer[i]
            for key, lins, wn, bn in ((("qkv", i), [at.query_projection, at.key_projection, at.value_projection], "w_qkv", "b_qkv"),
                                      (("out", i), [at.out_projection], "w_out", "b_out"),
                                      (("ff1", i), [layer.linear1], "w_ff1", "b_ff1"), (("ff2", i), [layer.linear2], "w_ff2", "b_ff2")):
                wc, _bc, _rows, _ms, _wt, b32 = enc._cache.get(key, lins, dt)
                setattr(L, wn, ptr(wc))
                setattr(L, bn, ptr(b32))
            L.ln1_g, L.ln1_b, L.ln2_g, L.ln2_b = ptr(layer.norm1.weight), ptr(layer.norm1.bias), ptr(layer.norm2.weight), ptr(layer.norm2.bias)
            L.S, L.Z = self.S[i].data_ptr(), self.Z[i].data_ptr()
        cfg.lnf_g, cfg.lnf_b = ptr(enc.norm.weight), ptr(enc.norm.bias)
        wc, _bc, _rows, _ms, _wt, b32 = m._cache.get("heads", m._heads(), dt, 8)
        cfg.w_heads, cfg.b_heads = ptr(wc), ptr(b32)
        for a, s0 in enumerate(m.seg):
            cfg.seg[a] = s0
        cfg.logits_ld = m.logits_width
        for a in range(A):
            cfg.temperature[a] = float(self.temperature[a])
            cfg.top_p[a] = float(self.top_p[a]) if self.top_p[a] is not None else 0.0
        cfg.greedy = int(self.greedy)
        cfg.ln_eps, cfg.attn_eps = enc.norm.eps, ops.EPS_ATTN
        cfg.seed, cfg.seq_base = self.seed, self.seq_base
        cfg.cur, cfg.logp = self.cur.data_ptr(), self.logp.data_ptr()
        cfg.hist_tok, cfg.hist_logp = self.hist_tok.data_ptr(), self.hist_logp.data_ptr()
        cfg.step_dev, cfg.max_steps = self.step_dev.data_ptr(), self.max_steps
        scratch = {n: torch.zeros(B, w, dtype=dt, device=dev) for n, w in
                   (("x0", d), ("x1", d), ("y", d), ("qkv", 3 * d), ("attn", d), ("g", cfg.d_ff), ("logits", m.logits_width))}
        for n, t in scratch.items():
            setattr(cfg, n, t.data_ptr())
        barrier = torch.zeros(1, dtype=torch.int64, device=dev)
        cfg.barrier, cfg.err_flag = barrier.data_ptr(), ops.IndexGuard.flag(dev).data_ptr()
        plan = torch.zeros(int(lib.cpm_rollout_plan_bytes()) + 256, dtype=torch.uint8, device=dev)
        plan_ptr = (plan.data_ptr() + 255) // 256 * 256
        handle = ctypes.c_void_p()
        check(lib.cpm_rollout_create(ctypes.byref(cfg), plan_ptr, ctypes.byref(handle)))
        self._plan = {"handle": handle, "plan": plan, "scratch": scratch, "barrier": barrier, "keep": keep, "seed": self.seed,
                      "stamp": tuple(t.data_ptr() for t in keep), "phases": int(lib.cpm_rollout_phases(handle))}

    def _drop_plan(self):
        if self._plan is not None:
            _lib.load().cpm_rollout_destroy(self._plan["handle"])
            self._plan = None

    def __del__(self):
        try:
            self._drop_plan()
        except Exception:
            pass

    def _run_persistent(self, n_steps):
        m = self.model
        m.refresh_packs()                    # in place: the plan reads the packed weights by address
        if self._plan is not None and (self._plan["seed"] != self.seed
                                       or self._plan["stamp"] != tuple(t.data_ptr() for t in self._plan["keep"])):
            self._drop_plan()                # new seed, or a pack / parameter was re-allocated (model.to(), set_compute_dtype())
        if self._plan is None:
            try:
                self._build_plan()
            except _lib.CpmError as e:
                if e.code != -7:             # CPM_ERR_UNSUPPORTED: shapes the persistent kernel does not take
                    raise
                self.mode = "chain"
                return False
        self._plan["barrier"].zero_()
        check(_lib.load().cpm_rollout_run(self._plan["handle"], n_steps, ops._st()))
        self.launches_per_step = self._plan["phases"]
        return True

    # ------------------------------------------------------------------ LayerNorm folded into the chain's Linear kernels
    def _fold_refresh(self):
        """(gamma o W) bf16, c1 = row sums of it, c2 = W beta + b for every Linear that consumes a LayerNorm: linear1 (norm1),
        the next layer's QKV projection (norm2) and the heads (the encoder's final norm).  Buffers are allocated once and
        refilled in place (the captured graph reads them by address)."""
        m, enc = self.model, self.model.transformer_encoder
        with torch.no_grad():
            if self._folded is None:
                self._folded = {}
            def fold(key, lins, norm, pad_to=1):
                W = torch.cat([l.weight for l in lins], 0).float()
                b = torch.cat([l.bias for l in lins], 0).float()
                rows = -(-W.shape[0] // pad_to) * pad_to
                ent = self._folded.get(key)
                if ent is None:
                    dev = W.device
                    ent = self._folded[key] = (torch.zeros(rows, W.shape[1], dtype=torch.bfloat16, device=dev),
                                               torch.zeros(rows, dtype=torch.float32, device=dev), torch.zeros(rows, dtype=torch.float32, device=dev))
                wf, c1, c2 = ent
                wf[:W.shape[0]].copy_(W * norm.weight.float()[None, :])
                c1[:W.shape[0]].copy_(wf[:W.shape[0]].float().sum(1))
                c2[:W.shape[0]].copy_((W * norm.bias.float()[None, :]).sum(1) + b)
            for i, layer in enumerate(enc.layers):
                at = layer.attention
                fold(("ff1", i), [layer.linear1], layer.norm1)
                if i > 0:
                    fold(("qkv", i), [at.query_projection, at.key_projection, at.value_projection], enc.layers[i - 1].norm2)
            fold("heads", m._heads(), enc.norm, 8)

    def _logits_fold(self):
        m, enc = self.model, self.model.transformer_encoder
        dt, c, H, E, N = torch.bfloat16, enc._cache, enc.n_heads, enc.d_head, self.N
        z = m._embed(self.cur[:, None, :], 0, self.step_dev if self.true_positions else None)
        x = z.view(N, m.d_model)                       # layer 0: the residual stream is a plain tensor
        y = stats2 = prev = None                       # later layers: pre-norm rows of the previous layer + their statistics
        for i, (layer, st) in enumerate(zip(enc.layers, self.state)):
            at = layer.attention
            if i == 0:
                pk, _ = c.get_gemm_pack(("qkv", i), [at.query_projection, at.key_projection, at.value_projection], dt)
                qkv = ops.gemm_nt_small(x, pk[0], pk[2])
            else:
                wf, c1, c2 = self._folded[("qkv", i)]
                stats2 = torch.empty(N, 2, dtype=torch.float32, device=x.device)
                qkv = ops.gemm_nt_small_ln(y, wf, c2, fold_c1=c1, stats_out=stats2, ln_eps=prev.norm2.eps)
            q, k, v = (qkv[:, j * H * E:(j + 1) * H * E].unflatten(-1, (H, E)) for j in range(3))
            a = ops.linattn_step(q, k, v, st[0], st[1]).view(N, H * E)
            pk, _ = c.get_gemm_pack(("out", i), [at.out_projection], dt)
            if i == 0:
                y1 = ops.gemm_nt_small_ln(a, pk[0], pk[2], resid=x)
            else:
                y1 = ops.gemm_nt_small_ln(a, pk[0], pk[2], resid=y, r_stats=stats2, r_gamma=prev.norm2.weight, r_beta=prev.norm2.bias)
            wf, c1, c2 = self._folded[("ff1", i)]
            stats1 = torch.empty(N, 2, dtype=torch.float32, device=x.device)
            g = ops.gemm_nt_small_ln(y1, wf, c2, gelu=True, fold_c1=c1, stats_out=stats1, ln_eps=layer.norm1.eps)
            pk, _ = c.get_gemm_pack(("ff2", i), [layer.linear2], dt)
            y = ops.gemm_nt_small_ln(g, pk[0], pk[2], resid=y1, r_stats=stats1, r_gamma=layer.norm1.weight, r_beta=layer.norm1.bias)
            prev = layer
        xl = ops.ln_residual(y, None, prev.norm2.weight, prev.norm2.bias, prev.norm2.eps, 0.0)      # the last norm2 stays a kernel
        wf, c1, c2 = self._folded["heads"]
        return ops.gemm_nt_small_ln(xl, wf, c2, fold_c1=c1, ln_eps=enc.norm.eps)

    def _logits(self):
        """Logits of the next token for every sequence, given self.cur and the recurrent state."""
        if self.fold:
            if self._folded is None:
                self._fold_refresh()
            return self._logits_fold()
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
        if self.mode == "persistent":
            self.reset(init_tokens)
            if self._run_persistent(n_steps):
                toks = torch.cat([init_tokens.to(self.cur.device, torch.int64)[None], self.hist_tok[:n_steps]], 0)
                return {"tokens": toks.permute(1, 0, 2).contiguous(), "logp": self.hist_logp[:n_steps].permute(1, 0, 2).contiguous()}
        was_training = self.model.training
        self.model.eval()
        if self.use_graph and self.graph is None:
            self._capture()
        self.reset(init_tokens)
        self.model.refresh_packs()          # the graph reads the packed weights by address
        if self.fold:
            self._fold_refresh()
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


