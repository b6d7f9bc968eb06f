"""Batched, device-resident recurrent generation (the rollout half of the PPO loop).

Replaces the reference's batch-1 host-driven loop (testing-no-type-cp.py:126-179: per token six
``.cpu()`` syncs, numpy sampling, one H2D copy) with: per-sequence recurrent state resident in HBM,
sampling on the device (Philox stream keyed by (seed, sequence id, step, attribute) so results do
not depend on how sequences are sharded over GPUs), a device-side step counter, and the whole
one-token step (embedding → 12 layers → heads → sample → bookkeeping) captured once in a CUDA graph
and replayed with no host synchronisation.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib, ops

_P = ctypes.c_void_p


class _MegaPhase(ctypes.Structure):          # mirrors cpm::MegaPhase (csrc/rollout_mega.cu)
    _fields_ = [("type", ctypes.c_int32), ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32),
                ("lda", ctypes.c_int32), ("ldy", ctypes.c_int32), ("ldr", ctypes.c_int32), ("pro", ctypes.c_int32),
                ("epi", ctypes.c_int32), ("H", ctypes.c_int32), ("eps", ctypes.c_float), ("pad0", ctypes.c_int32),
                ("A", _P), ("W", _P), ("bias", _P), ("R", _P), ("Y", _P), ("xout", _P),
                ("gamma", _P), ("beta", _P), ("gamma2", _P), ("beta2", _P), ("S", _P), ("Z", _P)]


class _MegaGlobals(ctypes.Structure):        # mirrors cpm::MegaGlobals
    _fields_ = [("batch", ctypes.c_int32), ("n_attr", ctypes.c_int32), ("emb_total", ctypes.c_int32), ("logits_ld", ctypes.c_int32),
                ("n_tokens", ctypes.c_int32 * 8), ("emb", ctypes.c_int32 * 8), ("emb_off", ctypes.c_int32 * 9), ("seg", ctypes.c_int32 * 9),
                ("emb_scale", ctypes.c_float * 8), ("temperature", ctypes.c_float * 8), ("top_p", ctypes.c_float * 8),
                ("greedy", ctypes.c_int32), ("true_positions", ctypes.c_int32), ("max_steps", ctypes.c_int32), ("pe_max", ctypes.c_int32),
                ("seed", ctypes.c_uint64), ("seq_base", ctypes.c_int64), ("tables", _P * 8), ("pe", _P),
                ("cur", _P), ("hist_tok", _P), ("logp", _P), ("hist_logp", _P), ("step_dev", _P), ("barrier", _P),
                ("n_phases", ctypes.c_int32), ("pad1", ctypes.c_int32)]


class RolloutEngine:
    def __init__(self, model, batch: int, max_steps: int, greedy: bool = False, true_positions: bool = True,
                 temperature=None, top_p=None, seed: int = 0, seq_base: int = 0, use_graph: bool = True,
                 fused: Optional[bool] = None, mode: Optional[str] = None, pdl: bool = False, lazy_state: bool = False,
                 split_state: bool = False, prefetch_state: int = 0):
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
        # deferred state write-back (unfused mode, opt-in): S goes back to HBM once per ops.LAZY_STATE_PERIOD tokens; the
        # updates in between live in a small ring and are re-applied in registers (bit-identical outputs, ~40 % less HBM
        # traffic).  Measured on B200 (tools/bench_lazy_step.py): 15.3 us + 1.7 us per pending entry vs 13.5 us eager -
        # re-applying the entries is shared-memory-bandwidth bound (every warp re-reads the 64-wide v vector), so it is
        # slower than simply streaming S both ways; kept as a tested experiment, not the default.
        self.lazy_state = bool(lazy_state)
        if self.lazy_state:
            self.ring = torch.zeros(nl, batch, H, ops.LAZY_STATE_PERIOD, 128, dtype=torch.float32, device=dev)
            self.state = [[self.S[i], self.Z[i], self.ring[i], self.step_dev] for i in range(nl)]
        # split step (unfused mode, opt-in): the S write-back runs as its own kernel on a side branch of the step graph
        self.split_state = bool(split_state)
        if self.split_state:
            if self.lazy_state:
                raise ValueError("choose one of lazy_state / split_state")
            self.kvp = torch.zeros(nl, batch, H, 128, dtype=torch.float32, device=dev)
            self.side = torch.cuda.Stream(device=dev)
            self.state = [[self.S[i], self.Z[i], self._split_hook(i)] for i in range(nl)]
        # L2 prefetch (unfused mode): layer i's step kernel also pulls layer i+1's state tiles into L2, so that their HBM
        # read overlaps the latency-bound GEMM / LayerNorm kernels in between (1: after the write-back, 2: first thing)
        self.prefetch_state = int(prefetch_state)
        if self.prefetch_state:
            if self.lazy_state or self.split_state:
                raise ValueError("choose one of lazy_state / split_state / prefetch_state")
            self.state = [[self.S[i], self.Z[i], self._prefetch_hook(i, nl)] for i in range(nl)]
        self.cur = torch.zeros(batch, A, dtype=torch.int64, device=dev)
        self.logp = torch.zeros(batch, A, dtype=torch.float32, device=dev)
        self.hist_tok = torch.zeros(max_steps, batch, A, dtype=torch.int64, device=dev)
        self.hist_logp = torch.zeros(max_steps, batch, A, dtype=torch.float32, device=dev)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_step = None
        # step implementation: "mega" (one persistent cooperative kernel per token), "fused" (one launch
        # per Linear) or "unfused" (library GEMMs + elementwise kernels).  Default: the fastest supported.
        if mode is None:
            # measured on B200 (profiles/): the CUDA-graph "unfused" step is currently the fastest
            # (409 us/token at 32 sequences vs 416 fused, 543 mega), so it is the default.
            mode = ("fused" if fused else "unfused") if fused is not None else "unfused"
        if mode in ("mega", "fused") and not self.fused_supported():
            raise ValueError("mega / fused rollout steps need bf16 compute, <= 32 sequences and 64-aligned widths <= 2048")
        if mode in ("tc", "fold") and not self.tc_supported():
            raise ValueError("the tcgen05 rollout step needs bf16 compute, widths that are multiples of 64 (inputs) / 32 (outputs)")
        if mode not in ("mega", "fused", "unfused", "tc", "fold"):
            raise ValueError(f"unknown rollout mode {mode!r}")
        if E != 64 and (mode != "unfused" or self.lazy_state or self.split_state or self.prefetch_state):
            raise ValueError("128-wide heads run the plain unfused rollout step only")
        if (self.lazy_state or self.split_state or self.prefetch_state) and mode != "unfused":
            raise ValueError("lazy_state / split_state / prefetch_state are implemented for the unfused step")
        self.mode = mode
        self.fused = mode == "fused"
        self.pdl = pdl
        import os
        self.chain_pdl = (mode == "unfused" and model.compute_dtype == torch.bfloat16 and ops.GEMM_IMPL == "own"
                          and os.environ.get("CPM_CHAIN_PDL", "1") == "1")
        self._mega = None
        self._tc = None

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
        w, b, _, _ = m._cache.get("in", [m.in_linear], dt)[:4]
        x = ops.skinny_linear(emb, w, b, epilogue=ops.EPI_PE, pe=m.pos_emb.pe,
                              pos_dev=self.step_dev if self.true_positions else None)
        s_prev, prev = None, None
        for i, layer in enumerate(enc.layers):
            at = layer.attention
            wq, bq, _, _ = enc._cache.get(("qkv", i), [at.query_projection, at.key_projection, at.value_projection], dt)[:4]
            if i == 0:
                xin, qkv = x, ops.skinny_linear(x, wq, bq)
            else:
                xin = torch.empty(N, m.d_model, dtype=dt, device=x.device)
                qkv = ops.skinny_linear(s_prev, wq, bq, ln=(prev.norm2.weight, prev.norm2.bias, prev.norm2.eps), xout=xin)
            q, k, v = (qkv[:, j * H * 64:(j + 1) * H * 64].unflatten(-1, (H, 64)) for j in range(3))
            a = ops.linattn_step(q, k, v, self.state[i][0], self.state[i][1]).view(N, H * 64)
            wo, bo, _, _ = enc._cache.get(("out", i), [at.out_projection], dt)[:4]
            s1 = ops.skinny_linear(a, wo, bo, epilogue=ops.EPI_RESIDUAL, residual=xin)
            w1, b1, _, _ = enc._cache.get(("ff1", i), [layer.linear1], dt)[:4]
            x1 = torch.empty(N, m.d_model, dtype=dt, device=x.device)
            hmid = ops.skinny_linear(s1, w1, b1, ln=(layer.norm1.weight, layer.norm1.bias, layer.norm1.eps), xout=x1,
                                     epilogue=ops.EPI_GELU)
            w2, b2, _, _ = enc._cache.get(("ff2", i), [layer.linear2], dt)[:4]
            s_prev = ops.skinny_linear(hmid, w2, b2, epilogue=ops.EPI_RESIDUAL, residual=x1)
            prev = layer
        xl = ops.ln_residual(s_prev, None, prev.norm2.weight, prev.norm2.bias, prev.norm2.eps, 0.0)
        wh, bh, _, _ = m._cache.get("heads", m._heads(), dt, 8)[:4]
        return ops.skinny_linear(xl, wh, bh, ln=(enc.norm.weight, enc.norm.bias, enc.norm.eps))

    # ---- tcgen05 step: every Linear is one cpm_tc_linear launch, no LayerNorm / GELU / residual kernels ----
    def tc_supported(self) -> bool:
        m = self.model
        ins = (m.d_model, m.d_inner, int(sum(m.emb_sizes)))
        outs = (m.d_model, 3 * m.d_model, m.d_inner)
        return m.compute_dtype == torch.bfloat16 and all(w % 64 == 0 for w in ins + outs)

    def _tc_stamp(self):
        from .encoder import _OPT_EPOCH                      # fused optimizers do not bump _version (see PackCache)
        return tuple(p._version for p in self.model.parameters()) + (_OPT_EPOCH[0],)

    def _tc_refresh(self):
        """(Re)builds, in place, the bf16 weights with the consumer-side LayerNorm folded in:
        W' = gamma (.) W,  c1 = rowsum(W'),  c2 = W beta + bias  (include/cpmusic.h, cpm_tc_linear)."""
        m, enc = self.model, self.model.transformer_encoder
        stamp = self._tc_stamp()
        if self._tc is not None and self._tc["stamp"] == stamp:
            return
        bf = torch.bfloat16
        first = self._tc is None
        packs = {} if first else self._tc["packs"]

        def put(key, w, c1, c2):
            if first:
                packs[key] = (w.to(bf).contiguous(), None if c1 is None else c1.float().contiguous(), c2.float().contiguous())
            else:                                          # same addresses: captured graphs stay valid
                packs[key][0].copy_(w)
                if c1 is not None:
                    packs[key][1].copy_(c1)
                packs[key][2].copy_(c2)

        def fold(key, linears, norm):
            w = torch.cat([l.weight for l in linears], 0).float()
            b = torch.cat([l.bias for l in linears], 0).float()
            if norm is None:
                put(key, w, None, b)
            else:
                wf = (w * norm.weight.float()[None, :]).to(bf)
                put(key, wf, wf.float().sum(1), w @ norm.bias.float() + b)

        with torch.no_grad():
            fold("in", [m.in_linear], None)
            prev = None
            for i, layer in enumerate(enc.layers):
                at = layer.attention
                fold(("qkv", i), [at.query_projection, at.key_projection, at.value_projection], None if prev is None else prev.norm2)
                fold(("out", i), [at.out_projection], None)
                fold(("ff1", i), [layer.linear1], layer.norm1)
                fold(("ff2", i), [layer.linear2], None)
                prev = layer
        self._tc = {"stamp": stamp, "packs": packs}

    def _logits_tc(self):
        m, enc, N = self.model, self.model.transformer_encoder, self.N
        dt, H, d, pdl = torch.bfloat16, enc.n_heads, m.d_model, self.pdl
        P = self._tc["packs"]
        parts = d // 64
        emb = ops.cp_embed(self.cur[:, None, :], m._tables(), dt).view(N, -1)
        w, _, b = P["in"]
        x0 = ops.tc_linear(emb, w, b, epilogue=ops.TL_PE, pe=m.pos_emb.pe, pos_dev=self.step_dev if self.true_positions else None,
                           block_n=64, pdl=pdl)
        s_prev, st_prev, prev = x0, None, None
        for i, layer in enumerate(enc.layers):
            w, c1, c2 = P[("qkv", i)]
            qkv = ops.tc_linear(s_prev, w, c2, c1=c1, stats_in=st_prev, eps=0.0 if prev is None else prev.norm2.eps, block_n=64, pdl=pdl)
            q, k, v = (qkv[:, j * H * 64:(j + 1) * H * 64].unflatten(-1, (H, 64)) for j in range(3))
            a = ops.linattn_step(q, k, v, self.state[i][0], self.state[i][1]).view(N, H * 64)
            w, _, c2 = P[("out", i)]
            st1 = torch.empty(N, parts, 2, dtype=torch.float32, device=a.device)
            if prev is None:
                s1 = ops.tc_linear(a, w, c2, epilogue=ops.TL_RES, residual=s_prev, stats_out=st1, block_n=64, pdl=pdl)
            else:
                s1 = ops.tc_linear(a, w, c2, epilogue=ops.TL_RES_LN, residual=s_prev, stats_r=st_prev, gamma_r=prev.norm2.weight,
                                   beta_r=prev.norm2.bias, eps=prev.norm2.eps, stats_out=st1, block_n=64, pdl=pdl)
            w, c1, c2 = P[("ff1", i)]
            h = ops.tc_linear(s1, w, c2, c1=c1, stats_in=st1, eps=layer.norm1.eps, epilogue=ops.TL_GELU, block_n=64, pdl=pdl)
            w, _, c2 = P[("ff2", i)]
            st2 = torch.empty(N, parts, 2, dtype=torch.float32, device=a.device)
            s2 = ops.tc_linear(h, w, c2, epilogue=ops.TL_RES_LN, residual=s1, stats_r=st1, gamma_r=layer.norm1.weight, beta_r=layer.norm1.bias,
                               eps=layer.norm1.eps, stats_out=st2, block_n=64, pdl=pdl)
            s_prev, st_prev, prev = s2, st2, layer
        xl = ops.ln_residual(s_prev, None, prev.norm2.weight, prev.norm2.bias, prev.norm2.eps, 0.0)
        xf = ops.ln_residual(xl, None, enc.norm.weight, enc.norm.bias, enc.norm.eps, 0.0)
        return m.logits_concat(xf)

    # ---- "fold" step: library GEMMs on raw pre-LayerNorm sums, no LayerNorm launches (csrc/rollout_fold.cu) ----
    def _logits_fold(self):
        m, enc, N = self.model, self.model.transformer_encoder, self.N
        H = enc.n_heads
        P = self._tc["packs"]
        x0 = m._embed(self.cur[:, None, :], 0, self.step_dev if self.true_positions else None).view(N, m.d_model)
        s_prev, prev = x0, None
        for i, layer in enumerate(enc.layers):
            w, c1, c2 = P[("qkv", i)]
            raw = s_prev @ w.t()
            bo = P[("out", i)][2]
            a, xres = ops.linattn_step_fold(raw, s_prev, c1, c2, None if prev is None else prev.norm2.weight,
                                            None if prev is None else prev.norm2.bias, bo, self.state[i][0], self.state[i][1], H,
                                            fold=prev is not None, eps_ln=0.0 if prev is None else prev.norm2.eps)
            s1 = torch.addmm(xres, a, P[("out", i)][0].t())
            w, c1, c2 = P[("ff1", i)]
            rawh = s1 @ w.t()
            h, xres2 = ops.gelu_fold(rawh, s1, c1, c2, layer.norm1.weight, layer.norm1.bias, P[("ff2", i)][2], layer.norm1.eps)
            s_prev = torch.addmm(xres2, h, P[("ff2", i)][0].t())
            prev = layer
        xl = ops.ln_residual(s_prev, None, prev.norm2.weight, prev.norm2.bias, prev.norm2.eps, 0.0)
        xf = ops.ln_residual(xl, None, enc.norm.weight, enc.norm.bias, enc.norm.eps, 0.0)
        return m.logits_concat(xf)

    # ---- persistent megakernel step -----------------------------------------------------------------
    def _build_mega(self):
        m, enc, N = self.model, self.model.transformer_encoder, self.N
        dev, dt = self.cur.device, torch.bfloat16
        lib = _lib.load()
        gb, pb = ctypes.c_int(), ctypes.c_int()
        lib.cpm_mega_sizes(ctypes.byref(gb), ctypes.byref(pb))
        if gb.value != ctypes.sizeof(_MegaGlobals) or pb.value != ctypes.sizeof(_MegaPhase):
            raise RuntimeError("megakernel struct layout mismatch between rollout.py and rollout_mega.cu")
        d, di, H, A = m.d_model, m.d_inner, enc.n_heads, len(m.attrs)
        E = int(sum(m.emb_sizes))
        buf = lambda w: torch.zeros(N, w, dtype=dt, device=dev)
        sc = {"x0": buf(d), "xb": buf(d), "qkv": buf(3 * d), "a": buf(d), "s1": buf(d), "x1": buf(d), "h": buf(di), "s2": buf(d),
              "logits": buf(m.logits_width)}
        keep = [sc]
        phases = []

        def gemm(A_, W, b, Y, Nn, K, pro=0, epi=0, ln=None, ln2=None, xout=None, R=None):
            ph = _MegaPhase()
            ph.type, ph.M, ph.N, ph.K = 0, N, Nn, K
            ph.lda = 0 if A_ is None else A_.stride(0)
            ph.ldy, ph.ldr = Y.stride(0), (0 if R is None else R.stride(0))
            ph.pro, ph.epi, ph.eps = pro, epi, (ln[2] if ln is not None else 0.0)
            ph.A = None if A_ is None else A_.data_ptr()
            ph.W, ph.bias, ph.Y = W.data_ptr(), b.data_ptr(), Y.data_ptr()
            ph.R = None if R is None else R.data_ptr()
            ph.xout = None if xout is None else xout.data_ptr()
            if ln is not None:
                ph.gamma, ph.beta = ln[0].data_ptr(), ln[1].data_ptr()
            if ln2 is not None:
                ph.gamma2, ph.beta2 = ln2[0].data_ptr(), ln2[1].data_ptr()
            phases.append(ph)

        w, b, _, _ = m._cache.get("in", [m.in_linear], dt)[:4]
        gemm(None, w, b, sc["x0"], d, E, pro=3, epi=3)
        prev = None
        for i, layer in enumerate(enc.layers):
            at = layer.attention
            wq, bq, _, _ = enc._cache.get(("qkv", i), [at.query_projection, at.key_projection, at.value_projection], dt)[:4]
            if i == 0:
                gemm(sc["x0"], wq, bq, sc["qkv"], 3 * d, d)
                xin = sc["x0"]
            else:
                gemm(sc["s2"], wq, bq, sc["qkv"], 3 * d, d, pro=1, ln=(prev.norm2.weight, prev.norm2.bias, prev.norm2.eps), xout=sc["xb"])
                xin = sc["xb"]
            ph = _MegaPhase()
            ph.type, ph.M, ph.H, ph.lda, ph.ldy, ph.eps = 1, N, H, 3 * d, d, ops.EPS_ATTN
            ph.A, ph.Y, ph.S, ph.Z = sc["qkv"].data_ptr(), sc["a"].data_ptr(), self.S[i].data_ptr(), self.Z[i].data_ptr()
            phases.append(ph)
            wo, bo, _, _ = enc._cache.get(("out", i), [at.out_projection], dt)[:4]
            gemm(sc["a"], wo, bo, sc["s1"], d, d, epi=2, R=xin)
            w1, b1, _, _ = enc._cache.get(("ff1", i), [layer.linear1], dt)[:4]
            gemm(sc["s1"], w1, b1, sc["h"], di, d, pro=1, epi=1, ln=(layer.norm1.weight, layer.norm1.bias, layer.norm1.eps), xout=sc["x1"])
            w2, b2, _, _ = enc._cache.get(("ff2", i), [layer.linear2], dt)[:4]
            gemm(sc["h"], w2, b2, sc["s2"], d, di, epi=2, R=sc["x1"])
            prev = layer
        wh, bh, _, _ = m._cache.get("heads", m._heads(), dt, 8)[:4]
        gemm(sc["s2"], wh, bh, sc["logits"], m.logits_width, d, pro=2, ln=(prev.norm2.weight, prev.norm2.bias, prev.norm2.eps),
             ln2=(enc.norm.weight, enc.norm.bias, enc.norm.eps))
        ph = _MegaPhase()
        ph.type, ph.M, ph.lda, ph.A = 2, N, m.logits_width, sc["logits"].data_ptr()
        phases.append(ph)

        g = _MegaGlobals()
        g.batch, g.n_attr, g.emb_total, g.logits_ld = N, A, E, m.logits_width
        off = 0
        for a in range(A):
            g.n_tokens[a], g.emb[a], g.emb_off[a] = m.n_token[a], m.emb_sizes[a], off
            g.emb_scale[a] = float(m.emb_sizes[a]) ** 0.5
            g.temperature[a] = float(self.temperature[a])
            g.top_p[a] = 0.0 if self.top_p[a] is None else float(self.top_p[a])
            g.tables[a] = m._tables()[a].data_ptr()
            g.seg[a] = m.seg[a]
            off += m.emb_sizes[a]
        g.emb_off[A], g.seg[A] = off, m.seg[A]
        g.greedy, g.true_positions, g.max_steps = int(self.greedy), int(self.true_positions), self.max_steps
        pe2 = m.pos_emb.pe.reshape(-1, d)
        g.pe_max, g.pe = pe2.shape[0], pe2.data_ptr()
        g.seed, g.seq_base = self.seed, self.seq_base
        self._mega_barrier = torch.zeros(1, dtype=torch.int32, device=dev)
        g.cur, g.hist_tok, g.logp, g.hist_logp = self.cur.data_ptr(), self.hist_tok.data_ptr(), self.logp.data_ptr(), self.hist_logp.data_ptr()
        g.step_dev, g.barrier, g.n_phases = self.step_dev.data_ptr(), self._mega_barrier.data_ptr(), len(phases)
        arr = (_MegaPhase * len(phases))(*phases)
        to_dev = lambda obj: torch.frombuffer(bytearray(bytes(obj)), dtype=torch.uint8).to(dev)
        self._mega = {"globals": to_dev(g), "phases": to_dev(arr), "keep": keep, "n_phases": len(phases), "seed": self.seed}

    def _step_mega(self):
        if self._mega is None or self._mega["seed"] != self.seed:
            self._build_mega()
        ops.check(_lib.load().cpm_rollout_step_mega(self._mega["globals"].data_ptr(), self._mega["phases"].data_ptr(),
                                                    torch.cuda.current_stream().cuda_stream))

    # one token for every sequence: reads self.cur, overwrites self.cur with the sampled token
    def _step(self):
        """One token for every sequence.  In the default mode every kernel of the step is a chain kernel (csrc/cpm_common.cuh):
        with ``chain_pdl`` they are launched with programmatic dependent launch, each overlapping its set-up (and the GEMMs
        their weight fetch) with its predecessor's tail."""
        if self.chain_pdl:
            ops.set_chain_pdl(True)
            try:
                return self._step_inner()
            finally:
                ops.set_chain_pdl(False)
        return self._step_inner()

    def _step_inner(self):
        m = self.model
        if self.mode == "mega":
            return self._step_mega()
        if self.mode in ("tc", "fold"):
            if self._tc is None:
                self._tc_refresh()
            lc = self._logits_tc() if self.mode == "tc" else self._logits_fold()
        else:
            lc = self._logits_fused() if self.fused else self._logits_unfused()
        ops.heads_sample(lc, m.seg, self.temperature, self.top_p, greedy=self.greedy, seed=self.seed,
                         seq_base=self.seq_base, step_dev=self.step_dev, tokens_out=self.cur, logp_out=self.logp)
        ops.rollout_advance(self.cur, self.hist_tok, self.logp, self.hist_logp, self.step_dev, self.max_steps)
        if self.split_state or self.prefetch_state == 4:
            torch.cuda.current_stream().wait_stream(self.side)   # join: all write-backs land before the next token step
            self._pf_pending = False

    def _prefetch_hook(self, i, nl):
        nxt = self.S[(i + 1) % nl]
        if self.prefetch_state in (1, 2):
            return lambda q, k, v, S, Z: ops.linattn_step(q, k, v, S, Z, prefetch=nxt, prefetch_when=self.prefetch_state)
        if self.prefetch_state == 3:                        # stand-alone prefetch kernel right after the step, same chain
            def hook3(q, k, v, S, Z):
                out = ops.linattn_step(q, k, v, S, Z)
                ops.l2_prefetch(nxt)
                return out
            return hook3
        if not hasattr(self, "side"):
            self.side = torch.cuda.Stream(device=self.S.device)
        self._pf_pending = False

        def hook4(q, k, v, S, Z):                           # stand-alone prefetch kernel on a side branch of the step graph
            main = torch.cuda.current_stream()
            if self._pf_pending:
                main.wait_stream(self.side)                 # join: the prefetch of THIS state was issued a layer ago
            out = ops.linattn_step(q, k, v, S, Z)
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                ops.l2_prefetch(nxt)
            self._pf_pending = True
            return out
        return hook4

    def _split_hook(self, i):
        def hook(q, k, v, S, Z):
            out = ops.linattn_step_out(q, k, v, S, Z, self.kvp[i])
            main = torch.cuda.current_stream()
            self.side.wait_stream(main)                     # fork: the write-back needs the parked [Kf | v]
            with torch.cuda.stream(self.side):
                ops.linattn_state_update(S, self.kvp[i])
            return out
        return hook

    def flush_state(self):
        """Brings S up to date when the write-back is deferred (no-op otherwise)."""
        if self.lazy_state:
            for st in self.state:
                ops.linattn_state_flush(st[0], st[1], st[2], st[3])

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
        mega = self.mode == "mega"
        if mega and (self._mega is None or self._mega["seed"] != self.seed):
            self._build_mega()
        if self.use_graph and not mega and self.graph is None:
            self._capture()
        self.reset(init_tokens)
        self.model.refresh_packs()          # graph / megakernel read the packed weights by address
        if self.mode in ("tc", "fold"):
            self._tc_refresh()
        if mega:
            for _ in range(n_steps):        # one cooperative launch per token step, queued back to back
                self._step_mega()
            self.launches_per_step = 1
        elif self.use_graph:
            for _ in range(n_steps):
                self.graph.replay()
            _lib.EXTRA_LAUNCHES[0] += n_steps * self.launches_per_step
        else:
            for _ in range(n_steps):
                self._step()
        self.flush_state()
        self.model.train(was_training)
        toks = torch.cat([init_tokens.to(self.cur.device, torch.int64)[None], self.hist_tok[:n_steps]], 0)
        return {"tokens": toks.permute(1, 0, 2).contiguous(), "logp": self.hist_logp[:n_steps].permute(1, 0, 2).contiguous()}


class GroupedRolloutEngine:
    """The same rollout with the batch cut into ``groups`` independent sub-batches whose token steps are
    captured as parallel branches of ONE CUDA graph (fork / join over side streams).  A token step is a
    chain of ~100 small dependent kernels, so a single chain is launch-latency bound; with several chains
    in flight the latency-bound kernels of one group overlap the bandwidth-bound state updates of the
    others.  Sampling streams are keyed by the global sequence id, so the generated tokens are identical
    to the ungrouped engine's for any ``groups`` (tests/test_gpu_model.py pins this)."""

    def __init__(self, model, batch: int, max_steps: int, groups: int = 4, steps_per_graph: int = 1, seq_base: int = 0, **kw):
        if batch % groups != 0:
            raise ValueError(f"batch {batch} is not divisible into {groups} groups")
        self.model, self.N, self.max_steps, self.groups = model, batch, max_steps, groups
        self.per = batch // groups
        self.steps_per_graph = max(1, int(steps_per_graph))
        self.engines = [RolloutEngine(model, self.per, max_steps, seq_base=seq_base + g * self.per, use_graph=False, **kw)
                        for g in range(groups)]
        self.streams = None
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_step = None

    def _fork_join_steps(self, n):
        main = torch.cuda.current_stream()
        for eng, st in zip(self.engines, self.streams):
            st.wait_stream(main)
            with torch.cuda.stream(st):
                for _ in range(n):
                    eng._step()
        for st in self.streams:
            main.wait_stream(st)

    def _capture(self):
        was_training = self.model.training
        self.model.eval()
        self.streams = [torch.cuda.Stream() for _ in self.engines]
        with torch.no_grad():
            for _ in range(2):                      # warm-up: cuBLAS handles / workspaces per stream, weight packs
                self._fork_join_steps(1)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        before = _lib.kernel_launches()
        with torch.no_grad(), ops.graph_capture(g):
            self._fork_join_steps(self.steps_per_graph)
        self.launches_per_step = (_lib.kernel_launches() - before) // self.steps_per_graph
        self.graph = g
        self.model.train(was_training)

    @torch.no_grad()
    def generate(self, init_tokens, n_steps: Optional[int] = None):
        n_steps = self.max_steps if n_steps is None else n_steps
        if n_steps > self.max_steps:
            raise ValueError(f"n_steps {n_steps} > max_steps {self.max_steps}")
        was_training = self.model.training
        self.model.eval()
        if self.graph is None:
            self._capture()
        init_tokens = init_tokens.to(self.engines[0].cur.device, torch.int64)
        for g, eng in enumerate(self.engines):
            eng.reset(init_tokens[g * self.per:(g + 1) * self.per])
        self.model.refresh_packs()
        full, rest = divmod(n_steps, self.steps_per_graph)
        for _ in range(full):
            self.graph.replay()
        if rest:                                    # tail shorter than one captured block: eager fork / join
            self._fork_join_steps(rest)
        _lib.EXTRA_LAUNCHES[0] += full * self.steps_per_graph * self.launches_per_step
        self.model.train(was_training)
        hist_tok = torch.cat([e.hist_tok[:n_steps] for e in self.engines], 1)
        hist_logp = torch.cat([e.hist_logp[:n_steps] for e in self.engines], 1)
        toks = torch.cat([init_tokens[None], hist_tok], 0)
        return {"tokens": toks.permute(1, 0, 2).contiguous(), "logp": hist_logp.permute(1, 0, 2).contiguous()}
