"""GPU parity tests, kernel level: every libcpmusic entry point (called through the C-ABI via
cpmusic.ops) against the CPU oracle on the same seeded inputs and against the committed golden
vectors.  Tolerances: fp32 storage -> 1e-4-class (fp32 accumulation order differs from the fp64
oracle); bf16 storage -> 2e-2-class relative to the output scale (8 mantissa bits), stated per test.
Integer outputs (greedy tokens, actions) are bit-exact."""
import math

import numpy as np
import pytest
import torch

from oracle import ft_oracle as ft, rl_oracle as rl, sampling_oracle as so

pytestmark = pytest.mark.gpu
VOCAB = [56, 135, 18, 87, 18, 25]


def _cmp(a, b, atol, rtol=0.0, what=""):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).double().cpu()
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    assert bool((err <= tol).all()), f"{what}: max err {err.max().item():.3e} (tol {atol}+{rtol}*|ref|), ref scale {b.abs().max().item():.3e}"


def _oracle_attn(q, k, v, go=None):
    q, k, v = (t.detach().double().cpu().requires_grad_() for t in (q, k, v))
    out = ft.causal_linear_attention(q, k, v)
    if go is None:
        return out
    out.backward(go.detach().double().cpu())
    return out, q.grad, k.grad, v.grad


# ------------------------------------------------------------------ linear attention
@pytest.mark.parametrize("impl", [1])
def test_linattn_fp32_golden(cuda, cpm, golden, impl):
    g = golden("linattn")
    q, k, v, go = (torch.from_numpy(g[n]).to(cuda).requires_grad_() for n in ("q", "k", "v", "go"))
    out = cpm.ops.causal_linear_attention(q, k, v, impl=impl)
    assert cpm.ops.linattn_last_impl() == "simt"
    out.backward(go.detach())
    _cmp(out, g["out"], 3e-5, 1e-5, "out")
    _cmp(q.grad, g["gq"], 2e-4, 1e-4, "gq")
    _cmp(k.grad, g["gk"], 2e-4, 1e-4, "gk")
    _cmp(v.grad, g["gv"], 2e-4, 1e-4, "gv")


@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 63, 3), (1, 64, 2), (3, 65, 1), (2, 300, 8), (1, 1024, 2), (1, 2048, 1), (160, 70, 1)])
def test_linattn_fp32_vs_oracle_ragged_and_segmented(cuda, cpm, shape):
    """Edge lengths (1, 63, 64, 65), non-multiples of the chunk, and few-(batch,head) long sequences
    that exercise the segment-total + scan path (N*H < 148)."""
    N, L, H = shape
    gen = torch.Generator().manual_seed(L * 7 + H)
    q, k, v, go = (torch.randn(N, L, H, 64, generator=gen).to(cuda).requires_grad_() for _ in range(4))
    out = cpm.ops.causal_linear_attention(q, k, v, impl=1)
    out.backward(go.detach())
    ro, rq, rk, rv = _oracle_attn(q, k, v, go)
    _cmp(out, ro, 5e-5, 2e-5, "out")
    scale = 1.0 + math.sqrt(L) * 0.02
    _cmp(q.grad, rq, 3e-4 * scale, 2e-4, "gq")
    _cmp(k.grad, rk, 3e-4 * scale, 2e-4, "gk")
    _cmp(v.grad, rv, 3e-4 * scale, 2e-4, "gv")


def test_linattn_fused_qkv_layout(cuda, cpm):
    """q,k,v as column slices of one (N,L,3*H*64) buffer (token stride 3*H*64), grads written into
    the matching slices of one buffer — the layout the encoder uses, no permute/contiguous copies."""
    N, L, H = 2, 130, 4
    gen = torch.Generator().manual_seed(5)
    qkv = torch.randn(N, L, 3 * H * 64, generator=gen).to(cuda).requires_grad_()
    go = torch.randn(N, L, H * 64, generator=gen).to(cuda)
    out = cpm.ops.causal_linear_attention_fused(qkv, H, impl=1)
    out.backward(go)
    q, k, v = (qkv.detach()[..., i * H * 64:(i + 1) * H * 64].reshape(N, L, H, 64) for i in range(3))
    ro, rq, rk, rv = _oracle_attn(q, k, v, go.view(N, L, H, 64))
    _cmp(out.view(N, L, H, 64), ro, 5e-5, 2e-5, "out")
    _cmp(qkv.grad, torch.cat([t.reshape(N, L, H * 64) for t in (rq, rk, rv)], -1), 4e-4, 2e-4, "gqkv")


@pytest.mark.parametrize("impl", [1])
def test_linattn_bf16_vs_oracle(cuda, cpm, impl):
    """bf16 storage: compare with the fp64 oracle evaluated on the SAME bf16-rounded inputs.
    Tolerance 2e-2 absolute on O(1) outputs (bf16 output rounding alone is 4e-3 relative)."""
    N, L, H = 2, 256, 4
    gen = torch.Generator().manual_seed(11)
    q, k, v, go = (torch.randn(N, L, H, 64, generator=gen).to(cuda).bfloat16().requires_grad_() for _ in range(4))
    out = cpm.ops.causal_linear_attention(q, k, v, impl=impl)
    out.backward(go.detach())
    ro, rq, rk, rv = _oracle_attn(q.float(), k.float(), v.float(), go.float())
    _cmp(out, ro, 2e-2, 1e-2, "out")
    _cmp(q.grad, rq, 3e-2, 2e-2, "gq")
    _cmp(k.grad, rk, 3e-2, 2e-2, "gk")
    _cmp(v.grad, rv, 3e-2, 2e-2, "gv")


def test_linattn_step_golden_and_equals_parallel(cuda, cpm, golden):
    g = golden("recurrent")
    q, k, v = (torch.from_numpy(g[n]).to(cuda) for n in ("q", "k", "v"))       # (T,N,H,E)
    T, N, H, E = q.shape
    S = torch.zeros(N, H, E, E, device=cuda)
    Z = torch.zeros(N, H, E, device=cuda)
    outs = [cpm.ops.linattn_step(q[t], k[t], v[t], S, Z) for t in range(T)]
    _cmp(torch.stack(outs), g["out"], 2e-5, 1e-5, "step out")
    _cmp(S, g["S"], 1e-5, 1e-5, "S")
    _cmp(Z, g["Z"], 1e-5, 1e-5, "Z")
    # chunked == recurrent on a longer sequence (SURVEY §4 equivalence row), incl. strided qkv rows
    T, N, H = 150, 2, 3
    gen = torch.Generator().manual_seed(12)
    qkv = torch.randn(T, N, 3 * H * 64, generator=gen).to(cuda)
    S = torch.zeros(N, H, 64, 64, device=cuda)
    Z = torch.zeros(N, H, 64, device=cuda)
    rec = []
    for t in range(T):
        qq, kk, vv = (qkv[t][:, i * H * 64:(i + 1) * H * 64].unflatten(-1, (H, 64)) for i in range(3))
        rec.append(cpm.ops.linattn_step(qq, kk, vv, S, Z))
    rec = torch.stack(rec, 1)                                                     # (N,T,H,64)
    par = cpm.ops.causal_linear_attention_fused(qkv.permute(1, 0, 2).contiguous(), H, impl=1).view(N, T, H, 64)
    _cmp(rec, par, 2e-5, 2e-5, "recurrent vs chunked")
    with pytest.raises(ValueError, match="batch size changed"):
        cpm.ops.linattn_step(qq[:1], kk[:1], vv[:1], S, Z)


def test_linattn_properties_full_size(cuda, cpm):
    """BASELINE cfg2 attention shape (32 x 512 x 8 x 64, bf16): size-independent properties.
    (a) causality: changing tokens >= t leaves outputs < t bit-identical; (b) linearity in v;
    (c) scale invariance: out(q,k,c*v) == c*out; (d) convex-combination bound: every output lies
    within [min v, max v] per channel prefix (weights are positive and normalised)."""
    N, L, H = 32, 512, 8
    gen = torch.Generator().manual_seed(13)
    q, k, v = (torch.randn(N, L, H, 64, generator=gen).to(cuda).bfloat16() for _ in range(3))
    base = cpm.ops.causal_linear_attention(q, k, v)
    t = 301
    q2, k2, v2 = q.clone(), k.clone(), v.clone()
    q2[:, t:], k2[:, t:], v2[:, t:] = 1.5, -0.25, 3.0
    pert = cpm.ops.causal_linear_attention(q2, k2, v2)
    assert torch.equal(base[:, :t], pert[:, :t]) and not torch.equal(base[:, t:], pert[:, t:])
    two = cpm.ops.causal_linear_attention(q, k, (2 * v))
    _cmp(two, 2 * base.float(), 2e-2, 1e-2, "scale")
    cmax = torch.cummax(v.float(), 1).values
    cmin = torch.cummin(v.float(), 1).values
    o = base.float()
    assert bool((o <= cmax + 3e-2).all()) and bool((o >= cmin - 3e-2).all())


def test_linattn_cfg5_long_sequence(cuda, cpm):
    """BASELINE cfg5 attention shape (1 x 8192 tokens, d_model 1024 = 16 heads x 64, bf16).  Forward and backward of
    the chunk-parallel tcgen05 kernels (64 chunks per head: the parallel state pre-pass + scan path) against the fp32-math
    SIMT kernels, plus exact causality of both directions: outputs before t are bit-identical when tokens >= t change,
    and gradients at positions >= t are exactly zero when the upstream gradient is zero there."""
    N, L, H = 1, 8192, 16
    gen = torch.Generator().manual_seed(5)
    q, k, v, go = (torch.randn(N, L, H, 64, generator=gen).to(cuda).bfloat16() for _ in range(4))
    saved = cpm.ops.linattn_saved(N, L, H, cuda)
    out, den = cpm.ops.linattn_fwd_raw(q, k, v, impl=3, saved=saved)
    assert cpm.ops.linattn_last_impl() == "tcgen05-cp"
    ref, den_ref = cpm.ops.linattn_fwd_raw(q, k, v, impl=1)
    _cmp(out, ref.float(), 3e-2, 2e-2, "cfg5 fwd vs simt")
    _cmp(den, den_ref, 1e-3, 1e-2, "cfg5 den vs simt")
    t = 5000
    q2, k2, v2 = q.clone(), k.clone(), v.clone()
    q2[:, t:], k2[:, t:], v2[:, t:] = -0.5, 0.75, 2.0
    out2, _ = cpm.ops.linattn_fwd_raw(q2, k2, v2, impl=3)
    assert torch.equal(out[:, :t], out2[:, :t]) and not torch.equal(out[:, t:], out2[:, t:])
    gq, gk, gv = (torch.empty_like(q) for _ in range(3))
    cpm.ops.linattn_bwd_raw(q, k, v, out, den, go, gq, gk, gv, impl=3, saved=saved)
    sq, sk, sv = (torch.empty_like(q) for _ in range(3))
    cpm.ops.linattn_bwd_raw(q, k, v, out, den, go, sq, sk, sv, impl=1)
    for name, a, b in (("gq", gq, sq), ("gk", gk, sk), ("gv", gv, sv)):
        _cmp(a, b.float(), 4e-2, 3e-2, f"cfg5 bwd vs simt {name}")
    go2 = go.clone()
    go2[:, t:] = 0
    cpm.ops.linattn_bwd_raw(q, k, v, out, den, go2, gq, gk, gv, impl=3, saved=saved)
    for g in (gq, gk, gv):
        assert not bool(g[:, t:].any()) and bool(g[:, :t].any())


@pytest.mark.parametrize("shape", [(3, 128, 2), (2, 384, 3), (1, 2048, 8), (100, 256, 1), (2, 200, 2)])
def test_linattn_native_128_wide_kernels_vs_oracle(cuda, cpm, shape):
    """The NATIVE 128-wide path (cpm_linattn_fwd / bwd with E = M = 128: the per-chunk kernels instantiated over the two feature
    halves of q / k and the two value halves of v / out - one K = 128 score tile, a 2 x 2 grid of state tiles per chunk): one
    chunk (no prefix state), several chunks, a long sequence, many chains (still the per-chunk state kernels: the streaming ones
    are 64-wide), a ragged length (zero-padded at the end).  Forward and all three gradients against the fp64 oracle at E = 128,
    with the prefix states kept from the forward and rebuilt by the backward; and against the two-pass 64-wide decomposition."""
    N, L, H = shape
    gen = torch.Generator().manual_seed(3 * L + H)
    qkv = torch.randn(N, L, 3 * H * 128, generator=gen).to(cuda).bfloat16().requires_grad_()
    go = torch.randn(N, L, H * 128, generator=gen).to(cuda).bfloat16()
    lib = cpm._lib.load()
    before = cpm._lib.COUNTS["cpm_linattn_fwd"]
    out = cpm.ops.causal_linear_attention_fused(qkv, H)
    assert cpm._lib.COUNTS["cpm_linattn_fwd"] == before + 1, "one kernel pass, not the two-pass decomposition"
    assert cpm.ops.linattn_last_impl() == "tcgen05-cp" and out.shape == (N, L, H * 128)
    out.backward(go)
    q, k, v = (qkv.detach()[..., i * H * 128:(i + 1) * H * 128].unflatten(-1, (H, 128)) for i in range(3))
    ro, rq, rk, rv = _oracle_attn(q.float(), k.float(), v.float(), go.unflatten(-1, (H, 128)).float())
    scale = 1.0 + math.sqrt(L / 1024.0)
    _cmp(out.unflatten(-1, (H, 128)), ro, 3e-2, 2e-2, "out vs oracle")
    g = qkv.grad
    for i, (name, ref) in enumerate((("gq", rq), ("gk", rk), ("gv", rv))):
        _cmp(g[..., i * H * 128:(i + 1) * H * 128].unflatten(-1, (H, 128)), ref, 4e-2 * scale, 3e-2, f"{name} vs oracle")
    if L % 128 == 0:                                   # raw calls: backward without the kept prefix states rebuilds them
        o2, den = cpm.ops.linattn_fwd_raw(q, k, v)
        assert torch.equal(o2.reshape(N, L, H * 128), out.detach())
        g2 = torch.empty_like(qkv.detach())
        gq, gk, gv = (g2[..., i * H * 128:(i + 1) * H * 128].unflatten(-1, (H, 128)) for i in range(3))
        cpm.ops.linattn_bwd_raw(q, k, v, o2, den, go.unflatten(-1, (H, 128)), gq, gk, gv, saved=None)
        assert torch.equal(g2, g), "gradients with rebuilt prefix states differ from those with the kept ones"
        assert lib.cpm_linattn_saved_bytes_wide(N, L, H, 128) == N * H * (L // 128) * (4 * 8192 + 512)
    two = cpm.ops._linattn_fused_e128(qkv.detach(), H, cpm.ops.EPS_ATTN, 0)
    _cmp(out, two, 3e-2, 2e-2, "native vs two 64-wide passes")


@pytest.mark.parametrize("dtype,impl,shape", [(torch.float32, 1, (2, 150, 3)), (torch.float32, 1, (1, 64, 1)),
                                              (torch.bfloat16, 0, (2, 256, 4)), (torch.bfloat16, 0, (1, 1024, 8))])
def test_linattn_head_width_128_vs_oracle(cuda, cpm, dtype, impl, shape):
    """128-wide heads (cfg5 as 8 heads x 128) through the ft-style (N,L,H,E) entry point: fp32 = two passes of the
    64-wide kernels over virtual heads, recombined with the joint normaliser; bf16 = the native 128-wide tensor-core kernels —
    forward and all three gradients against the fp64 oracle run at E = M = 128.
    Tolerances as for the 64-wide kernels (fp32 1e-4 class; bf16 storage 2e-2 class, on the same bf16-rounded inputs)."""
    N, L, H = shape
    gen = torch.Generator().manual_seed(L + H)
    q, k, v, go = (torch.randn(N, L, H, 128, generator=gen).to(cuda).to(dtype).requires_grad_() for _ in range(4))
    out = cpm.ops.causal_linear_attention(q, k, v, impl=impl)
    assert out.shape == (N, L, H, 128) and out.dtype == dtype
    if dtype == torch.bfloat16 and L % 128 == 0:
        assert cpm.ops.linattn_last_impl() == "tcgen05-cp"
    out.backward(go.detach())
    ro, rq, rk, rv = _oracle_attn(q.float(), k.float(), v.float(), go.float())
    if dtype == torch.float32:
        scale = 1.0 + math.sqrt(L) * 0.02
        _cmp(out, ro, 5e-5, 2e-5, "out")
        for name, a, b in (("gq", q.grad, rq), ("gk", k.grad, rk), ("gv", v.grad, rv)):
            _cmp(a, b, 3e-4 * scale, 2e-4, name)
    else:
        _cmp(out, ro, 2e-2, 1e-2, "out")
        for name, a, b in (("gq", q.grad, rq), ("gk", k.grad, rk), ("gv", v.grad, rv)):
            _cmp(a, b, 4e-2, 3e-2, name)


@pytest.mark.parametrize("E,M", [(128, 128), (32, 64), (96, 32)])
def test_linattn_step_other_head_widths(cuda, cpm, E, M):
    """The generic recurrent step (E != 64 or M != 64) against ft's recurrence (oracle, fp64) over 40 tokens, fp32 and
    bf16 inputs; E = M = 128 additionally equals the chunked path on the same sequence."""
    T, N, H = 40, 3, 2
    gen = torch.Generator().manual_seed(E + M)
    q, k = (torch.randn(T, N, H, E, generator=gen) for _ in range(2))
    v = torch.randn(T, N, H, M, generator=gen)
    for dtype, tol in ((torch.float32, 3e-5), (torch.bfloat16, 2e-2)):
        qd, kd, vd = (t.to(cuda).to(dtype) for t in (q, k, v))
        S = torch.zeros(N, H, E, M, device=cuda)
        Z = torch.zeros(N, H, E, device=cuda)
        lib, W, state = cpm._lib.load(), H * max(E, M), None
        for t in range(T):
            o = torch.empty(N, H, M, dtype=dtype, device=cuda)
            qp, kp, vp = (torch.zeros(N, W, dtype=dtype, device=cuda) for _ in range(3))      # one row stride for q, k and v
            qp[:, :H * E], kp[:, :H * E], vp[:, :H * M] = qd[t].reshape(N, -1), kd[t].reshape(N, -1), vd[t].reshape(N, -1)
            rc = lib.cpm_linattn_step(qp.data_ptr(), kp.data_ptr(), vp.data_ptr(), S.data_ptr(), Z.data_ptr(), o.data_ptr(),
                                      N, H, E, M, W, H * M, cpm.ops._dt(qd), cpm.ops.EPS_ATTN, torch.cuda.current_stream().cuda_stream)
            assert rc == 0, lib.cpm_last_error_string()
            ro, state = ft.recurrent_linear_attention(qd[t].double().cpu(), kd[t].double().cpu(), vd[t].double().cpu(), state)
            _cmp(o, ro, tol, tol, f"step {t} out ({dtype})")
        _cmp(S, state[0], 1e-4, 1e-5, "S")
        _cmp(Z, state[1], 1e-4, 1e-5, "Z")
    if E == M == 128:
        par = cpm.ops.causal_linear_attention(q.to(cuda).permute(1, 0, 2, 3).contiguous(), k.to(cuda).permute(1, 0, 2, 3).contiguous(),
                                              v.to(cuda).permute(1, 0, 2, 3).contiguous(), impl=1)
        S = torch.zeros(N, H, E, M, device=cuda)
        Z = torch.zeros(N, H, E, device=cuda)
        rec = torch.stack([cpm.ops.linattn_step(q[t].to(cuda), k[t].to(cuda), v[t].to(cuda), S, Z) for t in range(T)], 1)
        _cmp(rec, par, 3e-5, 3e-5, "recurrent vs chunked at E = M = 128")


# ------------------------------------------------------------------ embedding / PE / dropout
def test_embed_fwd_bwd(cuda, cpm):
    emb = [128, 256, 64, 512, 128, 128]
    gen = torch.Generator().manual_seed(20)
    tables = [torch.randn(n, e, generator=gen).to(cuda).requires_grad_() for n, e in zip(VOCAB, emb)]
    idx = torch.stack([torch.randint(0, n, (3, 77), generator=gen) for n in VOCAB], -1).to(cuda)
    ref = torch.cat([t.detach()[idx[..., a]] * math.sqrt(e) for a, (t, e) in enumerate(zip(tables, emb))], -1)
    out = cpm.ops.cp_embed(idx, tables, torch.float32)
    assert torch.equal(out, ref)                                        # pure gather * scale: bit-exact in fp32
    go = torch.randn(out.shape, generator=gen).to(cuda)
    out.backward(go)
    ref_t = [t.detach().clone().requires_grad_() for t in tables]
    torch.cat([t[idx[..., a]] * math.sqrt(e) for a, (t, e) in enumerate(zip(ref_t, emb))], -1).backward(go)
    for a in range(6):
        _cmp(tables[a].grad, ref_t[a].grad, 1e-4, 1e-5, f"gtable{a}")
    out16 = cpm.ops.cp_embed(idx, tables, torch.bfloat16)
    assert torch.equal(out16, ref.bfloat16())
    # out-of-range index: zeros + error flag, no fault
    bad = idx.clone()
    bad[0, 0, 1] = 9999
    o, err = cpm.ops.embed_fwd_raw(bad, [t.detach() for t in tables], torch.float32)
    assert int(err.item()) == 1 and float(o[0, 0, 128:384].abs().max()) == 0.0
    with pytest.raises(ValueError):
        cpm.ops.cp_embed(idx.int(), tables, torch.float32)
    empty = cpm.ops.cp_embed(idx[:0], tables, torch.float32)
    assert empty.shape == (0, 77, 1216)


def test_add_pe_and_dropout(cuda, cpm):
    pe = ft.sinusoidal_pe(300, 128).to(cuda)
    x = torch.randn(4, 50, 128, device=cuda)
    y = cpm.ops.add_pe(x, pe, 50, 0, None, 0.0)
    assert torch.equal(y, x + pe[:, :50])
    y = cpm.ops.add_pe(x, pe, 50, 7, None, 0.0)
    assert torch.equal(y, x + pe[:, 7:57])
    pos = torch.tensor([33], dtype=torch.int32, device=cuda)
    y1 = cpm.ops.add_pe(x[:, :1].contiguous(), pe, 1, 0, pos, 0.0)          # recurrent: one position for all rows
    assert torch.equal(y1, x[:, :1] + pe[:, 33:34])
    # dropout: keep-rate, scaling, and the backward mask equals the forward mask
    cpm.manual_seed(1234)
    big = torch.ones(64, 512, 128, device=cuda, requires_grad=True)
    pe0 = torch.zeros(1, 512, 128, device=cuda)
    yd = cpm.ops.add_pe(big, pe0, 512, 0, None, 0.1)
    keep = (yd != 0).float().mean().item()
    assert abs(keep - 0.9) < 2e-3
    kept_vals = yd[yd != 0]
    assert torch.allclose(kept_vals, torch.full_like(kept_vals, 65536.0 / (65536.0 - 6554.0)))
    yd.backward(torch.ones_like(yd))
    assert torch.equal(big.grad != 0, yd != 0)
    cpm.manual_seed(1234)
    yd2 = cpm.ops.add_pe(big.detach(), pe0, 512, 0, None, 0.1)
    assert torch.equal(yd2, yd.detach())                                   # reproducible from the seed


# ------------------------------------------------------------------ LayerNorm / GELU
@pytest.mark.parametrize("d,dtype,tol", [(512, torch.float32, 2e-5), (128, torch.float32, 2e-5), (1024, torch.float32, 2e-5),
                                         (512, torch.bfloat16, 3e-2)])
def test_ln_residual(cuda, cpm, d, dtype, tol):
    gen = torch.Generator().manual_seed(30 + d)
    rows = 777
    x, r, go = (torch.randn(rows, d, generator=gen).to(cuda).to(dtype) for _ in range(3))
    gamma = (1 + 0.1 * torch.randn(d, generator=gen)).to(cuda).requires_grad_()
    beta = (0.1 * torch.randn(d, generator=gen)).to(cuda).requires_grad_()
    x.requires_grad_(), r.requires_grad_()
    y = cpm.ops.ln_residual(x, r, gamma, beta)
    y.backward(go)
    xr, rr = x.detach().double().requires_grad_(), r.detach().double().requires_grad_()
    gr, br = gamma.detach().double().requires_grad_(), beta.detach().double().requires_grad_()
    yr = torch.nn.functional.layer_norm(xr + rr, (d,), gr, br, 1e-5)
    yr.backward(go.double())
    _cmp(y, yr, tol, tol, "y")
    _cmp(x.grad, xr.grad, tol * 3, tol * 3, "gx")
    _cmp(r.grad, rr.grad, tol * 3, tol * 3, "gres")
    gtol = 2e-3 if dtype == torch.float32 else 0.3
    _cmp(gamma.grad, gr.grad, gtol, 1e-3 if dtype == torch.float32 else 3e-2, "dgamma")
    _cmp(beta.grad, br.grad, gtol, 1e-3 if dtype == torch.float32 else 3e-2, "dbeta")
    # residual-branch bias folded in (the producing Linear runs bias-less); its gradient is a by-product of the backward
    rb = (0.5 * torch.randn(d, generator=gen)).to(cuda).requires_grad_()
    x2, r2 = x.detach().clone().requires_grad_(), r.detach().clone().requires_grad_()
    y3 = cpm.ops.ln_residual(x2, r2, gamma.detach(), beta.detach(), res_bias=rb)
    y3.backward(go)
    xr2, rr2, rbr = x.detach().double().requires_grad_(), r.detach().double().requires_grad_(), rb.detach().double().requires_grad_()
    yr3 = torch.nn.functional.layer_norm(xr2 + rr2 + rbr, (d,), gr.detach(), br.detach(), 1e-5)
    yr3.backward(go.double())
    _cmp(y3, yr3, tol, tol, "y with res_bias")
    _cmp(r2.grad, rr2.grad, tol * 3, tol * 3, "gres with res_bias")
    _cmp(rb.grad, rbr.grad, gtol, 1e-3 if dtype == torch.float32 else 3e-2, "dres_bias")
    # no residual (the encoder's final norm)
    y2 = cpm.ops.ln_residual(x.detach(), None, gamma.detach(), beta.detach())
    _cmp(y2, torch.nn.functional.layer_norm(x.detach().double(), (d,), gr.detach(), br.detach(), 1e-5), tol, tol, "y no-res")


def test_ln_residual_dropout_consistency(cuda, cpm):
    cpm.manual_seed(77)
    d, rows = 512, 4096
    x = torch.zeros(rows, d, device=cuda, requires_grad=True)
    r = torch.ones(rows, d, device=cuda, requires_grad=True)
    gamma, beta = torch.ones(d, device=cuda, requires_grad=True), torch.zeros(d, device=cuda, requires_grad=True)
    y = cpm.ops.ln_residual(x, r, gamma, beta, 1e-5, 0.1)
    dropped = y < 0                      # s is 0 (dropped) or 1.11 (kept): after LN dropped entries are negative
    assert abs(dropped.float().mean().item() - 0.1) < 3e-3
    y.backward(torch.randn_like(y))
    assert torch.equal(r.grad == 0, dropped)          # gradient is masked exactly where the forward dropped


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 2e-2)])
def test_gelu(cuda, cpm, dtype, tol):
    gen = torch.Generator().manual_seed(40)
    x = (3 * torch.randn(333, 2048, generator=gen)).to(cuda).to(dtype).requires_grad_()
    go = torch.randn(333, 2048, generator=gen).to(cuda).to(dtype)
    y = cpm.ops.gelu_dropout(x)
    y.backward(go)
    xr = x.detach().double().requires_grad_()
    yr = torch.nn.functional.gelu(xr)                 # exact erf form (ft activation='gelu')
    yr.backward(go.double())
    _cmp(y, yr, tol, tol, "gelu")
    _cmp(x.grad, xr.grad, tol * 2, tol * 2, "dgelu")
    # bias folded in (linear1 runs bias-less); the bias gradient comes out of the backward kernel's column sums
    b = torch.randn(2048, generator=gen).to(cuda).requires_grad_()
    x2 = x.detach().clone().requires_grad_()
    y2 = cpm.ops.gelu_dropout(x2, 0.0, bias=b)
    y2.backward(go)
    xr2, brf = x.detach().double().requires_grad_(), b.detach().double().requires_grad_()
    yr2 = torch.nn.functional.gelu(xr2 + brf)
    yr2.backward(go.double())
    _cmp(y2, yr2, tol, tol, "gelu(x + bias)")
    _cmp(x2.grad, xr2.grad, tol * 2, tol * 2, "dgelu with bias")
    _cmp(b.grad, brf.grad, 2e-3 if dtype == torch.float32 else 0.3, 1e-3 if dtype == torch.float32 else 3e-2, "dbias (fused column sums)")


# ------------------------------------------------------------------ heads: decode / logp / CE
def test_heads_greedy_bit_exact_and_sampling_golden(cuda, cpm, golden):
    g = golden("sampling")
    logits = torch.from_numpy(g["logits"]).to(cuda)
    seg = [int(s) for s in g["seg"]]
    tok, lp, ent = cpm.ops.heads_sample(logits, seg, greedy=True, want_logp=True, want_entropy=True)
    assert np.array_equal(tok.cpu().numpy(), g["greedy"])                    # bit-exact indices
    ref_lp = torch.stack([torch.log_softmax(logits[:, seg[a]:seg[a + 1]].double(), -1).max(-1).values for a in range(6)], -1)
    _cmp(lp, ref_lp, 1e-5, 1e-5, "greedy logp")
    ref_ent = torch.stack([torch.distributions.Categorical(logits=logits[:, seg[a]:seg[a + 1]].double()).entropy() for a in range(6)], -1)
    _cmp(ent, ref_ent, 1e-5, 1e-5, "entropy")
    t = [so.SAMPLING_CFG[a][0] for a in so.ATTRS]
    p = [so.SAMPLING_CFG[a][1] for a in so.ATTRS]
    tok, _, _ = cpm.ops.heads_sample(logits, seg, t, p, greedy=False, seed=int(g["seed"]), seq_base=int(g["seq_base"]), step=int(g["step"]))
    got = tok.cpu().numpy()
    safe = g["margin"] > 1e-5          # a draw within 1e-5 of a CDF edge may legitimately round either way in fp32
    assert safe.mean() > 0.95
    assert np.array_equal(got[safe], g["sampled"][safe])
    step_dev = torch.tensor([int(g["step"])], dtype=torch.int32, device=cuda)
    tok2, _, _ = cpm.ops.heads_sample(logits, seg, t, p, greedy=False, seed=int(g["seed"]), seq_base=int(g["seq_base"]), step_dev=step_dev)
    assert torch.equal(tok, tok2)                                            # device-side step counter path
    # ties resolve to the first maximal index, like torch.argmax / np.argmax
    tie = torch.zeros(2, seg[-1], device=cuda)
    tk, _, _ = cpm.ops.heads_sample(tie, seg, greedy=True)
    assert int(tk.abs().max()) == 0


def test_heads_sampling_distribution(cuda, cpm):
    """Nucleus / temperature draws follow the oracle's candidate distribution (chi-square)."""
    rng = np.random.RandomState(50)
    seg = [0, 18, 43]
    lg = (rng.randn(1, 43) * 1.5).astype(np.float32)
    rows = 40000
    logits = torch.from_numpy(lg).to(cuda).expand(rows, -1).contiguous()
    tok, _, _ = cpm.ops.heads_sample(logits, seg, [1.2, 2.0], [None, 0.9], greedy=False, seed=99, seq_base=0, step=3)
    tok = tok.cpu().numpy()
    for a, (t, p) in enumerate(((1.2, None), (2.0, 0.9))):
        probs = so.softmax_with_temperature(lg[0, seg[a]:seg[a + 1]], t)
        if p is None:
            cand, cp = np.arange(len(probs)), probs / probs.sum()
        else:
            cand, cp = so.nucleus_candidates(probs, p)
        counts = np.bincount(tok[:, a], minlength=len(probs)).astype(np.float64)
        assert counts[np.setdiff1d(np.arange(len(probs)), cand)].sum() == 0      # nothing outside the nucleus
        exp = np.zeros(len(probs))
        exp[cand] = cp * rows
        m = exp > 5
        chi2 = (((counts - exp) ** 2)[m] / exp[m]).sum()
        assert chi2 < 3.0 * m.sum() + 30, (chi2, m.sum())


def test_heads_logp_fwd_bwd(cuda, cpm):
    gen = torch.Generator().manual_seed(51)
    seg = cpm.ops.seg_offsets(VOCAB)
    W = 344
    logits = torch.randn(9, 25, W, generator=gen).to(cuda).requires_grad_()
    tokens = torch.stack([torch.randint(0, n, (9, 25), generator=gen) for n in VOCAB], -1).to(cuda)
    glp, gen_ = torch.randn(9, 25, 6, generator=gen).to(cuda), torch.randn(9, 25, 6, generator=gen).to(cuda)
    lp, ent = cpm.ops.heads_logp(logits, tokens, seg, True)
    ((lp * glp).sum() + (ent * gen_).sum()).backward()
    lr = logits.detach().double().requires_grad_()
    lps, ents = [], []
    for a in range(6):
        ls = torch.log_softmax(lr[..., seg[a]:seg[a + 1]], -1)
        lps.append(ls.gather(-1, tokens[..., a:a + 1])[..., 0])
        ents.append(-(ls.exp() * ls).sum(-1))
    lpr, entr = torch.stack(lps, -1), torch.stack(ents, -1)
    ((lpr * glp.double()).sum() + (entr * gen_.double()).sum()).backward()
    _cmp(lp, lpr, 1e-5, 1e-5, "logp")
    _cmp(ent, entr, 1e-5, 1e-5, "entropy")
    _cmp(logits.grad, lr.grad, 1e-5, 1e-4, "dlogits")
    assert float(logits.grad[..., seg[-1]:].abs().max()) == 0.0        # pad columns get zero gradient


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
def test_masked_ce(cuda, cpm, dtype, tol):
    gen = torch.Generator().manual_seed(52)
    seg = cpm.ops.seg_offsets(VOCAB)
    N, L, W = 3, 61, 344
    logits = (2 * torch.randn(N, L, W, generator=gen)).to(cuda).to(dtype).requires_grad_()
    tgt = torch.stack([torch.randint(0, n, (N, L), generator=gen) for n in VOCAB], -1).to(cuda)
    mask = (torch.rand(N, L, generator=gen) < 0.7).float().to(cuda)
    w = torch.randn(6, generator=gen).to(cuda)
    losses = cpm.ops.masked_ce(logits, tgt, mask, seg)
    (losses * w).sum().backward()
    lr = logits.detach().double().requires_grad_()
    ref = []
    for a in range(6):
        ce = torch.nn.functional.cross_entropy(lr[..., seg[a]:seg[a + 1]].permute(0, 2, 1), tgt[..., a], reduction="none")
        ref.append((ce * mask.double()).sum() / mask.double().sum())            # compute_loss, agent_pretrain.py:279-283
    ref = torch.stack(ref)
    (ref * w.double()).sum().backward()
    _cmp(losses, ref, tol, tol, "losses")
    _cmp(logits.grad, lr.grad, tol * 0.02 + 1e-7, 2e-2 if dtype == torch.bfloat16 else 1e-4, "dlogits")
    # long (int64) masks are what ppo_train.py:398 passes
    l2 = cpm.ops.masked_ce(logits.detach(), tgt, mask.long(), seg)
    _cmp(l2, ref, tol, tol, "long mask")


# ------------------------------------------------------------------ RL kernels
def test_returns_advantages_compat_golden(cuda, cpm, golden):
    g = golden("rl")
    rw, va = torch.from_numpy(g["rewards"]).to(cuda), torch.from_numpy(g["values"]).to(cuda)
    raw = cpm.rl.calculate_returns_compat(rw, 0.99, normalize=False)
    _cmp(raw, g["ret_raw"], 1e-5, 1e-5, "returns raw")
    ret = cpm.rl.calculate_returns_compat(rw, 0.99)
    _cmp(ret, g["ret"], 2e-5, 1e-5, "returns")
    adv = cpm.rl.calculate_advantages_compat(ret, va)
    _cmp(adv, g["adv"], 3e-5, 1e-5, "advantages")
    assert ret.shape == (30, 1) and adv.shape == (30, 1)


@pytest.mark.parametrize("B,T", [(4, 77), (1, 1), (3, 32), (2, 33), (5, 1024)])
def test_scans_standard(cuda, cpm, golden, B, T):
    if (B, T) == (4, 77):
        g = golden("rl")
        r, v, d, lv = (torch.from_numpy(g[n]).to(cuda) for n in ("r2", "v2", "d2", "lv"))
        adv, ret = cpm.ops.returns_scan(r, 0.99, "gae", values=v, dones=d, last_value=lv, lam=0.95)
        _cmp(adv, g["gae_adv"], 2e-5, 1e-5, "gae adv")
        _cmp(ret, g["gae_ret"], 2e-5, 1e-5, "gae ret")
        _cmp(cpm.ops.returns_scan(r, 0.99, "togo", dones=d), g["togo"], 2e-5, 1e-5, "togo")
        return
    gen = torch.Generator().manual_seed(B * 1000 + T)
    r, v = torch.rand(B, T, generator=gen), torch.randn(B, T, generator=gen)
    d = (torch.rand(B, T, generator=gen) < 0.1).float()
    lv = torch.randn(B, generator=gen)
    adv, ret = cpm.ops.returns_scan(r.to(cuda), 0.99, "gae", values=v.to(cuda), dones=d.to(cuda), last_value=lv.to(cuda), lam=0.95)
    ra, rr = rl.gae_standard(r.double(), v.double(), d.double(), lv.double(), 0.99, 0.95)
    _cmp(adv, ra, 1e-4, 1e-5, "gae adv")
    _cmp(ret, rr, 1e-4, 1e-5, "gae ret")
    _cmp(cpm.ops.returns_scan(r.to(cuda), 0.99, "togo", dones=d.to(cuda)), rl.rewards_to_go_standard(r.double(), d.double(), 0.99), 1e-4, 1e-5, "togo")
    _cmp(cpm.ops.returns_scan(r[:1].to(cuda), 0.97, "compat").reshape(-1, 1), rl.calculate_returns_compat(r[0].double(), 0.97, False), 1e-4, 1e-5, "compat")


def test_ppo_losses_golden(cuda, cpm, golden):
    g = golden("rl")
    nl = torch.from_numpy(g["new_logp"]).to(cuda).requires_grad_()
    loss = cpm.rl.ppo_policy_loss_compat(nl, torch.from_numpy(g["old_long"]).to(cuda).long(), torch.from_numpy(g["adv"]).to(cuda))
    loss.backward()
    _cmp(loss, g["ploss"], 1e-6, 1e-5, "compat loss")
    _cmp(nl.grad, g["dnew"], 1e-7, 1e-4, "compat dnew")
    n2, en, va = (torch.from_numpy(g[k]).to(cuda).requires_grad_() for k in ("nl", "en", "va"))
    out = cpm.ops.ppo_loss_standard(n2, torch.from_numpy(g["ol"]).to(cuda), torch.from_numpy(g["ad"]).to(cuda), va,
                                    torch.from_numpy(g["rt"]).to(cuda), en)
    out[0].backward()
    _cmp(out, g["std_losses"], 1e-5, 1e-5, "standard losses")
    _cmp(n2.grad, g["d_nl"], 1e-7, 1e-4, "d new_logp")
    _cmp(en.grad, g["d_en"], 1e-8, 1e-4, "d entropy")
    _cmp(va.grad, g["d_va"], 1e-7, 1e-4, "d value")


def test_dqn_td_golden(cuda, cpm, golden):
    g = golden("rl")
    seg = [int(s) for s in g["seg"]]
    act, rw, dn = (torch.from_numpy(g[k]).to(cuda) for k in ("action", "rw", "dn"))
    nq = torch.from_numpy(g["nq"]).to(cuda)
    for compat, lk, gk in ((True, "td_compat", "gq_compat"), (False, "td_standard", "gq_standard")):
        ql = torch.from_numpy(g["ql"]).to(cuda).requires_grad_()
        loss, tg = cpm.ops.dqn_td_loss(ql, nq, act, rw, dn, seg, 25, 0.95, compat)
        loss.backward()
        _cmp(loss, g[lk], 1e-5, 1e-5, lk)
        _cmp(ql.grad, g[gk], 1e-7, 1e-4, gk)
        assert tg.shape == (6, 25, 6)
    with pytest.raises(ValueError, match="B"):        # the reference's gather needs B <= L too
        big = torch.zeros(60, 50, seg[-1], device=cuda)
        cpm.ops.dqn_td_loss(big, big, torch.zeros(60, 25, 6, dtype=torch.long, device=cuda), torch.zeros(60, device=cuda),
                            torch.zeros(60, device=cuda), seg, 25, 0.95, True)


def test_dqn_td_replay_batch_1024(cuda, cpm):
    """BASELINE cfg4 size (1024 windows of 50) in the standard mode + bf16; property: the loss
    equals the mean of per-sample losses computed one sample at a time (batch independence)."""
    gen = torch.Generator().manual_seed(60)
    seg = cpm.ops.seg_offsets(VOCAB)
    B, L, W = 1024, 50, 344
    ql = torch.randn(B, L, W, generator=gen).to(cuda).bfloat16()
    nq = torch.randn(B, L, W, generator=gen).to(cuda).bfloat16()
    act = torch.stack([torch.randint(0, n, (B, 25), generator=gen) for n in VOCAB], -1).to(cuda)
    rw, dn = torch.rand(B, generator=gen).to(cuda), (torch.rand(B, generator=gen) < 0.2).float().to(cuda)
    loss, _ = cpm.ops.dqn_td_loss(ql, nq, act, rw, dn, seg, 25, 0.95, False)
    sub = [cpm.ops.dqn_td_loss(ql[i:i + 64], nq[i:i + 64], act[i:i + 64], rw[i:i + 64], dn[i:i + 64], seg, 25, 0.95, False)[0]
           for i in range(0, B, 64)]
    _cmp(loss, torch.stack(sub).mean(), 1e-4, 1e-4, "batch independence")
    split = lambda t: [t[..., seg[i]:seg[i + 1]].double().cpu() for i in range(6)]
    ref = rl.dqn_td_loss_standard(split(ql[:32]), split(nq[:32]), act[:32].cpu(), rw[:32, None].double().cpu(), dn[:32, None].double().cpu())
    _cmp(cpm.ops.dqn_td_loss(ql[:32], nq[:32], act[:32], rw[:32], dn[:32], seg, 25, 0.95, False)[0], ref, 1e-4, 1e-4, "vs oracle")


# ------------------------------------------------------------------ tcgen05 / TMA path
@pytest.mark.parametrize("impl", [3])
@pytest.mark.parametrize("shape", [(1, 128, 1), (2, 256, 4), (3, 384, 2), (1, 2048, 2), (32, 512, 8)])
def test_linattn_tc_fwd(cuda, cpm, shape, impl):
    """tcgen05 forward (impl=3 chunk-parallel; bf16) vs the fp64 oracle on the same bf16-rounded inputs and vs the
    SIMT kernel.  The tensor-core path rounds P (intra-chunk scores) and the carried state to bf16
    before the second MMA, so its tolerance is a little wider than the fp32-math SIMT path:
    3e-2 absolute on O(1) outputs; the normaliser `den` within 1e-2 relative."""
    N, L, H = shape
    gen = torch.Generator().manual_seed(L + H)
    q, k, v = (torch.randn(N, L, H, 64, generator=gen).to(cuda).bfloat16() for _ in range(3))
    out, den = cpm.ops.linattn_fwd_raw(q, k, v, impl=impl)
    assert cpm.ops.linattn_last_impl() == ("tcgen05" if impl == 2 else ("tcgen05-cp-stream" if N * H >= 96 else "tcgen05-cp"))
    torch.cuda.synchronize()
    ref_simt, den_simt = cpm.ops.linattn_fwd_raw(q, k, v, impl=1)
    _cmp(out, ref_simt.float(), 3e-2, 2e-2, "tc vs simt")
    _cmp(den, den_simt, 1e-3, 1e-2, "den tc vs simt")
    if N * L * H <= 8192:
        _cmp(out, _oracle_attn(q.float(), k.float(), v.float()), 3e-2, 2e-2, "tc vs oracle")


def test_linattn_tc_fwd_fused_layout_and_autograd(cuda, cpm):
    """Default dispatch (impl=0) takes the tensor-core forward for bf16, L%128==0, on the fused QKV
    layout (token stride 3*H*64), and the backward (SIMT) consumes its saved out/den."""
    N, L, H = 2, 256, 8
    gen = torch.Generator().manual_seed(77)
    qkv = torch.randn(N, L, 3 * H * 64, generator=gen).to(cuda).bfloat16().requires_grad_()
    go = torch.randn(N, L, H * 64, generator=gen).to(cuda).bfloat16()
    out = cpm.ops.causal_linear_attention_fused(qkv, H)
    out.backward(go)
    q, k, v = (qkv.detach()[..., i * H * 64:(i + 1) * H * 64].reshape(N, L, H, 64).float() for i in range(3))
    ro, rq, rk, rv = _oracle_attn(q, k, v, go.view(N, L, H, 64).float())
    _cmp(out.view(N, L, H, 64), ro, 3e-2, 2e-2, "out")
    _cmp(qkv.grad, torch.cat([t.reshape(N, L, H * 64) for t in (rq, rk, rv)], -1), 5e-2, 3e-2, "gqkv")


@pytest.mark.parametrize("impl", [3])
@pytest.mark.parametrize("shape", [(1, 128, 1), (2, 256, 4), (3, 384, 2), (1, 2048, 2), (32, 512, 8)])
def test_linattn_tc_bwd(cuda, cpm, shape, impl):
    """tcgen05 backward (impl=3 chunk-parallel, once rebuilding the prefix
    states and once reading the ones the forward saved) vs the fp64 oracle (small shapes) and vs the SIMT backward fed the
    same saved out/den.  Tolerance: gradients are O(0.1-1); 4e-2 absolute + 3e-2 relative (bf16
    rounding of G', of the masked score tiles and of the carried state)."""
    N, L, H = shape
    gen = torch.Generator().manual_seed(L * 3 + H)
    q, k, v, go = (torch.randn(N, L, H, 64, generator=gen).to(cuda).bfloat16() for _ in range(4))
    saved = cpm.ops.linattn_saved(N, L, H, cuda) if impl == 3 else None
    out, den = cpm.ops.linattn_fwd_raw(q, k, v, impl=impl, saved=saved)
    gq, gk, gv = (torch.empty_like(q) for _ in range(3))
    cpm.ops.linattn_bwd_raw(q, k, v, out, den, go, gq, gk, gv, impl=impl)
    assert cpm.ops.linattn_last_impl() == ("tcgen05" if impl == 2 else ("tcgen05-cp-stream" if N * H >= 96 else "tcgen05-cp"))
    torch.cuda.synchronize()
    if impl == 3:          # same result when backward reads the forward's saved prefix states
        tq, tk, tv = (torch.empty_like(q) for _ in range(3))
        cpm.ops.linattn_bwd_raw(q, k, v, out, den, go, tq, tk, tv, impl=3, saved=saved)
        for a, b in ((gq, tq), (gk, tk), (gv, tv)):
            assert torch.equal(a, b)
    sq, sk, sv = (torch.empty_like(q) for _ in range(3))
    cpm.ops.linattn_bwd_raw(q, k, v, out, den, go, sq, sk, sv, impl=1)
    for name, a, b in (("gq", gq, sq), ("gk", gk, sk), ("gv", gv, sv)):
        _cmp(a, b.float(), 4e-2, 3e-2, f"tc vs simt {name}")
    if N * L * H <= 8192:
        _, rq, rk, rv = _oracle_attn(q.float(), k.float(), v.float(), go.float())
        for name, a, b in (("gq", gq, rq), ("gk", gk, rk), ("gv", gv, rv)):
            _cmp(a, b, 4e-2, 3e-2, f"tc vs oracle {name}")


def _oracle_chains(q, k, v, go, pairs):
    """fp64 oracle (ft CausalLinearAttention restated, oracle/ft_oracle.py) on selected (batch, head) chains only: chains are
    independent, so any subset can be checked at sizes where the whole tensor would not fit the quadratic restatement."""
    sel = lambda t: torch.stack([t[n, :, h] for n, h in pairs], 0)[:, :, None, :].float()      # (P, L, 1, 64)
    return _oracle_attn(sel(q), sel(k), sel(v), sel(go))


@pytest.mark.parametrize("shape", [(16, 256, 8), (12, 384, 8), (100, 128 * 3, 1)])
def test_linattn_streaming_kernels_vs_oracle(cuda, cpm, shape):
    """The DEFAULT bench path: with N*H >= 96 chains the forward prefix and backward suffix states come from the STREAMING
    kernels (cp_prefix_stream_fwd / cp_suffix_stream_bwd: one CTA per chain, S and z carried in tensor memory), not from the
    per-chunk state kernels + scan the small shapes above exercise.  Held directly to the fp64 oracle on every chain."""
    N, L, H = shape
    gen = torch.Generator().manual_seed(N * 5 + L)
    q, k, v, go = (torch.randn(N, L, H, 64, generator=gen).to(cuda).bfloat16() for _ in range(4))
    saved = cpm.ops.linattn_saved(N, L, H, cuda)
    out, den = cpm.ops.linattn_fwd_raw(q, k, v, impl=0, saved=saved)
    assert cpm.ops.linattn_last_impl() == "tcgen05-cp-stream"
    gq, gk, gv = (torch.empty_like(q) for _ in range(3))
    cpm.ops.linattn_bwd_raw(q, k, v, out, den, go, gq, gk, gv, impl=0, saved=saved)
    assert cpm.ops.linattn_last_impl() == "tcgen05-cp-stream"
    pairs = [(n, h) for n in range(N) for h in range(H)]
    for i in range(0, len(pairs), 32):
        pp = pairs[i:i + 32]
        ro, rq, rk, rv = _oracle_chains(q, k, v, go, pp)
        pick = lambda t: torch.stack([t[n, :, h] for n, h in pp], 0)[:, :, None, :]
        _cmp(pick(out), ro, 3e-2, 2e-2, "out vs oracle")
        for name, a, b in (("gq", gq, rq), ("gk", gk, rk), ("gv", gv, rv)):
            _cmp(pick(a), b, 4e-2, 3e-2, f"{name} vs oracle")


@pytest.mark.parametrize("shape,n_chains", [((128, 1024, 8), 12), ((32, 512, 8), 12), ((1, 8192, 16), 2)])
def test_linattn_full_size_sampled_chains_vs_oracle(cuda, cpm, shape, n_chains):
    """BASELINE shapes as the bench runs them - the update minibatch 128 x 1024 x 8 (cfg3), the pretraining batch 32 x 512 x 8
    (cfg2) and the long sequence 8192 x 16 heads (cfg5) - compared with the fp64 oracle on a seeded sample of (batch, head)
    chains (forward and all three gradients).  cfg5 has 16 chains -> per-chunk state kernels + scan; the others stream."""
    N, L, H = shape
    gen = torch.Generator().manual_seed(L + 17 * N)
    q, k, v, go = (torch.randn(N, L, H, 64, generator=gen).to(cuda).bfloat16() for _ in range(4))
    saved = cpm.ops.linattn_saved(N, L, H, cuda)
    out, den = cpm.ops.linattn_fwd_raw(q, k, v, impl=0, saved=saved)
    gq, gk, gv = (torch.empty_like(q) for _ in range(3))
    cpm.ops.linattn_bwd_raw(q, k, v, out, den, go, gq, gk, gv, impl=0, saved=saved)
    assert cpm.ops.linattn_last_impl() == ("tcgen05-cp-stream" if N * H >= 96 else "tcgen05-cp")
    idx = torch.randperm(N * H, generator=gen)[:n_chains].tolist()
    scale = 1.0 + math.sqrt(L / 1024.0)          # gradients accumulate over the sequence: absolute tolerance grows ~ sqrt(L)
    for i in idx:
        pp = [(i // H, i % H)]
        ro, rq, rk, rv = _oracle_chains(q, k, v, go, pp)
        pick = lambda t: t[pp[0][0], :, pp[0][1]][None, :, None, :]
        _cmp(pick(out), ro, 3e-2, 2e-2, f"out chain {pp}")
        for name, a, b in (("gq", gq, rq), ("gk", gk, rk), ("gv", gv, rv)):
            _cmp(pick(a), b, 4e-2 * scale, 3e-2, f"{name} chain {pp}")


def test_rowdot_fwd_bwd(cuda, cpm):
    """out = h . u + c (fp32) and its gradients vs float64 (the collapsed critic read-out)."""
    gen = torch.Generator().manual_seed(4)
    for dtype, tol in ((torch.float32, 1e-4), (torch.bfloat16, 2e-2)):
        h = torch.randn(3, 700, 512, generator=gen).to(cuda).to(dtype).requires_grad_()
        u = (torch.randn(512, generator=gen) / 20).to(cuda).requires_grad_()
        c = torch.tensor(0.3, device=cuda, requires_grad=True)
        g = torch.randn(3, 700, generator=gen).to(cuda)
        out = cpm.ops.rowdot(h, u, c)
        out.backward(g)
        hd, ud, cd = h.detach().double().cpu().requires_grad_(), u.detach().double().cpu().requires_grad_(), c.detach().double().cpu().requires_grad_()
        ref = hd @ ud + cd
        ref.backward(g.double().cpu())
        _cmp(out, ref, 1e-4, 1e-5, "rowdot")
        _cmp(h.grad, hd.grad, tol * 0.1, tol, "dh")
        _cmp(u.grad, ud.grad, 2e-3, 1e-4, "du")
        _cmp(c.grad, cd.grad, 1e-3, 1e-5, "dc")


# ------------------------------------------------------------------ reward head
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 5e-3)])
def test_reward_head(cuda, cpm, dtype, tol):
    """Fused reward read-out vs the reference formula written out (ppo_policy/model.py:474-493) in fp64 on the same
    hidden states: N=30 windows of 50 tokens, d 512 (ppo_train.py:491)."""
    torch.manual_seed(8)
    head = cpm.rl.RewardHead([49, 19, 19, 89, 67, 25], d_model=512).to(cuda)
    h = torch.randn(30, 50, 512, device=cuda).to(dtype)
    reward, scores = head(h, want_scores=True)
    hd, ref_scores = h.double(), []
    for a in head.ATTRS:
        proj, ev = getattr(head, f"proj_{a}").double(), getattr(head, f"eval_{a}").double()
        ref_scores.append(torch.sigmoid(ev(proj(hd)).mean(dim=1)))
    head.float()
    ref_scores = torch.cat(ref_scores, -1)
    _cmp(scores, ref_scores, tol, 0.0, "scores")
    _cmp(reward, ref_scores.mean(-1), tol, 0.0, "reward")
    assert reward.shape == (30,) and bool(((reward > 0) & (reward < 1)).all())


@pytest.mark.parametrize("rows,width,dtype", [(65536, 2048, torch.bfloat16), (1000, 512, torch.bfloat16), (37, 344, torch.float32),
                                              (1, 8, torch.float32), (5000, 1536, torch.bfloat16)])
def test_colsum_bias_gradient(cuda, cpm, rows, width, dtype):
    """cpm_colsum (bias gradients of the Linear layers) against an fp64 sum of the same (rounded) values, incl. a strided
    view (column slice of a wider matrix) and row counts that do not divide the row slabs."""
    gen = torch.Generator().manual_seed(rows + width)
    x = torch.randn(rows, width + 16, generator=gen).to(cuda).to(dtype)
    for view in (x[:, :width].contiguous(), x[:, 8:8 + width]):
        got = cpm.ops.colsum(view)
        ref = view.double().sum(0)
        _cmp(got, ref, 1e-5 * math.sqrt(rows) + 1e-6, 1e-5, "colsum")
        assert torch.equal(got, cpm.ops.colsum(view))            # deterministic


@pytest.mark.parametrize("shape", [(3, 50, 4), (2, 130, 2), (1, 1, 1), (5, 300, 3), (16, 200, 8), (13, 127, 8), (1024, 50, 8)])
def test_linattn_bf16_ragged_lengths_run_on_tensor_cores(cuda, cpm, shape):
    """bf16 sequences that are not a multiple of 128 tokens (the 50-token DQN windows) take the tcgen05 kernels as they are: the
    tile of a short last chunk runs on into the next sequence's rows (or past the tensor), and the kernels keep those rows out -
    G' = 0, gd = 0, nothing stored.  Forward and gradients against the fp64 oracle at the TRUE length (one short chunk, several
    chunks with a short last one, the streaming state kernels at >= 96 chains, a single token) and against the CUDA-core kernels
    at the replay-batch size.  The buffers are surrounded by NaN canaries: a row written past a sequence's end, or a neighbour's
    rows leaking into a result, would show."""
    N, L, H = shape
    gen = torch.Generator().manual_seed(L * 3 + H)
    qkv = torch.randn(N, L, 3 * H * 64, generator=gen).to(cuda).bfloat16().requires_grad_()
    go = torch.randn(N, L, H * 64, generator=gen).to(cuda).bfloat16()
    out = cpm.ops.causal_linear_attention_fused(qkv, H)
    want = "tcgen05-cp-stream" if (N * H >= 96 and L > 128) else "tcgen05-cp"
    assert cpm.ops.linattn_last_impl() == want and out.shape == (N, L, H * 64)
    out.backward(go)
    assert bool(torch.isfinite(out).all()) and bool(torch.isfinite(qkv.grad).all())
    if N * L <= 4096:
        q, k, v = (qkv.detach()[..., i * H * 64:(i + 1) * H * 64].reshape(N, L, H, 64).float() for i in range(3))
        ro, rq, rk, rv = _oracle_attn(q, k, v, go.view(N, L, H, 64).float())
        _cmp(out.view(N, L, H, 64), ro, 2e-2, 1e-2, "out")
        _cmp(qkv.grad, torch.cat([t.reshape(N, L, H * 64) for t in (rq, rk, rv)], -1), 4e-2, 3e-2, "gqkv")
    else:
        ref_in = qkv.detach().clone().requires_grad_()
        ref = cpm.ops.causal_linear_attention_fused(ref_in, H, impl=1)
        assert cpm.ops.linattn_last_impl() == "simt"
        ref.backward(go)
        _cmp(out, ref.float(), 3e-2, 2e-2, "out vs simt")
        _cmp(qkv.grad, ref_in.grad.float(), 4e-2, 3e-2, "gqkv vs simt")
    # raw calls on views INSIDE larger NaN-filled buffers: nothing outside the (N, L) block may be touched, nothing of it read into a result
    if N * L <= 4096:
        W = 3 * H * 64
        big = torch.full((N * L + 300, W), float("nan"), device=cuda).bfloat16()
        big[150:150 + N * L] = qkv.detach().view(N * L, W)
        blk = big[150:150 + N * L].view(N, L, W)
        qq, kk, vv = (blk[..., i * H * 64:(i + 1) * H * 64].unflatten(-1, (H, 64)) for i in range(3))
        o2, den2 = cpm.ops.linattn_fwd_raw(qq, kk, vv)
        assert torch.equal(o2.view(N, L, H * 64), out.detach()), "neighbouring rows (NaN) leaked into the forward"
        gbig = torch.full((N * L + 300, W), float("nan"), device=cuda).bfloat16()
        gblk = gbig[150:150 + N * L].view(N, L, W)
        gq, gk, gv = (gblk[..., i * H * 64:(i + 1) * H * 64].unflatten(-1, (H, 64)) for i in range(3))
        cpm.ops.linattn_bwd_raw(qq, kk, vv, o2, den2, go.view(N, L, H, 64), gq, gk, gv)
        assert torch.equal(gblk, qkv.grad), "neighbouring rows (NaN) leaked into the gradients"
        assert bool(torch.isnan(gbig[:150]).all()) and bool(torch.isnan(gbig[150 + N * L:]).all()), "rows outside the block were written"


def test_out_of_range_token_ids_raise_index_error(cuda, cpm):
    """nn.Embedding raises IndexError on an id outside the vocabulary (agent_pretrain.py:185-196); the gather kernel flags it
    and the host raises at the next poll / check instead of training silently on zero rows."""
    m = cpm.TransformerModel([56, 135, 18, 87, 18, 25], d_model=128, n_layer=1, n_head=2, d_inner=256, dropout=0.0).to(cuda)
    x = torch.zeros(2, 8, 6, dtype=torch.int64, device=cuda)
    cpm.ops.IndexGuard.reset(cuda)                       # earlier tests feed bad ids to the raw kernel on purpose
    m.train_step(x, x, torch.ones(2, 8, device=cuda))
    cpm.ops.IndexGuard.check(cuda)                       # clean input: nothing raised
    x[1, 3, 1] = 135
    m.train_step(x, x.clamp(max=17), torch.ones(2, 8, device=cuda))
    with pytest.raises(IndexError):
        cpm.ops.IndexGuard.check(cuda)
    x[1, 3, 1] = 0
    m.train_step(x, x, torch.ones(2, 8, device=cuda))
    torch.cuda.synchronize()
    with pytest.raises(IndexError):                      # the non-blocking poll reports a flag raised by an earlier call
        x[0, 0, 0] = -1
        m.train_step(x, x.clamp(min=0), torch.ones(2, 8, device=cuda))
        torch.cuda.synchronize()
        for _ in range(3):
            cpm.ops.IndexGuard.poll(cuda)
            torch.cuda.synchronize()
    cpm.ops.IndexGuard.check(cuda)


def test_row_stores_fall_back_when_rows_are_not_32_byte_aligned(cuda, cpm):
    """The per-chunk attention kernels and the token-step GEMM write a lane's 64 bytes of a row as two 256-bit stores where the
    address allows it and as four 128-bit stores otherwise (row strides that are multiples of 8 but not of 16 elements; tiles cut
    off at N).  Both paths must give the same values: gradients into a buffer of odd row stride equal those into the dense one,
    a ragged length (plain stores instead of the bulk tensor stores) included; same for the small GEMM into a strided output."""
    gen = torch.Generator().manual_seed(9)
    for (N, L, H) in ((3, 256, 2), (2, 200, 2)):
        W = 3 * H * 64
        qkv = torch.randn(N, L, W, generator=gen).to(cuda).bfloat16()
        go = torch.randn(N, L, H, 64, generator=gen).to(cuda).bfloat16()
        q, k, v = (qkv[..., i * H * 64:(i + 1) * H * 64].unflatten(-1, (H, 64)) for i in range(3))
        out, den = cpm.ops.linattn_fwd_raw(q, k, v)
        res = []
        for width in (W, W + 8):                       # W + 8: rows of 3088 / 784 bytes - every other row starts 16 bytes off a 32-byte line
            gbuf = torch.zeros(N, L, width, device=cuda, dtype=torch.bfloat16)
            gq, gk, gv = (gbuf[..., i * H * 64:(i + 1) * H * 64].unflatten(-1, (H, 64)) for i in range(3))
            cpm.ops.linattn_bwd_raw(q, k, v, out, den, go, gq, gk, gv)
            assert not bool(gbuf[..., W:].any())
            res.append(gbuf[..., :W].clone())
        assert torch.equal(res[0], res[1])
    a = torch.randn(200, 512, generator=gen).to(cuda).bfloat16()
    w = (torch.randn(344, 512, generator=gen) / 16).to(cuda).bfloat16()
    b = torch.randn(344, generator=gen).to(cuda)
    dense = cpm.ops.gemm_nt_small(a, w, b)
    wide = torch.zeros(200, 344 + 8, device=cuda, dtype=torch.bfloat16)
    cpm.ops.gemm_nt_small(a, w, b, out=wide[:, :344])
    assert torch.equal(wide[:, :344], dense) and not bool(wide[:, 344:].any())
    ref = (a.float() @ w.float().t() + b)
    assert (dense.float() - ref).abs().max() < 0.06
