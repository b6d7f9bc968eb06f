"""GPU parity tests of the tcgen05 GEMM family (csrc/tc_gemm.cu) through the C-ABI (cpmusic.ops.gemm_nt / gemm_tn_acc).

Oracle: the same contraction in float64 on the CPU from the SAME bf16 operand values (numpy / torch CPU matmul — the
restatement of what nn.Linear computes for the reference's Linear layers, agent_pretrain.py:239,244-253,360-375).  The GEMM
accumulates in fp32 and rounds once to bf16, so the tolerance is bf16's half-ulp (2^-8 relative) plus fp32 accumulation noise.
The GELU epilogues are held to the exact-erf GELU (torch.nn.functional.gelu in float64, what ft's activation='gelu' calls) and
their dropout masks to the standalone cpm_gelu_fwd / cpm_gelu_bwd kernels, which draw from the same Philox streams."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _bf16(shape, gen, scale=1.0):
    return (torch.randn(*shape, generator=gen) * scale).to(torch.bfloat16)


def _assert_close_bf16(got, ref, what, extra_abs=0.0):
    got, ref = got.double().cpu(), ref.double().cpu()
    tol = ref.abs() * 2.0 ** -8 + 1e-3 + extra_abs
    err = (got - ref).abs()
    bad = err > tol
    assert not bool(bad.any()), f"{what}: {int(bad.sum())} of {bad.numel()} elements off, max err {err.max().item():.4e}, ref scale {ref.abs().max().item():.3e}"


# (M, N, K): the agent's layer shapes at small token counts, ragged edges in every dimension, multi-tile / multi-wave cases
NT_SHAPES = [(256, 256, 64), (256 * 76 + 100, 512, 512), (256 * 75, 1536, 448), (256 * 74 + 8, 512, 2048), (256 * 80, 344, 1216), (256, 512, 512), (512, 1536, 512), (384, 2048, 512), (300, 512, 2048), (1000, 344, 512), (130, 512, 1216),
             (77, 264, 344), (8, 512, 512), (256 * 80, 512, 512), (256 * 150 + 40, 256, 128)]


@pytest.fixture(params=[0, 1, 3], ids=["auto", "stream", "wide"])
def gemm_mode(request, cpm, cuda):
    """Every NT test runs under each schedule (requests a shape cannot honour fall back to auto inside the library)."""
    cpm.ops.gemm_set_mode(request.param)
    yield request.param
    cpm.ops.gemm_set_mode(0)


@pytest.mark.parametrize("shape", NT_SHAPES)
def test_gemm_nt_bias_vs_fp64(cuda, cpm, shape, gemm_mode):
    M, N, K = shape
    gen = torch.Generator().manual_seed(M * 31 + N * 7 + K)
    a, b = _bf16((M, K), gen), _bf16((N, K), gen, 1.0 / math.sqrt(K))
    bias = torch.randn(N, generator=gen)
    ref = a.double() @ b.double().t() + bias.double()
    got = cpm.ops.gemm_nt(a.to(cuda), b.to(cuda), bias.to(cuda))
    assert got.shape == (M, N) and got.dtype == torch.bfloat16
    _assert_close_bf16(got, ref, f"gemm_nt {shape}")
    got2 = cpm.ops.gemm_nt(a.to(cuda), b.to(cuda))                         # no bias
    _assert_close_bf16(got2, ref - bias.double(), f"gemm_nt {shape} (no bias)")


def test_gemm_nt_strided_operands_and_output_slice(cuda, cpm, gemm_mode):
    """A as a column slice of a wider buffer (row stride > K) and D written into a column slice of a wider buffer: how q,k,v
    slices of the fused projection output and their gradients are addressed."""
    gen = torch.Generator().manual_seed(5)
    M, N, K = 640, 512, 512
    wide = _bf16((M, 3 * K), gen).to(cuda)
    b = _bf16((N, K), gen, 1.0 / math.sqrt(K)).to(cuda)
    out = torch.full((M, 2 * N), 7.0, dtype=torch.bfloat16, device=cuda)
    a = wide[:, K:2 * K]
    cpm.ops.gemm_nt(a, b, out=out[:, N:])
    ref = a.double().cpu() @ b.double().cpu().t()
    _assert_close_bf16(out[:, N:], ref, "strided gemm_nt")
    assert bool((out[:, :N] == 7.0).all()), "columns outside the output slice were touched"


@pytest.mark.parametrize("p_drop", [0.0, 0.1])
@pytest.mark.parametrize("shape", [(256, 2048, 512), (1000, 2048, 512), (96, 256, 64)])
def test_gemm_nt_gelu_epilogue(cuda, cpm, shape, p_drop, gemm_mode):
    """h = bf16(a W^T + b) and g = dropout(gelu(h)) in one launch == the GEMM followed by cpm_gelu_fwd (same Philox stream ->
    same mask), and gelu itself == the exact-erf GELU in float64."""
    M, N, K = shape
    gen = torch.Generator().manual_seed(N + K)
    a, b = _bf16((M, K), gen).to(cuda), _bf16((N, K), gen, 1.0 / math.sqrt(K)).to(cuda)
    bias = torch.randn(N, generator=gen).to(cuda)
    seed, off = 1234, 96
    h, g = cpm.ops.gemm_nt(a, b, bias, epilogue=cpm.ops.GEMM_GELU, p_drop=p_drop, seed=seed, rng_offset=off)
    ref_h = a.double().cpu() @ b.double().cpu().t() + bias.double().cpu()
    _assert_close_bf16(h, ref_h, "pre-activation")
    # the standalone kernel on the stored pre-activation, same (seed, offset)
    g_ref = torch.empty_like(h)
    lib = cpm._lib.load()
    cpm._lib.check(lib.cpm_gelu_fwd(h.data_ptr(), None, g_ref.data_ptr(), M, N, p_drop, seed, off, cpm._lib.BF16, torch.cuda.current_stream().cuda_stream))
    assert torch.equal(g, g_ref), f"fused GELU epilogue differs from cpm_gelu_fwd in {int((g != g_ref).sum())} elements"
    if p_drop == 0.0:
        exact = torch.nn.functional.gelu(h.double().cpu())
        _assert_close_bf16(g, exact, "gelu vs exact erf")
    else:
        kept = (g != 0).float().mean().item()
        assert abs(kept - 230 / 256) < 0.02, f"kept fraction {kept}"           # thr8 = 26 of 256 for p = 0.1


@pytest.mark.parametrize("p_drop", [0.0, 0.1])
def test_gemm_nt_dgelu_epilogue(cuda, cpm, p_drop, gemm_mode):
    """dgrad of linear2 with the GELU backward fused: (dy W) * gelu'(h) * mask == GEMM then cpm_gelu_bwd."""
    M, N, K = 700, 2048, 512
    gen = torch.Generator().manual_seed(9)
    dy, wt = _bf16((M, K), gen).to(cuda), _bf16((N, K), gen, 1.0 / math.sqrt(K)).to(cuda)
    h = _bf16((M, N), gen).to(cuda)
    seed, off = 77, 1 << 20
    got = cpm.ops.gemm_nt(dy, wt, epilogue=cpm.ops.GEMM_DGELU, aux=h, p_drop=p_drop, seed=seed, rng_offset=off)
    acc = dy.double().cpu() @ wt.double().cpu().t()
    hd = h.double().cpu()
    dgelu = 0.5 * (1 + torch.erf(hd / math.sqrt(2))) + hd * torch.exp(-0.5 * hd * hd) / math.sqrt(2 * math.pi)
    ref = acc * dgelu
    if p_drop > 0:
        # the mask of the standalone backward kernel on a gradient of ones: nonzero <=> kept
        ones, gx = torch.ones_like(h), torch.empty_like(h)
        lib = cpm._lib.load()
        cpm._lib.check(lib.cpm_gelu_bwd(h.data_ptr(), None, ones.data_ptr(), gx.data_ptr(), None, None, M, N, p_drop, seed, off,
                                        cpm._lib.BF16, torch.cuda.current_stream().cuda_stream))
        keep = (gx != 0).double().cpu()
        # gelu'(h) is exactly 0 only where h is far negative; treat those as kept-or-dropped alike
        scale = 256.0 / (256.0 - round(p_drop * 256.0))
        ref = ref * keep * scale
        sure = (dgelu.abs() > 1e-3)
        _assert_close_bf16(torch.where(sure.to(cuda), got, torch.zeros_like(got)), torch.where(sure, ref, torch.zeros_like(ref)), "dgelu+dropout")
    else:
        _assert_close_bf16(got, ref, "dgelu", extra_abs=2e-3)


TN_SHAPES = [(256, 256, 256), (4096, 512, 512), (5000, 1536, 512), (3000, 2048, 512), (2000, 512, 2048), (1500, 344, 512), (900, 512, 1216),
             (100, 264, 344), (70000, 512, 512), (3000, 768, 1024)]


@pytest.mark.parametrize("shape", TN_SHAPES)
def test_gemm_tn_weight_gradient(cuda, cpm, shape):
    """dW += dY^T X accumulated onto existing values (gradient accumulation), float64 reference from the same bf16 operands."""
    T, N, K = shape
    gen = torch.Generator().manual_seed(T + N + K)
    dy, x = _bf16((T, N), gen), _bf16((T, K), gen)
    dw0 = torch.randn(N, K, generator=gen)
    dw = dw0.clone().to(cuda)
    cpm.ops.gemm_tn_acc(dy.to(cuda), x.to(cuda), dw)
    ref_w = dw0.double() + dy.double().t() @ x.double()
    tol = 1e-5 * math.sqrt(T) * 4 + 1e-4            # fp32 accumulation over T terms of O(1) products
    err = (dw.double().cpu() - ref_w).abs().max().item()
    assert err <= tol * max(1.0, ref_w.abs().max().item() / 8), f"dW max err {err:.3e} (tol {tol:.1e})"


def test_gemm_tn_routes_row_blocks_to_separate_masters(cuda, cpm):
    """One GEMM over the fused q/k/v gradient (T, 1536), three separate (512, 512) master gradients accumulated in place; dY and X
    as column slices of wider buffers."""
    gen = torch.Generator().manual_seed(3)
    T = 2304
    gbuf, xbuf = _bf16((T, 2048), gen).to(cuda), _bf16((T, 1024), gen).to(cuda)
    dy, x = gbuf[:, 512:], xbuf[:, 512:]
    flat = torch.randn(3 * 512 * 512 + 4096, generator=gen).to(cuda)
    base = flat.clone()
    dws = [flat[1024 + i * 512 * 512:1024 + (i + 1) * 512 * 512].view(512, 512) for i in range(3)]       # views into one flat bucket
    cpm.ops.gemm_tn_acc(dy, x, dws)
    ref = dy.double().cpu().t() @ x.double().cpu()
    for i in range(3):
        want = base[1024 + i * 512 * 512:1024 + (i + 1) * 512 * 512].view(512, 512).double().cpu() + ref[i * 512:(i + 1) * 512]
        assert (dws[i].double().cpu() - want).abs().max().item() < 3e-3
    assert torch.equal(flat[:1024], base[:1024]) and torch.equal(flat[1024 + 3 * 512 * 512:], base[1024 + 3 * 512 * 512:])


@pytest.mark.parametrize("shape", [(256, 1536, 512), (256, 512, 512), (256, 2048, 512), (256, 512, 2048), (256, 512, 1216), (256, 344, 512),
                                   (32, 1536, 512), (100, 344, 512), (1, 512, 512), (70, 64, 64), (300, 96, 1216)])
@pytest.mark.parametrize("pdl", [False, True])
def test_gemm_nt_small_token_step_shapes(cuda, cpm, shape, pdl):
    """The recurrent step's Linear layers (M = songs in flight) on the 64 x 32-tile kernel, with and without programmatic
    dependent launch (back-to-back launches exercise the early-launch path), bias and bias + GELU epilogues."""
    M, N, K = shape
    gen = torch.Generator().manual_seed(M + 3 * N + K)
    a, w = _bf16((M, K), gen).to(cuda), _bf16((N, K), gen, 1.0 / math.sqrt(K)).to(cuda)
    bias = torch.randn(N, generator=gen).to(cuda)
    ref = a.double().cpu() @ w.double().cpu().t() + bias.double().cpu()
    cpm.ops.set_chain_pdl(pdl)
    try:
        outs = [cpm.ops.gemm_nt_small(a, w, bias) for _ in range(4)]          # a chain of dependent-launch kernels
        g = cpm.ops.gemm_nt_small(a, w, bias, gelu=True)
        nb = cpm.ops.gemm_nt_small(a, w)
    finally:
        cpm.ops.set_chain_pdl(False)
    for o in outs:
        _assert_close_bf16(o, ref, f"gemm_nt_small {shape}")
    _assert_close_bf16(nb, ref - bias.double().cpu(), "no bias")
    exact = torch.nn.functional.gelu(outs[0].double().cpu())
    _assert_close_bf16(g, exact, "gelu epilogue vs exact erf of the bf16 pre-activation")


@pytest.mark.parametrize("shape", [(256, 512, 2048), (256, 512, 1216), (70, 96, 2048), (300, 344, 512)])
@pytest.mark.parametrize("pdl", [False, True])
def test_gemm_nt_small_layernorm_forms(cuda, cpm, shape, pdl):
    """cpm_gemm_nt_small_ln: the residual and LayerNorm-rebuilt-residual epilogues against the plain product plus the residual
    (fp64), and the FOLD form - raw rows times gamma o W with the row statistics taken from the activation tile - against
    LayerNorm followed by the product, statistics included."""
    M, N, K = shape
    gen = torch.Generator().manual_seed(7 * M + N + K)
    a, w = _bf16((M, K), gen).to(cuda), _bf16((N, K), gen, 1.0 / math.sqrt(K)).to(cuda)
    bias = torch.randn(N, generator=gen).to(cuda)
    res = _bf16((M, N), gen).to(cuda)
    cpm.ops.set_chain_pdl(pdl)
    try:
        plain = cpm.ops.gemm_nt_small(a, w, bias)
        r1 = cpm.ops.gemm_nt_small_ln(a, w, bias, resid=res)
        stats = torch.stack([res.float().mean(1), torch.rsqrt(res.float().var(1, unbiased=False) + 1e-5)], 1).contiguous()
        gam, bet = torch.randn(N, generator=gen).to(cuda), torch.randn(N, generator=gen).to(cuda)
        r2 = cpm.ops.gemm_nt_small_ln(a, w, bias, resid=res, r_stats=stats, r_gamma=gam, r_beta=bet)
        # FOLD: LayerNorm(y) W^T + b on raw rows (K = 512: the statistics come from the resident activation tile)
        y = (_bf16((M, 512), gen) * 1.7 + 0.4).to(cuda)
        W = _bf16((N, 512), gen, 1.0 / math.sqrt(512)).to(cuda)
        g2, b2 = (torch.randn(512, generator=gen) * 0.2 + 1.0).to(cuda), (torch.randn(512, generator=gen) * 0.1).to(cuda)
        wf = (W.float() * g2[None, :]).bfloat16()
        c1, c2 = wf.float().sum(1), (W.float() * b2[None, :]).sum(1) + bias
        st = torch.empty(M, 2, device=cuda)
        out = cpm.ops.gemm_nt_small_ln(y, wf, c2, fold_c1=c1, stats_out=st, ln_eps=1e-5)
        outg = cpm.ops.gemm_nt_small_ln(y, wf, c2, gelu=True, fold_c1=c1, ln_eps=1e-5)
    finally:
        cpm.ops.set_chain_pdl(False)
    torch.cuda.synchronize()
    _assert_close_bf16(r1, plain.double().cpu() + res.double().cpu(), "residual epilogue")
    ln = torch.nn.functional.layer_norm(res.float(), (N,), gam, bet, 1e-5).bfloat16()
    # the rebuilt LayerNorm value is rounded to bf16 like the LayerNorm kernel's output: one ulp of it may differ from torch's
    _assert_close_bf16(r2, plain.double().cpu() + ln.double().cpu(), "LayerNorm-rebuilt residual epilogue", extra_abs=ln.float().abs().max().item() * 2.0 ** -7)
    want = torch.nn.functional.layer_norm(y.double().cpu(), (512,), g2.double().cpu(), b2.double().cpu(), 1e-5) @ W.double().cpu().t() + bias.double().cpu()
    assert (out.double().cpu() - want).abs().max() < 0.06 and (out.double().cpu() - want).abs().mean() < 6e-3
    _assert_close_bf16(outg, torch.nn.functional.gelu(out.double().cpu()), "fold + gelu")
    assert (st[:, 0].cpu() - y.float().mean(1).cpu()).abs().max() < 1e-4
    assert ((st[:, 1].cpu() - torch.rsqrt(y.float().var(1, unbiased=False) + 1e-5).cpu()).abs() / st[:, 1].cpu()).max() < 1e-3


@pytest.mark.parametrize("shape", [(256, 512, 2048), (70, 96, 2048), (300, 512, 4096), (256, 512, 1024)])
def test_gemm_nt_small_split_k_cluster(cuda, cpm, shape):
    """Long K on the token-step GEMM: two K slices per output tile as a thread-block cluster, the leader adds the other slice's
    fp32 tile from its shared memory.  Against the fp64 product, against the unsplit kernel (same values up to fp32 summation
    order: equal after bf16 rounding almost everywhere), identical from launch to launch, under PDL back to back."""
    M, N, K = shape
    gen = torch.Generator().manual_seed(11 * M + N + K)
    a, w = _bf16((M, K), gen).to(cuda), _bf16((N, K), gen, 1.0 / math.sqrt(K)).to(cuda)
    bias = torch.randn(N, generator=gen).to(cuda)
    res = _bf16((M, N), gen).to(cuda)
    ref = a.double().cpu() @ w.double().cpu().t() + bias.double().cpu()
    lib = cpm._lib.load()
    lib.cpm_gemm_small_set_split(0)
    try:
        unsplit = cpm.ops.gemm_nt_small(a, w, bias)
    finally:
        lib.cpm_gemm_small_set_split(1)
    cpm.ops.set_chain_pdl(True)
    try:
        outs = [cpm.ops.gemm_nt_small(a, w, bias) for _ in range(6)]
        r = cpm.ops.gemm_nt_small_ln(a, w, bias, resid=res)
    finally:
        cpm.ops.set_chain_pdl(False)
    torch.cuda.synchronize()
    for o in outs:
        _assert_close_bf16(o, ref, f"split-K gemm_nt_small {shape}")
        assert torch.equal(o, outs[0])
    assert (outs[0] != unsplit).float().mean() < 0.02
    _assert_close_bf16(r, outs[0].double().cpu() + res.double().cpu(), "split K + residual epilogue")
