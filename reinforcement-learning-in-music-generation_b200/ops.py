"""torch-facing wrappers of the libcpmusic C-ABI (include/cpmusic.h).

Every function here launches hand-written sm_100a kernels through ctypes on the
current torch CUDA stream; torch only owns memory and streams.  Plain dense
GEMMs (the Linear layers) are the one thing delegated to the vendor library
(cuBLASLt through ``torch.addmm``/``torch.mm``).  Nothing falls back to the CPU:
tensors must be CUDA tensors and the library must be loadable.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import contextlib
import os
import gc

import torch

from . import _lib
from ._lib import check

EPS_ATTN = 1e-6     # ft CausalLinearAttention eps
EPS_LN = 1e-5       # torch.nn.LayerNorm default used by ft


# --------------------------------------------------------------------------- helpers
def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.bfloat16:
        return _lib.BF16
    raise ValueError(f"cpmusic kernels take float32 or bfloat16 activations, got {t.dtype}")


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("cpmusic ops need CUDA tensors: there is no CPU fallback "
                               "(the CPU restatement lives in oracle/ and is test-only)")


def _st() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class _Rng:
    """Counter-based dropout RNG state: (seed, running offset in units of 8 elements)."""
    seed = 0x5EED
    offset = 0
    base_dev = None          # optional device-side counter added by the kernels (enable_device_rng)

    @classmethod
    def take(cls, n_elems: int) -> Tuple[int, int]:
        off = cls.offset
        cls.offset += (n_elems + 7) // 8
        return cls.seed, off


def manual_seed(seed: int) -> None:
    _Rng.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    _Rng.offset = 0
    if _Rng.base_dev is not None:
        _Rng.base_dev.zero_()


def enable_device_rng(device) -> torch.Tensor:
    """Installs (once) the device-side counter every dropout kernel adds to its host-side offset (cpm_set_rng_base), so
    that CUDA graphs captured over dropout kernels draw fresh masks per replay; returns the counter (int64, 1 element)."""
    if _Rng.base_dev is None:
        _Rng.base_dev = torch.zeros(1, dtype=torch.int64, device=device)
        check(_lib.load().cpm_set_rng_base(_Rng.base_dev.data_ptr()))
    elif _Rng.base_dev.device != torch.device(device):
        raise RuntimeError("the device RNG base lives on one device per process")
    return _Rng.base_dev


def rng_advance(n_counters: int) -> None:
    """base += n (a device-side add: capturable; call it at the end of a captured step with the offsets that step used)."""
    if _Rng.base_dev is None:
        raise RuntimeError("call enable_device_rng() first")
    _Rng.base_dev.add_(int(n_counters))


# --------------------------------------------------------------------------- in-situ kernel timing
class KernelTimer:
    """CUDA-event timing of selected kernels on the launching stream while a real step runs
    (bench.py's live roofline measurement).  Disabled by default; zero cost when off."""
    enabled = False
    events = {}
    exclude_streams = set()          # cuda_stream handles whose launches are not timed (work overlapped with another stream's)

    @classmethod
    def start(cls):
        cls.enabled, cls.events = True, {}

    @classmethod
    def stop(cls):
        """-> {name: (calls, total_ms)}; synchronises."""
        cls.enabled = False
        torch.cuda.synchronize()
        out = {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in cls.events.items()}
        cls.events = {}
        return out

    @classmethod
    def span(cls, name):
        if not cls.enabled or (cls.exclude_streams and torch.cuda.current_stream().cuda_stream in cls.exclude_streams):
            return _NULL
        return _Span(name)


class _Span:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        self.a = torch.cuda.Event(enable_timing=True)
        self.b = torch.cuda.Event(enable_timing=True)
        self.a.record()

    def __exit__(self, *exc):
        self.b.record()
        KernelTimer.events.setdefault(self.name, []).append((self.a, self.b))


class _Null:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NULL = _Null()


# --------------------------------------------------------------------------- CUDA graph capture
@contextlib.contextmanager
def graph_capture(graph, **kw):
    """``torch.cuda.graph`` with Python's cyclic garbage collector held off: a collection that runs in the middle of a capture
    can destroy an older CUDAGraph / free device memory (objects kept alive only by reference cycles, e.g. a discarded
    rollout engine), and such calls invalidate the capture in progress (cudaErrorStreamCaptureInvalidated)."""
    gc.collect()
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        with torch.cuda.graph(graph, **kw):
            yield
    finally:
        if was_enabled:
            gc.enable()


# --------------------------------------------------------------------------- linear attention
def linattn_workspace(N: int, L: int, H: int, device, E: int = 64) -> torch.Tensor:
    nbytes = _lib.load().cpm_linattn_workspace_bytes_wide(N, L, H, E)
    return torch.empty(max(nbytes, 16), dtype=torch.uint8, device=device)


def _check_qkv_layout(q, k, v):
    N, L, H, E = q.shape
    for t in (q, k, v):
        if t.shape != q.shape or t.stride(3) != 1 or t.stride(2) != E or t.stride(0) != L * t.stride(1):
            raise ValueError("q,k,v must be (N,L,H,E) with E contiguous, heads packed and a common token stride")
    if not (q.stride(1) == k.stride(1) == v.stride(1)):
        raise ValueError("q,k,v must share one token stride")
    return N, L, H, E, q.stride(1)


def linattn_saved(N: int, L: int, H: int, device, E: int = 64) -> Optional[torch.Tensor]:
    """Buffer for the chunk-parallel kernels' per-chunk prefix states (kept from fwd to bwd); E = head width (64 | 128)."""
    nbytes = _lib.load().cpm_linattn_saved_bytes_wide(N, L, H, E)
    return torch.empty(nbytes, dtype=torch.uint8, device=device) if nbytes > 0 else None


def _linattn_extra_launches(N, L, H, bwd=False, E=64, bf16=True):
    """Kernels of one chunk-parallel call beyond the nominal two (launch accounting of bench.py).  Many (batch, head) chains of
    64-wide heads: streaming state kernel + per-chunk kernel.  Otherwise: per-chunk state kernel, scan, per-chunk kernel.  A single
    chunk has no prefix state to build (the backward still runs its pre-pass for the per-token normaliser gradients)."""
    if not bf16:
        return 0                                  # CUDA-core path (fp32 parity mode)
    if L <= 128:
        return 0 if bwd else -1
    return 1 if (N * H < 96 or E != 64) else 0      # 128-wide heads: always per-chunk state kernel + scan


def linattn_fwd_raw(q, k, v, eps=EPS_ATTN, impl=0, need_den=True, saved=None):
    """q,k,v: (N,L,H,E) views (may be column slices of one fused QKV buffer), E = 64, or 128 on the tensor-core path (bf16).  `saved`: optional buffer from linattn_saved() that receives the prefix states for linattn_bwd_raw."""
    _cuda(q, k, v)
    N, L, H, E, ld = _check_qkv_layout(q, k, v)
    out = torch.empty(N, L, H, E, dtype=q.dtype, device=q.device)
    den = torch.empty(N, L, H, dtype=torch.float32, device=q.device) if need_den else None
    ws = linattn_workspace(N, L, H, q.device, E)
    _lib.EXTRA_LAUNCHES[0] += _linattn_extra_launches(N, L, H, E=E, bf16=q.dtype == torch.bfloat16)
    with KernelTimer.span("linattn_fwd"):
        check(_lib.load().cpm_linattn_fwd(_p(q), _p(k), _p(v), _p(out), _p(den), N, L, H, E, E, ld, H * E,
                                          _dt(q), eps, impl, _p(ws), ws.numel(), _p(saved),
                                          0 if saved is None else saved.numel(), _st()))
    return out, den


def linattn_bwd_raw(q, k, v, out, den, gout, gq, gk, gv, eps=EPS_ATTN, impl=0, saved=None):
    N, L, H, E, ld = _check_qkv_layout(q, k, v)
    _, _, _, _, ldg = _check_qkv_layout(gq, gk, gv)
    gout = gout.contiguous()
    ws = linattn_workspace(N, L, H, q.device, E)
    _lib.EXTRA_LAUNCHES[0] += _linattn_extra_launches(N, L, H, bwd=True, E=E, bf16=q.dtype == torch.bfloat16)
    with KernelTimer.span("linattn_bwd"):
        check(_lib.load().cpm_linattn_bwd(_p(q), _p(k), _p(v), _p(out), _p(den), _p(gout), _p(gq), _p(gk), _p(gv),
                                          N, L, H, E, E, ld, H * E, ldg, _dt(q), eps, impl, _p(ws), ws.numel(),
                                          _p(saved), 0 if saved is None else saved.numel(), _st()))


class _LinAttnFused(torch.autograd.Function):
    """qkv (N,L,3*H*E) fused projection output -> (N,L,H*E); E = 64, or 128 for bf16 on the tensor-core kernels.

    Any length: a sequence's last 128-token chunk may be short (e.g. the 50-token DQN windows, IRL_dqn_train.py:55-59) - the
    tensor-core kernels mask its surplus rows themselves (no padded copies of the activations or gradients)."""

    @staticmethod
    def forward(ctx, qkv, H, eps, impl, want_den=False):
        N, L, W = qkv.shape
        E = W // (3 * H)
        qkv = qkv.contiguous()
        q, k, v = (qkv[..., i * H * E:(i + 1) * H * E].unflatten(-1, (H, E)) for i in range(3))
        need = ctx.needs_input_grad[0]
        saved = linattn_saved(N, L, H, qkv.device, E) if (need and impl in (0, 3) and qkv.dtype == torch.bfloat16) else None
        out, den = linattn_fwd_raw(q, k, v, eps, impl, saved=saved)
        ctx.save_for_backward(qkv, out, den, saved)
        ctx.cfg = (H, E, eps, impl, L)
        res = out.view(N, L, H * E)
        if want_den:                          # the kernel's normaliser as a second, non-differentiable output
            den_out = den.clone()
            ctx.mark_non_differentiable(den_out)
            return res, den_out
        return res

    @staticmethod
    def backward(ctx, gout, *_unused):
        qkv, out, den, saved = ctx.saved_tensors
        H, E, eps, impl, L = ctx.cfg
        N = qkv.shape[0]
        q, k, v = (qkv[..., i * H * E:(i + 1) * H * E].unflatten(-1, (H, E)) for i in range(3))
        gqkv = torch.empty_like(qkv)
        gq, gk, gv = (gqkv[..., i * H * E:(i + 1) * H * E].unflatten(-1, (H, E)) for i in range(3))
        linattn_bwd_raw(q, k, v, out, den, gout.reshape(N, L, H, E), gq, gk, gv, eps, impl, saved=saved)
        return gqkv, None, None, None, None


class _LinAttn(torch.autograd.Function):
    """Separate q,k,v (N,L,H,64) -> (N,L,H,64) (the ft CausalLinearAttention signature)."""

    @staticmethod
    def forward(ctx, q, k, v, eps, impl):
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        out, den = linattn_fwd_raw(q, k, v, eps, impl)
        ctx.save_for_backward(q, k, v, out, den)
        ctx.cfg = (eps, impl)
        return out

    @staticmethod
    def backward(ctx, gout):
        q, k, v, out, den = ctx.saved_tensors
        eps, impl = ctx.cfg
        gq, gk, gv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        linattn_bwd_raw(q, k, v, out, den, gout, gq, gk, gv, eps, impl)
        return gq, gk, gv, None, None


def causal_linear_attention(q, k, v, eps=EPS_ATTN, impl=0):
    """q,k,v (N,L,H,E) -> (N,L,H,E) (ft's CausalLinearAttention signature); E = 64 or 128."""
    if q.shape[-1] == 128:
        N, L, H, _ = q.shape
        qkv = torch.cat([t.reshape(N, L, H * 128) for t in (q, k, v)], -1)
        return causal_linear_attention_fused(qkv, H, eps, impl).view(N, L, H, 128)
    return _LinAttn.apply(q, k, v, eps, impl)


def _linattn_fused_e128(qkv, H, eps, impl):
    """128-wide heads (cfg5 may be 8 heads x 128) on the 64-wide kernels.  With the feature-mapped queries /
    keys split into halves a and the values into halves b,
        A_ij = A0_ij + A1_ij,   out_i[b] = (num^{0,b}_i + num^{1,b}_i) / (den^0_i + den^1_i - eps),
    where num^{a,b} = out^{a,b} * den^a is what one 64-wide attention over (q^a, k^a, v^b) returns.  Two kernel passes over
    2H virtual heads give all four blocks: the buffer as it is (a = b) and with the value halves swapped (a != b).  The
    per-half normalisers carry their VALUE from the kernels and their GRADIENT from a differentiable restatement
    (phi(q^a) . cumsum phi(k^a)), so that the recombination is exact in the forward and autograd handles the rest."""
    N, L, W = qkv.shape
    d = H * 128
    q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
    v_sw = v.unflatten(-1, (H, 2, 64)).flip(-2).flatten(-3)
    same, den_s = _LinAttnFused.apply(qkv, 2 * H, eps, impl, True)
    cross, _ = _LinAttnFused.apply(torch.cat([q, k, v_sw], -1), 2 * H, eps, impl, True)
    # feature-major copies: the running key sum is then a scan along the contiguous axis (an outer-dimension cumsum over
    # L = 8192 costs 2 ms per call, 10x everything else here)
    qf = torch.nn.functional.elu(q.transpose(1, 2).contiguous().float()) + 1.0               # (N, H*128, L)
    kf = torch.nn.functional.elu(k.transpose(1, 2).contiguous().float()) + 1.0
    den = (qf * kf.cumsum(-1)).unflatten(1, (H, 2, 64)).sum(3) + eps                         # (N,H,2,L)
    den = den.permute(0, 3, 1, 2).unsqueeze(-1)                                              # (N,L,H,2,1)
    den = den + (den_s.view(N, L, H, 2, 1) - den).detach()
    same, cross = same.float().unflatten(-1, (H, 2, 64)), cross.float().unflatten(-1, (H, 2, 64))
    num = same * den + (cross * den).flip(-2)            # [h, b] = out^{b,b} den^b + out^{1-b,b} den^{1-b}
    tot = den.sum(-2, keepdim=True) - eps
    return (num / tot).flatten(-3).to(qkv.dtype)


KEY_MASK_FILL = -30000.0        # elu(x) + 1 = exp(x) underflows to exactly 0 in fp32 (and so does its derivative)


def causal_linear_attention_fused(qkv, n_heads, eps=EPS_ATTN, impl=0, key_mask=None):
    """qkv (N,L,3*H*E) -> (N,L,H*E); E = 64 or 128.  128-wide heads run natively on the tensor-core kernels in bf16 (templates of
    the 64-wide kernels over the two halves of a head); fp32 takes two passes over 64-wide virtual heads (_linattn_fused_e128).
    ``key_mask`` (N,L) bool / 0-1: ft's key-padding mask, ``K = K * k_len.float_matrix``.  A padded key must
    feed neither the KV state nor the normaliser; the kernels apply the feature map themselves, so the padded keys are sent in as
    a large negative number, whose feature value (and derivative) is exactly zero - the same outputs and the same (zero) key
    gradients as the multiplication."""
    E = qkv.shape[-1] // (3 * n_heads)
    if qkv.shape[-1] != 3 * n_heads * E or E not in (64, 128):
        raise ValueError(f"causal linear attention: head width {E} (supported: 64, 128)")
    if key_mask is not None:
        if key_mask.shape != qkv.shape[:2]:
            raise ValueError(f"key_mask {tuple(key_mask.shape)} against qkv {tuple(qkv.shape)}")
        HE = n_heads * E
        is_key = torch.zeros(3 * HE, dtype=torch.bool, device=qkv.device)
        is_key[HE:2 * HE] = True
        drop = (~key_mask.to(device=qkv.device).bool())[..., None] & is_key
        qkv = torch.where(drop, torch.full((), KEY_MASK_FILL, dtype=qkv.dtype, device=qkv.device), qkv)
    if E == 128 and not (qkv.dtype == torch.bfloat16 and qkv.is_cuda and impl in (0, 3)):
        return _linattn_fused_e128(qkv, n_heads, eps, impl)      # fp32 parity mode: two passes over 64-wide virtual heads
    return _LinAttnFused.apply(qkv, n_heads, eps, impl)


def linattn_last_impl() -> str:
    return _lib.load().cpm_linattn_last_impl().decode()


def linattn_step(q, k, v, S, Z, eps=EPS_ATTN):
    """Recurrent step. q,k,v: (N,H,64) views sharing a row stride; S (N,H,64,64), Z (N,H,64) fp32
    are updated in place; returns (N,H,64)."""
    _cuda(q, k, v, S, Z)
    N, H, E = q.shape
    ld = q.stride(0)
    if not (k.stride(0) == ld and v.stride(0) == ld and q.stride(2) == 1 and q.stride(1) == E):
        raise ValueError("q,k,v must be (N,H,E) with a common row stride and packed heads")
    if S.shape[0] != N:
        raise ValueError("The batch size changed during iteration")     # ft's message
    if S.dtype != torch.float32 or Z.dtype != torch.float32 or not S.is_contiguous() or not Z.is_contiguous():
        raise ValueError("recurrent state must be contiguous float32")
    out = torch.empty(N, H, E, dtype=q.dtype, device=q.device)
    check(_lib.load().cpm_linattn_step(_p(q), _p(k), _p(v), _p(S), _p(Z), _p(out), N, H, E, E, ld, H * E,
                                       _dt(q), eps, _st()))
    return out


# --------------------------------------------------------------------------- dense linear: own tcgen05 GEMMs (csrc/tc_gemm.cu)
GEMM_BIAS, GEMM_GELU, GEMM_DGELU = 0, 1, 2
SMALL_GEMM_ROWS = int(os.environ.get("CPM_SMALL_GEMM_ROWS", "1024"))      # below: cpm_gemm_nt_small


def _gemm_operand(t, what):
    if t.dtype != torch.bfloat16 or t.dim() != 2:
        raise ValueError(f"{what} must be a 2-D bfloat16 matrix")
    if t.stride(1) != 1 or t.stride(0) % 8 or t.data_ptr() % 16:
        t = t.contiguous()
    return t


def gemm_nt(a, b, bias=None, epilogue=GEMM_BIAS, aux=None, p_drop=0.0, seed=0, rng_offset=0, out=None):
    """epilogue(a @ b.T): a (M,K), b (N,K) bf16 -> (M,N) bf16 on the 2-CTA tcgen05 GEMM (cpm_gemm_nt).  bias: fp32 (N,).
    GEMM_GELU returns (h, dropout(gelu(h))); GEMM_DGELU multiplies by gelu'(aux) and the regenerated dropout mask."""
    _cuda(a, b, bias, aux)
    a, b = _gemm_operand(a, "a"), _gemm_operand(b, "b")
    M, K = a.shape
    N = b.shape[0]
    if b.shape[1] != K:
        raise ValueError(f"gemm_nt: a (M,{K}) against b {tuple(b.shape)}")
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != N or not bias.is_contiguous()):
        raise ValueError("gemm_nt: bias must be contiguous float32 (N,)")
    d = out if out is not None else torch.empty(M, N, dtype=torch.bfloat16, device=a.device)
    if M < SMALL_GEMM_ROWS and epilogue == GEMM_BIAS:      # few rows (token step, small batches): the 64 x 32-tile kernel
        check(_lib.load().cpm_gemm_nt_small(_p(a), a.stride(0), _p(b), b.stride(0), _p(d), d.stride(0), M, N, K, _p(bias), GEMM_BIAS, _st()))
        return d
    d2 = torch.empty(M, N, dtype=torch.bfloat16, device=a.device) if epilogue == GEMM_GELU else None
    if aux is not None:
        aux = _gemm_operand(aux, "aux")
    check(_lib.load().cpm_gemm_nt(_p(a), a.stride(0), _p(b), b.stride(0), _p(d), d.stride(0), _p(d2), 0 if d2 is None else d2.stride(0),
                                  M, N, K, _p(bias), epilogue, _p(aux), 0 if aux is None else aux.stride(0), p_drop, seed, rng_offset, _st()))
    return (d, d2) if epilogue == GEMM_GELU else d


def gemm_set_mode(mode: int) -> None:
    """cpm_gemm_set_mode: 0 auto | 1 stream | 2 A-stationary | 3 wide (tests and A/B measurements)."""
    check(_lib.load().cpm_gemm_set_mode(int(mode)))


def gemm_nt_small(a, w, bias=None, gelu=False, out=None):
    """epilogue(a @ w.T + bias) for the recurrent token step (cpm_gemm_nt_small): a (M,K), w (N,K) bf16; bias fp32 (N,)."""
    _cuda(a, w, bias)
    a, w = _gemm_operand(a, "a"), _gemm_operand(w, "w")
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise ValueError(f"gemm_nt_small: a (M,{K}) against w {tuple(w.shape)}")
    d = out if out is not None else torch.empty(M, N, dtype=torch.bfloat16, device=a.device)
    check(_lib.load().cpm_gemm_nt_small(_p(a), a.stride(0), _p(w), w.stride(0), _p(d), d.stride(0), M, N, K, _p(bias),
                                        GEMM_GELU if gelu else GEMM_BIAS, _st()))
    return d


def gemm_nt_small_ln(a, w, bias, gelu=False, fold_c1=None, stats_out=None, ln_eps=EPS_LN, resid=None, r_stats=None, r_gamma=None,
                     r_beta=None, out=None):
    """cpm_gemm_nt_small_ln: the token-step Linear with the LayerNorm around it folded in.  FOLD form (``fold_c1``): ``a`` holds raw
    pre-norm rows, ``w`` = gamma o W, ``bias`` = c2, the row statistics are computed in the kernel (and written to ``stats_out``
    (M,2) fp32 if given).  RESIDUAL form (``resid``): out = bf16(a w^T + bias) + resid, resid optionally passed through the
    LayerNorm described by (r_stats, r_gamma, r_beta)."""
    _cuda(a, w, bias, fold_c1, stats_out, resid, r_stats, r_gamma, r_beta)
    a, w = _gemm_operand(a, "a"), _gemm_operand(w, "w")
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise ValueError(f"gemm_nt_small_ln: a (M,{K}) against w {tuple(w.shape)}")
    d = out if out is not None else torch.empty(M, N, dtype=torch.bfloat16, device=a.device)
    if resid is not None:
        resid = _gemm_operand(resid, "resid")
    check(_lib.load().cpm_gemm_nt_small_ln(_p(a), a.stride(0), _p(w), w.stride(0), _p(d), d.stride(0), M, N, K, _p(bias),
                                           GEMM_GELU if gelu else GEMM_BIAS, _p(fold_c1), _p(stats_out), ln_eps, _p(resid),
                                           0 if resid is None else resid.stride(0), _p(r_stats), _p(r_gamma), _p(r_beta), _st()))
    return d


def set_chain_pdl(on: bool) -> None:
    check(_lib.load().cpm_set_chain_pdl(1 if on else 0))


def gemm_tn_acc(dy, x, dws):
    """dW (N,K) fp32 += dy (T,N).T @ x (T,K): the weight gradient of a Linear layer, accumulated in place (cpm_gemm_tn: operands
    read MN-major from the row-major activations, no transposes).  ``dws``: one (N,K) fp32 matrix, or a list of up to 4 equally
    tall matrices that receive consecutive row blocks (the q / k / v masters behind one fused projection)."""
    if torch.is_tensor(dws):
        dws = [dws]
    _cuda(dy, x, *dws)
    dy, x = _gemm_operand(dy, "dy"), _gemm_operand(x, "x")
    T, N = dy.shape
    K = x.shape[1]
    rows = dws[0].shape[0]
    if x.shape[0] != T or len(dws) > 4 or rows * len(dws) != N:
        raise ValueError(f"gemm_tn_acc: dy {tuple(dy.shape)}, x {tuple(x.shape)}, {len(dws)} destinations of {rows} rows")
    ldw = dws[0].stride(0)
    for w in dws:
        if w.shape != (rows, K) or w.dtype != torch.float32 or w.stride(1) != 1 or w.stride(0) != ldw:
            raise ValueError(f"gemm_tn_acc: destinations must be float32 ({rows},{K}) with a common row stride")
    check(_lib.load().cpm_gemm_tn(_p(dy), dy.stride(0), _p(x), x.stride(0), _lib.ptr_array([w.data_ptr() for w in dws]), len(dws), rows,
                                  ldw, T, N, K, _st()))
    return dws


# How the Linear layers run: "own" (default) = the tcgen05 GEMMs above for bf16 activations (the 2-CTA kernels from
# SMALL_GEMM_ROWS rows up, the 64 x 32-tile kernel below that: the rollout token step and small batches); "lib" = cuBLASLt
# through torch, kept for A/B measurements and as the fp32 parity mode's GEMM.
GEMM_IMPL = os.environ.get("CPM_GEMM", "own")


def use_own_gemm(x) -> bool:
    return GEMM_IMPL == "own" and x.dtype == torch.bfloat16 and x.is_cuda


def _fire_grad_hooks(p):
    """A gradient was accumulated into ``p.grad`` outside autograd's AccumulateGrad node: run the parameter's
    post-accumulate-grad hooks (the bucketed all-reduce of dist.py counts gradients through them)."""
    hooks = getattr(p, "_post_accumulate_grad_hooks", None)
    if hooks:
        for h in list(hooks.values()):
            h(p)


def _wgrad_into_masters(gy2, x2, rows, masters, has_bias):
    """dW_i (+)= gy[:, rows_i]^T x for the masters packed behind one GEMM, bias gradients as column sums of gy (cpm_colsum).
    Weight masters whose ``.grad`` already exists (gradient accumulation; the flat buckets of dist.BucketedGradAllReduce) are
    accumulated IN PLACE by the kernel's fp32 atomics - no separate add kernels - and report ``None`` to autograd; fresh
    gradients are returned."""
    n_w = len(rows)
    out = [None] * (2 * n_w if has_bias else n_w)
    N, K = gy2.shape[1], x2.shape[1]
    ws = masters[:n_w]
    in_place = (len(set(rows)) == 1 and rows[0] % 8 == 0 and n_w <= 4 and sum(rows) == N
                and all(w.grad is not None and w.grad.dtype == torch.float32 and w.grad.is_contiguous() and w.grad.data_ptr() % 16 == 0 for w in ws))
    if in_place:
        gemm_tn_acc(gy2, x2, [w.grad for w in ws])
        for w in ws:
            _fire_grad_hooks(w)
    else:                                 # fresh gradients (or ragged row counts, e.g. the six output heads): one packed buffer
        gw = torch.zeros(N, K, dtype=torch.float32, device=gy2.device)
        gemm_tn_acc(gy2, x2, gw)
        r0 = 0
        for i, r in enumerate(rows):
            out[i] = gw[r0:r0 + r]
            r0 += r
    if has_bias:
        gb = colsum(gy2)
        r0 = 0
        for i, r in enumerate(rows):
            out[n_w + i] = gb[r0:r0 + r]
            r0 += r
    return out


class _TcLinear(torch.autograd.Function):
    """y = x @ Wc^T + b on the own GEMMs.  ``pack`` = (wc (N,K) bf16, wt (K,N) bf16, b32 (N,) fp32 | None, rows): the packed
    compute copies of the fp32 masters (encoder.PackCache)."""

    @staticmethod
    def forward(ctx, x, pack, use_bias, *masters):
        wc, wt, b32, rows = pack
        x2 = x.reshape(-1, x.shape[-1])
        y = gemm_nt(x2, wc, b32 if use_bias else None)
        ctx.save_for_backward(x2)
        ctx.pack, ctx.use_bias, ctx.xshape, ctx.masters = pack, use_bias, x.shape, masters
        return y.view(*x.shape[:-1], wc.shape[0])

    @staticmethod
    def backward(ctx, gy):
        (x2,) = ctx.saved_tensors
        wc, wt, b32, rows = ctx.pack
        gy2 = gy.reshape(-1, gy.shape[-1])
        gx = gemm_nt(gy2, wt).view(ctx.xshape) if ctx.needs_input_grad[0] else None
        grads = _wgrad_into_masters(gy2, x2, rows, ctx.masters, ctx.use_bias)
        grads += [None] * (len(ctx.masters) - len(grads))
        return (gx, None, None, *grads)


def tc_linear_packed(x, pack, masters, use_bias=True):
    return _TcLinear.apply(x, pack, use_bias, *masters)


class _TcFFN(torch.autograd.Function):
    """ft's feed-forward block  linear2(dropout(gelu(linear1(x))))  (linear1/linear2 of
    agent_pretrain.py:244-253's builder) as two GEMM launches forward and four backward: bias + exact GELU + dropout run in
    linear1's epilogue (which also keeps the pre-activation), GELU backward + the regenerated dropout mask in the epilogue of
    linear2's data gradient - no element-wise pass over the (T, 2048) tensors in either direction."""

    @staticmethod
    def forward(ctx, x, pack1, pack2, p_drop, w1, b1, w2, b2):
        x2 = x.reshape(-1, x.shape[-1])
        seed, off = _Rng.take(x2.shape[0] * pack1[0].shape[0]) if p_drop > 0 else (0, 0)
        h, g = gemm_nt(x2, pack1[0], pack1[2], epilogue=GEMM_GELU, p_drop=p_drop, seed=seed, rng_offset=off)
        f = gemm_nt(g, pack2[0], pack2[2])
        ctx.save_for_backward(x2, h, g)
        ctx.cfg = (pack1, pack2, p_drop, seed, off, x.shape, (w1, b1, w2, b2))
        return f.view(*x.shape[:-1], pack2[0].shape[0])

    @staticmethod
    def backward(ctx, gf):
        x2, h, g = ctx.saved_tensors
        pack1, pack2, p_drop, seed, off, xshape, (w1, b1, w2, b2) = ctx.cfg
        gf2 = gf.reshape(-1, gf.shape[-1])
        dh = gemm_nt(gf2, pack2[1], epilogue=GEMM_DGELU, aux=h, p_drop=p_drop, seed=seed, rng_offset=off)
        g2 = _wgrad_into_masters(gf2, g, pack2[3], (w2, b2), True)
        gx = gemm_nt(dh, pack1[1]).view(xshape) if ctx.needs_input_grad[0] else None
        g1 = _wgrad_into_masters(dh, x2, pack1[3], (w1, b1), True)
        return gx, None, None, None, g1[0], g1[1], g2[0], g2[1]


def tc_ffn(x, pack1, pack2, p_drop, lin1, lin2):
    return _TcFFN.apply(x, pack1, pack2, p_drop, lin1.weight, lin1.bias, lin2.weight, lin2.bias)


# --------------------------------------------------------------------------- dense linear (vendor GEMM)
def _mm_f32_out(a, b):
    """a @ b with fp32 output (bf16 inputs accumulate in fp32 inside cuBLAS either way)."""
    if a.dtype == torch.float32:
        return a @ b
    try:
        return torch.mm(a, b, out_dtype=torch.float32)
    except (TypeError, RuntimeError):
        return (a @ b).float()


class _PackedLinear(torch.autograd.Function):
    """y = x @ Wc^T + bc where (Wc, bc) is a cached compute-dtype packing (row-concatenation,
    optionally zero-padded) of several fp32 master weights/biases.  Gradients are routed back to
    the masters as row slices."""

    @staticmethod
    def forward(ctx, x, wc, bc, rows, *masters):
        x2 = x.reshape(-1, x.shape[-1])
        y = torch.addmm(bc, x2, wc.t()) if bc is not None else x2 @ wc.t()
        ctx.save_for_backward(x2, wc)
        ctx.rows = rows
        ctx.n_w = len(rows)
        ctx.has_bias = bc is not None
        ctx.xshape = x.shape
        ctx.master_dtype = masters[0].dtype
        return y.view(*x.shape[:-1], wc.shape[0])

    @staticmethod
    def backward(ctx, gy):
        x2, wc = ctx.saved_tensors
        gy2 = gy.reshape(-1, gy.shape[-1])
        gx = (gy2 @ wc).view(ctx.xshape) if ctx.needs_input_grad[0] else None
        gw = _mm_f32_out(gy2.t(), x2)
        gb = colsum(gy2) if ctx.has_bias else None
        grads, r0 = [], 0
        for r in ctx.rows:
            grads.append(gw[r0:r0 + r])
            r0 += r
        if ctx.has_bias:
            r0 = 0
            for r in ctx.rows:
                grads.append(gb[r0:r0 + r])
                r0 += r
        return (gx, None, None, None, *grads)


class PackTable:
    """Device-side item table of cpm_pack_weights for a fixed set of (master, packing) buffers; built once, launched per refresh."""

    def __init__(self, entries, device):
        """entries: (w, b, wc, wt, bc, b32, r0) per master - fp32 contiguous w (rows, cols) and b (rows) or None; bf16 wc (padded, cols),
        wt (cols, padded) or None, bc (padded,) or None; fp32 b32 (padded,) or None; r0 = the master's first row in the packing."""
        import struct
        lib = _lib.load()
        if lib.cpm_pack_item_bytes() != 64:
            raise RuntimeError("cpm_pack_weights: unexpected item layout")
        blob, tile0 = b"", 0
        for (w, bias, wc, wt, bc, b32, r0) in entries:
            rows, cols = w.shape
            ptr = lambda t, off=0: 0 if t is None else t.data_ptr() + off * t.element_size()
            blob += struct.pack("<6Q4i", ptr(w), ptr(bias), ptr(wc, r0 * cols), ptr(wt, r0), ptr(bc, r0), ptr(b32, r0), rows, cols,
                                0 if wt is None else wt.shape[1], tile0)
            tile0 += -(-rows // 32) * -(-cols // 32)
        self.n_items, self.n_tiles = len(entries), tile0
        self.table = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(device)
        self.keep = entries                                   # the buffers the table points at

    def launch(self):
        check(_lib.load().cpm_pack_weights(_p(self.table), self.n_items, self.n_tiles, _st()))


def colsum(x):
    """fp32 column sums of a 2-D bf16 / fp32 CUDA matrix (row stride free, unit column stride): cpm_colsum."""
    _cuda(x)
    rows, width = x.shape
    if (x.stride(1) != 1 or width % (32 // x.element_size()) or x.stride(0) % 8 or x.dtype not in (torch.float32, torch.bfloat16)
            or x.data_ptr() % 16):
        return x.sum(0, dtype=torch.float32)            # odd layouts (e.g. the 344-wide logits are 8-aligned; anything else): library reduction
    lib = _lib.load()
    out = torch.empty(width, dtype=torch.float32, device=x.device)
    partials = torch.empty(lib.cpm_colsum_partials_rows(width), width, dtype=torch.float32, device=x.device)
    check(lib.cpm_colsum(_p(x), rows, width, x.stride(0), _p(out), _p(partials), _dt(x), _st()))
    return out


def packed_linear(x, wc, bc, rows, masters):
    return _PackedLinear.apply(x, wc, bc, tuple(rows), *masters)


# --------------------------------------------------------------------------- embedding
class IndexGuard:
    """Out-of-range token ids.  nn.Embedding raises IndexError on them (the reference: agent_pretrain.py:185-196); the gather
    kernel instead writes a zero row and sets a sticky per-device flag.  The flag is read back WITHOUT stalling the stream: each
    ``poll()`` queues an asynchronous copy into pinned memory and raises for any earlier copy that has landed with the flag set,
    so a corrupted dataset or a vocabulary mismatch surfaces one step late instead of never.  ``check()`` is the blocking form
    for natural sync points (end of a rollout, end of an epoch, tests)."""
    _flags = {}
    _pending = {}

    @classmethod
    def flag(cls, device) -> torch.Tensor:
        device = torch.device(device)
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        f = cls._flags.get(key)
        if f is None:
            f = cls._flags[key] = torch.zeros(1, dtype=torch.int32, device=device)
        return f

    @classmethod
    def _raise(cls, f):
        f.zero_()
        raise IndexError("index out of range in self: a CP token id lies outside its attribute's vocabulary (cpm_embed_fwd)")

    @classmethod
    def poll(cls, device) -> None:
        if torch.device(device).type != "cuda" or os.environ.get("CPM_CHECK_INDICES", "1") == "0" or torch.cuda.is_current_stream_capturing():
            return
        f = cls.flag(device)
        prev = cls._pending.get(id(f))
        if prev is not None and prev[1].query():
            host, _ = cls._pending.pop(id(f))
            if int(host[0]) != 0:
                cls._raise(f)
            prev = None
        if prev is None:
            host = torch.empty(1, dtype=torch.int32, pin_memory=True)
            host.copy_(f, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            cls._pending[id(f)] = (host, ev)

    @classmethod
    def reset(cls, device) -> None:
        f = cls.flag(device)
        cls._pending.pop(id(f), None)
        f.zero_()

    @classmethod
    def check(cls, device) -> None:
        f = cls.flag(device)
        cls._pending.pop(id(f), None)
        if int(f.item()) != 0:
            cls._raise(f)


def _embed_meta(tables):
    n_tok = [int(t.shape[0]) for t in tables]
    emb = [int(t.shape[1]) for t in tables]
    return n_tok, emb


def embed_fwd_raw(idx, tables, dtype):
    _cuda(idx, *tables)
    idx = idx.contiguous()
    if idx.dtype != torch.int64:
        raise ValueError("token indices must be int64 (.long()), like the reference")
    n_attr = idx.shape[-1]
    if n_attr != len(tables):
        raise ValueError(f"idx has {n_attr} attributes but {len(tables)} tables were given")
    for t in tables:
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("embedding tables must be contiguous float32")
    n_tok, emb = _embed_meta(tables)
    T = idx.numel() // n_attr
    out = torch.empty(*idx.shape[:-1], sum(emb), dtype=dtype, device=idx.device)
    err = IndexGuard.flag(idx.device)
    check(_lib.load().cpm_embed_fwd(_p(idx), _lib.ptr_array([t.data_ptr() for t in tables]), _lib.int_array(n_tok),
                                    _lib.int_array(emb), n_attr, T, _p(out), _dt(out), _p(err), _st()))
    return out, err


class _Embed(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idx, dtype, *tables):
        out, err = embed_fwd_raw(idx, tables, dtype)
        ctx.save_for_backward(idx)
        ctx.meta = [(t.shape, t.device) for t in tables]
        ctx.err = err
        return out

    @staticmethod
    def backward(ctx, gout):
        (idx,) = ctx.saved_tensors
        gout = gout.contiguous()
        grads = [torch.zeros(s, dtype=torch.float32, device=d) for s, d in ctx.meta]
        n_tok = [int(s[0]) for s, _ in ctx.meta]
        emb = [int(s[1]) for s, _ in ctx.meta]
        T = idx.numel() // len(grads)
        check(_lib.load().cpm_embed_bwd(_p(idx.contiguous()), _p(gout), _lib.ptr_array([g.data_ptr() for g in grads]),
                                        _lib.int_array(n_tok), _lib.int_array(emb), len(grads), T, _dt(gout), _st()))
        return (None, None, *grads)


def cp_embed(idx, tables, dtype=torch.bfloat16):
    """CP gather: idx (...,A) int64, tables[a] (n_a, e_a) fp32 -> (..., sum e_a) * sqrt(e_a)."""
    return _Embed.apply(idx, dtype, *tables)


# --------------------------------------------------------------------------- PE / dropout
class _AddPE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pe, L, pos_offset, pos_dev, p_drop):
        _cuda(x, pe)
        x = x.contiguous()
        d = x.shape[-1]
        rows = x.numel() // d
        pe2 = pe.reshape(-1, d)
        seed, off = _Rng.take(x.numel()) if p_drop > 0 else (0, 0)
        y = torch.empty_like(x)
        check(_lib.load().cpm_add_pe(_p(x), _p(pe2), _p(y), rows, L, d, pos_offset, _p(pos_dev), pe2.shape[0],
                                     p_drop, seed, off, _dt(x), _st()))
        ctx.rng = (p_drop, seed, off)
        return y

    @staticmethod
    def backward(ctx, gy):
        p_drop, seed, off = ctx.rng
        if p_drop <= 0:
            return gy, None, None, None, None, None
        gy = gy.contiguous()
        gx = torch.empty_like(gy)
        check(_lib.load().cpm_dropout(_p(gy), _p(gx), gy.numel(), p_drop, seed, off, _dt(gy), _st()))
        return gx, None, None, None, None, None


def add_pe(x, pe, L, pos_offset=0, pos_dev=None, p_drop=0.0):
    return _AddPE.apply(x, pe, L, pos_offset, pos_dev, p_drop)


# --------------------------------------------------------------------------- residual + LayerNorm
class _LNResidual(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, res, gamma, beta, eps, p_drop, res_bias):
        _cuda(x, gamma, beta)
        x = x.contiguous()
        res = res.contiguous() if res is not None else None
        if res_bias is not None and (res is None or res_bias.dtype != torch.float32):
            raise ValueError("res_bias needs a residual branch and must be float32")
        d = x.shape[-1]
        rows = x.numel() // d
        need = any(ctx.needs_input_grad[:4]) or ctx.needs_input_grad[6]
        y = torch.empty_like(x)
        s = torch.empty_like(x) if need else None
        mean = torch.empty(rows, dtype=torch.float32, device=x.device) if need else None
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device) if need else None
        use_drop = p_drop > 0 and res is not None
        seed, off = _Rng.take(x.numel()) if use_drop else (0, 0)
        check(_lib.load().cpm_ln_residual_fwd(_p(x), _p(res), _p(res_bias), _p(gamma), _p(beta), _p(y), _p(s), _p(mean), _p(rstd), rows, d,
                                              eps, p_drop if use_drop else 0.0, seed, off, _dt(x), _st()))
        if need:
            ctx.save_for_backward(s, mean, rstd, gamma)
        ctx.cfg = (p_drop if use_drop else 0.0, seed, off, res is not None, res_bias is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        s, mean, rstd, gamma = ctx.saved_tensors
        p_drop, seed, off, has_res, has_rb = ctx.cfg
        gy = gy.contiguous()
        d = s.shape[-1]
        rows = s.numel() // d
        lib = _lib.load()
        gs = torch.empty_like(s)
        gres = torch.empty_like(s) if (has_res and p_drop > 0) else None
        acc = torch.zeros(3, d, dtype=torch.float32, device=s.device)          # dgamma | dbeta | dres_bias
        partials = torch.empty(lib.cpm_ln_partials_rows() * 3 * d, dtype=torch.float32, device=s.device)
        check(lib.cpm_ln_residual_bwd(_p(gy), _p(s), _p(mean), _p(rstd), _p(gamma), _p(gs), _p(gres), _p(acc[0]), _p(acc[1]),
                                      _p(acc[2]) if has_rb else None, _p(partials), rows, d, p_drop, seed, off, _dt(s), _st()))
        g_res = (gres if gres is not None else gs) if has_res else None
        return gs, g_res, acc[0].to(gamma.dtype), acc[1].to(gamma.dtype), None, None, (acc[2] if has_rb else None)


def ln_residual(x, res, gamma, beta, eps=EPS_LN, p_drop=0.0, res_bias=None):
    """LayerNorm(x + dropout(res + res_bias)) with fp32 affine parameters.  res_bias (fp32 (d), optional) is the bias of the
    Linear that produced ``res``: passing it here lets that GEMM run bias-less and makes the bias gradient a by-product of
    the backward kernel instead of a separate reduction over the gradient tensor."""
    return _LNResidual.apply(x, res, gamma, beta, eps, p_drop, res_bias)


# --------------------------------------------------------------------------- GELU
class _Gelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p_drop, bias):
        _cuda(x)
        x = x.contiguous()
        d = x.shape[-1]
        rows = x.numel() // d
        if bias is not None and (bias.dtype != torch.float32 or bias.numel() != d):
            raise ValueError("gelu bias must be float32 of the row width")
        seed, off = _Rng.take(x.numel()) if p_drop > 0 else (0, 0)
        y = torch.empty_like(x)
        check(_lib.load().cpm_gelu_fwd(_p(x), _p(bias), _p(y), rows, d, p_drop, seed, off, _dt(x), _st()))
        ctx.save_for_backward(x, bias)
        ctx.rng = (p_drop, seed, off)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, bias = ctx.saved_tensors
        p_drop, seed, off = ctx.rng
        gy = gy.contiguous()
        d = x.shape[-1]
        gx = torch.empty_like(x)
        lib = _lib.load()
        dbias = partials = None
        fused = bias is not None and ctx.needs_input_grad[2] and lib.cpm_gelu_bwd_partials_rows(d) > 0
        if fused:
            dbias = torch.zeros(d, dtype=torch.float32, device=x.device)
            partials = torch.empty(lib.cpm_gelu_bwd_partials_rows(d) * d, dtype=torch.float32, device=x.device)
        check(lib.cpm_gelu_bwd(_p(x), _p(bias), _p(gy), _p(gx), _p(dbias), _p(partials), x.numel() // d, d, p_drop, seed, off, _dt(x), _st()))
        if bias is not None and ctx.needs_input_grad[2] and not fused:
            dbias = gx.reshape(-1, d).sum(0, dtype=torch.float32)
        return gx, None, dbias


def gelu_dropout(x, p_drop=0.0, bias=None):
    """dropout(gelu(x + bias)) with the exact-erf GELU ft uses (activation='gelu').  bias (fp32, optional): the bias of the
    Linear that produced x — its gradient is then a by-product of the backward kernel."""
    return _Gelu.apply(x, p_drop, bias)


# --------------------------------------------------------------------------- heads: decode / logp / CE
def seg_offsets(n_token: Sequence[int]) -> List[int]:
    seg = [0]
    for n in n_token:
        seg.append(seg[-1] + int(n))
    return seg


def heads_sample(logits, seg, temperature=None, top_p=None, greedy=True, seed=0, seq_base=0, step=0, step_dev=None,
                 want_logp=False, want_entropy=False, tokens_out=None, logp_out=None):
    """Decode every row of the concatenated logits (rows, ld). Returns (tokens int64 (rows,A), logp, entropy)."""
    _cuda(logits)
    ld = logits.shape[-1]
    logits2 = logits.reshape(-1, ld)
    if logits2.stride(-1) != 1:
        logits2 = logits2.contiguous()
    rows, A = logits2.shape[0], len(seg) - 1
    tokens = tokens_out if tokens_out is not None else torch.empty(rows, A, dtype=torch.int64, device=logits.device)
    logp = logp_out if logp_out is not None else (torch.empty(rows, A, dtype=torch.float32, device=logits.device) if want_logp else None)
    ent = torch.empty(rows, A, dtype=torch.float32, device=logits.device) if want_entropy else None
    t_arr = _lib.float_array(temperature if temperature is not None else [1.0] * A)
    p_arr = _lib.float_array([0.0 if p is None else p for p in (top_p if top_p is not None else [None] * A)])
    check(_lib.load().cpm_heads_sample(_p(logits2), rows, logits2.stride(0), _lib.int_array(seg), A, t_arr, p_arr,
                                       0 if greedy else 1, seed, seq_base, step, _p(step_dev), _p(tokens), _p(logp), _p(ent),
                                       _dt(logits2), _st()))
    return tokens, logp, ent


class _HeadsLogp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, tokens, seg, want_entropy):
        _cuda(logits, tokens)
        ld = logits.shape[-1]
        l2 = logits.reshape(-1, ld).contiguous()
        tok = tokens.reshape(-1, len(seg) - 1).contiguous()
        rows, A = tok.shape
        logp = torch.empty(rows, A, dtype=torch.float32, device=logits.device)
        ent = torch.empty(rows, A, dtype=torch.float32, device=logits.device) if want_entropy else None
        check(_lib.load().cpm_heads_logp(_p(l2), rows, ld, _lib.int_array(seg), A, _p(tok), _p(logp), _p(ent), _dt(l2), _st()))
        ctx.save_for_backward(l2, tok)
        ctx.cfg = (seg, logits.shape, want_entropy)
        out_shape = tokens.shape
        if want_entropy:
            return logp.view(out_shape), ent.view(out_shape)
        return logp.view(out_shape), None

    @staticmethod
    def backward(ctx, glogp, gent):
        l2, tok = ctx.saved_tensors
        seg, shape, want_entropy = ctx.cfg
        rows, A = tok.shape
        glogp = glogp.reshape(rows, A).float().contiguous() if glogp is not None else None
        gent = gent.reshape(rows, A).float().contiguous() if (gent is not None and want_entropy) else None
        dl = torch.empty_like(l2)
        check(_lib.load().cpm_heads_logp_bwd(_p(l2), rows, l2.shape[-1], _lib.int_array(seg), A, _p(tok), _p(glogp), _p(gent),
                                             _p(dl), _dt(l2), _st()))
        return dl.view(shape), None, None, None


def heads_logp(logits, tokens, seg, want_entropy=False):
    """log softmax(logits[seg a])[token] (+ entropy) per attribute, differentiable w.r.t. logits."""
    return _HeadsLogp.apply(logits, tokens, tuple(seg), want_entropy)


class _MaskedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, mask, seg, group):
        _cuda(logits, targets, mask)
        ld = logits.shape[-1]
        l2 = logits.reshape(-1, ld).contiguous()
        A = len(seg) - 1
        tg = targets.reshape(-1, A).contiguous()
        m = mask.reshape(-1).to(torch.float32).contiguous()
        T = tg.shape[0]
        num = torch.zeros(A, dtype=torch.float32, device=logits.device)
        msum = torch.zeros(1, dtype=torch.float32, device=logits.device)
        lse = torch.empty(T, A, dtype=torch.float32, device=logits.device)
        check(_lib.load().cpm_masked_ce_fwd(_p(l2), T, ld, _lib.int_array(seg), A, _p(tg), _p(m), _p(num), _p(msum), _p(lse),
                                            _dt(l2), _st()))
        if group is not None:               # data parallel: the denominator is the GLOBAL mask sum
            import torch.distributed as dist
            packed = torch.cat([num, msum])
            dist.all_reduce(packed, group=group)
            num, msum = packed[:A].clone(), packed[A:].clone()
        # The returned loss is the GLOBAL loss; each rank back-propagates its own tokens' share, so
        # the data-parallel gradient all-reduce must SUM (see dist.py), not average.
        ctx.save_for_backward(l2, tg, m, lse, msum)
        ctx.cfg = (seg, logits.shape)
        return num / msum

    @staticmethod
    def backward(ctx, gloss):
        l2, tg, m, lse, msum = ctx.saved_tensors
        seg, shape = ctx.cfg
        A = len(seg) - 1
        gscale = gloss.float().contiguous()
        dl = torch.empty_like(l2)
        check(_lib.load().cpm_masked_ce_bwd(_p(l2), tg.shape[0], l2.shape[-1], _lib.int_array(seg), A, _p(tg), _p(m), _p(lse),
                                            _p(gscale), _p(msum), _p(dl), _dt(l2), _st()))
        return dl.view(shape), None, None, None, None


def masked_ce(logits, targets, mask, seg, group=None):
    """Per-attribute masked-mean cross-entropy (A,) fp32: sum_t m_t CE_ta / sum_t m_t
    (reference compute_loss, agent_pretrain.py:279-283)."""
    return _MaskedCE.apply(logits, targets, mask, tuple(seg), group)


# --------------------------------------------------------------------------- RL maths
def returns_scan(rewards, gamma, mode="compat", values=None, dones=None, last_value=None, lam=0.95):
    """rewards (B,T) fp32.  mode: 'compat' | 'togo' | 'gae'.  Returns ret (and adv for gae)."""
    _cuda(rewards)
    r = rewards.to(torch.float32).contiguous()
    B, T = r.shape
    code = {"compat": _lib.RET_COMPAT, "togo": _lib.RET_TOGO, "gae": _lib.RET_GAE}[mode]
    f = lambda t: None if t is None else t.to(torch.float32).contiguous()
    values, dones, last_value = f(values), f(dones), f(last_value)
    ret = torch.empty_like(r)
    adv = torch.empty_like(r) if code == _lib.RET_GAE else None
    check(_lib.load().cpm_returns_scan(_p(r), _p(values), _p(dones), _p(last_value), _p(ret), _p(adv), B, T, gamma, lam, code, _st()))
    return (adv, ret) if code == _lib.RET_GAE else ret


def zscore(x, sub=None, unbiased=True, eps=0.0, group=None):
    """((x - sub) - mean)/std over all elements; moments are all-reduced over `group` if given."""
    _cuda(x)
    x = x.to(torch.float32).contiguous()
    sub = None if sub is None else sub.to(torch.float32).contiguous().expand_as(x).contiguous()
    m3 = torch.zeros(3, dtype=torch.float64, device=x.device)
    lib = _lib.load()
    check(lib.cpm_moments(_p(x), _p(sub), x.numel(), _p(m3), _st()))
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(m3, group=group)
    out = torch.empty_like(x)
    check(lib.cpm_zscore(_p(x), _p(sub), _p(out), x.numel(), _p(m3), 1 if unbiased else 0, eps, _st()))
    return out


class _PPOLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, new_logp, old_logp, adv, entropy, value, ret, clip, vf_coef, ent_coef, mode):
        _cuda(new_logp, old_logp, adv)
        f = lambda t: None if t is None else t.detach().to(torch.float32).contiguous()
        nl, ol, ad, en, va, re = f(new_logp), f(old_logp), f(adv), f(entropy), f(value), f(ret)
        dev = nl.device
        out = torch.zeros(4, dtype=torch.float32, device=dev)
        dnew = torch.empty_like(nl)
        dent = torch.empty_like(en) if en is not None else None
        dval = torch.empty_like(va) if va is not None else None
        if mode == _lib.PPO_COMPAT:
            T = ol.shape[0]
            C = ol.numel() // T
            if nl.numel() != C or ad.numel() != T:
                raise ValueError("compat PPO loss: new_logp (C), old_logp (T,C), adv (T)")
            nv = 0
        else:
            T, C = nl.numel(), 1
            if ol.numel() != T or ad.numel() != T:
                raise ValueError("standard PPO loss: new_logp, old_logp, adv must have equal numel")
            nv = 0 if va is None else va.numel()
        check(_lib.load().cpm_ppo_loss_fwd_bwd(_p(nl), _p(ol), _p(ad), _p(en), _p(va), _p(re), _p(out), _p(dnew), _p(dent), _p(dval),
                                               T, C, nv, clip, vf_coef, ent_coef, 1.0, mode, _st()))
        ctx.save_for_backward(dnew, dent, dval)
        ctx.shapes = (new_logp.shape, None if entropy is None else entropy.shape, None if value is None else value.shape)
        return out

    @staticmethod
    def backward(ctx, gout):
        dnew, dent, dval = ctx.saved_tensors
        s_new, s_ent, s_val = ctx.shapes
        g = gout[0]
        return (dnew.view(s_new) * g, None, None, None if dent is None else dent.view(s_ent) * g,
                None if dval is None else dval.view(s_val) * g, None, None, None, None, None)


def ppo_loss_compat(new_logp, old_logp, adv, clip=0.2):
    """Reference surrogate (ppo_train.py:388-396). Returns the scalar loss (differentiable in new_logp)."""
    return _PPOLoss.apply(new_logp, old_logp, adv, None, None, None, clip, 0.0, 0.0, _lib.PPO_COMPAT)[0]


def ppo_loss_standard(new_logp, old_logp, adv, value=None, ret=None, entropy=None, clip=0.2, vf_coef=0.5, ent_coef=0.01):
    """Clipped surrogate + vf_coef*MSE - ent_coef*entropy. Returns tensor [loss, policy, value, entropy];
    only element 0 carries gradient."""
    return _PPOLoss.apply(new_logp, old_logp, adv, entropy, value, ret, clip, vf_coef, ent_coef, _lib.PPO_STANDARD)


class _DQNTD(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q_logits, next_logits, action, reward, done, seg, A, gamma, mode):
        _cuda(q_logits, next_logits, action)
        B, L, ld = q_logits.shape
        ql, nl = q_logits.contiguous(), next_logits.detach().contiguous()
        if nl.dtype != ql.dtype:
            nl = nl.to(ql.dtype)
        act = action.to(torch.int64).contiguous()
        rw = reward.reshape(-1).to(torch.float32).contiguous()
        dn = done.reshape(-1).to(torch.float32).contiguous()
        out = torch.zeros(1, dtype=torch.float32, device=ql.device)
        dq = torch.empty_like(ql)
        n_attr = len(seg) - 1
        tg = torch.empty(B, A, n_attr, dtype=torch.float32, device=ql.device)
        check(_lib.load().cpm_dqn_td_fwd_bwd(_p(ql), _p(nl), _p(act), _p(rw), _p(dn), _p(out), _p(dq), _p(tg), B, L, ld,
                                             _lib.int_array(seg), n_attr, A, gamma, 1.0, mode, _dt(ql), _st()))
        ctx.save_for_backward(dq)
        ctx.mark_non_differentiable(tg)
        return out[0], tg

    @staticmethod
    def backward(ctx, gloss, _gtg):
        (dq,) = ctx.saved_tensors
        return dq * gloss.to(dq.dtype), None, None, None, None, None, None, None, None


def dqn_td_loss(q_logits, next_logits, action, reward, done, seg, n_actions=25, gamma=0.95, compat=True):
    """Mean over attributes of MSE(Q(s,a), r + gamma(1-done) target) over concatenated logits (B,L,ld).
    Returns (loss, targets (B,A,n_attr)).  The target net's logits are treated as constants (the
    reference never steps the target net's parameters)."""
    return _DQNTD.apply(q_logits, next_logits, action, reward, done, tuple(seg), n_actions, gamma,
                        _lib.TD_COMPAT if compat else _lib.TD_STANDARD)


def reward_head(h, u, c, want_scores=False):
    """reward (N,) = mean_a sigmoid(mean_l h[n,l,:] . u_a + c_a) in one launch (cpm_reward_head).  h (N,L,d) bf16/fp32,
    u (A,d) fp32, c (A,) fp32.  No autograd: the reference only reads rewards (ppo_train.py:491)."""
    _cuda(h, u, c)
    h = h.contiguous()
    N, L, d = h.shape
    A = u.shape[0]
    reward = torch.empty(N, dtype=torch.float32, device=h.device)
    scores = torch.empty(N, A, dtype=torch.float32, device=h.device) if want_scores else None
    check(_lib.load().cpm_reward_head(_p(h), _p(u.float().contiguous()), _p(c.float().contiguous()), _p(reward), _p(scores), N, L, d, A,
                                      _dt(h), _st()))
    return (reward, scores) if want_scores else reward


class _RowDot(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, u, c):
        _cuda(h, u, c)
        h = h.contiguous()
        d = h.shape[-1]
        rows = h.numel() // d
        u32 = u.detach().to(torch.float32).contiguous()
        c32 = None if c is None else c.detach().to(torch.float32).reshape(1).contiguous()
        out = torch.empty(h.shape[:-1], dtype=torch.float32, device=h.device)
        check(_lib.load().cpm_rowdot_fwd(_p(h), _p(u32), _p(c32), _p(out), rows, d, _dt(h), _st()))
        ctx.save_for_backward(h, u32)
        ctx.has_c = c is not None
        return out

    @staticmethod
    def backward(ctx, g):
        h, u32 = ctx.saved_tensors
        d = h.shape[-1]
        rows = h.numel() // d
        g = g.to(torch.float32).contiguous()
        lib = _lib.load()
        dh = torch.empty_like(h) if ctx.needs_input_grad[0] else None
        du = torch.empty(d, dtype=torch.float32, device=h.device)
        partials = torch.empty(lib.cpm_rowdot_partials_rows() * d, dtype=torch.float32, device=h.device)
        check(lib.cpm_rowdot_bwd(_p(h), _p(g), _p(u32), _p(dh), _p(du), _p(partials), rows, d, _dt(h), _st()))
        return dh, du, (g.sum().reshape(()) if ctx.has_c else None)


def rowdot(h, u, c=None):
    """out[...] = h[..., :] . u + c in fp32 (cpm_rowdot_fwd/bwd): differentiable in h, u (d,) and the scalar tensor c."""
    return _RowDot.apply(h, u, c)


def rollout_advance(tokens, history_tok, vals, history_f, step_dev, max_steps):
    n_tok = tokens.numel() if tokens is not None else 0
    n_f = vals.numel() if vals is not None else 0
    check(_lib.load().cpm_rollout_advance(_p(tokens), _p(history_tok), n_tok, _p(vals), _p(history_f), n_f, _p(step_dev),
                                          max_steps, _st()))
