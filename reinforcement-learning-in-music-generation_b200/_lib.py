"""ctypes binding of libcpmusic.so (the C-ABI declared in include/cpmusic.h).

There is NO CPU fallback: if the library is missing or cannot be loaded every
compute entry point raises.  ``load()`` is lazy so that importing the package on
a CPU-only box (module definitions, state_dict handling, tests that only check
exported symbols) works.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CPM_LIB_PATH", os.path.join(_HERE, "libcpmusic.so"))      # override: A/B runs of kernel variants

F32, BF16 = 0, 1
RET_COMPAT, RET_TOGO, RET_GAE = 0, 1, 2
PPO_COMPAT, PPO_STANDARD = 0, 1
TD_COMPAT, TD_STANDARD = 0, 1
MAX_ATTR = 8

ERR_NAMES = {0: "CPM_OK", -1: "CPM_ERR_BAD_SHAPE", -2: "CPM_ERR_BAD_ALIGN", -3: "CPM_ERR_BAD_DTYPE",
             -4: "CPM_ERR_NULL", -5: "CPM_ERR_WORKSPACE", -6: "CPM_ERR_CUDA", -7: "CPM_ERR_UNSUPPORTED"}

_P = c_void_p
_FPP = POINTER(c_void_p)
_IP = POINTER(c_int)
_FP = POINTER(c_float)

# name -> (restype, argtypes); mirrors include/cpmusic.h one to one
SIGNATURES = {
    "cpm_version": (c_int, []),
    "cpm_last_error_string": (c_char_p, []),
    "cpm_error_name": (c_char_p, [c_int]),
    "cpm_linattn_last_impl": (c_char_p, []),
    "cpm_linattn_workspace_bytes": (c_int64, [c_int, c_int, c_int]),
    "cpm_linattn_saved_bytes": (c_int64, [c_int, c_int, c_int]),
    "cpm_linattn_workspace_bytes_wide": (c_int64, [c_int, c_int, c_int, c_int]),
    "cpm_linattn_saved_bytes_wide": (c_int64, [c_int, c_int, c_int, c_int]),
    "cpm_debug_linattn_timing": (c_int, [_P]),
    "cpm_linattn_fwd": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int64, c_int64,
                                c_int, c_float, c_int, _P, c_int64, _P, c_int64, _P]),
    "cpm_linattn_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int,
                                c_int64, c_int64, c_int64, c_int, c_float, c_int, _P, c_int64, _P, c_int64, _P]),
    "cpm_linattn_step": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int64, c_int64,
                                 c_int, c_float, _P]),
    "cpm_pack_item_bytes": (c_int, []),
    "cpm_pack_weights": (c_int, [_P, c_int, c_int, _P]),
    "cpm_colsum_partials_rows": (c_int, [c_int]),
    "cpm_colsum": (c_int, [_P, c_int64, c_int, c_int64, _P, _P, c_int, _P]),
    "cpm_reward_head": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "cpm_set_rng_base": (c_int, [_P]),
    "cpm_rowdot_partials_rows": (c_int, []),
    "cpm_rowdot_fwd": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, _P]),
    "cpm_rowdot_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, _P]),
    "cpm_gemm_nt": (c_int, [_P, c_int64, _P, c_int64, _P, c_int64, _P, c_int64, c_int, c_int, c_int, _P, c_int, _P, c_int64,
                            c_float, c_uint64, c_uint64, _P]),
    "cpm_gemm_set_mode": (c_int, [c_int]),
    "cpm_set_chain_pdl": (c_int, [c_int]),
    "cpm_gemm_nt_small": (c_int, [_P, c_int64, _P, c_int64, _P, c_int64, c_int, c_int, c_int, _P, c_int, _P]),
    "cpm_gemm_nt_small_ln": (c_int, [_P, c_int64, _P, c_int64, _P, c_int64, c_int, c_int, c_int, _P, c_int, _P, _P, c_float, _P, c_int64,
                                     _P, _P, _P, _P]),
    "cpm_debug_small_timing": (c_int, [_P, c_int]),
    "cpm_gemm_small_set_split": (c_int, [c_int]),
    "cpm_gemm_tn": (c_int, [_P, c_int64, _P, c_int64, _FPP, c_int, c_int, c_int64, c_int, c_int, c_int, _P]),
    "cpm_embed_fwd": (c_int, [_P, _FPP, _IP, _IP, c_int, c_int64, _P, c_int, _P, _P]),
    "cpm_embed_bwd": (c_int, [_P, _P, _FPP, _IP, _IP, c_int, c_int64, c_int, _P]),
    "cpm_add_pe": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_int, _P, c_int, c_float, c_uint64, c_uint64, c_int, _P]),
    "cpm_dropout": (c_int, [_P, _P, c_int64, c_float, c_uint64, c_uint64, c_int, _P]),
    "cpm_ln_residual_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_float, c_float,
                                    c_uint64, c_uint64, c_int, _P]),
    "cpm_ln_partials_rows": (c_int, []),
    "cpm_ln_residual_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_float,
                                    c_uint64, c_uint64, c_int, _P]),
    "cpm_gelu_fwd": (c_int, [_P, _P, _P, c_int64, c_int, c_float, c_uint64, c_uint64, c_int, _P]),
    "cpm_gelu_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_float, c_uint64, c_uint64, c_int, _P]),
    "cpm_gelu_bwd_partials_rows": (c_int, [c_int]),
    "cpm_heads_sample": (c_int, [_P, c_int64, c_int64, _IP, c_int, _FP, _FP, c_int, c_uint64, c_int64, c_int,
                                 _P, _P, _P, _P, c_int, _P]),
    "cpm_heads_logp": (c_int, [_P, c_int64, c_int64, _IP, c_int, _P, _P, _P, c_int, _P]),
    "cpm_heads_logp_bwd": (c_int, [_P, c_int64, c_int64, _IP, c_int, _P, _P, _P, _P, c_int, _P]),
    "cpm_masked_ce_fwd": (c_int, [_P, c_int64, c_int64, _IP, c_int, _P, _P, _P, _P, _P, c_int, _P]),
    "cpm_masked_ce_bwd": (c_int, [_P, c_int64, c_int64, _IP, c_int, _P, _P, _P, _P, _P, _P, c_int, _P]),
    "cpm_returns_scan": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_float, c_float, c_int, _P]),
    "cpm_moments": (c_int, [_P, _P, c_int64, _P, _P]),
    "cpm_zscore": (c_int, [_P, _P, _P, c_int64, _P, c_int, c_float, _P]),
    "cpm_ppo_loss_fwd_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int64,
                                     c_float, c_float, c_float, c_float, c_int, _P]),
    "cpm_dqn_td_fwd_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int64, _IP, c_int, c_int,
                                   c_float, c_float, c_int, c_int, _P]),
    "cpm_rollout_advance": (c_int, [_P, _P, c_int64, _P, _P, c_int64, _P, c_int32, _P]),
    "cpm_rollout_plan_bytes": (c_int64, []),
    "cpm_rollout_create": (c_int, [_P, _P, POINTER(c_void_p)]),
    "cpm_rollout_run": (c_int, [_P, c_int, _P]),
    "cpm_rollout_phases": (c_int, [_P]),
    "cpm_debug_rollout_timing": (c_int, [_P]),
    "cpm_rollout_destroy": (c_int, [_P]),
}

ROLLOUT_MAX_LAYERS = 32


class RolloutLayer(ctypes.Structure):
    """CpmRolloutLayer (include/cpmusic.h)."""
    _fields_ = [(n, c_void_p) for n in ("w_qkv", "b_qkv", "w_out", "b_out", "ln1_g", "ln1_b", "w_ff1", "b_ff1", "w_ff2", "b_ff2",
                                        "ln2_g", "ln2_b", "S", "Z")]


class RolloutConfig(ctypes.Structure):
    """CpmRolloutConfig (include/cpmusic.h), field for field."""
    _fields_ = [("batch", c_int), ("d_model", c_int), ("n_heads", c_int), ("d_ff", c_int), ("n_layers", c_int), ("n_attr", c_int),
                ("n_tokens", c_int * MAX_ATTR), ("emb", c_int * MAX_ATTR), ("tables", c_void_p * MAX_ATTR),
                ("w_in", c_void_p), ("b_in", c_void_p), ("pe", c_void_p), ("pe_len", c_int), ("true_positions", c_int),
                ("layer", RolloutLayer * ROLLOUT_MAX_LAYERS), ("lnf_g", c_void_p), ("lnf_b", c_void_p),
                ("w_heads", c_void_p), ("b_heads", c_void_p), ("seg", c_int * (MAX_ATTR + 1)), ("logits_ld", c_int),
                ("temperature", c_float * MAX_ATTR), ("top_p", c_float * MAX_ATTR), ("greedy", c_int),
                ("ln_eps", c_float), ("attn_eps", c_float), ("seed", c_uint64), ("seq_base", c_int64),
                ("cur", c_void_p), ("logp", c_void_p), ("hist_tok", c_void_p), ("hist_logp", c_void_p),
                ("step_dev", c_void_p), ("max_steps", c_int32),
                ("x0", c_void_p), ("x1", c_void_p), ("y", c_void_p), ("qkv", c_void_p), ("attn", c_void_p), ("g", c_void_p),
                ("logits", c_void_p), ("barrier", c_void_p), ("err_flag", c_void_p)]

_lib = None

# ---- launch accounting (bench.py's `gpu_launches`): C-ABI calls made, and kernels per call -------
import collections
COUNTS = collections.Counter()
KERNELS_PER_CALL = collections.defaultdict(lambda: 1, {
    "cpm_version": 0, "cpm_last_error_string": 0, "cpm_error_name": 0, "cpm_linattn_last_impl": 0,
    "cpm_linattn_workspace_bytes": 0, "cpm_linattn_saved_bytes": 0, "cpm_linattn_workspace_bytes_wide": 0, "cpm_linattn_saved_bytes_wide": 0, "cpm_set_rng_base": 0, "cpm_gelu_bwd_partials_rows": 0, "cpm_debug_linattn_timing": 0, "cpm_ln_partials_rows": 0, "cpm_colsum_partials_rows": 0, "cpm_pack_item_bytes": 0, "cpm_rowdot_partials_rows": 0, "cpm_gemm_set_mode": 0, "cpm_set_chain_pdl": 0,
    "cpm_rollout_plan_bytes": 0, "cpm_debug_rollout_timing": 0, "cpm_debug_small_timing": 0, "cpm_gemm_small_set_split": 0, "cpm_rollout_create": 0, "cpm_rollout_phases": 0, "cpm_rollout_destroy": 0,

    "cpm_linattn_fwd": 2, "cpm_linattn_bwd": 2, "cpm_ln_residual_bwd": 2, "cpm_colsum": 2, "cpm_rowdot_bwd": 2,      # chunk-parallel path: streaming state kernel + per-chunk
                                                                                # kernel (ops adds the scan when the per-chunk state path runs)
})
EXTRA_LAUNCHES = [0]          # segment-total / scan kernels of segmented linear attention, graph replays


def kernel_launches() -> int:
    return sum(n * KERNELS_PER_CALL[name] for name, n in COUNTS.items()) + EXTRA_LAUNCHES[0]


def reset_counts() -> None:
    COUNTS.clear()
    EXTRA_LAUNCHES[0] = 0


class _Counted:
    """The loaded library with a per-entry-point call counter in front of every function."""

    def __init__(self, cdll):
        self._cdll = cdll
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(cdll, name)      # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, self._wrap(name, fn))

    @staticmethod
    def _wrap(name, fn):
        def call(*a):
            COUNTS[name] += 1
            return fn(*a)
        call.__name__ = name
        return call


class CpmError(RuntimeError):
    """A libcpmusic entry point returned a negative code."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


def load() -> ctypes.CDLL:
    """Load libcpmusic.so or raise — never falls back to a CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python __graft_entry__.py` (or "
                f"`python {os.path.join(_HERE, 'build.py')}`); this package has no CPU fallback.")
        import torch  # noqa: F401  (its libcudart.so.12 is the runtime the library binds to; the toolkit's copy otherwise)
        _lib = _Counted(ctypes.CDLL(LIB_PATH))
    return _lib


def check(rc: int) -> None:
    """Map an error code to the exception type the reference stack would raise
    (SURVEY §8b): shape/dtype problems -> ValueError, the rest -> RuntimeError."""
    if rc == 0:
        return
    msg = load().cpm_last_error_string().decode()
    if rc in (-1, -3):
        raise ValueError(f"{ERR_NAMES.get(rc, rc)}: {msg}")
    raise CpmError(rc, msg)


def int_array(vals):
    return (c_int * len(vals))(*[int(v) for v in vals])


def float_array(vals):
    return (c_float * len(vals))(*[float(v) for v in vals])


def ptr_array(ptrs):
    return (c_void_p * len(ptrs))(*[int(p) for p in ptrs])
