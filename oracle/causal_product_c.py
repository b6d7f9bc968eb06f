"""ctypes binding + autograd wrapper for ``oracle/causal_product.c``.

TEST INFRASTRUCTURE ONLY — see ``oracle/__init__.py``.  Mirrors how ft binds its
CPU kernel: ``causal_dot_product(Q,K,V,product)`` /
``causal_dot_product_backward(Q,K,V,grad_out,gQ,gK,gV)`` on (N,H,L,E) fp32
contiguous tensors with outputs zero-initialised by Python.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcausal_product_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "causal_product.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        base = ["-O3", "-fPIC", "-shared", "-o", _SO, src]
        for cc, extra in (("/usr/bin/gcc", ["-fopenmp"]), ("gcc", ["-fopenmp"]), ("/usr/bin/gcc", []), ("gcc", [])):
            try:
                subprocess.run([cc] + extra + base, check=True, capture_output=True)
                break
            except (subprocess.CalledProcessError, FileNotFoundError):
                continue
        else:
            raise RuntimeError("could not compile oracle/causal_product.c")
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        fp = ctypes.POINTER(ctypes.c_float)
        _lib.oracle_causal_dot_product.argtypes = [fp] * 4 + [ctypes.c_int] * 5
        _lib.oracle_causal_dot_product_backward.argtypes = [fp] * 7 + [ctypes.c_int] * 5
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def _p(t):
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_float))


class CausalDotProductC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Q, K, V):
        Q, K, V = (t.detach().float().contiguous() for t in (Q, K, V))
        N, H, L, E = Q.shape
        M = V.shape[-1]
        out = torch.zeros(N, H, L, M)
        lib().oracle_causal_dot_product(_p(Q), _p(K), _p(V), _p(out), N, H, L, E, M)
        ctx.save_for_backward(Q, K, V)
        return out

    @staticmethod
    def backward(ctx, G):
        Q, K, V = ctx.saved_tensors
        G = G.float().contiguous()
        N, H, L, E = Q.shape
        M = V.shape[-1]
        gQ, gK, gV = torch.zeros_like(Q), torch.zeros_like(K), torch.zeros_like(V)
        lib().oracle_causal_dot_product_backward(_p(Q), _p(K), _p(V), _p(G), _p(gQ), _p(gK), _p(gV), N, H, L, E, M)
        return gQ, gK, gV


def causal_dot_product_c(Q, K, V):
    return CausalDotProductC.apply(Q, K, V)


def num_threads() -> int:
    return int(lib().oracle_num_threads())
