import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, cpmusic
from cpmusic import ops
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_gemm import timed
dev = torch.device("cuda:0"); T = 131072
x = torch.randn(T, 512, device=dev).bfloat16(); w = (torch.randn(2048, 512, device=dev) / 512 ** 0.5).bfloat16(); bias = torch.randn(2048, device=dev)
w2t = (torch.randn(2048, 512, device=dev) / 45).bfloat16()       # dgrad of linear2: (T,512) x (2048,512)^T
dyo = torch.randn(T, 512, device=dev).bfloat16(); h = torch.randn(T, 2048, device=dev).bfloat16()
for mode, tag in ((1, "stream"), (3, "wide")):
    ops.gemm_set_mode(mode)
    r = {"mode": tag,
         "ff1_fwd_plain_us": round(timed(lambda: ops.gemm_nt(x, w, bias), 20), 1),
         "ff1_fwd_gelu_drop_us": round(timed(lambda: ops.gemm_nt(x, w, bias, epilogue=ops.GEMM_GELU, p_drop=0.1, seed=1, rng_offset=0), 20), 1),
         "ff1_fwd_gelu_nodrop_us": round(timed(lambda: ops.gemm_nt(x, w, bias, epilogue=ops.GEMM_GELU, p_drop=0.0), 20), 1),
         "ff2_dgrad_plain_us": round(timed(lambda: ops.gemm_nt(dyo, w2t), 20), 1),
         "ff2_dgrad_dgelu_us": round(timed(lambda: ops.gemm_nt(dyo, w2t, epilogue=ops.GEMM_DGELU, aux=h, p_drop=0.1, seed=1, rng_offset=0), 20), 1)}
    print(json.dumps(r), flush=True)
ops.gemm_set_mode(0)
hh = ops.gemm_nt(x, w, bias)
print(json.dumps({"gelu_fwd_kernel_us": round(timed(lambda: ops.gelu_dropout(hh, 0.1), 20), 1)}))
