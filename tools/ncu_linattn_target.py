"""Target program for an ncu capture of the linear-attention kernels at the update minibatch shape (128 x 1024 x 8):
two forward + backward calls (the first is the warm-up the capture skips with --launch-skip).
    ncu --set full --clock-control none --import-source on -k regex:cp_ --launch-skip <K> --launch-count <K> \\
        -o gpurun_out/linattn python tools/ncu_linattn_target.py
K = kernels per forward + backward call = 4 (streaming state kernel + per-chunk kernel, each direction)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cpmusic

N, L, H = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (128, 1024, 8)))
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
qkv = torch.randn(N, L, 3 * H * 64, generator=g).to(dev).bfloat16()
q, k, v = (qkv[..., i * H * 64:(i + 1) * H * 64].unflatten(-1, (H, 64)) for i in range(3))
go = torch.randn(N, L, H, 64, generator=g).to(dev).bfloat16()
gqkv = torch.empty_like(qkv)
gq, gk, gv = (gqkv[..., i * H * 64:(i + 1) * H * 64].unflatten(-1, (H, 64)) for i in range(3))
saved = cpmusic.ops.linattn_saved(N, L, H, dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for _ in range(2):
    flush.zero_()
    out, den = cpmusic.ops.linattn_fwd_raw(q, k, v, impl=0, saved=saved)
    flush.zero_()
    cpmusic.ops.linattn_bwd_raw(q, k, v, out, den, go, gq, gk, gv, impl=0, saved=saved)
torch.cuda.synchronize()
print("kernels per fwd+bwd: 4", cpmusic.ops.linattn_last_impl())
