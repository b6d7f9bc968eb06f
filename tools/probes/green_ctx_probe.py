"""Can a stream be confined to a subset of the SMs (CUDA green contexts, driver API through cuda-python) and used from torch?
Times a large bf16 matmul on the default stream and on a green-context stream of `sms` SMs.
    python tools/probes/green_ctx_probe.py [sms]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from cuda.bindings import driver as drv

def green_stream(sms, priority=0, device=0):
    def ck(r):
        if r[0] != drv.CUresult.CUDA_SUCCESS:
            raise RuntimeError(f"driver call failed: {r[0]}")
        return r[1:] if len(r) > 2 else r[1]
    torch.cuda.init()
    torch.zeros(1, device=f"cuda:{device}")
    dev = ck(drv.cuDeviceGet(device))
    res = ck(drv.cuDeviceGetDevResource(dev, drv.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
    print("device SMs:", res.sm.smCount)
    groups, nb, remaining = ck(drv.cuDevSmResourceSplitByCount(1, res, 0, sms))
    print("group SMs:", groups[0].sm.smCount, "remaining:", remaining.sm.smCount)
    desc = ck(drv.cuDevResourceGenerateDesc([groups[0]], 1))
    gctx = ck(drv.cuGreenCtxCreate(desc, dev, drv.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
    st = ck(drv.cuGreenCtxStreamCreate(gctx, drv.CUstream_flags.CU_STREAM_NON_BLOCKING, priority))
    return torch.cuda.ExternalStream(int(st), device=f"cuda:{device}"), gctx, groups[0].sm.smCount

if __name__ == "__main__":
    sms = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    s, gctx, got = green_stream(sms)
    a = torch.randn(8192, 8192, device="cuda").bfloat16(); b = torch.randn(8192, 8192, device="cuda").bfloat16()
    def timeit(stream):
        with torch.cuda.stream(stream):
            for _ in range(3): a @ b
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): a @ b
            e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 10
    print("matmul 8192^3 on the default stream: %.3f ms" % timeit(torch.cuda.current_stream()))
    print("matmul 8192^3 on the green stream (%d SMs): %.3f ms" % (got, timeit(s)))
    import cpmusic
    x = torch.randn(131072, 512, device="cuda").bfloat16(); w = torch.randn(1536, 512, device="cuda").bfloat16()
    def own(stream):
        with torch.cuda.stream(stream):
            for _ in range(3): cpmusic.ops.gemm_nt(x, w)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): cpmusic.ops.gemm_nt(x, w)
            e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 10
    print("own QKV GEMM on the default stream: %.3f ms; on the green stream: %.3f ms" % (own(torch.cuda.current_stream()), own(s)))
