"""Effective SM clock around the PPO phases: a fixed-cycle spin kernel (torch.cuda._sleep) timed with CUDA events right before and
right after the rollout of every iteration.  NVML keeps reporting the nominal 1965 MHz; this shows what the SMs really run at.
    python tools/clock_probe.py            (one GPU;   torchrun ... tools/clock_probe.py for the data-parallel case)"""
import os, sys
if os.environ.get("PROBE_MASK", "0") == "1" and "LOCAL_RANK" in os.environ:      # one visible GPU per process, set before CUDA starts
    os.environ["CUDA_VISIBLE_DEVICES"] = os.environ["LOCAL_RANK"]
    os.environ["LOCAL_RANK"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
import ctypes
_rt = ctypes.CDLL("libcudart.so.12")


def get_limit(which):
    v = ctypes.c_size_t(0)
    rc = _rt.cudaDeviceGetLimit(ctypes.byref(v), ctypes.c_int(which))
    _rt.cudaGetLastError()
    return v.value if rc == 0 else -1


torch.zeros(1, device=dev)
limits_before = {n: get_limit(i) for i, n in ((0, "stack"), (2, "malloc_heap"), (5, "max_l2_fetch"), (6, "persisting_l2"))}
MODE = os.environ.get("PROBE_MODE", "dp")          # dp | nccl-idle (communicator built, never used by the loop) | independent
if world > 1 and MODE not in ("independent", "nccl-late"):
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
    t = torch.ones(1, device=dev); dist.all_reduce(t)
limits_after = {n: get_limit(i) for i, n in ((0, "stack"), (2, "malloc_heap"), (5, "max_l2_fetch"), (6, "persisting_l2"))}
if rank == 0:
    print("device limits before NCCL:", limits_before, "\n              after NCCL:", limits_after, flush=True)
if MODE == "nccl-idle-reset":
    for i, n in ((0, "stack"),):
        _rt.cudaDeviceSetLimit(ctypes.c_int(i), ctypes.c_size_t(limits_before[n]))
    print("stack limit reset to", get_limit(0), flush=True)
if MODE != "dp":
    world_used = 1
else:
    world_used = world
CYC = 4_000_000


def spin_mhz():
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); torch.cuda._sleep(CYC); b.record()
    return a, b


it = bench.PPOIteration(rank, world_used, dev)
for _ in range(2):
    it.step(it.init_dev)
torch.cuda.synchronize()
for i in range(6 if MODE == "nccl-late" else 4):
    if MODE == "nccl-late" and i == 3 and world > 1:      # communicator built AFTER every buffer of the loop exists
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        t = torch.ones(1, device=dev); dist.all_reduce(t); torch.cuda.synchronize()
        print(f"rank {rank}: NCCL communicator built", flush=True)
    pre = spin_mhz()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time as _t
    h0 = _t.perf_counter(); r0.record(); roll = it.engine.generate(it.init_dev); r1.record(); h1 = _t.perf_counter()
    post = spin_mhz()
    it.step(it.init_dev)                      # a full iteration (rollout + update) to re-create the heavy phase
    after_update = spin_mhz()
    torch.cuda.synchronize()
    f = lambda ev: CYC / (ev[0].elapsed_time(ev[1]) * 1e3)
    print(f"[{MODE}] rank {rank}/{world} iter {i}: SM MHz before rollout {f(pre):.0f}, after rollout {f(post):.0f}, after update {f(after_update):.0f}; "
          f"rollout {r0.elapsed_time(r1) * 1e3 / 1024:.1f} us/token (host enqueue {(h1 - h0) * 1e6 / 1024:.1f} us/token)", flush=True)
