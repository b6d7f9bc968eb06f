"""Whole-step CUDA graphs for teacher-forced training.

The reference's pretraining loop (agent_pretrain.py:535-577) is one host-driven step per batch; at its batch sizes
(4 x 512 in the reference, 32 x 512 in BASELINE cfg2) a step is ~1500 kernel launches of a few microseconds each and the
host launch rate, not the GPU, bounds it.  ``GraphedTrainStep`` captures forward (``train_step``), backward, gradient
clipping and the Adam update ONCE into a CUDA graph over static input buffers and replays it per batch:

* the bf16 weight packings are marked stale right before capture, so their in-place refresh from the fp32 masters is part
  of the graph and every replay sees the previous replay's optimizer update;
* dropout masks stay fresh: the kernels add a device-side counter to their (captured, hence frozen) RNG offsets and the
  graph's last node advances that counter (``ops.enable_device_rng`` / ``cpm_set_rng_base``);
* the optimizer must be created with ``capturable=True`` (its step counters live on the device).

Single process / single GPU: the data-parallel gradient all-reduce hooks are not captured.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


class GraphedTrainStep:
    def __init__(self, model, optimizer, batch_size: int, seq_len: int, max_grad_norm: Optional[float] = 3.0, warmup: int = 3):
        if not all(g.get("capturable", False) for g in optimizer.param_groups):
            raise ValueError("GraphedTrainStep needs an optimizer built with capturable=True (e.g. torch.optim.Adam(..., fused=True, "
                             "capturable=True)): its step counters must live on the device")
        self.model, self.opt, self.max_grad_norm = model, optimizer, max_grad_norm
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("cpmusic ops need CUDA tensors: there is no CPU fallback")
        A = len(model.attrs)
        self.x = torch.zeros(batch_size, seq_len, A, dtype=torch.int64, device=dev)
        self.y = torch.zeros(batch_size, seq_len, A, dtype=torch.int64, device=dev)
        self.mask = torch.ones(batch_size, seq_len, dtype=torch.float32, device=dev)
        self.losses = torch.zeros(A, dtype=torch.float32, device=dev)
        self.grad_norm = torch.zeros((), dtype=torch.float32, device=dev)
        ops.enable_device_rng(dev)
        model.train()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):           # allocates gradients / optimizer state, warms cuBLAS and the caches
                self._step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        model.invalidate_packs()                      # the refresh of every weight packing becomes part of the graph
        self.graph = torch.cuda.CUDAGraph()
        start = ops._Rng.offset
        with ops.graph_capture(self.graph):
            self._step()
            ops.rng_advance(ops._Rng.offset - start)  # next replay: same frozen offsets + an advanced device base
        self.rng_counters_per_step = ops._Rng.offset - start

    def _step(self):
        losses = self.model.train_step(self.x, self.y, self.mask)
        loss = sum(losses) / len(losses)
        self.opt.zero_grad(set_to_none=False)
        loss.backward()
        if self.max_grad_norm is not None:
            self.grad_norm.copy_(torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.max_grad_norm, foreach=True))
        self.opt.step()
        self.losses.copy_(torch.stack([l.detach().float() for l in losses]))

    def __call__(self, x, target, loss_mask):
        """Copies the batch into the static buffers (device or pinned-host tensors, no sync), replays the step and returns
        the per-attribute losses (a static device tensor: clone it to keep a history)."""
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(target, non_blocking=True)
        self.mask.copy_(loss_mask, non_blocking=True)
        self.graph.replay()
        self.model.invalidate_packs()                 # the masters moved under the host-side version stamps
        return self.losses


# --------------------------------------------------------------------------- SM partitions for side streams
_PARTITIONS = []        # (green context, stream handle) kept alive for the life of the process


def sm_partition_stream(n_sms: int, device=None, priority: int = 0):
    """A torch stream whose kernels run on ``n_sms`` SMs of the device only (a CUDA green context: driver API through
    cuda-python, CUDA >= 12.4), sharing the primary context's memory.  For bulk work queued UNDERNEATH a latency-bound chain -
    the critic update under the next rollout in bench.py: persistent GEMM kernels otherwise hold every SM for 150-300 us at a
    time, and the token step's small dependent kernels, whatever their stream priority, cannot start until one of them ends
    (measured, tools/probes/overlap_timeline.py: rollout 450 -> 533 ms with the update alongside, update 111 -> 190 ms - the two
    streams time-slice).  Returns ``(stream, sms_granted)``; raises RuntimeError when green contexts are not available.
    Kernels sized for the whole device still run correctly on the partition (their grids queue in waves)."""
    from cuda.bindings import driver as drv

    def ck(r):
        if r[0] != drv.CUresult.CUDA_SUCCESS:
            raise RuntimeError(f"green context: driver call failed with {r[0]}")
        return r[1:] if len(r) > 2 else r[1]

    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    torch.zeros(1, device=dev)                                   # the primary context exists
    cudev = ck(drv.cuDeviceGet(dev.index or 0))
    res = ck(drv.cuDeviceGetDevResource(cudev, drv.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
    groups, _, _ = ck(drv.cuDevSmResourceSplitByCount(1, res, 0, int(n_sms)))
    desc = ck(drv.cuDevResourceGenerateDesc([groups[0]], 1))
    gctx = ck(drv.cuGreenCtxCreate(desc, cudev, drv.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
    handle = ck(drv.cuGreenCtxStreamCreate(gctx, drv.CUstream_flags.CU_STREAM_NON_BLOCKING, int(priority)))
    _PARTITIONS.append((gctx, handle))
    return torch.cuda.ExternalStream(int(handle), device=dev), int(groups[0].sm.smCount)
