"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch on
B200; gloo in the CPU tests), rollout batches sharded by song, bucketed gradient all-reduce
overlapped with backward.

The reference has no distributed code at all; this layer exists because
BASELINE.json's north_star partitions the box data-parallel.  Exactness rules:
* losses are defined over the GLOBAL batch (the masked-CE denominator and advantage moments are
  all-reduced inside the ops), every rank back-propagates its own tokens' share, therefore the
  gradient all-reduce SUMS (no averaging);
* per-sequence Philox streams are keyed by global sequence id, so sampled tokens are identical for
  any GPU count.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None):
    """Initialise the default process group from RANK / WORLD_SIZE / MASTER_* (torchrun)."""
    if dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world == 1:
        return 0, 1
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n_items for `rank` (songs / sequences are independent units)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class BucketedGradAllReduce:
    """Flat fp32 gradient buckets (~bucket_mb each, reverse parameter order so the first bucket to
    fill is the last layer's); ``p.grad`` are views into the buckets, so autograd accumulates in
    place and a bucket is all-reduced (SUM, async on the communication stream) the moment its last
    gradient lands — overlapping with the rest of backward.

    Gradient accumulation: tell the reducer how many backward passes make up one optimizer step
    (``zero_grad(n_micro=k)`` or ``begin(k)``).  A bucket is reduced when its last gradient of the
    LAST micro-batch lands; the earlier passes only accumulate locally.  (Reducing at the end of the
    first pass - what a plain per-backward counter does - would race later local accumulation with
    the in-flight collective and never reduce the later micro-batches.)"""

    def __init__(self, params, bucket_mb: float = 25.0, group=None):
        self.group = group
        self.params = [p for p in params if p.requires_grad]
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        cap = int(bucket_mb * 1024 * 1024 / 4)
        self.buckets: List[dict] = []
        cur, cur_n = [], 0
        for p in reversed(self.params):
            if cur and cur_n + self._padded(p.numel()) > cap:
                self._close(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += self._padded(p.numel())
        if cur:
            self._close(cur)
        self._handles = []
        self._hooks = []
        self.n_micro = 1
        for bi, b in enumerate(self.buckets):
            for p in b["params"]:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(bi)))

    def attach(self, group=None):
        """(Re)binds the reducer to a process group created AFTER it was constructed (buckets and hooks are unchanged)."""
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        return self

    ALIGN = 32          # floats: every gradient view starts on a 128-byte boundary (the GEMM kernels accumulate into them with
                        # 16-byte vector atomics; odd-sized parameters such as a 135-wide value head would misalign the rest)

    @classmethod
    def _padded(cls, n):
        return -(-n // cls.ALIGN) * cls.ALIGN

    def _close(self, plist):
        n = sum(self._padded(p.numel()) for p in plist)
        flat = torch.zeros(n, dtype=torch.float32, device=plist[0].device)
        off = 0
        for p in plist:
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += self._padded(p.numel())
        self.buckets.append({"params": plist, "flat": flat, "pending": len(plist), "launched": False, "round": 0})

    def _make_hook(self, bi):
        def hook(_p):
            b = self.buckets[bi]
            b["pending"] -= 1
            if b["pending"] == 0:
                b["round"] += 1
                if b["round"] >= self.n_micro:
                    self._launch(b)
                else:                                   # more micro-batches to come: keep accumulating locally
                    b["pending"] = len(b["params"])
        return hook

    def _launch(self, b):
        if b["launched"]:
            return
        b["launched"] = True
        if self.world > 1:
            self._handles.append(dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Call after backward: flush buckets whose parameters got no gradient, wait for all."""
        for b in self.buckets:
            self._launch(b)
        for h in self._handles:
            h.wait()
        self._handles.clear()
        for b in self.buckets:
            b["pending"], b["launched"], b["round"] = len(b["params"]), False, 0

    def begin(self, n_micro: int = 1):
        """Number of backward passes (micro-batches) that accumulate into the buckets before ``finish()``."""
        if n_micro < 1:
            raise ValueError("n_micro must be >= 1")
        self.n_micro = int(n_micro)
        for b in self.buckets:
            b["pending"], b["launched"], b["round"] = len(b["params"]), False, 0
        return self

    def zero_grad(self, n_micro: int = None):
        if n_micro is not None:
            self.begin(n_micro)
        for b in self.buckets:
            b["flat"].zero_()
            off = 0
            for p in b["params"]:          # re-attach views if an optimizer dropped them
                if p.grad is None or p.grad.data_ptr() != b["flat"].data_ptr() + 4 * off:
                    p.grad = b["flat"][off:off + p.numel()].view_as(p)
                off += self._padded(p.numel())

    def grad_bytes(self) -> int:
        return sum(b["flat"].numel() * 4 for b in self.buckets)

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks.clear()
