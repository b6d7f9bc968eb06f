"""CP data formats and device-resident experience buffers.

* ``load_cp_npz`` / ``CPBatches``: the reference's ``train_data_linear.npz`` (``x``, ``y`` ``(songs, L, 7)`` int with
  the ``type`` attribute in column 3, ``mask`` ``(songs, L)``; agent_pretrain.py:491-526, IRL_dqn_train.py:418-434) →
  6-attribute int64 tensors in pinned host memory, and an iterator that keeps one batch in flight to the GPU on a side
  stream (the reference does a blocking ``torch.from_numpy(...).long().cuda()`` per step, agent_pretrain.py:552-554).
* ``AgentMemory`` / ``ExpertMemory``: the numpy float64 ring buffers of ppo_train.py:69-212 and
  IRL_dqn_train.py:78-204 as preallocated device tensors with the same method and field names; storing a transition is
  a handful of device-side copies (the reference does seven ``.detach().cpu().numpy()`` syncs per transition).
"""
from __future__ import annotations

from typing import Dict, Iterator, Optional, Tuple

import numpy as np
import torch

TYPE_COLUMN = 3            # the 'type' attribute the reference deletes (agent_pretrain.py:525-526)


def load_cp_npz(path, drop_type: bool = True, pin: bool = True) -> Dict[str, torch.Tensor]:
    """-> {'x','y': (songs, L, 6) int64, 'mask': (songs, L) float32} on the host (pinned when possible)."""
    with np.load(path) as z:
        x, y, mask = z["x"], z["y"], z["mask"]
    if x.ndim != 3 or x.shape != y.shape or mask.shape != x.shape[:2]:
        raise ValueError(f"unexpected CP npz shapes x{x.shape} y{y.shape} mask{mask.shape}")
    if drop_type:
        if x.shape[2] != 7:
            raise ValueError(f"drop_type expects the 7-attribute layout, got {x.shape[2]} columns")
        keep = [c for c in range(7) if c != TYPE_COLUMN]
        x, y = x[:, :, keep], y[:, :, keep]
    out = {"x": torch.from_numpy(np.ascontiguousarray(x)).long(), "y": torch.from_numpy(np.ascontiguousarray(y)).long(),
           "mask": torch.from_numpy(np.ascontiguousarray(mask)).float()}
    if pin and torch.cuda.is_available():
        out = {k: v.pin_memory() for k, v in out.items()}
    return out


class CPBatches:
    """Batches ``(x, y, mask)`` of ``batch_size`` songs (optionally truncated to ``seq_len``) on ``device``; the next
    batch's host→device copy runs on a side stream while the caller computes on the current one.  ``lo:hi`` selects this
    rank's shard of the songs (``cpmusic.dist.shard_range``)."""

    def __init__(self, data: Dict[str, torch.Tensor], batch_size: int, device, seq_len: Optional[int] = None,
                 lo: int = 0, hi: Optional[int] = None, shuffle: bool = False, seed: int = 0):
        self.data, self.bs, self.device, self.seq_len = data, int(batch_size), torch.device(device), seq_len
        n = data["x"].shape[0]
        self.lo, self.hi = lo, n if hi is None else hi
        self.shuffle, self.seed, self.epoch = shuffle, seed, 0
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None

    def __len__(self):
        return (self.hi - self.lo) // self.bs

    def _host(self, idx):
        sl = slice(None) if self.seq_len is None else slice(0, self.seq_len)
        return tuple(self.data[k][idx][:, sl] for k in ("x", "y", "mask"))

    def _to_device(self, host):
        if self.stream is None:
            return tuple(t.to(self.device) for t in host), None
        with torch.cuda.stream(self.stream):
            dev = tuple(t.pin_memory().to(self.device, non_blocking=True) if not t.is_pinned() else t.to(self.device, non_blocking=True)
                        for t in host)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return dev, ev

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        order = torch.arange(self.lo, self.hi)
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            order = order[torch.randperm(len(order), generator=g)]
        self.epoch += 1
        batches = [order[i * self.bs:(i + 1) * self.bs] for i in range(len(self))]
        nxt = self._to_device(self._host(batches[0])) if batches else None
        for i in range(len(batches)):
            cur, ev = nxt
            nxt = self._to_device(self._host(batches[i + 1])) if i + 1 < len(batches) else None
            if ev is not None:
                torch.cuda.current_stream(self.device).wait_event(ev)
                for t in cur:
                    t.record_stream(torch.cuda.current_stream(self.device))
            yield cur


class _Memory:
    """Device ring buffer with the reference's field layout: states / next_states (cap, n_states, n_feat) int64, actions
    (cap, n_actions, n_feat) int64, log_actions (cap, n_actions, n_feat) float32, value / rewards / dones (cap, 1)."""
    suffix = "agent"

    def __init__(self, capacity: int, n_states: int = 50, n_actions: int = 25, n_features: int = 6, device="cuda",
                 log_prob_long_compat: bool = True, seed: int = 0):
        dev = torch.device(device)
        z = lambda *s, dt=torch.int64: torch.zeros(*s, dtype=dt, device=dev)
        s = self.suffix
        setattr(self, f"states_{s}", z(capacity, n_states, n_features))
        setattr(self, f"next_states_{s}", z(capacity, n_states, n_features))
        setattr(self, f"actions_{s}", z(capacity, n_actions, n_features))
        setattr(self, f"log_actions_{s}", z(capacity, n_actions, n_features, dt=torch.float32))
        setattr(self, f"value_{s}", z(capacity, 1, dt=torch.float32))
        setattr(self, f"rewards_{s}", z(capacity, 1, dt=torch.float32))
        setattr(self, f"dones_{s}", z(capacity, 1, dt=torch.float32))
        self.capacity, self.device, self.memory_counter = capacity, dev, 0
        # the reference reads log-probs back with .long() (ppo_train.py:135): values truncate to {0,-1,-2,...}
        self.log_prob_long_compat = log_prob_long_compat
        self.gen = torch.Generator(device=dev).manual_seed(seed)

    def _f(self, name):
        return getattr(self, f"{name}_{self.suffix}")

    def store_transition(self, state, action, log_action, value_state, reward, next_state, done):
        i = self.memory_counter % self.capacity
        self._f("states")[i].copy_(torch.as_tensor(state, device=self.device).reshape(self._f("states")[i].shape))
        self._f("actions")[i].copy_(torch.as_tensor(action, device=self.device).reshape(self._f("actions")[i].shape))
        self._f("log_actions")[i].copy_(torch.as_tensor(log_action, device=self.device).reshape(self._f("log_actions")[i].shape))
        self._f("value")[i].copy_(torch.as_tensor(value_state, device=self.device).reshape(1))
        self._f("rewards")[i].copy_(torch.as_tensor(reward, device=self.device).reshape(1))
        self._f("next_states")[i].copy_(torch.as_tensor(next_state, device=self.device).reshape(self._f("next_states")[i].shape))
        self._f("dones")[i].copy_(torch.as_tensor(done, device=self.device).reshape(1))
        self.memory_counter += 1

    def _pack(self, idx):
        lp = self._f("log_actions")[idx]
        if self.log_prob_long_compat:
            lp = lp.long()
        return (self._f("states")[idx], self._f("actions")[idx], lp, self._f("value")[idx], self._f("rewards")[idx],
                self._f("next_states")[idx], self._f("dones")[idx].long())

    def sampling(self, batch_size: int):
        """Uniform over the WHOLE buffer like the reference (``np.random.choice(BUFFER_SIZE, batch_size)``)."""
        idx = torch.randint(0, self.capacity, (batch_size,), device=self.device, generator=self.gen)
        return self._pack(idx)

    def get(self):
        s, a, lp, v, r, ns, d = self._pack(slice(None))
        return {"states": s, "actions": a, "log_actions": lp, "values": v, "rewards": r, "next_states": ns, "dones": d}


class AgentMemory(_Memory):
    suffix = "agent"


class ExpertMemory(_Memory):
    suffix = "expert"
