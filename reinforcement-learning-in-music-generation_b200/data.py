"""CP data formats and device-resident experience buffers.

* ``load_cp_npz`` / ``CPBatches``: the reference's ``train_data_linear.npz`` (``x``, ``y`` ``(songs, L, 7)`` int with
  the ``type`` attribute in column 3, ``mask`` ``(songs, L)``; agent_pretrain.py:491-526, IRL_dqn_train.py:418-434) →
  6-attribute int64 tensors in pinned host memory, and an iterator that keeps one batch in flight to the GPU on a side
  stream (the reference does a blocking ``torch.from_numpy(...).long().cuda()`` per step, agent_pretrain.py:552-554).
* ``AgentMemory`` / ``ExpertMemory`` (ppo_train.py:69-212) and ``DQNAgentMemory`` / ``DQNExpertMemory``
  (IRL_dqn_train.py:78-204, where the two classes carry the same names with fewer fields and tuple returns): the numpy
  float64 ring buffers as preallocated device tensors with the reference's attribute names, ``store_transition``
  argument orders and ``sampling`` / ``get`` return layouts; storing a transition is a handful of device-side copies
  (the reference does up to seven ``.detach().cpu().numpy()`` syncs per transition).
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


class _Ring:
    """Preallocated device ring buffer.  ``FIELDS`` lists (reference attribute name, per-slot shape key, dtype) in the
    order ``store_transition`` takes them; shape keys: 's' (n_states, n_features), 'a' (n_actions, n_features), '1' (1,),
    'm' (n_states,).  Integer-valued fields are int64 and real-valued ones float32 on the device (the reference keeps
    float64 numpy arrays and converts on every read)."""
    FIELDS = ()

    def __init__(self, capacity: int, n_states: int = 50, n_actions: int = 25, n_features: int = 6, device="cuda", seed: int = 0):
        dev = torch.device(device)
        shapes = {"s": (n_states, n_features), "a": (n_actions, n_features), "1": (1,), "m": (n_states,)}
        for name, key, dt in self.FIELDS:
            setattr(self, name, torch.zeros((capacity,) + shapes[key], dtype=dt, device=dev))
        self.capacity, self.device, self.memory_counter = capacity, dev, 0
        self.gen = torch.Generator(device=dev).manual_seed(seed)

    def store_transition(self, *values):
        if len(values) != len(self.FIELDS):
            raise TypeError(f"store_transition takes {len(self.FIELDS)} values ({', '.join(f[0] for f in self.FIELDS)})")
        i = self.memory_counter % self.capacity
        for (name, _, _), v in zip(self.FIELDS, values):
            slot = getattr(self, name)[i]
            slot.copy_(torch.as_tensor(v, device=self.device).reshape(slot.shape))
        self.memory_counter += 1

    def _indices(self, batch_size, idx):
        """Uniform WITH replacement over the WHOLE buffer, filled or not, like the reference
        (``np.random.choice(BUFFER_SIZE, batch_size)``); ``idx`` injects the draw (tests, replays)."""
        if idx is not None:
            return torch.as_tensor(idx, device=self.device, dtype=torch.int64)
        return torch.randint(0, self.capacity, (batch_size,), device=self.device, generator=self.gen)


_I, _F = torch.int64, torch.float32


class AgentMemory(_Ring):
    """PPO trajectory buffer, ppo_train.py:69-144.  ``sampling`` -> (states, actions, log_actions, values, rewards,
    next_states, dones); ``get`` -> dict with the reference's keys.  The reference reads the stored log-probs back with
    ``.long()`` (ppo_train.py:118,135: values truncate toward zero); ``log_prob_long_compat=False`` returns them as stored."""
    FIELDS = (("states_agent", "s", _I), ("actions_agent", "a", _I), ("log_actions_agent", "a", _F), ("value_agent", "1", _F),
              ("rewards_agent", "1", _F), ("next_states_agent", "s", _I), ("dones_agent", "1", _I))

    def __init__(self, capacity: int, *a, log_prob_long_compat: bool = True, **kw):
        super().__init__(capacity, *a, **kw)
        self.log_prob_long_compat = log_prob_long_compat

    def _pack(self, idx):
        lp = self.log_actions_agent[idx]
        return (self.states_agent[idx], self.actions_agent[idx], lp.long() if self.log_prob_long_compat else lp, self.value_agent[idx],
                self.rewards_agent[idx], self.next_states_agent[idx], self.dones_agent[idx])

    def sampling(self, batch_size: int, idx=None):
        return self._pack(self._indices(batch_size, idx))

    def get(self):
        s, a, lp, v, r, ns, d = self._pack(slice(None))
        return {"states": s, "actions": a, "log_actions": lp, "values": v, "rewards": r, "next_states": ns, "dones": d}


class ExpertMemory(_Ring):
    """PPO expert buffer, ppo_train.py:147-212: no log-probs / values, two loss masks per transition.  ``sampling`` ->
    (states, actions, rewards, next_states, dones, mask_state, mask_next_state) with float masks; ``get`` -> dict without
    ``dones`` and with int64 masks (what ``train_step`` is then handed, ppo_train.py:207,398)."""
    FIELDS = (("states_exp", "s", _I), ("actions_exp", "a", _I), ("rewards_exp", "1", _F), ("next_states_exp", "s", _I),
              ("dones_exp", "1", _I), ("mask_state", "m", _F), ("mask_next_state", "m", _F))

    def sampling(self, batch_size: int, idx=None):
        i = self._indices(batch_size, idx)
        return (self.states_exp[i], self.actions_exp[i], self.rewards_exp[i], self.next_states_exp[i], self.dones_exp[i],
                self.mask_state[i], self.mask_next_state[i])

    def get(self):
        return {"states": self.states_exp, "actions": self.actions_exp, "rewards": self.rewards_exp, "next_states": self.next_states_exp,
                "mask_state": self.mask_state.long(), "mask_next_state": self.mask_next_state.long()}


class DQNAgentMemory(_Ring):
    """DQN replay buffer, IRL_dqn_train.py:78-134 (the script's ``AgentMemory``): five fields, ``sampling`` and ``get``
    both return (states, actions, rewards, next_states, dones) tuples."""
    FIELDS = (("states_agent", "s", _I), ("actions_agent", "a", _I), ("rewards_agent", "1", _F), ("next_states_agent", "s", _I),
              ("dones_agent", "1", _I))

    def _pack(self, idx):
        return (self.states_agent[idx], self.actions_agent[idx], self.rewards_agent[idx], self.next_states_agent[idx], self.dones_agent[idx])

    def sampling(self, batch_size: int, idx=None):
        return self._pack(self._indices(batch_size, idx))

    def get(self):
        return self._pack(slice(None))


class DQNExpertMemory(ExpertMemory):
    """DQN expert buffer, IRL_dqn_train.py:136-204 (the script's ``ExpertMemory``): as the PPO one, but ``get`` returns the
    7-tuple (states, actions, rewards, next_states, dones, mask_state, mask_next_state) with int64 masks."""

    def get(self):
        return (self.states_exp, self.actions_exp, self.rewards_exp, self.next_states_exp, self.dones_exp, self.mask_state.long(),
                self.mask_next_state.long())
