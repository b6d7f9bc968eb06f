"""Deterministic, name-keyed parameter values shared by ``make_ref_golden.py`` (which loads them
into the REAL reference modules) and ``tests/test_ref_golden.py`` (which loads them into the
oracle / the CUDA-backed modules).  Independent of module construction order and of torch's
default-init RNG consumption, so the same state dict can be rebuilt wherever the key names and
shapes agree."""
import zlib

import torch


def tensor_for(name: str, shape, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed((seed * 1_000_003 + zlib.crc32(name.encode())) % (2 ** 31))
    shape = tuple(shape)
    if name.endswith("lut.weight"):                       # embeddings
        return torch.randn(shape, generator=g) * 0.5
    if "norm" in name and name.endswith("weight"):        # LayerNorm gains
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if len(shape) >= 2:                                   # Linear weights (out, in)
        return torch.randn(shape, generator=g) / (shape[-1] ** 0.5)
    return 0.1 * torch.randn(shape, generator=g)          # biases


def fill_(module: torch.nn.Module, seed: int) -> None:
    """Overwrites every PARAMETER of ``module`` in place (buffers such as ``pos_emb.pe`` keep the
    value the module computed for itself)."""
    with torch.no_grad():
        for name, p in module.named_parameters():
            p.copy_(tensor_for(name, p.shape, seed).to(p.dtype))


def dqn_td_inputs(seed: int, vocab, B: int = 30, L: int = 50, n_actions: int = 25):
    """Inputs of the DQN TD golden (too large to commit: 2 x B x L x sum(vocab) floats), rebuilt from
    the seed by the generator script and by the tests."""
    g = torch.Generator().manual_seed(seed)
    q = [torch.randn(B, L, n, generator=g).requires_grad_() for n in vocab]
    nx = [torch.randn(B, L, n, generator=g).requires_grad_() for n in vocab]
    action = torch.stack([torch.randint(0, n, (B, n_actions), generator=g) for n in vocab], -1)
    reward = torch.rand(B, 1, generator=g)
    done = (torch.rand(B, 1, generator=g) < 0.2).long()
    return q, nx, action, reward, done


def synthetic_dictionary():
    """An AIlabs-Pop1K7-shaped ``word2event`` without the ``type`` class (sizes 56,135,18,87,18,25 as in
    IRL_dqn_train.py:403): event 0 is the padding / ignore event, 'CONTI' continues tempo / chord, bar-beat 1 is 'Bar'."""
    roots = ["C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "A", "A#", "B"]
    quals = ["M", "m", "o", "+", "MM7", "Mm7", "mM7", "mm7", "o7", "%7", "+7"]
    chords = [f"{r}_{q}" for r in roots for q in quals] + ["N_N"]
    w2e = {
        "tempo": {0: 0, 1: "CONTI", **{i + 2: f"Tempo_{32 + 3 * i}" for i in range(54)}},
        "chord": {0: 0, 1: "CONTI", **{i + 2: chords[i] for i in range(133)}},
        "bar-beat": {0: 0, 1: "Bar", **{i + 2: f"Beat_{i}" for i in range(16)}},
        "pitch": {0: 0, **{i + 1: f"Note_Pitch_{22 + i}" for i in range(86)}},
        "duration": {0: 0, **{i + 1: f"Note_Duration_{120 * i}" for i in range(17)}},
        "velocity": {0: 0, **{i + 1: f"Note_Velocity_{40 + 3 * i}" for i in range(24)}},
    }
    e2w = {k: {e: w for w, e in v.items()} for k, v in w2e.items()}
    return e2w, w2e


def pretrain_corpus(n_songs: int = 8, L: int = 48, seed: int = 61):
    """A tiny ``train_data_linear.npz``-shaped corpus: x, y (n_songs, L, 7) with the ``type`` class at column 3
    (dropped by the training script, agent_pretrain.py:525-526), mask (n_songs, L) with ragged lengths."""
    import numpy as np
    g = torch.Generator().manual_seed(seed)
    sizes = [56, 135, 18, 3, 87, 18, 25]
    x = torch.stack([torch.randint(0, n, (n_songs, L + 1), generator=g) for n in sizes], -1)
    lens = torch.randint(L // 2, L + 1, (n_songs,), generator=g)
    mask = (torch.arange(L)[None, :] < lens[:, None]).float()
    return dict(x=x[:, :-1].numpy(), y=x[:, 1:].numpy(), mask=mask.numpy().astype(np.float32))


def transition_stream(n: int, n_states: int = 50, n_actions: int = 25, n_features: int = 6, seed: int = 81):
    """n synthetic transitions as the RL scripts produce them (ppo_train.py:455-480, IRL_dqn_train.py:455-480): int64 token
    windows, float log-probs, scalar-shaped value / reward / done tensors, float loss masks."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for t in range(n):
        out.append(dict(
            state=torch.randint(0, 18, (1, n_states, n_features), generator=g),
            action=torch.randint(0, 18, (n_actions, n_features), generator=g),
            log_action=-3.0 * torch.rand(n_actions, n_features, generator=g),
            value=torch.randn(1, 1, generator=g),
            reward=torch.rand(1, generator=g),
            next_state=torch.randint(0, 18, (1, n_states, n_features), generator=g),
            done=torch.tensor([float(t % 7 == 6)]),
            mask_state=(torch.rand(n_states, generator=g) > 0.2).float(),
            mask_next_state=(torch.rand(n_states, generator=g) > 0.2).float()))
    return out


def rl_update_batches(n_updates: int, vocab, seed: int, B: int = 30, L: int = 50, n_actions: int = 25):
    """Replay batches as ``DQN.update`` / ``PPO.update_policy`` receive them: token windows, greedy-action-shaped indices,
    rewards in (0,1), sparse dones, ragged loss masks."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_updates):
        tok = lambda rows: torch.stack([torch.randint(0, n, (B, rows), generator=g) for n in vocab], -1)      # noqa: E731
        lens = torch.randint(L // 2, L + 1, (B,), generator=g)
        out.append(dict(state=tok(L), nextstate=tok(L), action=tok(n_actions), reward=torch.rand(B, 1, generator=g),
                        done=(torch.rand(B, 1, generator=g) < 0.15).long(),
                        mask=(torch.arange(L)[None, :] < lens[:, None]).float()))
    return out


def fill_ppo_buffers(abuf, ebuf, b, seed: int = 97):
    """Stores the 30 transitions of replay batch ``b`` through ``store_transition`` — the reference's buffers and the
    device-resident ones take the same arguments in the same order."""
    g = torch.Generator().manual_seed(seed)
    for i in range(b["state"].shape[0]):
        logp = -3.0 * torch.rand(25, 6, generator=g)
        abuf.store_transition(b["state"][i], b["action"][i], logp, torch.randn(1, generator=g), b["reward"][i], b["nextstate"][i],
                              b["done"][i].float())
        ebuf.store_transition(b["nextstate"][i], b["action"][i], b["reward"][i], b["state"][i], b["done"][i].float(), b["mask"][i],
                              b["mask"][i])
