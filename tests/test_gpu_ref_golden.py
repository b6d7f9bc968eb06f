"""GPU parity against vectors produced by EXECUTING THE REFERENCE'S OWN PYTHON
(tests/golden/make_ref_golden.py; see tests/test_ref_golden.py for what those vectors pin).
The CUDA-backed modules are loaded with the same name-keyed weights the reference modules ran with
and compared in fp32 compute mode (tight) and, for the loss, in the bf16 production mode.
Nothing here reads /root/reference at run time."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import ref_weights  # noqa: E402
from oracle import model_oracle as mo  # noqa: E402

pytestmark = pytest.mark.gpu
VOCAB_DQN = [56, 135, 18, 87, 18, 25]
VOCAB_PPO = [49, 19, 19, 89, 67, 25]
SMALL = dict(d_model=128, n_layer=2, n_head=2, d_inner=2048)
ATTRS = ("tempo", "chord", "barbeat", "pitch", "duration", "velocity")


def T(a):
    return torch.from_numpy(np.asarray(a))


def _cmp(a, b, atol, rtol=0.0, what=""):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).double().cpu()
    err = (a - b).abs()
    assert bool((err <= atol + rtol * b.abs()).all()), f"{what}: max err {err.max().item():.3e}, ref scale {b.abs().max().item():.3e}"


def _weights(vocab, seed, variant="dqn", critic=False):
    """State dict with the values the reference module was given (names are the contract, App. A.3)."""
    o = mo.OracleCritic(vocab, **SMALL) if critic else mo.OracleCPModel(vocab, variant=variant, **SMALL)
    ref_weights.fill_(o, seed)
    return o.state_dict()


def test_teacher_forced_surface_vs_reference_run(cuda, cpm, golden):
    g = golden("ref_model")
    m = cpm.LinearTransformer(VOCAB_DQN, True, compute_dtype=torch.float32, dropout=0.0, **SMALL)
    m.load_state_dict(_weights(VOCAB_DQN, 11))
    m = m.to(cuda).train()
    x, y, mask = (T(g[k]).to(cuda) for k in ("dqn_x", "dqn_y", "dqn_mask"))
    h = m.forward_hidden(x)
    _cmp(h, g["dqn_h"], 5e-4, 2e-4, "hidden")
    for a, lg in zip(ATTRS, m.forward_output(h, y)):
        _cmp(lg, g[f"dqn_logits_{a}"], 1e-3, 2e-4, f"logits {a}")
    losses = torch.stack(m.train_step(x, y, mask))
    _cmp(losses, g["dqn_losses"], 3e-4, 1e-4, "losses")
    (losses.sum() / 6).backward()
    for name, key in (("in_linear.weight", "dqn_grad_in_linear"), ("word_emb_pitch.lut.weight", "dqn_grad_lut_pitch"),
                      ("transformer_encoder.layers.0.attention.query_projection.weight", "dqn_grad_q0"),
                      ("proj_tempo.weight", "dqn_grad_proj_tempo")):
        ref = g[key]
        _cmp(dict(m.named_parameters())[name].grad, ref, 2e-5 + 3e-3 * np.abs(ref).max(), 3e-3, key)
    # the PPO script's int64 mask (ppo_train.py:207,398)
    _cmp(torch.stack(m.train_step(x, y, mask.long())), g["dqn_losses_longmask"], 3e-4, 1e-4, "losses, int64 mask")


def test_bf16_losses_vs_reference_run(cuda, cpm, golden):
    """Production dtype; tolerance as in test_gpu_model.test_train_step_bf16."""
    g = golden("ref_model")
    m = cpm.LinearTransformer(VOCAB_DQN, True, compute_dtype=torch.bfloat16, dropout=0.0, **SMALL)
    m.load_state_dict(_weights(VOCAB_DQN, 11))
    m = m.to(cuda).train()
    x, y, mask = (T(g[k]).to(cuda) for k in ("dqn_x", "dqn_y", "dqn_mask"))
    _cmp(torch.stack(m.train_step(x, y, mask)), g["dqn_losses"], 3e-2, 2e-2, "losses bf16")


def test_recurrent_protocol_vs_reference_run(cuda, cpm, golden):
    g = golden("ref_model")
    m = cpm.LinearTransformer(VOCAB_DQN, False, compute_dtype=torch.float32, dropout=0.0, **SMALL)
    m.load_state_dict(_weights(VOCAB_DQN, 11))
    m = m.to(cuda).eval()
    x = T(g["dqn_x"]).to(cuda)
    mem, hs = None, []
    with torch.no_grad():
        for t in range(g["dqn_rec_h"].shape[0]):
            h, mem = m.forward_hidden(x[:1, t:t + 1], mem, is_training=False)    # testing-no-type-cp.py:157-167
            hs.append(h)
    _cmp(torch.stack(hs), g["dqn_rec_h"], 5e-4, 2e-4, "recurrent hidden")
    _cmp(mem[-1][0], g["dqn_rec_S_last"], 5e-4, 5e-4, "S of the last layer")
    _cmp(mem[-1][1], g["dqn_rec_Z_last"], 5e-4, 5e-4, "Z of the last layer")


def test_actor_critic_and_readouts_vs_reference_run(cuda, cpm, golden):
    g, gr = golden("ref_model"), golden("ref_rl")
    x = T(g["ppo_x"]).to(cuda)
    a = cpm.Actor_Transformer(VOCAB_PPO, compute_dtype=torch.float32, dropout=0.0, **SMALL)
    a.load_state_dict(_weights(VOCAB_PPO, 21, variant="actor"))
    a = a.to(cuda).eval()
    c = cpm.Critic_Transformer(VOCAB_PPO, compute_dtype=torch.float32, dropout=0.0, **SMALL)
    c.load_state_dict(_weights(VOCAB_PPO, 22, critic=True))
    c = c.to(cuda).eval()
    with torch.no_grad():
        h = a.forward_hidden(x)
        _cmp(h, g["ppo_h"], 5e-4, 2e-4, "actor hidden")
        _cmp(a.value_funtion(h), g["ppo_value_funtion"], 5e-4, 2e-4, "value_funtion")
        _cmp(c.value_produce(x), g["ppo_value_produce"], 5e-4, 2e-4, "value_produce")
        act, lp = cpm.rl.ppo_choose_action(a, x[:1])
        assert torch.equal(act.cpu(), T(gr["ppo_choose_action"]))
        _cmp(lp, gr["ppo_choose_logp"], 2e-3, 1e-3, "choose_action log-probs")
        act, lp = cpm.rl.ppo_select_update(a, x)
        assert torch.equal(act.cpu(), T(gr["ppo_select_action"]))
        _cmp(lp, gr["ppo_select_logp"], 2e-3, 1e-3, "select_udpate log-probs")


def test_reward_head_kernel_vs_reference_run(cuda, cpm, golden):
    gr = golden("ref_rl")
    head = cpm.rl.RewardHead(VOCAB_PPO, d_model=64)
    ref_weights.fill_(head, seed=31)
    head = head.to(cuda)
    reward = head(T(gr["rw_hidden"]).to(cuda))
    assert reward.shape == (5,)
    _cmp(reward, gr["rw_ppo_score"].reshape(-1), 2e-5, 1e-5, "reward score")


def test_generation_drivers_on_the_cuda_model(cuda, cpm, golden):
    """``inference_from_scratch`` (testing-no-type-cp.py:126-179) and its batched device-resident form on the CUDA model:
    the device sampler draws from Philox, not numpy, so the words differ from the reference run; the protocol does not —
    priming bar token, value ranges, and the stop at exactly ``bar_cond`` bars."""
    _, w2e = ref_weights.synthetic_dictionary()
    m = cpm.LinearTransformer(VOCAB_DQN, False, compute_dtype=torch.float32, dropout=0.0, **SMALL)
    m.load_state_dict(_weights(VOCAB_DQN, 11))
    m = m.to(cuda).eval()

    def bars(words):
        return 1 + sum(1 for w in words[1:] if w2e["bar-beat"][int(w[2])] == "Bar")

    words = cpm.midi.inference_from_scratch(m, w2e, 3, max_tokens=400)
    assert words.shape[1] == 6 and tuple(words[0]) == cpm.midi.BAR_TOKEN
    assert all(0 <= int(words[:, a].max()) < VOCAB_DQN[a] and int(words[:, a].min()) >= 0 for a in range(6))
    assert len(words) == 400 or (bars(words) == 3 and bars(words[:-1]) == 2)
    songs = cpm.midi.batched_generate(m, w2e, 3, n_songs=4, max_tokens=256, seed=7)
    assert len(songs) == 4
    for s in songs:
        assert tuple(s[0]) == cpm.midi.BAR_TOKEN and (len(s) == 256 or (bars(s) == 3 and bars(s[:-1]) == 2))


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 3e-3), (torch.bfloat16, 8e-2)])
def test_pretraining_loss_curve_vs_reference_train_loop(cuda, cpm, golden, tmp_path, dtype, tol):
    """Reference-matching loss curves (north_star): the first 10 'batch loss' values logged by the REFERENCE's own
    ``train()`` (agent_pretrain.py:485-565 run on files in its formats, dropout off; ``pre_losses_eval``) against the CUDA
    model fed by ``CPBatches`` from the same npz and stepped by the same loop (mean of six losses, clip 3, Adam 1e-4).
    Tolerances as test_gpu_model.test_pretraining_loss_curve_matches_oracle, slightly widened for the larger weights."""
    ref = golden("ref_rl")["pre_losses_eval"]
    np.savez(tmp_path / "train_data_linear.npz", **ref_weights.pretrain_corpus())
    d = cpm.data.load_cp_npz(tmp_path / "train_data_linear.npz")
    m = cpm.TransformerModel(VOCAB_DQN, True, compute_dtype=dtype, dropout=0.0, **SMALL)
    m.load_state_dict(_weights(VOCAB_DQN, 13))
    m = m.to(cuda).train()
    opt = torch.optim.Adam(m.parameters(), lr=0.0001, fused=True)
    curve = []
    while len(curve) < len(ref):
        for x, y, mask in cpm.data.CPBatches(d, batch_size=4, device=cuda):
            loss = sum(m.train_step(x, y, mask)) / 6
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(m.parameters(), 3.0)
            opt.step()
            curve.append(loss.item())
    curve = curve[:len(ref)]
    _cmp(torch.tensor(curve), ref, tol, 0.0, f"loss curve {curve} vs {ref.tolist()}")
    assert curve[-1] < curve[0] - 0.3


def test_dqn_td_kernel_vs_reference_dqn_update(cuda, cpm, golden):
    """The fused TD kernel on the logits the reference's ``DQN.update`` was fed (stub nets, seed-rebuilt), model layout
    (six segments concatenated, padded to a multiple of 8 columns): MSE and dQ against the executed reference."""
    gr = golden("ref_rl")
    q, nx, action, reward, done = ref_weights.dqn_td_inputs(int(gr["dqnrl_seed"]), VOCAB_DQN)
    seg = [0]
    for n in VOCAB_DQN:
        seg.append(seg[-1] + n)
    pad = (-seg[-1]) % 8
    cat = lambda parts: torch.nn.functional.pad(torch.cat([p.detach() for p in parts], -1), (0, pad)).to(cuda)      # noqa: E731
    ql = cat(q).requires_grad_()
    loss, _ = cpm.ops.dqn_td_loss(ql, cat(nx), action.to(cuda), reward.to(cuda), done.to(cuda), seg, 25, 0.95, True)
    (0.3 * loss).backward()                                                     # IRL_dqn_train.py:335: alpha = 0.3
    _cmp(loss, float(gr["dqnrl_mse"]), 1e-4, 1e-5, "TD MSE")
    _cmp(ql.grad[..., seg[3]:seg[4]], gr["dqnrl_grad_q_pitch"], 1e-8, 1e-4, "dQ (pitch segment)")


def test_dqn_update_loop_vs_reference_run(cuda, cpm, golden):
    """Four whole ``DQN.update`` calls (IRL_dqn_train.py:267-345, executed from the reference with real networks) against the
    CUDA path: ``rl.dqn_td_loss`` + ``train_step`` + Adam 0.01.  Update 0 is tight; later ones carry three Adam steps of
    lr 0.01 of fp32 divergence."""
    ref = golden("ref_rl")["loop_dqn_mse_ce_total"]
    ev = cpm.LinearTransformer(VOCAB_DQN, True, compute_dtype=torch.float32, dropout=0.0, **SMALL)
    tg = cpm.LinearTransformer(VOCAB_DQN, True, compute_dtype=torch.float32, dropout=0.0, **SMALL)
    ev.load_state_dict(_weights(VOCAB_DQN, 14))
    tg.load_state_dict(_weights(VOCAB_DQN, 15))
    ev, tg = ev.to(cuda).train(), tg.to(cuda).train()
    opt = torch.optim.Adam(ev.parameters(), lr=0.01)
    rows = []
    for u, b in enumerate(ref_weights.rl_update_batches(4, VOCAB_DQN, seed=95)):
        b = {k: v.to(cuda) for k, v in b.items()}
        if u % 50 == 0:
            tg.load_state_dict(ev.state_dict())
        mse = cpm.rl.dqn_td_loss(ev, tg, b["state"], b["nextstate"], b["action"], b["reward"], b["done"], gamma=0.95)
        ce = sum(ev.train_step(b["state"], b["nextstate"], b["mask"])) / 6
        total = 0.3 * mse + 0.7 * ce
        opt.zero_grad()
        total.backward()
        opt.step()
        rows.append([mse.item(), ce.item(), total.item()])
    _cmp(torch.tensor(rows[0]), ref[0], 2e-3, 2e-4, f"update 0: {rows[0]} vs {ref[0].tolist()}")
    _cmp(torch.tensor(rows), ref, 5e-2, 3e-2, f"DQN loss curve {rows} vs {ref.tolist()}")


def test_ppo_update_loop_vs_reference_run(cuda, cpm, golden):
    """Three ``PPO.update_policy`` epochs (ppo_train.py:365-416, executed from the reference with real networks and its own
    buffers) against the CUDA path: device-resident AgentMemory / ExpertMemory, ``rl.ppo_select_update``, the compat
    surrogate kernel, ``train_step`` with the int64 mask, ``value_produce``, two Adam optimizers at 0.01."""
    gr = golden("ref_rl")
    actor = cpm.Actor_Transformer(VOCAB_PPO, compute_dtype=torch.float32, dropout=0.0, **SMALL)
    critic = cpm.Critic_Transformer(VOCAB_PPO, compute_dtype=torch.float32, dropout=0.0, **SMALL)
    actor.load_state_dict(_weights(VOCAB_PPO, 16, variant="actor"))
    critic.load_state_dict(_weights(VOCAB_PPO, 17, critic=True))
    actor, critic = actor.to(cuda).train(), critic.to(cuda).train()
    abuf, ebuf = cpm.data.AgentMemory(30, device=cuda), cpm.data.ExpertMemory(30, device=cuda)
    ref_weights.fill_ppo_buffers(abuf, ebuf, ref_weights.rl_update_batches(1, VOCAB_PPO, seed=96)[0])
    agent_all, expert_all = abuf.get(), ebuf.get()
    returns = cpm.rl.calculate_returns_compat(agent_all["rewards"], 0.99)
    adv = cpm.rl.calculate_advantages_compat(returns, agent_all["values"])
    _cmp(returns, gr["loop_ppo_returns"], 1e-4, 1e-4, "returns")
    _cmp(adv, gr["loop_ppo_adv"], 1e-4, 1e-4, "advantages")
    a_opt, c_opt = torch.optim.Adam(actor.parameters(), lr=0.01), torch.optim.Adam(critic.parameters(), lr=0.01)
    actor_losses, value_losses = [], []
    for _ in range(3):
        states = agent_all["states"]
        _, new_logp = cpm.rl.ppo_select_update(actor, states)
        value_pred = critic.value_produce(states)
        policy_loss = cpm.rl.ppo_policy_loss_compat(new_logp, agent_all["log_actions"], adv)
        ce = sum(actor.train_step(states, expert_all["states"], expert_all["mask_state"])) / 6
        actor_loss = policy_loss + ce
        value_loss = cpm.rl.value_loss_compat(returns.detach(), value_pred)
        a_opt.zero_grad()
        actor_loss.backward()
        a_opt.step()
        c_opt.zero_grad()
        value_loss.backward()
        c_opt.step()
        actor_losses.append(actor_loss.item())
        value_losses.append(value_loss.item())
    _cmp(torch.tensor(actor_losses[:1]), gr["loop_ppo_actor_loss"][:1], 2e-3, 2e-4, "epoch 0 actor loss")
    _cmp(torch.tensor(value_losses[:1]), gr["loop_ppo_value_loss"][:1], 2e-3, 2e-4, "epoch 0 value loss")
    _cmp(torch.tensor(actor_losses), gr["loop_ppo_actor_loss"], 5e-2, 3e-2, f"actor losses {actor_losses}")
    _cmp(torch.tensor(value_losses), gr["loop_ppo_value_loss"], 5e-2, 8e-2, f"value losses {value_losses}")


def test_policy_classes_vs_reference_run(cuda, cpm, golden):
    """``cpmusic.rl.PPO`` / ``cpmusic.rl.DQN`` (the scripts' classes on the fused kernels, same method names) against the losses
    the reference's own classes produced: three ``update_policy`` epochs, four ``update`` calls."""
    gr = golden("ref_rl")
    actor = cpm.Actor_Transformer(VOCAB_PPO, compute_dtype=torch.float32, dropout=0.0, **SMALL)
    critic = cpm.Critic_Transformer(VOCAB_PPO, compute_dtype=torch.float32, dropout=0.0, **SMALL)
    actor.load_state_dict(_weights(VOCAB_PPO, 16, variant="actor"))
    critic.load_state_dict(_weights(VOCAB_PPO, 17, critic=True))
    abuf, ebuf = cpm.data.AgentMemory(30, device=cuda), cpm.data.ExpertMemory(30, device=cuda)
    ref_weights.fill_ppo_buffers(abuf, ebuf, ref_weights.rl_update_batches(1, VOCAB_PPO, seed=96)[0])
    ppo = cpm.rl.PPO(actor.to(cuda).train(), critic.to(cuda).train(), abuf, ebuf, lr=0.01)
    returns = ppo.calculate_returns(abuf.get()["rewards"], 0.99)
    adv = ppo.calculate_advantages(returns, abuf.get()["values"])
    actor_losses, value_losses = [], []
    for _ in range(3):
        actor_losses.append(ppo.update_policy(1, 0.2, adv, returns))
        value_losses.append(ppo.last_value_loss.item())
    _cmp(torch.tensor(actor_losses), gr["loop_ppo_actor_loss"], 5e-2, 3e-2, f"actor losses {actor_losses}")
    _cmp(torch.tensor(value_losses), gr["loop_ppo_value_loss"], 5e-2, 8e-2, f"value losses {value_losses}")
    act, logp = ppo.choose_action(abuf.get()["states"][:1])
    assert act.shape == (25, 6) and logp.shape == (25, 6)

    ev = cpm.LinearTransformer(VOCAB_DQN, True, compute_dtype=torch.float32, dropout=0.0, **SMALL)
    tg = cpm.LinearTransformer(VOCAB_DQN, True, compute_dtype=torch.float32, dropout=0.0, **SMALL)
    ev.load_state_dict(_weights(VOCAB_DQN, 14))
    tg.load_state_dict(_weights(VOCAB_DQN, 15))
    dqn = cpm.rl.DQN(ev.to(cuda).train(), tg.to(cuda).train(), lr=0.01)
    rows = []
    for b in ref_weights.rl_update_batches(4, VOCAB_DQN, seed=95):
        tr = {"state": b["state"], "nextstate": b["nextstate"], "action": b["action"], "reward": b["reward"], "done": b["done"]}
        rows.append([t.item() for t in dqn.update(tr, {"state": b["state"], "nextstate": b["nextstate"]}, b["mask"])])
    _cmp(torch.tensor(rows), gr["loop_dqn_mse_ce_total"], 5e-2, 3e-2, f"DQN loss curve {rows}")
    _cmp(dqn.total_val, gr["loop_dqn_mse_ce_total"][:, 2].sum(), 1e-1, 3e-2, "running total like the script's total_val")
    assert dqn.cnt_update == 4 and dqn.choose_action(b["state"][:1].to(cuda)).shape == (25, 6)
