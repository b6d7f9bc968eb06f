"""The oracle against vectors produced by EXECUTING THE REFERENCE'S OWN PYTHON
(tests/golden/make_ref_golden.py → ref_model.npz / ref_rl.npz; reference files imported unmodified
from /root/reference in the build container, fast_transformers supplied by oracle/ft_oracle.py).

This pins everything the reference itself wrote around the encoder — embeddings and their
sqrt(d) scale, positional encoding, concat + in_linear, the teacher-forced and the recurrent call
protocol, the 6 heads, the masked CE (float and int64 masks), the numpy sampling functions under
the global RNG, Critic_Transformer.value_produce, Actor value_funtion, PPO.choose_action /
select_udpate / calculate_returns / calculate_advantages / update_policy and DQN.choose_action /
DQN.update.  The encoder internals (fast_transformers 0.4.0, absent) stay "restated, unpinned".
Nothing here reads /root/reference at run time."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import ref_weights  # noqa: E402
from oracle import model_oracle as mo, rl_oracle as rl, sampling_oracle as so  # noqa: E402

VOCAB_DQN = [56, 135, 18, 87, 18, 25]
VOCAB_PPO = [49, 19, 19, 89, 67, 25]
SMALL = dict(d_model=128, n_layer=2, n_head=2, d_inner=2048)
ATTRS = ("tempo", "chord", "barbeat", "pitch", "duration", "velocity")
TOL = dict(rtol=2e-5, atol=2e-5)


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(scope="module")
def gm(golden):
    return golden("ref_model")


@pytest.fixture(scope="module")
def gr(golden):
    return golden("ref_rl")


@pytest.fixture(scope="module")
def dqn_small():
    m = mo.OracleCPModel(VOCAB_DQN, is_training=True, variant="dqn", dropout=0.1, **SMALL).eval()
    ref_weights.fill_(m, seed=11)
    return m


def test_teacher_forced_surface_matches_reference_run(gm, dqn_small):
    m = dqn_small
    assert sorted(m.state_dict().keys()) == list(gm["dqn_state_keys"])
    x, y, mask = T(gm["dqn_x"]), T(gm["dqn_y"]), T(gm["dqn_mask"])
    m.zero_grad()
    h = m.forward_hidden(x)
    torch.testing.assert_close(h, T(gm["dqn_h"]), **TOL)
    for a, lg in zip(ATTRS, m.forward_output(h, y)):
        torch.testing.assert_close(lg, T(gm[f"dqn_logits_{a}"]), **TOL)
    losses = torch.stack(m.train_step(x, y, mask))
    torch.testing.assert_close(losses, T(gm["dqn_losses"]), **TOL)
    (losses.sum() / 6).backward()
    torch.testing.assert_close(m.in_linear.weight.grad, T(gm["dqn_grad_in_linear"]), **TOL)
    torch.testing.assert_close(m.word_emb_pitch.lut.weight.grad, T(gm["dqn_grad_lut_pitch"]), **TOL)
    torch.testing.assert_close(m.transformer_encoder.layers[0].attention.query_projection.weight.grad,
                               T(gm["dqn_grad_q0"]), **TOL)
    torch.testing.assert_close(m.proj_tempo.weight.grad, T(gm["dqn_grad_proj_tempo"]), **TOL)
    # ppo_train.py:207,398 passes an int64 mask
    torch.testing.assert_close(torch.stack(m.train_step(x, y, mask.long())), T(gm["dqn_losses_longmask"]), **TOL)


def test_recurrent_protocol_and_sampled_words_match_reference_run(gm, dqn_small):
    r = mo.OracleCPModel(VOCAB_DQN, is_training=False, variant="dqn", **SMALL).eval()
    r.load_state_dict(dqn_small.state_dict())
    x = T(gm["dqn_x"])
    memory, hs, words = None, [], []
    np.random.seed(77)
    with torch.no_grad():
        for t in range(gm["dqn_rec_h"].shape[0]):
            h, memory = r.forward_hidden(x[:1, t:t + 1], memory, is_training=False)
            hs.append(h)
            logits = {a: lg.squeeze().numpy() for a, lg in zip(ATTRS, r.forward_output(h))}
            words.append(so.forward_output_sampling(logits))
    torch.testing.assert_close(torch.stack(hs), T(gm["dqn_rec_h"]), **TOL)
    torch.testing.assert_close(memory[-1][0], T(gm["dqn_rec_S_last"]), **TOL)
    torch.testing.assert_close(memory[-1][1], T(gm["dqn_rec_Z_last"]), **TOL)
    assert np.array_equal(np.stack(words), gm["dqn_rec_words"])       # bit-exact under the same RNG


def test_full_geometry_matches_reference_run(gm):
    m = mo.OracleCPModel(VOCAB_DQN, is_training=True, variant="dqn").eval()     # 12 x 512 x 8, ff 2048
    ref_weights.fill_(m, seed=12)
    assert len(m.state_dict()) == int(gm["dqn_full_n_keys"]) == 217
    x = T(gm["dqn_full_x"])
    with torch.no_grad():
        torch.testing.assert_close(m.forward_hidden(x), T(gm["dqn_full_h"]), rtol=1e-4, atol=1e-4)
        losses = torch.stack(m.train_step(x, x.roll(-1, 1), torch.ones(2, 24)))
    torch.testing.assert_close(losses, T(gm["dqn_full_losses"]), rtol=1e-4, atol=1e-4)


def test_sampling_functions_match_reference_run(gr):
    logits, ps, ts = gr["samp_logits"], gr["samp_p"], gr["samp_t"]
    np.random.seed(123)
    got = []
    for i, lg in enumerate(logits):
        p, t = ps[i % len(ps)], ts[i % len(ts)]
        got.append(so.sampling(lg[None], p=None if p < 0 else float(p), t=float(t)))
    assert np.array_equal(np.array(got), gr["samp_words"])
    np.testing.assert_allclose(so.softmax_with_temperature(logits[0], 1.3), gr["samp_softmax_t13"], rtol=1e-6)


def test_actor_and_critic_match_reference_run(gm):
    actor = mo.OracleCPModel(VOCAB_PPO, is_training=True, variant="actor", **SMALL).eval()
    critic = mo.OracleCritic(VOCAB_PPO, **SMALL).eval()
    ref_weights.fill_(actor, seed=21)
    ref_weights.fill_(critic, seed=22)
    assert sorted(actor.state_dict().keys()) == list(gm["ppo_actor_keys"])
    assert sorted(critic.state_dict().keys()) == list(gm["ppo_critic_keys"])
    x = T(gm["ppo_x"])
    with torch.no_grad():
        h = actor.forward_hidden(x)
        torch.testing.assert_close(h, T(gm["ppo_h"]), **TOL)
        for a, lg in zip(ATTRS, actor.forward_output(h)):
            torch.testing.assert_close(lg, T(gm[f"ppo_logits_{a}"]), **TOL)
        torch.testing.assert_close(actor.value_funtion(h), T(gm["ppo_value_funtion"]), **TOL)
        torch.testing.assert_close(critic.value_produce(x), T(gm["ppo_value_produce"]), **TOL)


def test_ppo_class_arithmetic_matches_reference_run(gm, gr):
    logits = [T(gm[f"ppo_logits_{a}"]) for a in ATTRS]
    act, lp = rl.ppo_choose_action_compat([lg[:1] for lg in logits])
    assert torch.equal(act, T(gr["ppo_choose_action"]))
    torch.testing.assert_close(lp, T(gr["ppo_choose_logp"]), rtol=1e-5, atol=1e-5)
    act, lp = rl.ppo_select_update_compat(logits)
    assert torch.equal(act, T(gr["ppo_select_action"]))
    torch.testing.assert_close(lp, T(gr["ppo_select_logp"]), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(T(gr["ppo_select_value"]), T(gm["ppo_value_produce"]))
    rewards, values = T(gr["ppo_rewards"]), T(gr["ppo_values"])
    ret = rl.calculate_returns_compat(rewards, 0.99)
    torch.testing.assert_close(ret, T(gr["ppo_returns"]), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(rl.calculate_returns_compat(rewards, 0.99, normalize=False), T(gr["ppo_returns_raw"]),
                               rtol=1e-6, atol=1e-6)
    adv = rl.calculate_advantages_compat(ret, values)
    torch.testing.assert_close(adv, T(gr["ppo_advantages"]), rtol=1e-5, atol=1e-5)
    # update_policy: actor_loss = policy_loss + mean of the 6 CE terms (stubbed to 0.25 each)
    new_logp = T(gr["ppo_new_logp"]).clone().requires_grad_()
    vpred = T(gr["ppo_vpred"]).clone().requires_grad_()
    pl = rl.ppo_policy_loss_compat(new_logp, T(gr["ppo_old_logp_long"]), T(gr["ppo_advantages"]))
    np.testing.assert_allclose(float(pl.detach()) + float(gr["ppo_ce_stub"]), float(gr["ppo_actor_loss"]), rtol=1e-5)
    pl.backward()
    rl.value_loss_compat(T(gr["ppo_returns"]), vpred).backward()
    torch.testing.assert_close(new_logp.grad, T(gr["ppo_new_logp_grad"]), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(vpred.grad, T(gr["ppo_vpred_grad"]), rtol=1e-5, atol=1e-7)


def test_dqn_class_arithmetic_matches_reference_run(gr):
    q, nx, action, reward, done = ref_weights.dqn_td_inputs(int(gr["dqnrl_seed"]), VOCAB_DQN)
    assert torch.equal(action, T(gr["dqnrl_action"])) and torch.equal(done, T(gr["dqnrl_done"]))
    torch.testing.assert_close(reward, T(gr["dqnrl_reward"]))
    act = rl.dqn_choose_action_compat([lg[:1] for lg in q])
    assert torch.equal(act, T(gr["dqnrl_choose_action"]))
    mse = rl.dqn_td_loss_compat(q, nx, action, reward, done, gamma=0.95)
    np.testing.assert_allclose(float(mse.detach()), float(gr["dqnrl_mse"]), rtol=1e-5)
    total = 0.3 * mse + 0.7 * float(gr["dqnrl_ce"])                    # IRL_dqn_train.py:335-336
    np.testing.assert_allclose(float(total.detach()), float(gr["dqnrl_total"]), rtol=1e-5)
    (0.3 * mse).backward()
    torch.testing.assert_close(q[3].grad, T(gr["dqnrl_grad_q_pitch"]), rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(nx[3].grad, T(gr["dqnrl_grad_next_pitch"]), rtol=1e-5, atol=1e-8)


def test_reward_read_outs_match_reference_run(gr, cpm):
    """PPO reward model ``token_forward`` and the AIRL discriminator ``forward`` were executed from the reference with the
    HF Longformer body stubbed to return ``rw_hidden``; the product heads carry the reference's parameter names, so the
    same name-keyed weights load."""
    hidden = T(gr["rw_hidden"])
    head = cpm.rl.RewardHead(VOCAB_PPO, d_model=64)
    ref_weights.fill_(head, seed=31)
    u, c = head.collapsed()                                   # what the fused kernel evaluates (GPU twin in test_gpu_model)
    got = torch.sigmoid(hidden.mean(1) @ u.t() + c).mean(-1, keepdim=True)
    torch.testing.assert_close(got, T(gr["rw_ppo_score"]), rtol=1e-5, atol=1e-6)
    disc = cpm.rl.DiscriminatorHead(d_model=64)
    ref_weights.fill_(disc, seed=32)
    disc.train()
    torch.testing.assert_close(disc(hidden), T(gr["rw_dqn_score_train"]), rtol=1e-5, atol=1e-6)
    bn = disc.score_classifier[1]
    torch.testing.assert_close(bn.running_mean, T(gr["rw_dqn_bn_mean"]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(bn.running_var, T(gr["rw_dqn_bn_var"]), rtol=1e-5, atol=1e-6)
    disc.eval()
    with torch.no_grad():
        torch.testing.assert_close(disc(hidden), T(gr["rw_dqn_score_eval"]), rtol=1e-5, atol=1e-6)


def _pretrain_curve(model, batches, n, seed=None):
    """The body of the reference's training loop (agent_pretrain.py:536-565): batches in file order, epoch after epoch."""
    opt = torch.optim.Adam(model.parameters(), lr=0.0001)
    if seed is not None:
        torch.manual_seed(seed)
    out = []
    while len(out) < n:
        for x, y, mask in batches:
            losses = model.train_step(x, y, mask)
            loss = sum(losses) / 6
            model.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 3)
            opt.step()
            out.append(float(loss.detach()))
            if len(out) == n:
                break
    return np.asarray(out)


@pytest.mark.parametrize("tag,live_dropout", [("drop", True), ("eval", False)])
def test_pretraining_loss_curve_matches_reference_train_loop(gr, cpm, tmp_path, tag, live_dropout):
    """10 optimizer steps of the REFERENCE's own ``train()`` (run from agent_pretrain.py on files in its formats) against
    the oracle driven by the loop restated above, with the corpus read back through the product's npz loader.  With dropout
    live the oracle consumes torch's CPU RNG stream exactly as the reference model did (same modules in the same order)."""
    np.savez(tmp_path / "train_data_linear.npz", **ref_weights.pretrain_corpus())
    d = cpm.data.load_cp_npz(tmp_path / "train_data_linear.npz", pin=False)
    batches = list(cpm.data.CPBatches(d, batch_size=4, device="cpu"))
    assert len(batches) == 2 and batches[0][0].shape == (4, 48, 6)
    m = mo.OracleCPModel(VOCAB_DQN, is_training=True, variant="dqn", dropout=0.1, **SMALL)
    ref_weights.fill_(m, seed=13)
    m.train(live_dropout)
    got = _pretrain_curve(m, batches, 10, seed=71)
    np.testing.assert_allclose(got, gr[f"pre_losses_{tag}"], rtol=2e-5, atol=2e-5)
    assert got[-1] < got[0] - 0.3


def test_dqn_update_loop_matches_reference_run(gr):
    """Four whole ``DQN.update`` calls of the reference (real eval / target networks, Adam 0.01 + MultiStepLR, target sync at
    update 0) against the oracle driven by the update restated from IRL_dqn_train.py:267-345."""
    ev = mo.OracleCPModel(VOCAB_DQN, variant="dqn", **SMALL).eval()
    tg = mo.OracleCPModel(VOCAB_DQN, variant="dqn", **SMALL).eval()
    ref_weights.fill_(ev, seed=14)
    ref_weights.fill_(tg, seed=15)
    opt = torch.optim.Adam(ev.parameters(), lr=0.01)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[20, 40], gamma=0.1)
    rows = []
    for u, b in enumerate(ref_weights.rl_update_batches(4, VOCAB_DQN, seed=95)):
        if u % 50 == 0:
            tg.load_state_dict(ev.state_dict())
        mse = rl.dqn_td_loss_compat(ev(b["state"]), tg(b["nextstate"]), b["action"], b["reward"], b["done"], gamma=0.95)
        ce = sum(ev.train_step(b["state"], b["nextstate"], b["mask"])) / 6
        total = 0.3 * mse + 0.7 * ce
        opt.zero_grad()
        total.backward()
        opt.step()
        sched.step()
        rows.append([float(mse.detach()), float(ce.detach()), float(total.detach())])
    np.testing.assert_allclose(np.asarray(rows), gr["loop_dqn_mse_ce_total"], rtol=2e-4, atol=2e-4)


def test_ppo_update_loop_matches_reference_run(gr, cpm):
    """Three ``PPO.update_policy`` epochs of the reference (real actor / critic, the script's own AgentMemory / ExpertMemory)
    against the oracle + the product's device-resident buffers (on the CPU here) driven by the update restated from
    ppo_train.py:365-416."""
    actor = mo.OracleCPModel(VOCAB_PPO, variant="actor", **SMALL).eval()
    critic = mo.OracleCritic(VOCAB_PPO, **SMALL).eval()
    ref_weights.fill_(actor, seed=16)
    ref_weights.fill_(critic, seed=17)
    abuf, ebuf = cpm.data.AgentMemory(30, device="cpu"), cpm.data.ExpertMemory(30, device="cpu")
    ref_weights.fill_ppo_buffers(abuf, ebuf, ref_weights.rl_update_batches(1, VOCAB_PPO, seed=96)[0])
    agent_all, expert_all = abuf.get(), ebuf.get()
    returns = rl.calculate_returns_compat(agent_all["rewards"], 0.99)
    adv = rl.calculate_advantages_compat(returns, agent_all["values"])
    torch.testing.assert_close(returns, T(gr["loop_ppo_returns"]), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(adv, T(gr["loop_ppo_adv"]), rtol=1e-5, atol=1e-5)
    a_opt, c_opt = torch.optim.Adam(actor.parameters(), lr=0.01), torch.optim.Adam(critic.parameters(), lr=0.01)
    actor_losses, value_losses = [], []
    for _ in range(3):
        states = agent_all["states"]
        _, new_logp = rl.ppo_select_update_compat(actor.forward_output(actor.forward_hidden(states)))
        value_pred = critic.value_produce(states)
        policy_loss = rl.ppo_policy_loss_compat(new_logp, agent_all["log_actions"], adv)
        ce = sum(actor.train_step(states, expert_all["states"], expert_all["mask_state"])) / 6
        actor_loss = policy_loss + ce
        value_loss = rl.value_loss_compat(returns, value_pred)
        a_opt.zero_grad()
        actor_loss.backward()
        a_opt.step()
        c_opt.zero_grad()
        value_loss.backward()
        c_opt.step()
        actor_losses.append(float(actor_loss.detach()))
        value_losses.append(float(value_loss.detach()))
    np.testing.assert_allclose(actor_losses, gr["loop_ppo_actor_loss"], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(value_losses, gr["loop_ppo_value_loss"], rtol=2e-3, atol=2e-4)
