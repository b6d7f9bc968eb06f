"""The product's HOST LOGIC on a machine without a GPU: every test body of tests/test_gpu_ref_golden.py (the GPU twins of
the executed-reference vectors) is run here on the CPU with the CUDA entry points of ``cpmusic.ops`` replaced by the
plain-torch stand-ins of tests/emulated_ops.py.  What this holds to the reference vectors is everything between the
reference-shaped API and the kernel calls — weight packing caches, embedding / logits segment layout, position offsets,
the recurrent state protocol, read-out index arithmetic, the replay buffers, the generation driver, the training loops;
the kernels themselves are only exercised by ``-m gpu``."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import emulated_ops  # noqa: E402
import test_gpu_ref_golden as twins  # noqa: E402

CPU = torch.device("cpu")
pytestmark = pytest.mark.skipif(torch.cuda.is_available(), reason="GPU present: the real kernels are tested by -m gpu, nothing stands in for them")


@pytest.fixture()
def emu(cpm):
    with emulated_ops.emulated(cpm):
        yield cpm


def test_teacher_forced_surface(emu, golden):
    twins.test_teacher_forced_surface_vs_reference_run(CPU, emu, golden)


def test_bf16_losses(emu, golden):
    twins.test_bf16_losses_vs_reference_run(CPU, emu, golden)


def test_recurrent_protocol(emu, golden):
    twins.test_recurrent_protocol_vs_reference_run(CPU, emu, golden)


def test_actor_critic_and_readouts(emu, golden):
    twins.test_actor_critic_and_readouts_vs_reference_run(CPU, emu, golden)


def test_reward_head(emu, golden):
    twins.test_reward_head_kernel_vs_reference_run(CPU, emu, golden)


def test_inference_from_scratch_driver(emu, golden):
    """The per-token driver only: ``batched_generate`` runs through the CUDA-graph rollout engine (GPU test)."""
    import ref_weights
    _, w2e = ref_weights.synthetic_dictionary()
    m = emu.LinearTransformer(twins.VOCAB_DQN, False, compute_dtype=torch.float32, dropout=0.0, **twins.SMALL)
    m.load_state_dict(twins._weights(twins.VOCAB_DQN, 11))
    m.eval()
    words = emu.midi.inference_from_scratch(m, w2e, 3, max_tokens=200)
    bars = 1 + sum(1 for w in words[1:] if w2e["bar-beat"][int(w[2])] == "Bar")
    assert tuple(words[0]) == emu.midi.BAR_TOKEN and (len(words) == 200 or bars == 3)
    assert all(0 <= int(words[:, a].min()) and int(words[:, a].max()) < twins.VOCAB_DQN[a] for a in range(6))


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 3e-3)])
def test_pretraining_loss_curve(emu, golden, tmp_path, dtype, tol):
    twins.test_pretraining_loss_curve_vs_reference_train_loop(CPU, emu, golden, tmp_path, dtype, tol)


def test_dqn_td_layout(emu, golden):
    twins.test_dqn_td_kernel_vs_reference_dqn_update(CPU, emu, golden)


def test_dqn_update_loop(emu, golden):
    twins.test_dqn_update_loop_vs_reference_run(CPU, emu, golden)


def test_ppo_update_loop(emu, golden):
    twins.test_ppo_update_loop_vs_reference_run(CPU, emu, golden)


# The module-level GPU tests that predate the executed-reference vectors (oracle / golden comparisons): the same bodies on the
# CPU stand-ins keep the host logic under regression test between GPU runs.
@pytest.mark.parametrize("name", ["test_train_step_fp32_golden", "test_train_step_bf16", "test_module_vs_live_oracle_random_weights",
                                  "test_module_with_128_wide_heads_vs_live_oracle", "test_recurrent_forward_hidden",
                                  "test_ppo_and_dqn_readouts", "test_actor_critic_value_paths",
                                  "test_fast_transformers_shim_runs_reference_style_model"])
def test_module_level_gpu_tests_on_stand_ins(emu, golden, name):
    import inspect
    import test_gpu_model as tm
    fn = getattr(tm, name)
    have = {"cuda": CPU, "cpm": emu, "golden": golden}
    fn(**{p: have[p] for p in inspect.signature(fn).parameters})


# ------------------------------------------------------------------------------------------------------------------------
# The reference's own training-script classes, UNMODIFIED, driving the product's models (CPU stand-ins for the kernels).
# Needs the reference tree, which only exists in the build container.
needs_reference = pytest.mark.skipif(not os.path.isdir("/root/reference/ppo_policy"), reason="reference tree only exists in the build container")


def _golden_tools():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_ref_golden as mk
    import ref_weights
    return mk, ref_weights


@needs_reference
def test_reference_dqn_class_drives_the_product_models(emu, golden, monkeypatch):
    """``DQN.update`` lifted from IRL_dqn_train.py, with ``eval_net`` / ``target_net`` = cpmusic.LinearTransformer: the four
    updates reproduce the losses the same class produced on the reference's own networks."""
    import contextlib
    import io
    import numpy as np
    from tqdm import tqdm
    import torch.nn as nn
    import torch.nn.functional as F
    mk, rw = _golden_tools()
    ns = dict(torch=torch, nn=nn, F=F, np=np, N_ACTIONS=25, GAMMA=0.95, Target_update=50, object=object, tqdm=tqdm,
              wandb=mk._Obj(log=lambda *a, **k: None), NUM_SONGS=1500, EPISODES=50, num=0)
    DQN = mk.lift_class("/root/reference/dqn_policy/IRL_dqn_train.py", "DQN", ns)
    ev = emu.LinearTransformer(twins.VOCAB_DQN, True, compute_dtype=torch.float32, dropout=0.0, **twins.SMALL)
    tg = emu.LinearTransformer(twins.VOCAB_DQN, True, compute_dtype=torch.float32, dropout=0.0, **twins.SMALL)
    ev.load_state_dict(twins._weights(twins.VOCAB_DQN, 14))
    tg.load_state_dict(twins._weights(twins.VOCAB_DQN, 15))
    agent = object.__new__(DQN)
    agent.eval_net, agent.target_net = ev.train(), tg.train()
    agent.optim = torch.optim.Adam(ev.parameters(), lr=0.01)
    agent.scheduler = torch.optim.lr_scheduler.MultiStepLR(agent.optim, milestones=[20, 40], gamma=0.1)
    agent.target_count = agent.cnt_update = agent.record_fore_epoch = 0
    agent.mse_val = agent.ce_val = agent.total_val = 0.0
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    rows, prev = [], (0.0, 0.0, 0.0)
    for b in rw.rl_update_batches(4, twins.VOCAB_DQN, seed=95):
        tr = {"state": b["state"], "nextstate": b["nextstate"], "action": b["action"], "reward": b["reward"], "done": b["done"]}
        with contextlib.redirect_stdout(io.StringIO()):
            agent.update(tr, {"state": b["state"], "nextstate": b["nextstate"]}, b["mask"], False, 0)
        cur = (agent.mse_val, agent.ce_val, agent.total_val)
        rows.append([c - p for c, p in zip(cur, prev)])
        prev = cur
    np.testing.assert_allclose(np.asarray(rows), golden("ref_rl")["loop_dqn_mse_ce_total"], rtol=3e-4, atol=3e-4)
    assert torch.equal(agent.choose_action(b["state"][:1], None), emu.rl.dqn_choose_action(ev, b["state"][:1]))


@needs_reference
def test_reference_ppo_class_drives_the_product_models(emu, golden):
    """``PPO.choose_action`` / ``select_udpate`` / ``update_policy`` and the script's own AgentMemory / ExpertMemory lifted
    from ppo_train.py, with ``actor_net`` / ``critic_net`` = cpmusic.Actor_Transformer / Critic_Transformer."""
    import contextlib
    import io
    import numpy as np
    from tqdm import tqdm
    import torch.nn as nn
    import torch.nn.functional as F
    mk, rw = _golden_tools()
    gr = golden("ref_rl")
    actor = emu.Actor_Transformer(twins.VOCAB_PPO, compute_dtype=torch.float32, dropout=0.0, **twins.SMALL)
    critic = emu.Critic_Transformer(twins.VOCAB_PPO, compute_dtype=torch.float32, dropout=0.0, **twins.SMALL)
    actor.load_state_dict(twins._weights(twins.VOCAB_PPO, 16, variant="actor"))
    critic.load_state_dict(twins._weights(twins.VOCAB_PPO, 17, critic=True))
    value_losses = []

    def recording_mse(a, b, *args, **kw):
        v = F.mse_loss(a, b, *args, **kw)
        value_losses.append(float(v.detach()))
        return v

    script = "/root/reference/ppo_policy/ppo_train.py"
    ns = dict(torch=torch, nn=nn, F=mk._Obj(mse_loss=recording_mse), np=np, device=CPU, N_ACTIONS=25, tqdm=tqdm, Load_Pretrain=False,
              object=object, BUFFER_SIZE=30, N_STATES=50, N_FEATURES=6)
    PPO = mk.lift_class(script, "PPO", ns)
    abuf, ebuf = mk.lift_class(script, "AgentMemory", ns)(), mk.lift_class(script, "ExpertMemory", ns)()
    rw.fill_ppo_buffers(abuf, ebuf, rw.rl_update_batches(1, twins.VOCAB_PPO, seed=96)[0])
    ppo = object.__new__(PPO)
    ppo.actor_net, ppo.critic_net = actor.train(), critic.train()
    ppo.actor_optim = torch.optim.Adam(actor.parameters(), lr=0.01)
    ppo.critic_optim = torch.optim.Adam(critic.parameters(), lr=0.01)
    ns.update(AgentBuffer=abuf, ExpertBuffer=ebuf, Agent=ppo)
    states = abuf.get()["states"]
    with torch.no_grad():                                              # the script's own read-outs == the fused ones
        act, logp = ppo.choose_action(states[:1])
        act2, logp2 = emu.rl.ppo_choose_action(actor, states[:1])
        assert torch.equal(act, act2)
        torch.testing.assert_close(logp, logp2, rtol=1e-5, atol=1e-5)
        act, logp, value = ppo.select_udpate(states)
        act2, logp2 = emu.rl.ppo_select_update(actor, states)
        assert torch.equal(act, act2)
        torch.testing.assert_close(logp, logp2, rtol=1e-5, atol=1e-5)
    returns = ppo.calculate_returns(abuf.get()["rewards"], 0.99)
    adv = ppo.calculate_advantages(returns, abuf.get()["values"])
    actor_losses = []
    for _ in range(3):
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            actor_losses.append(ppo.update_policy(1, 0.2, adv, returns))
    np.testing.assert_allclose(actor_losses, gr["loop_ppo_actor_loss"], rtol=3e-4, atol=3e-4)
    np.testing.assert_allclose(value_losses, gr["loop_ppo_value_loss"], rtol=3e-3, atol=3e-4)


@needs_reference
def test_reference_train_and_generation_functions_drive_the_product_model(emu, golden, tmp_path, monkeypatch):
    """``train()`` from agent_pretrain.py and ``inference_from_scratch`` from testing-no-type-cp.py, lifted unmodified, with
    ``TransformerModel`` / ``model`` = the cpmusic classes: the training loop logs the loss curve it logged on the
    reference's own model; the generation loop runs to its bar condition on the product's recurrent surface."""
    import ast
    import contextlib
    import datetime
    import io
    import math
    import pickle
    import time
    import numpy as np
    import torch.nn as nn
    import torch.nn.functional as F
    mk, rw = _golden_tools()
    path = "/root/reference/dqn_policy/agent_pretrain.py"
    tree = ast.parse(open(path).read())
    nodes = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("network_paras", "train")]
    losses = []

    class Saver:
        first = True

        def __init__(self, *a, **k):
            pass

        def add_summary_msg(self, msg):
            if Saver.first:
                Saver.first = False
                net = sys._getframe(1).f_locals["net"]
                rw.fill_(net, seed=13)

        def add_summary(self, key, val, *a, **k):
            if key == "batch loss":
                losses.append(val)
                if len(losses) == 10:
                    raise mk._StopTraining

        def global_step_increment(self):
            pass

    e2w, w2e = rw.synthetic_dictionary()
    full = lambda d: {"tempo": d["tempo"], "chord": d["chord"], "bar-beat": d["bar-beat"], "type": {0: "EOS", 1: "Metrical", 2: "Note"},    # noqa: E731
                      "pitch": d["pitch"], "duration": d["duration"], "velocity": d["velocity"]}
    np.savez(tmp_path / "train_data_linear.npz", **rw.pretrain_corpus())
    (tmp_path / "dictionary.pkl").write_bytes(pickle.dumps((full(e2w), full(w2e))))
    (tmp_path / "ckpt").mkdir()
    product = lambda n_class: emu.TransformerModel(n_class, compute_dtype=torch.float32, dropout=0.0, **twins.SMALL)    # noqa: E731
    ns = dict(torch=torch, nn=nn, F=F, np=np, os=os, sys=sys, math=math, time=time, pickle=pickle, optim=torch.optim, datetime=datetime,
              clip_grad_norm_=torch.nn.utils.clip_grad_norm_, Saver=Saver, TransformerModel=product, batch_size=4, init_lr=0.0001,
              path_exp=str(tmp_path / "exp"), path_train_data=str(tmp_path / "train_data_linear.npz"),
              path_dictionary=str(tmp_path / "dictionary.pkl"))
    exec(compile(ast.Module(nodes, []), path, "exec"), ns)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)
    monkeypatch.chdir(tmp_path)
    with contextlib.redirect_stdout(io.StringIO()), pytest.raises(mk._StopTraining):
        ns["train"]()
    np.testing.assert_allclose(losses, golden("ref_rl")["pre_losses_eval"], rtol=3e-4, atol=3e-4)
    # generation loop of the reference on the product's recurrent model (device sampler: Philox, not numpy)
    fn = mk.lift_function("/root/reference/dqn_policy/testing-no-type-cp.py", "inference_from_scratch", dict(np=np, torch=torch))
    m = emu.LinearTransformer(twins.VOCAB_DQN, False, compute_dtype=torch.float32, dropout=0.0, **twins.SMALL)
    m.load_state_dict(twins._weights(twins.VOCAB_DQN, 11))
    np.random.seed(3)
    with contextlib.redirect_stdout(io.StringIO()):
        res = fn(m.eval(), w2e, 3)
    bars = 1 + sum(1 for w in res[1:] if w2e["bar-beat"][int(w[2])] == "Bar")
    assert res.shape[1] == 6 and bars == 3 and tuple(res[0]) == emu.midi.BAR_TOKEN
    np.random.seed(3)
    m._sample_step = 0
    assert np.array_equal(emu.midi.inference_from_scratch(m, w2e, 3), res)       # the product's driver is the same loop


# ------------------------------------------------------------------------------------------------------------------------
# Data parallel on the product model, world_size 2 over gloo: rank shards of one batch, the loss defined over the
# GLOBAL mask sum, gradients summed by BucketedGradAllReduce == the single-process full-batch gradients.
def _group_masked_ce(logits, targets, mask, seg, group=None):
    """Stand-in with the contract of ops.masked_ce under a process group: the returned value is the GLOBAL loss, the gradient
    is this rank's tokens' share of it (numerator local, denominator global)."""
    import torch.distributed as dist
    import torch.nn.functional as F
    A = len(seg) - 1
    l2 = logits.reshape(-1, logits.shape[-1]).float()
    tg, m = targets.reshape(-1, A), mask.reshape(-1).float()
    ce = torch.stack([F.cross_entropy(l2[:, seg[a]:seg[a + 1]], tg[:, a], reduction="none") for a in range(A)], -1)
    num, msum = (ce * m[:, None]).sum(0), m.sum()
    if group is None:
        return num / msum
    packed = torch.cat([num.detach(), msum[None]])
    dist.all_reduce(packed, group=group)
    local = num / packed[A]
    return local + (packed[:A] / packed[A] - local).detach()


def _dp_model_worker(rank, world, port, q):
    import numpy as np
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tests"), os.path.join(root, "tests", "golden")]
    import cpmusic
    import emulated_ops as emo
    import ref_weights
    from oracle import model_oracle as mo
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    vocab, small = [56, 135, 18, 87, 18, 25], dict(d_model=128, n_layer=2, n_head=2, d_inner=256)
    o = mo.OracleCPModel(vocab, **small)
    ref_weights.fill_(o, seed=19)
    g = np.load(os.path.join(root, "tests", "golden", "ref_model.npz"))
    x, y, mask = (torch.from_numpy(g[k]) for k in ("dqn_x", "dqn_y", "dqn_mask"))             # 3 ragged sequences
    with emo.emulated(cpmusic):
        cpmusic.ops.masked_ce = _group_masked_ce
        m = cpmusic.LinearTransformer(vocab, True, compute_dtype=torch.float32, dropout=0.0, **small)
        m.load_state_dict(o.state_dict())
        m.train()
        red = cpmusic.dist.BucketedGradAllReduce(m.parameters(), bucket_mb=0.25)           # several buckets
        cpmusic.dist.init_from_env("gloo")
        red.attach()
        lo, hi = cpmusic.dist.shard_range(x.shape[0], rank, world)                          # 2 + 1 sequences
        red.zero_grad()
        losses = torch.stack(m.train_step(x[lo:hi], y[lo:hi], mask[lo:hi], group=dist.group.WORLD))
        (losses.sum() / 6).backward()
        red.finish()
        got = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
        # single process, whole batch
        full = cpmusic.LinearTransformer(vocab, True, compute_dtype=torch.float32, dropout=0.0, **small)
        full.load_state_dict(o.state_dict())
        full.train()
        ref_losses = torch.stack(full.train_step(x, y, mask))
        (ref_losses.sum() / 6).backward()
    ok = torch.allclose(losses, ref_losses, atol=1e-5) and len(red.buckets) > 1
    worst = 0.0
    for n, p in full.named_parameters():
        if p.grad is None:
            ok = ok and float(got[n].abs().max()) == 0.0 if n in got else ok
            continue
        err = float((got[n] - p.grad).abs().max())
        worst = max(worst, err)
        ok = ok and err <= 1e-5 + 1e-4 * float(p.grad.abs().max())
    q.put((rank, bool(ok), worst))
    dist.destroy_process_group()


def test_data_parallel_product_model_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 23
    procs = [ctx.Process(target=_dp_model_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(60)
    assert [r[:2] for r in res] == [(0, True), (1, True)], res


def test_rollout_engine_host_logic(emu, golden):
    """RolloutEngine without its CUDA graph (``use_graph=False``) on the stand-ins: device-side step counter driving the
    positions, history buffers, per-song Philox streams keyed by the GLOBAL song id (shard invariance), greedy roll-out ==
    teacher-forced argmax on the generated tokens, and ``midi.batched_generate`` on top of it."""
    import ref_weights
    m = emu.LinearTransformer(twins.VOCAB_DQN, False, compute_dtype=torch.float32, dropout=0.0, reference_compat=False, **twins.SMALL)
    m.load_state_dict(twins._weights(twins.VOCAB_DQN, 11))
    m.eval()
    init = torch.from_numpy(golden("ref_model")["dqn_x"][:, 0]).repeat(2, 1)[:4]                 # 4 songs
    full = emu.RolloutEngine(m, 4, 12, greedy=False, true_positions=True, seed=5, seq_base=0, use_graph=False).generate(init, 12)
    assert full["tokens"].shape == (4, 13, 6) and full["logp"].shape == (4, 12, 6)
    for lo in (0, 2):                                                    # two "ranks" of two songs each
        part = emu.RolloutEngine(m, 2, 12, greedy=False, true_positions=True, seed=5, seq_base=lo, use_graph=False).generate(init[lo:lo + 2], 12)
        assert torch.equal(part["tokens"], full["tokens"][lo:lo + 2])
        torch.testing.assert_close(part["logp"], full["logp"][lo:lo + 2])
    greedy = emu.RolloutEngine(m, 4, 10, greedy=True, true_positions=True, use_graph=False).generate(init, 10)["tokens"]
    par = emu.LinearTransformer(twins.VOCAB_DQN, True, compute_dtype=torch.float32, dropout=0.0, **twins.SMALL)
    par.load_state_dict(m.state_dict())
    with torch.no_grad():
        logits = par.eval().forward_output(par.forward_hidden(greedy[:, :-1]))
    assert torch.equal(torch.stack([lg.argmax(-1) for lg in logits], -1), greedy[:, 1:])
    _, w2e = ref_weights.synthetic_dictionary()
    songs = emu.midi.batched_generate(m, w2e, 3, n_songs=3, max_tokens=40, seed=7, use_graph=False)
    for s in songs:
        bars = 1 + sum(1 for w in s[1:] if w2e["bar-beat"][int(w[2])] == "Bar")
        assert tuple(s[0]) == emu.midi.BAR_TOKEN and (len(s) == 40 or bars == 3)


def test_policy_classes(emu, golden):
    twins.test_policy_classes_vs_reference_run(CPU, emu, golden)


def test_policy_classes_tight(emu, golden):
    """The same classes at CPU-fp32 tolerances (the shared body above carries the GPU tolerances)."""
    import numpy as np
    import ref_weights
    gr = golden("ref_rl")
    ev = emu.LinearTransformer(twins.VOCAB_DQN, True, compute_dtype=torch.float32, dropout=0.0, **twins.SMALL)
    tg = emu.LinearTransformer(twins.VOCAB_DQN, True, compute_dtype=torch.float32, dropout=0.0, **twins.SMALL)
    ev.load_state_dict(twins._weights(twins.VOCAB_DQN, 14))
    tg.load_state_dict(twins._weights(twins.VOCAB_DQN, 15))
    dqn = emu.rl.DQN(ev.train(), tg.train(), lr=0.01)
    rows = []
    for b in ref_weights.rl_update_batches(4, twins.VOCAB_DQN, seed=95):
        tr = {k: b[k] for k in ("state", "nextstate", "action", "reward", "done")}
        rows.append([float(t) for t in dqn.update(tr, {"state": b["state"], "nextstate": b["nextstate"]}, b["mask"])])
    np.testing.assert_allclose(rows, gr["loop_dqn_mse_ce_total"], rtol=3e-4, atol=3e-4)
    actor = emu.Actor_Transformer(twins.VOCAB_PPO, compute_dtype=torch.float32, dropout=0.0, **twins.SMALL)
    critic = emu.Critic_Transformer(twins.VOCAB_PPO, compute_dtype=torch.float32, dropout=0.0, **twins.SMALL)
    actor.load_state_dict(twins._weights(twins.VOCAB_PPO, 16, variant="actor"))
    critic.load_state_dict(twins._weights(twins.VOCAB_PPO, 17, critic=True))
    abuf, ebuf = emu.data.AgentMemory(30, device="cpu"), emu.data.ExpertMemory(30, device="cpu")
    ref_weights.fill_ppo_buffers(abuf, ebuf, ref_weights.rl_update_batches(1, twins.VOCAB_PPO, seed=96)[0])
    ppo = emu.rl.PPO(actor.train(), critic.train(), abuf, ebuf, lr=0.01)
    returns = ppo.calculate_returns(abuf.get()["rewards"], 0.99)
    adv = ppo.calculate_advantages(returns, abuf.get()["values"])
    got = [(ppo.update_policy(1, 0.2, adv, returns), float(ppo.last_value_loss.detach())) for _ in range(3)]
    np.testing.assert_allclose([g[0] for g in got], gr["loop_ppo_actor_loss"], rtol=3e-4, atol=3e-4)
    np.testing.assert_allclose([g[1] for g in got], gr["loop_ppo_value_loss"], rtol=3e-3, atol=3e-4)
    assert ppo.update_policy(2, 0.2, adv, returns) > 0                  # several epochs in one call, mean returned
