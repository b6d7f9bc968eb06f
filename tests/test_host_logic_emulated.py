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
