"""CPU: host-side logic — the C-ABI library loads and exports every symbol include/cpmusic.h declares
(no compute calls without a GPU), module surface / state_dict contract, fast_transformers shim,
loud failure without CUDA, data-parallel helpers over gloo (world_size 2)."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VOCAB = [56, 135, 18, 87, 18, 25]


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "cpmusic.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cpm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(cpm):
    lib = cpm._lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 28
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/cpmusic.h but not exported"
        assert name in cpm._lib.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.cpm_version() == 100
    assert lib.cpm_error_name(-1) == b"CPM_ERR_BAD_SHAPE"
    nhc = 32 * 8 * 4                                    # chunk-parallel path: increments + 2 state regions + gd
    assert lib.cpm_linattn_workspace_bytes(32, 512, 8) == nhc * 4160 * 4 + 2 * nhc * (8192 + 256) + 32 * 512 * 8 * 4
    assert lib.cpm_linattn_saved_bytes(32, 512, 8) == nhc * (8192 + 256)
    # a ragged length takes the chunk-parallel path too (one short chunk here): max(segment states, increments + 2 regions + gd)
    assert lib.cpm_linattn_workspace_bytes(4, 100, 8) == max(2 * 4 * 8 * 1 * (64 * 64 + 64) * 4, 32 * 4160 * 4 + 2 * 32 * (8192 + 256) + 4 * 100 * 8 * 4)
    assert lib.cpm_linattn_saved_bytes(4, 100, 8) == 32 * (8192 + 256) and lib.cpm_linattn_saved_bytes(4, 300, 8) == 32 * 3 * (8192 + 256)
    nhc = 16 * 65                                       # 8192 + 64 tokens: 65 chunks, the last one short; the SIMT plan (17 segments) is smaller
    assert lib.cpm_linattn_workspace_bytes(1, 8192 + 64, 16) == nhc * 4160 * 4 + 2 * nhc * (8192 + 256) + (8192 + 64) * 16 * 4 > 2 * 16 * 17 * (64 * 64 + 64) * 4
    assert lib.cpm_ln_partials_rows() == 296


def test_argument_validation_without_gpu(cpm):
    """Validation happens before any CUDA call, so error codes are testable on CPU."""
    lib = cpm._lib.load()
    rc = lib.cpm_linattn_fwd(None, None, None, None, None, 1, 64, 1, 64, 64, 64, 64, 0, 1e-6, 0, None, 0, None, 0, None)
    assert rc == -4 and b"non-NULL" in lib.cpm_last_error_string()
    buf = ctypes.create_string_buffer(1 << 16)
    p = ctypes.addressof(buf)
    p += (-p) % 16
    rc = lib.cpm_linattn_fwd(p, p, p, p, None, 1, 64, 1, 32, 32, 64, 64, 0, 1e-6, 0, None, 0, None, 0, None)
    assert rc == -1 and b"E = M = 64 or 128" in lib.cpm_last_error_string()
    # 128-wide heads: tensor-core kernels only (bf16, whole 128-token chunks), never the CUDA-core path
    rc = lib.cpm_linattn_fwd(p, p, p, p, None, 1, 128, 1, 128, 128, 128, 128, 0, 1e-6, 0, None, 0, None, 0, None)
    assert rc == -7 and b"128-wide heads" in lib.cpm_last_error_string()            # fp32
    ws = lib.cpm_linattn_workspace_bytes_wide(1, 100, 1, 128)
    rc = lib.cpm_linattn_fwd(p, p, p, p, None, 1, 100, 1, 128, 128, 128, 128, 1, 1e-6, 1, p, ws, None, 0, None)
    assert rc == -7 and b"64-wide" in lib.cpm_last_error_string()                   # bf16 but the CUDA-core path (impl 1) asked for
    nhc = 2 * 8 * 4
    assert lib.cpm_linattn_saved_bytes_wide(2, 512, 8, 128) == nhc * (4 * 8192 + 512)
    assert lib.cpm_linattn_saved_bytes_wide(2, 512, 8, 64) == lib.cpm_linattn_saved_bytes(2, 512, 8)
    assert lib.cpm_linattn_workspace_bytes_wide(2, 512, 8, 64) == lib.cpm_linattn_workspace_bytes(2, 512, 8)
    assert lib.cpm_linattn_workspace_bytes_wide(2, 512, 8, 128) == nhc * (4 * 4096 + 128) * 4 + 2 * nhc * (4 * 8192 + 512) + 2 * 512 * 8 * 4
    assert lib.cpm_linattn_workspace_bytes_wide(2, 500, 8, 128) == lib.cpm_linattn_workspace_bytes_wide(2, 512, 8, 128) - 2 * 12 * 8 * 4
    assert lib.cpm_linattn_workspace_bytes_wide(2, 512, 8, 96) == 0
    rc = lib.cpm_linattn_fwd(p, p, p, p, None, 1, 128, 1, 64, 64, 64, 64, 0, 1e-6, 3, p, 1 << 20, None, 0, None)
    assert rc == -7                                                   # tcgen05 path refuses fp32
    rc = lib.cpm_linattn_fwd(p + 2, p, p, p, None, 1, 64, 1, 64, 64, 64, 64, 0, 1e-6, 0, p, 1 << 20, None, 0, None)
    assert rc == -2
    with pytest.raises(ValueError):
        cpm._lib.check(-1)
    with pytest.raises(cpm._lib.CpmError):
        cpm._lib.check(-6)


def test_model_surface_and_state_dict(cpm):
    from oracle import model_oracle as mo
    for cls, ocls, kw, vocab in ((cpm.LinearTransformer, mo.OracleCPModel, {}, VOCAB),
                                 (cpm.TransformerModel, mo.OracleCPModel, {}, VOCAB),
                                 (cpm.Actor_Transformer, mo.OracleCPModel, {"variant": "actor"}, [49, 19, 19, 89, 67, 25])):
        m, o = cls(vocab), ocls(vocab, **kw)
        assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in o.state_dict().items()}
        m.load_state_dict(o.state_dict())                               # strict
        for meth in ("train_step", "forward_hidden", "forward_output", "forward", "forward_output_sampling",
                     "compute_loss", "inference"):
            assert callable(getattr(m, meth))
    assert len(cpm.LinearTransformer(VOCAB).state_dict()) == 217
    # ... and against the key lists of the REFERENCE's own modules (tests/golden/make_ref_golden.py)
    gm = np.load(os.path.join(ROOT, "tests", "golden", "ref_model.npz"))
    small = dict(d_model=128, n_layer=2, n_head=2, d_inner=2048)
    ppo_vocab = [49, 19, 19, 89, 67, 25]
    assert sorted(cpm.LinearTransformer(VOCAB, **small).state_dict()) == list(gm["dqn_state_keys"])
    assert sorted(cpm.Actor_Transformer(ppo_vocab, **small).state_dict()) == list(gm["ppo_actor_keys"])
    assert sorted(cpm.Critic_Transformer(ppo_vocab, **small).state_dict()) == list(gm["ppo_critic_keys"])
    c = cpm.Critic_Transformer([49, 19, 19, 89, 67, 25])
    assert set(c.state_dict()) == set(mo.OracleCritic([49, 19, 19, 89, 67, 25]).state_dict())
    a = cpm.Actor_Transformer(VOCAB)
    assert hasattr(a, "value_funtion") and not hasattr(a, "project_concat_type")
    seven = cpm.LinearTransformer([56, 135, 18, 3, 87, 18, 25])       # upstream 7-type CP layout
    assert len(seven.attrs) == 7 and seven.seg[-1] == 342 and seven.logits_width == 344
    with pytest.raises(ValueError):
        cpm.LinearTransformer([1, 2, 3])
    # checkpoint wrapper formats of the reference (agent_pretrain.py:601-605 / ppo_train.py:514)
    m = cpm.LinearTransformer(VOCAB, d_model=128, n_layer=1, n_head=2, d_inner=128)
    ck = {"epoch": 3, "model_state_dict": m.state_dict(), "optimizer_state_dict": {}}
    m.load_state_dict(ck["model_state_dict"])


def test_no_cpu_fallback(cpm):
    m = cpm.LinearTransformer(VOCAB, d_model=128, n_layer=1, n_head=2, d_inner=128)
    x = torch.zeros(1, 8, 6, dtype=torch.long)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.train_step(x, x, torch.ones(1, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.forward_hidden(x)
    # the product package never imports the oracle
    pkg_dir = os.path.dirname(cpm.__file__)
    for f in os.listdir(pkg_dir):
        if f.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg_dir, f)).read().replace("oracle/", "").replace("the oracle", "") \
                or f in ("ops.py",), f


def test_fast_transformers_shim(cpm):
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.startswith("fast_transformers")}
    try:
        cpm.install_fast_transformers_shim()
        from fast_transformers.builders import TransformerEncoderBuilder, RecurrentEncoderBuilder
        from fast_transformers.masking import TriangularCausalMask
        enc = TransformerEncoderBuilder.from_kwargs(n_layers=2, n_heads=2, query_dimensions=64, value_dimensions=64,
                                                    feed_forward_dimensions=256, activation="gelu", dropout=0.1,
                                                    attention_type="causal-linear").get()
        from oracle import ft_oracle
        ref = ft_oracle.TransformerEncoderBuilder.from_kwargs(n_layers=2, n_heads=2, query_dimensions=64, value_dimensions=64,
                                                              feed_forward_dimensions=256, activation="gelu", dropout=0.1,
                                                              attention_type="causal-linear").get()
        assert set(enc.state_dict()) == set(ref.state_dict())
        assert TriangularCausalMask(5).lower_triangular
        with pytest.raises(RuntimeError, match="lower triangular"):
            enc(torch.zeros(1, 4, 128), None)
        with pytest.raises(ValueError):
            TransformerEncoderBuilder.from_kwargs(attention_type="full", activation="gelu").get()
        rec = RecurrentEncoderBuilder.from_kwargs(n_layers=1, n_heads=2, query_dimensions=64, value_dimensions=64,
                                                  feed_forward_dimensions=128, activation="gelu", attention_type="causal-linear").get()
        assert hasattr(rec, "new_state")
    finally:
        for k in list(sys.modules):
            if k.startswith("fast_transformers"):
                sys.modules.pop(k)
        sys.modules.update(saved)


def test_segment_plan_matches_header_contract(cpm):
    lib = cpm._lib.load()
    per = (64 * 64 + 64) * 4 * 2
    for N, L, H in ((4, 512, 8), (32, 512, 8), (1, 8192, 16), (1, 8192, 8), (2, 100, 1), (1, 50, 8)):
        b = lib.cpm_linattn_workspace_bytes(N, L, H)
        assert b >= per * N * H
        nhc = N * H * -(-L // 128)               # chunk-parallel tcgen05 path (any length): increments + 2 state regions + gd
        assert b >= nhc * 4160 * 4 + 2 * nhc * 8448 + N * L * H * 4
        assert lib.cpm_linattn_saved_bytes(N, L, H) == nhc * 8448


def test_shard_range(cpm):
    for n, w in ((256, 8), (10, 4), (3, 8), (1625, 8)):
        spans = [cpm.dist.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _dp_worker(rank, world, port, q, late=False):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import cpmusic
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    if not late:
        cpmusic.dist.init_from_env("gloo")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4), torch.nn.Linear(4, 4))
    for p in net[3].parameters():          # a parameter that never receives a gradient in this loss
        p.requires_grad_(True)
    red = cpmusic.dist.BucketedGradAllReduce(net.parameters(), bucket_mb=0.001)
    if late:                               # bench.py's order: buffers first, process group afterwards, then attach()
        assert red.world == 1
        cpmusic.dist.init_from_env("gloo")
        red.attach()
    assert red.world == world
    g = torch.Generator().manual_seed(1)
    X, Y = torch.randn(8, 16, generator=g), torch.randn(8, 4, generator=g)
    lo, hi = cpmusic.dist.shard_range(8, rank, world)
    red.zero_grad()
    loss = ((net[2](net[1](net[0](X[lo:hi]))) - Y[lo:hi]) ** 2).sum() / 8      # share of the GLOBAL mean loss
    loss.backward()
    red.finish()
    grads = [p.grad.clone() for p in net.parameters()]
    # single-process reference on the full batch
    ref = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))
    ref.load_state_dict({k: v for k, v in net.state_dict().items() if not k.startswith("3.")})
    (((ref(X) - Y) ** 2).sum() / 8).backward()
    ok = all(torch.allclose(a, b.grad, atol=1e-6) for a, b in zip(grads[:4], ref.parameters()))
    ok = ok and all(float(g_.abs().max()) == 0.0 for g_ in grads[4:]) and len(red.buckets) > 1
    # moments all-reduce path of zscore is GPU-only; check the collective convention (SUM) directly
    t = torch.ones(3) * (rank + 1)
    dist.all_reduce(t)
    ok = ok and float(t[0]) == sum(range(1, world + 1))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("late", [False, True])
def test_bucketed_allreduce_gloo_world2(late):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (7 if late else 0)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q, late)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]


def _dp_accum_worker(rank, world, port, q):
    """Two micro-batches per optimizer step (what bench.py's update does): the buckets must be reduced ONCE, after the last
    micro-batch, and equal the single-process full-batch gradient."""
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import cpmusic
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    cpmusic.dist.init_from_env("gloo")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))
    red = cpmusic.dist.BucketedGradAllReduce(net.parameters(), bucket_mb=0.001)
    g = torch.Generator().manual_seed(1)
    X, Y = torch.randn(16, 16, generator=g), torch.randn(16, 4, generator=g)
    lo, hi = cpmusic.dist.shard_range(16, rank, world)
    launches = []
    orig = red._launch
    red._launch = lambda b: ((launches.append(b["round"]) if not b["launched"] else None), orig(b))[1]
    ok = True
    for n_micro in (2, 4, 1):                              # also: the counter state is clean again after finish()
        red.zero_grad(n_micro=n_micro)
        launches.clear()
        step = (hi - lo) // n_micro
        for m in range(n_micro):
            sl = slice(lo + m * step, lo + (m + 1) * step)
            (((net(X[sl]) - Y[sl]) ** 2).sum() / 16).backward()
            if m < n_micro - 1:
                ok = ok and not launches                   # nothing is reduced before the last micro-batch
        red.finish()
        ok = ok and all(r == n_micro for r in launches) and len(launches) == len(red.buckets)
        grads = [p.grad.clone() for p in net.parameters()]
        ref = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))
        ref.load_state_dict(net.state_dict())
        (((ref(X) - Y) ** 2).sum() / 16).backward()
        ok = ok and all(torch.allclose(a, b.grad, atol=1e-6) for a, b in zip(grads, ref.parameters()))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_bucketed_allreduce_gradient_accumulation_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 13
    procs = [ctx.Process(target=_dp_accum_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]


def _literal_reward(head, h):
    """ppo_policy/model.py:474-493 written out: six proj, six eval, mean over the sequence, sigmoid, average."""
    scores = []
    for a in head.ATTRS:
        y = getattr(head, f"proj_{a}")(h)
        scores.append(torch.sigmoid(getattr(head, f"eval_{a}")(y).mean(dim=1)))
    return sum(scores) / len(scores), torch.cat(scores, -1)


def test_reward_head_collapse_equals_reference_formula(cpm):
    torch.manual_seed(3)
    head = cpm.rl.RewardHead([49, 19, 19, 89, 67, 25], d_model=64).double()
    h = torch.randn(5, 50, 64, dtype=torch.float64)
    ref, ref_scores = _literal_reward(head, h)
    u, c = head.collapsed()
    got = torch.sigmoid(h.mean(1) @ u.double().t() + c.double())
    assert torch.allclose(got, ref_scores, atol=1e-6) and torch.allclose(got.mean(-1, keepdim=True), ref, atol=1e-6)
    assert {k.split(".")[0] for k in head.state_dict()} == {f"{p}_{a}" for p in ("proj", "eval") for a in head.ATTRS}


def test_cp_npz_loader_and_batches(cpm, tmp_path):
    rng = np.random.RandomState(0)
    x = rng.randint(0, 18, size=(10, 64, 7)).astype(np.int32)
    y = np.roll(x, -1, axis=1)
    mask = (rng.rand(10, 64) > 0.2).astype(np.float64)
    np.savez(tmp_path / "train_data_linear.npz", x=x, y=y, mask=mask)
    d = cpm.data.load_cp_npz(tmp_path / "train_data_linear.npz", pin=False)
    assert d["x"].shape == (10, 64, 6) and d["x"].dtype == torch.int64 and d["mask"].dtype == torch.float32
    assert np.array_equal(d["x"].numpy(), np.concatenate((x[:, :, :3], x[:, :, 4:]), axis=2))      # agent_pretrain.py:525-526
    it = cpm.data.CPBatches(d, batch_size=3, device="cpu", seq_len=50, lo=2, hi=10)
    got = list(it)
    assert len(got) == len(it) == 2 and got[0][0].shape == (3, 50, 6) and got[0][2].shape == (3, 50)
    assert torch.equal(got[1][0], d["x"][5:8, :50]) and torch.equal(got[1][1], d["y"][5:8, :50])
    with pytest.raises(ValueError):
        np.savez(tmp_path / "bad.npz", x=x[:, :, :6], y=y[:, :, :6], mask=mask)
        cpm.data.load_cp_npz(tmp_path / "bad.npz", pin=False)


def test_device_resident_memory_mirrors_reference_buffers(cpm, golden):
    """The four ring buffers against the REFERENCE's own classes (lifted from ppo_train.py:69-212 and
    IRL_dqn_train.py:78-204 by tests/golden/make_ref_golden.py, driven with the same 37 transitions into 30 slots):
    attribute names, store_transition argument orders, get() / sampling() layouts, dtypes and values, incl. the ring
    wrap-around and the .long() truncation of stored log-probs."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import ref_weights
    g = golden("ref_rl")
    stream = ref_weights.transition_stream(37)
    kinds = {("ppo", "agent"): cpm.data.AgentMemory, ("ppo", "expert"): cpm.data.ExpertMemory,
             ("dqn", "agent"): cpm.data.DQNAgentMemory, ("dqn", "expert"): cpm.data.DQNExpertMemory}
    for (tag, name), cls in kinds.items():
        mem = cls(30, device="cpu")
        for tr in stream:
            if name == "expert":
                mem.store_transition(tr["state"], tr["action"], tr["reward"], tr["next_state"], tr["done"], tr["mask_state"], tr["mask_next_state"])
            elif tag == "ppo":
                mem.store_transition(tr["state"], tr["action"], tr["log_action"], tr["value"], tr["reward"], tr["next_state"], tr["done"])
            else:
                mem.store_transition(tr["state"], tr["action"], tr["reward"], tr["next_state"], tr["done"])
        assert mem.memory_counter == int(g[f"mem_{tag}_{name}_counter"]) == 37
        got = mem.get()
        items = list(got.items() if isinstance(got, dict) else enumerate(got))
        want = [k for k in g.files if k.startswith(f"mem_{tag}_{name}_get_")]
        assert [f"mem_{tag}_{name}_get_{k}" for k, _ in items] == want, (tag, name)
        for k, v in items:
            ref = torch.from_numpy(g[f"mem_{tag}_{name}_get_{k}"])
            assert v.dtype == ref.dtype and v.shape == ref.shape and torch.equal(v, ref), (tag, name, k)
        batch = mem.sampling(8, idx=g["mem_sample_idx"])
        assert len(batch) == len([k for k in g.files if k.startswith(f"mem_{tag}_{name}_sample_")])
        for k, v in enumerate(batch):
            ref = torch.from_numpy(g[f"mem_{tag}_{name}_sample_{k}"])
            assert v.dtype == ref.dtype and torch.equal(v, ref), (tag, name, "sample", k)
        own = mem.sampling(16)                                       # own draw: device generator, whole buffer, with replacement
        assert own[0].shape == (16, 50, 6)
    for attr in ("states_agent", "value_agent", "actions_agent", "log_actions_agent", "rewards_agent", "next_states_agent", "dones_agent"):
        assert hasattr(cpm.data.AgentMemory(2, device="cpu"), attr)
    for attr in ("states_exp", "actions_exp", "rewards_exp", "next_states_exp", "dones_exp", "mask_state", "mask_next_state"):
        assert hasattr(cpm.data.ExpertMemory(2, device="cpu"), attr)
    raw = cpm.data.AgentMemory(2, n_states=5, n_actions=2, device="cpu", log_prob_long_compat=False)
    raw.store_transition(torch.zeros(5, 6), torch.zeros(2, 6), torch.full((2, 6), -0.25), 0.0, 1.0, torch.zeros(5, 6), 0.0)
    assert raw.get()["log_actions"][0, 0, 0].item() == -0.25
    with pytest.raises(TypeError):
        raw.store_transition(torch.zeros(5, 6), torch.zeros(2, 6))


def test_pack_cache_notices_fused_optimizer_steps(cpm):
    """torch's fused Adam rewrites parameters without bumping Tensor._version; the packed compute copies must still be
    rebuilt after such a step (PackCache stamps carry an optimizer-step epoch per parameter), while the packings of a model
    that no optimizer touched - a frozen target network, the critic while the actor steps - stay valid."""
    from cpmusic.encoder import PackCache
    lin = torch.nn.Linear(8, 4)
    cache = PackCache()
    w0 = cache.get("k", [lin], torch.float32)[0].clone()
    opt = torch.optim.Adam(lin.parameters(), lr=0.1, fused=True)
    lin.weight.grad, lin.bias.grad = torch.ones_like(lin.weight), torch.ones_like(lin.bias)
    opt.step()
    wc, bc = cache.get("k", [lin], torch.float32)[:2]
    assert not torch.equal(wc, w0)
    assert torch.equal(wc, lin.weight.detach()) and torch.equal(bc, lin.bias.detach())
    # another model's optimizer stepping does not stale this cache: same stamp, no refill
    other = torch.nn.Linear(8, 4)
    stamp = cache._store["k"][0]
    opt2 = torch.optim.Adam(other.parameters(), lr=0.1, fused=True)
    other.weight.grad, other.bias.grad = torch.ones_like(other.weight), torch.ones_like(other.bias)
    opt2.step()
    cache.get("k", [lin], torch.float32)
    assert cache._store["k"][0] == stamp
    # in-place refresh keeps the buffer (graphs captured over it stay valid), and invalidate() forces a refill
    ptr = wc.data_ptr()
    with torch.no_grad():
        lin.weight.data.add_(1.0)                      # a version-less write the stamps cannot see
    cache.invalidate()
    wc2 = cache.get("k", [lin], torch.float32)[0]
    assert wc2.data_ptr() == ptr and torch.equal(wc2, lin.weight.detach())


def test_seven_attribute_surface_matches_oracle_keys(cpm):
    """The seven-attribute layout (`type` at column 3): same parameter names and shapes as the oracle restatement, embedding
    widths sum to 1248, and the oracle's train_step returns seven finite losses on the CPU."""
    from oracle import model_oracle as mo
    vocab7 = [56, 135, 18, 4, 87, 18, 25]
    cfg = dict(d_model=64, n_layer=1, n_head=1, d_inner=64, dropout=0.0)
    o = mo.OracleCPModel(vocab7, **cfg)
    m = cpm.TransformerModel(vocab7, **cfg)
    assert m.attrs == mo.ATTRS7 and sum(m.emb_sizes) == 1248 == m.in_linear.in_features
    so, sm = o.state_dict(), m.state_dict()
    assert set(so) == set(sm) and all(so[k].shape == sm[k].shape for k in so)
    g = torch.Generator().manual_seed(0)
    x = torch.stack([torch.randint(0, n, (2, 20), generator=g) for n in vocab7], -1)
    losses = o.train_step(x, x.roll(-1, 1), torch.ones(2, 20))
    assert len(losses) == 7 and all(torch.isfinite(l) for l in losses)
    with pytest.raises(ValueError, match="6 .*or 7"):
        cpm.TransformerModel([5, 5, 5])


def test_head_width_validation(cpm):
    """Head widths: 64 (the reference) and 128 build; anything else is rejected at construction with a clear message."""
    from cpmusic.encoder import TransformerEncoderBuilder
    for ok in (64, 128):
        enc = TransformerEncoderBuilder.from_kwargs(n_layers=1, n_heads=2, query_dimensions=ok, value_dimensions=ok,
                                                    feed_forward_dimensions=64, activation="gelu", dropout=0.0,
                                                    attention_type="causal-linear").get()
        assert enc.d_head == ok and enc.layers[0].attention.query_projection.out_features == 2 * ok
        assert [tuple(t.shape) for t in enc.new_state(3, "cpu")[0]] == [(3, 2, ok, ok), (3, 2, ok)]
    with pytest.raises(ValueError, match="64 .*or 128"):
        TransformerEncoderBuilder.from_kwargs(n_layers=1, n_heads=2, query_dimensions=32, value_dimensions=32,
                                              feed_forward_dimensions=64, activation="gelu", dropout=0.0, attention_type="causal-linear").get()


def test_discriminator_head_matches_reference_formula(cpm):
    """DQN-side AIRL read-out (AIRL_model.py:91-98,117-120): sequence mean over all positions, then the score classifier;
    parameter names as in the reference, training-mode BatchNorm statistics included."""
    torch.manual_seed(5)
    head = cpm.rl.DiscriminatorHead(d_model=48).double()
    assert sorted(k for k in head.state_dict() if k.endswith("weight")) == [f"score_classifier.{i}.weight" for i in (0, 1, 3, 5)]
    h = torch.randn(9, 50, 48, dtype=torch.float64)
    sc = head.score_classifier
    for mode in (True, False):
        head.train(mode)
        m = h.mean(dim=1)                                      # literal restatement, BatchNorm written out
        z = m @ sc[0].weight.T + sc[0].bias
        if mode:
            mu, var = z.mean(0), z.var(0, unbiased=False)
        else:
            mu, var = sc[1].running_mean, sc[1].running_var
        z = (z - mu) / torch.sqrt(var + sc[1].eps) * sc[1].weight + sc[1].bias
        z = torch.tanh(torch.tanh(z) @ sc[3].weight.T + sc[3].bias)
        ref = torch.sigmoid(z @ sc[5].weight.T + sc[5].bias)
        got = head(h)
        assert got.shape == (9, 1) and torch.allclose(got, ref, atol=1e-6), (mode, (got - ref).abs().max())
    head.train()
    h.requires_grad_()
    torch.nn.functional.binary_cross_entropy(head(h), torch.ones(9, 1, dtype=torch.float64)).backward()      # AIRL.py trains it with BCE
    assert h.grad is not None and sc[0].weight.grad is not None


@pytest.mark.skipif(not os.path.isdir("/root/reference/dqn_policy"), reason="reference tree only exists in the build container")
def test_unmodified_reference_model_files_import_the_shim(cpm):
    """with ``install_fast_transformers_shim()`` the reference's own model files import and construct
    unmodified, their encoders ARE the cpmusic encoders, checkpoints move both ways between the reference classes and
    the native ones, and a forward without a GPU fails loudly instead of computing on the CPU."""
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.startswith("fast_transformers") or k in ("config", "model")}
    try:
        cpm.install_fast_transformers_shim()
        for sub, cls_name, native, vocab in (("dqn_policy", "LinearTransformer", cpm.LinearTransformer, VOCAB),
                                             ("ppo_policy", "Actor_Transformer", cpm.Actor_Transformer, [49, 19, 19, 89, 67, 25]),
                                             ("ppo_policy", "Critic_Transformer", cpm.Critic_Transformer, [49, 19, 19, 89, 67, 25])):
            for k in ("config", "model"):
                sys.modules.pop(k, None)
            sys.path.insert(0, os.path.join("/root/reference", sub))
            try:
                import importlib
                mod = importlib.import_module("model")
            finally:
                sys.path.pop(0)
            ref = getattr(mod, cls_name)(vocab)
            assert type(ref.transformer_encoder).__module__.startswith(cpm.__name__)
            nat = native(vocab)
            assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == {k: tuple(v.shape) for k, v in nat.state_dict().items()}
            nat.load_state_dict(ref.state_dict())
            ref.load_state_dict(nat.state_dict())
            x = torch.zeros(1, 8, 6, dtype=torch.long)
            with pytest.raises(RuntimeError, match="no CPU fallback"):
                ref.value_produce(x) if cls_name == "Critic_Transformer" else ref.forward_hidden(x)
        rec = mod.Actor_Transformer([49, 19, 19, 89, 67, 25], is_training=False)           # RecurrentEncoderBuilder path
        assert hasattr(rec.transformer_encoder, "new_state")
    finally:
        for k in list(sys.modules):
            if k.startswith("fast_transformers") or k in ("config", "model"):
                sys.modules.pop(k)
        sys.modules.update(saved)


def test_ctypes_signatures_match_the_header_prototypes(cpm):
    """ABI drift guard: every prototype in include/cpmusic.h is parsed (return type, parameter count, parameter C types)
    and compared with the ctypes signature the Python host binds it with."""
    src = open(os.path.join(ROOT, "include", "cpmusic.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    protos = re.findall(r"([A-Za-z_][A-Za-z0-9_ ]*?[\s\*]+)(cpm_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src)
    assert len(protos) >= 40

    def ctype_of(decl):
        decl = decl.strip()
        if "*" in decl:
            base = decl.replace("const", "").split("*")[0].strip()
            if decl.count("*") == 2 or "* const *" in decl:
                return "ptrptr"
            return {"int": "ptr:int", "float": "ptr:float"}.get(base, "ptr")
        words = [w for w in decl.replace("const", "").split() if w]
        t = " ".join(words[:-1]) if len(words) > 1 else words[0]
        return {"int": "int", "int32_t": "int32", "int64_t": "int64", "uint64_t": "uint64", "float": "float", "double": "double",
                "long long": "int64", "unsigned long long": "uint64"}[t]

    names = {ctypes.c_int: "int", ctypes.c_int32: "int32", ctypes.c_int64: "int64", ctypes.c_uint64: "uint64", ctypes.c_float: "float",
             ctypes.c_double: "double", ctypes.c_void_p: "ptr", ctypes.c_char_p: "ptr"}
    seen = set()
    for ret, name, params in protos:
        seen.add(name)
        res, args = cpm._lib.SIGNATURES[name]
        params = [p for p in (q.strip() for q in params.split(",")) if p and p != "void"]
        assert len(params) == len(args), f"{name}: header has {len(params)} parameters, ctypes {len(args)}"
        for i, (p, a) in enumerate(zip(params, args)):
            want = ctype_of(p)
            if a in names:
                got = names[a]
            else:                                             # POINTER(c_int) / POINTER(c_float) / POINTER(c_void_p)
                got = {ctypes.c_int: "ptr:int", ctypes.c_float: "ptr:float", ctypes.c_void_p: "ptrptr"}[a._type_]
            ok = want == got or (want.startswith("ptr") and got == "ptr") or {want, got} == {"int", "int32"}
            assert ok, f"{name} parameter {i} ({p!r}): header {want}, ctypes {got}"
        rwant = "ptr" if "*" in ret else ctype_of(ret + " x")
        assert rwant == names[res] or {rwant, names[res]} == {"int", "int32"}, f"{name}: return {ret!r} vs {res}"
    assert seen == set(_declared_symbols()) <= set(cpm._lib.SIGNATURES)
