#!/usr/bin/env python
"""bench.py — CP tokens/s of one PPO iteration (recurrent rollout + update) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cpmusic|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2]: "rollout of 256 songs x 1024 CP tokens with nucleus sampling,
followed by a GAE plus clipped-loss update"; it fits one GPU, so it is the per-GPU workload and N GPUs
run N such shards — weak scaling): autoregressive RECURRENT rollout of 256 songs x 1024 compound-word
tokens with the reference's per-attribute temperature / nucleus sampling, critic values, GAE(lambda)
with globally normalised advantages, then one clipped-PPO update of the 12-layer / d512 / 8-head actor
and critic over those 256x1024 tokens (2 minibatches of 128x1024 with gradient accumulation, dropout
0.1, grad-clip 3, Adam, bucketed NCCL gradient all-reduce).  Synthetic data: random-init
weights, random initial tokens, and a synthetic reward (the reference's Longformer reward model is
out of scope).  One "step" = one such iteration; value = tokens generated and trained
on per second over all GPUs.

Prints ONE JSON line (rank 0).  `--impl reference` times the oracle port of the reference's CPU
path (oracle/, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

VOCAB = [56, 135, 18, 87, 18, 25]            # AIlabs-Pop1K7 dictionary without 'type' (IRL_dqn_train.py:403)
SONGS_PER_GPU, ROLLOUT_LEN = 256, 1024
# sequences per update minibatch (gradient accumulation over 256/MINIBATCH).  Measured on one B200: 64 -> 276 ms per update phase,
# 128 -> 261 ms, 256 -> 254 ms but +36 ms of allocator churn outside it (63 GB of activations per model), so 128.
MINIBATCH = int(os.environ.get("CPM_MINIBATCH", "128"))
# critic update of iteration i under the rollout of i+1 (second stream; the rollout reads only the actor, and its chain of small
# dependent kernels leaves most of the GPU idle).  Round 1 (library GEMMs, 103 launches per token): 333.6 k tokens/s vs 344.9 k
# sequential - off.  Round 2 (own kernels, 67 launches per token, rollout graph on a high-priority stream): 366 k vs 348 k - ON.
# The in-situ roofline timing of the attention kernels leaves the overlapped stream out (its launches share the GPU with the
# rollout); CPM_OVERLAP_CRITIC=0 gives the strictly sequential iteration.
OVERLAP_CRITIC = os.environ.get("CPM_OVERLAP_CRITIC", "1") == "1"
# SMs the overlapped critic update may use (a CUDA green context, graphs.sm_partition_stream); 0 = the whole device (default).
# Measured (profiles/r02_summary.md section O): the rollout takes 450 ms alone and 520-533 ms with the critic update alongside
# whether that update may use all 148 SMs (and then lasts 190 ms instead of 111) or is confined to 48 (438 ms instead of 279) or 80 -
# the loss follows the update's WORK, not the SMs it holds, so confining it buys nothing: 684.8 ms per iteration with 0,
# 773.9 / 697.8 / 684.6 / 678.7 with 32 / 48 / 64 / 80 SMs.  Kept as a knob for A/B runs.
CRITIC_SMS = int(os.environ.get("CPM_CRITIC_SMS", "0"))
# DRAM bytes of one linear-attention fwd+bwd launch group measured with ncu (cold L2), keyed by the update minibatch shape
# -> (bytes, the committed ncu log).  (128, 1024), final kernels of round 2: fwd 234.9+41.3 (streaming prefix) + 463.2+120.0 (per-chunk
# output) MB, bwd 390.1+53.1 (streaming suffix) + 666.6+361.3 (main) MB.  (64, 1024): round-1 capture (same passes over the data).
LINATTN_DRAM_BYTES_PER_PAIR = {(64, 1024): (1_098_400_000, "profiles/r01_ncu_linattn_cp_final_64x1024x8.csv"),
                               (128, 1024): (2_330_500_000, "profiles/r02_ncu_linattn_fwd_bwd_128x1024x8_final.csv")}
METRIC = "CP tokens/s, PPO rollout+update"
UNIT = "tokens/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path, bounded sample
# --------------------------------------------------------------------------------------------
def cpu_reference_sample(roll_steps=6, upd_seqs=4, upd_len=512, seed=0):
    """Times, on the host cores, (a) batched recurrent rollout steps with host numpy sampling as
    the reference does (testing-no-type-cp.py:157-167), (b) a teacher-forced value pass + fwd + bwd
    (actor with the C causal-product clone + critic) on upd_seqs x upd_len tokens and (c) one Adam
    step per network; returns the tokens/s of one PPO iteration of the bench workload extrapolated
    from the per-token costs plus the once-per-iteration optimizer cost."""
    import numpy as np
    import torch
    from oracle import model_oracle as mo, sampling_oracle as so
    from oracle.causal_product_c import causal_dot_product_c, num_threads
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(seed)
    actor_r = mo.OracleCPModel(VOCAB, is_training=False).eval()
    actor = mo.OracleCPModel(VOCAB, is_training=True).train()
    critic = mo.OracleCritic(VOCAB).train()
    actor.transformer_encoder.product = causal_dot_product_c
    critic.transformer_encoder.product = causal_dot_product_c
    rng = np.random.RandomState(seed)
    B = SONGS_PER_GPU
    cur = torch.stack([torch.randint(0, n, (B,)) for n in VOCAB], -1)
    # (a) rollout
    mem = None
    with torch.no_grad():
        t0 = None
        for s in range(roll_steps + 1):
            if s == 1:
                t0 = time.perf_counter()          # first step is warm-up
            z = actor_r.pos_emb(actor_r.embed(cur[:, None, :]), 0).squeeze(1)
            h, mem = actor_r.transformer_encoder(z, memory=mem)
            logits = actor_r.forward_output(h)
            nxt = np.zeros((B, 6), np.int64)
            for b in range(B):
                nxt[b] = so.forward_output_sampling({a: logits[i][b].numpy() for i, a in enumerate(so.ATTRS)}, rng)
            cur = torch.from_numpy(nxt)
        t_roll = (time.perf_counter() - t0) / (roll_steps * B)                   # s per generated token
    # (b) update: forward + backward cost per token, and the two Adam steps once per iteration (the GPU arm accumulates the
    #     whole iteration's minibatches into ONE optimizer step per network, so their cost must not be charged per sample token)
    x = torch.stack([torch.randint(0, n, (upd_seqs, upd_len)) for n in VOCAB], -1)
    opt_a = torch.optim.Adam(actor.parameters(), lr=1e-5)
    opt_c = torch.optim.Adam(critic.parameters(), lr=1e-5)

    def fwd_bwd(xs):
        with torch.no_grad():
            critic.value_produce(xs)                                              # rollout-time critic values
        logits = actor.forward_output(actor.forward_hidden(xs))
        lp = sum(torch.log_softmax(l, -1).gather(-1, xs[..., i:i + 1]).mean() for i, l in enumerate(logits))
        opt_a.zero_grad()
        (-lp).backward()
        v = critic.value_produce(xs)
        opt_c.zero_grad()
        (v ** 2).mean().backward()

    fwd_bwd(x)                                                                    # warm-up at full sample size: primitive creation,
    opt_a.step()                                                                  # allocator growth, Adam state
    opt_c.step()
    t0 = time.perf_counter()
    fwd_bwd(x)
    t_fb = (time.perf_counter() - t0) / (upd_seqs * upd_len)
    t0 = time.perf_counter()
    opt_a.step()
    opt_c.step()
    t_adam = time.perf_counter() - t0                                             # s per iteration
    tokens_iter = SONGS_PER_GPU * ROLLOUT_LEN
    tokens_per_s = tokens_iter / (tokens_iter * (t_roll + t_fb) + t_adam)
    return {"value": tokens_per_s, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{roll_steps} batch-{B} recurrent steps + numpy sampling ({t_roll * 1e3:.2f} ms/token), one actor+critic "
                      f"value pass + fwd + bwd on {upd_seqs}x{upd_len} tokens ({t_fb * 1e3:.2f} ms/token) and one Adam step per network "
                      f"({t_adam:.2f} s, charged once per {tokens_iter}-token iteration); oracle port with C/OpenMP "
                      f"causal_product ({num_threads()} OMP threads), fp32"}


def workload_config(world):
    tokens_step = SONGS_PER_GPU * ROLLOUT_LEN * world
    return {"workload": f"cfg3: PPO rollout ({SONGS_PER_GPU} songs x {ROLLOUT_LEN} CP tokens per GPU, recurrent, per-attribute "
                        f"temperature/nucleus sampling) + critic values + GAE + one clipped-PPO update (actor+critic, "
                        f"{SONGS_PER_GPU // MINIBATCH} minibatches of {MINIBATCH}x{ROLLOUT_LEN}, dropout 0.1, grad-clip 3, Adam)",
            "agent": "CP linear transformer 12L d512 h8 ff2048 (38.98M params) x2 (actor, critic)",
            "tokens_per_step": tokens_step, "parallelism": f"dp{world}", "l2": "inputs larger than L2 (working set > 1 GB/step)",
            "reward": "synthetic (Longformer reward model out of scope)",
            "schedule": ("critic update of iteration i runs on a second stream under the rollout of iteration i+1 (the rollout reads only the "
                         "actor); the last one is drained inside the timed region" if OVERLAP_CRITIC else "sequential")}


def run_reference(args, rank):
    if rank != 0:
        return
    t0 = time.perf_counter()
    vals = []
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_reference_sample(roll_steps=2, upd_seqs=1, upd_len=128)
    for _ in range(max(1, min(args.steps, 3))):
        r = cpu_reference_sample()
        vals.append(r["value"])
    v = sum(vals) / len(vals)
    r["value"] = v
    tokens = SONGS_PER_GPU * ROLLOUT_LEN
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": tokens / v * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": dict(workload_config(1), parallelism="cpu", note="CPU arm: per-token costs measured on a bounded sample of this "
                           "workload (cpu_baseline.sample) and extrapolated to the full iteration"),
            "cpu_baseline": r, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
class PPOIteration:
    def __init__(self, rank, world, dev, dropout=0.1, lr=1e-5, seed=0, songs=None, minibatch=None):
        import torch
        import cpmusic
        self.torch, self.cpm = torch, cpmusic
        self.rank, self.world, self.dev = rank, world, dev
        self.songs = SONGS_PER_GPU if songs is None else songs
        self.mb = min(MINIBATCH if minibatch is None else minibatch, self.songs)
        torch.manual_seed(seed)                       # identical initial weights on every rank
        self.actor = cpmusic.LinearTransformer(VOCAB, dropout=dropout).to(dev)
        self.critic = cpmusic.Critic_Transformer(VOCAB, dropout=dropout).to(dev)
        self.opt_a = torch.optim.Adam(self.actor.parameters(), lr=lr, fused=True)
        self.opt_c = torch.optim.Adam(self.critic.parameters(), lr=lr, fused=True)
        self.red_a = cpmusic.dist.BucketedGradAllReduce(self.actor.parameters(), 25.0)
        self.red_c = cpmusic.dist.BucketedGradAllReduce(self.critic.parameters(), 25.0)
        self.engine = cpmusic.RolloutEngine(self.actor, self.songs, ROLLOUT_LEN, greedy=False, true_positions=True,
                                            seed=1234, seq_base=rank * self.songs)
        self.group = None                              # attach_group() once the process group exists
        g = torch.Generator().manual_seed(1234 + rank)
        self.init_host = torch.stack([torch.randint(0, n, (self.songs,), generator=g) for n in VOCAB], -1).pin_memory()
        self.init_dev = self.init_host.to(dev)
        self.phase_ms = {"rollout": 0.0, "update": 0.0}
        self.host_ms = {"generate": 0.0, "critic_enqueue": 0.0, "steps": 0}      # host time inside the enqueue calls
        self.cstream, self.critic_sms = None, 0
        if OVERLAP_CRITIC and CRITIC_SMS > 0:
            try:
                self.cstream, self.critic_sms = cpmusic.graphs.sm_partition_stream(CRITIC_SMS, dev)
            except Exception as e:                      # no green contexts on this driver: an ordinary side stream
                print(f"bench: SM partition for the critic stream unavailable ({e}); using a plain stream", file=sys.stderr)
        if self.cstream is None:
            self.cstream = torch.cuda.Stream(device=dev)
        if OVERLAP_CRITIC:
            cpmusic.ops.KernelTimer.exclude_streams.add(self.cstream.cuda_stream)
        self.pending, self._inflight = None, None
        self.vstat = torch.zeros((), device=dev)

    def attach_group(self):
        import torch.distributed as dist
        self.group = dist.group.WORLD
        self.red_a.attach()
        self.red_c.attach()

    # The critic's update needs the iteration's returns and trajectories but nothing of it feeds the NEXT rollout (which only
    # reads the actor): with OVERLAP_CRITIC it is queued on a second stream and runs underneath the next iteration's rollout,
    # whose dependent chain of small kernels leaves most of the GPU idle.  Same arithmetic in the same order for both
    # networks (the critic update still completes before the next value pass); flush() drains the last one.
    def _critic_update(self, x, ret, scale, n_mb):
        torch = self.torch
        self.critic.train()
        self.red_c.zero_grad(n_micro=-(-x.shape[0] // self.mb))        # reduce after the LAST micro-batch's gradients
        vstat = torch.zeros((), device=self.dev)
        for i in range(0, x.shape[0], self.mb):
            sl = slice(i, i + self.mb)
            v = self.critic.value_per_position(x[sl])
            vloss = torch.nn.functional.mse_loss(v, ret[sl])
            (vloss * scale).backward()
            vstat += vloss.detach() * (1.0 / n_mb)
        self.red_c.finish()
        self.opt_c.step()
        return vstat

    def _launch_pending(self):
        if self.pending is None:
            return
        torch = self.torch
        x, ret, scale, n_mb, ready = self.pending
        self.pending = None
        self.cstream.wait_event(ready)
        with torch.cuda.stream(self.cstream):
            self.vstat = self._critic_update(x, ret, scale, n_mb)
        self._inflight = (x, ret)                      # keep the inputs alive until the main stream has joined

    def flush(self):
        """Runs a deferred critic update (if any) and joins it: call before reading the critic or stopping the clock."""
        self._launch_pending()
        self.torch.cuda.current_stream().wait_stream(self.cstream)
        self._inflight = None

    def step(self, init_tokens, time_phases=False):
        torch, cpm = self.torch, self.cpm
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if time_phases else None
        if ev:
            self.flush()
            ev[0].record()
        torch.cuda.nvtx.range_push("rollout")
        h0 = time.perf_counter()
        roll = self.engine.generate(init_tokens)                       # tokens (B,T+1,A), logp (B,T,A)
        h1 = time.perf_counter()
        torch.cuda.nvtx.range_pop()
        self._launch_pending()                                         # last iteration's critic update, under this rollout
        h2 = time.perf_counter()
        self.host_ms["generate"] += (h1 - h0) * 1e3
        self.host_ms["critic_enqueue"] += (h2 - h1) * 1e3
        self.host_ms["steps"] += 1
        if ev:
            ev[1].record()
        torch.cuda.nvtx.range_push("update")
        tokens, old_logp = roll["tokens"], roll["logp"]
        x, act = tokens[:, :-1].contiguous(), tokens[:, 1:].contiguous()
        B, T, A = act.shape
        reward = ((act.sum(-1) % 7).float() / 7.0)                     # synthetic per-token reward
        dones = torch.zeros(B, T, device=self.dev)
        dones[:, -1] = 1.0
        torch.cuda.current_stream().wait_stream(self.cstream)          # the value pass reads the updated critic
        self._inflight = None
        self.critic.eval()
        with torch.no_grad():
            values = torch.cat([self.critic.value_per_position(x[i:i + self.mb]) for i in range(0, B, self.mb)], 0)
        adv, ret = cpm.rl.gae(reward, values, dones, torch.zeros(B, device=self.dev), 0.99, 0.95, True, self.group)
        self.actor.train()
        n_mb = B // self.mb
        self.red_a.zero_grad(n_micro=n_mb)
        scale = 1.0 / (n_mb * self.world)                               # mean over the GLOBAL batch; grads are SUM-reduced
        if OVERLAP_CRITIC:
            ready = torch.cuda.Event()
            ready.record()
            self.pending = (x, ret, scale, n_mb, ready)
        stats = torch.zeros(4, device=self.dev)
        for i in range(0, B, self.mb):
            sl = slice(i, i + self.mb)
            lc = self.actor.logits_concat(self.actor.hidden(x[sl]))
            new_logp, ent = cpm.ops.heads_logp(lc, act[sl], self.actor.seg, True)
            out = cpm.ops.ppo_loss_standard(new_logp, old_logp[sl], adv[sl, :, None].expand(-1, -1, A), None, None, ent,
                                            clip=0.2, vf_coef=0.0, ent_coef=0.01)
            (out[0] * scale).backward()
            stats += torch.stack([out[0].detach(), out[1].detach(), torch.zeros((), device=self.dev), out[3].detach()]) * (1.0 / n_mb)
        self.red_a.finish()
        torch.nn.utils.clip_grad_norm_(self.actor.parameters(), 3.0, foreach=True)      # reference: clip 3 (agent_pretrain.py:563)
        self.opt_a.step()
        if not OVERLAP_CRITIC:
            self.vstat = self._critic_update(x, ret, scale, n_mb)
        stats[2] = self.vstat                        # value loss of the most recent COMPLETED critic update (one iteration late when overlapped)
        torch.cuda.nvtx.range_pop()
        if ev:
            ev[2].record()
            torch.cuda.synchronize()
            self.phase_ms["rollout"] += ev[0].elapsed_time(ev[1])
            self.phase_ms["update"] += ev[1].elapsed_time(ev[2])
        return tokens, stats


def tokens_per_s_weak(ms, steps, world):
    return SONGS_PER_GPU * ROLLOUT_LEN * world * steps / (ms * 1e-3)


def time_recurrent_step_kernel(dev, peak):
    """Standalone HBM roofline of the recurrent attention step kernel at the rollout shape
    (256 sequences x 8 heads x 12 layers' states = 830 MB touched per pass, L2 flushed between)."""
    import torch
    import cpmusic
    N, H, layers = SONGS_PER_GPU, 8, 12
    S = torch.zeros(layers, N, H, 64, 64, device=dev)
    Z = torch.zeros(layers, N, H, 64, device=dev)
    qkv = torch.randn(N, 3 * H * 64, device=dev).bfloat16()
    q, k, v = (qkv[:, j * H * 64:(j + 1) * H * 64].unflatten(-1, (H, 64)) for j in range(3))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for l in range(layers):
        cpmusic.ops.linattn_step(q, k, v, S[l], Z[l])
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()                       # 12 launches captured: GPU time, not Python launch time
    with cpmusic.ops.graph_capture(graph):
        for l in range(layers):
            cpmusic.ops.linattn_step(q, k, v, S[l], Z[l])
    tot = 0.0
    reps = 5
    for r in range(reps + 2):
        flush.zero_()
        a.record()
        graph.replay()
        b.record()
        torch.cuda.synchronize()
        if r >= 2:
            tot += a.elapsed_time(b)
    ms = tot / reps / layers
    bytes_alg = N * 270336                     # per (sequence, layer)
    return {"kernel": "linattn_step_kernel", "bound": "hbm", "achieved": bytes_alg / (ms * 1e-3) / 1e9, "peak": peak,
            "unit": "GB/s", "frac": bytes_alg / (ms * 1e-3) / 1e9 / peak, "us_per_launch": ms * 1e3,
            "note": "standalone CUDA graph of 12 launches (one per layer state), L2 flushed before each replay"}


def time_linattn_standalone(dev, peak):
    """The chunked attention kernels alone at the update shape (MINIBATCH x ROLLOUT_LEN x 8 heads, fused QKV layout, prefix states
    kept for the backward as in training): one CUDA graph per direction, L2 flushed before each replay.  Reported next to the
    in-situ figure because the two differ by the SM clock: inside the update phase the GEMMs before and after hold the chip at
    its power cap (SM clocks near 1.3-1.5 GHz instead of 1.95), and these kernels - a chain of short phases per tile - scale
    with it (tools/bench_linattn.py --hot 40 reproduces the in-situ times)."""
    import torch
    import cpmusic
    N, L, H = MINIBATCH, ROLLOUT_LEN, 8
    gen = torch.Generator().manual_seed(0)
    qkv = torch.randn(N, L, 3 * H * 64, generator=gen).to(dev).bfloat16()
    q, k, v = (qkv[..., i * H * 64:(i + 1) * H * 64].unflatten(-1, (H, 64)) for i in range(3))
    go = torch.randn(N, L, H, 64, generator=gen).to(dev).bfloat16()
    gqkv = torch.empty_like(qkv)
    gq, gk, gv = (gqkv[..., i * H * 64:(i + 1) * H * 64].unflatten(-1, (H, 64)) for i in range(3))
    saved = cpmusic.ops.linattn_saved(N, L, H, dev)
    out, den = cpmusic.ops.linattn_fwd_raw(q, k, v, saved=saved)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    res = {}
    for name, fn in (("ms_fwd", lambda: cpmusic.ops.linattn_fwd_raw(q, k, v, saved=saved)),
                     ("ms_bwd", lambda: cpmusic.ops.linattn_bwd_raw(q, k, v, out, den, go, gq, gk, gv, saved=saved))):
        fn()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with cpmusic.ops.graph_capture(graph):
            fn()
        tot, reps = 0.0, 10
        for r in range(reps + 2):
            flush.zero_()
            a.record()
            graph.replay()
            b.record()
            torch.cuda.synchronize()
            if r >= 2:
                tot += a.elapsed_time(b)
        res[name] = tot / reps
    byt = N * L * H * 1408
    ach = byt / ((res["ms_fwd"] + res["ms_bwd"]) * 1e-3) / 1e9
    res.update({"achieved": ach, "frac": ach / peak,
                "note": "kernels alone (one CUDA graph per direction, L2 flushed, chip not power-capped by surrounding GEMMs)"})
    return res


def run_gpu(args, rank, world):
    import torch
    import cpmusic
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpmusic._lib.load()
    peak, peak_src = load_peaks()
    it = PPOIteration(rank, world, dev)
    if world > 1:
        # Build the NCCL communicator only AFTER the models, the recurrent state, the weight packings and the captured rollout
        # graph exist.  Measured (tools/clock_probe.py, profiles/r01_summary.md section R): with the communicator created
        # before anything else the very same graph replays take 8 % longer (505 vs 466 us per token); created after three full
        # iterations they do not.  Creating it after the first rollout (all we can run before the ranks must stay in step)
        # recovers part of it: 2 GPUs 665.7 k -> 675.6 k tokens/s.
        if os.environ.get("CPM_BENCH_PREWARM", "0") == "1":
            # Opt-in experiment (not yet measured): one whole throw-away iteration on this rank alone, so
            # that the optimizer state, cuBLAS workspaces and the allocator's activation segments of the update phase exist before
            # the communicator too; parameters are then put back and the Adam state zeroed IN PLACE (addresses stay stable for the
            # packed weights and the captured graph), so every rank still starts the measured run from identical weights.
            keep = [p.detach().clone() for m in (it.actor, it.critic) for p in m.parameters()]
            it.step(it.init_dev)
            it.flush()
            torch.cuda.synchronize()
            with torch.no_grad():
                for p, k in zip((p for m in (it.actor, it.critic) for p in m.parameters()), keep):
                    p.copy_(k)
                for opt in (it.opt_a, it.opt_c):
                    for st in opt.state.values():
                        for v in st.values():
                            if torch.is_tensor(v):
                                v.zero_()
            del keep
        else:
            it.engine.generate(it.init_dev)
        torch.cuda.synchronize()
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout to the one JSON line (NCCL prints its version there)
        cpmusic.dist.init_from_env("nccl")
        it.attach_group()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        it.step(it.init_dev)
    it.flush()
    barrier()
    if args.profile_step:                              # ncu --profile-from-start off: one whole iteration, every thread's launches
        if args.profile_step == "rollout":
            torch.cuda.profiler.start()
            it.engine.generate(it.init_dev, n_steps=8)
        else:
            if args.profile_step == "update":          # the update phase alone: the rollout it consumes is generated beforehand
                roll = it.engine.generate(it.init_dev)
                torch.cuda.synchronize()
                it.engine.generate = lambda *_a, **_k: roll
            torch.cuda.profiler.start()
            it.step(it.init_dev)
            it.flush()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    # ---- device-resident timed region ------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    cpmusic._lib.reset_counts()
    cpmusic.ops.KernelTimer.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        it.step(it.init_dev)
    it.flush()                                          # the last critic update belongs to the timed region
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ktimes = cpmusic.ops.KernelTimer.stop()
    launches = cpmusic._lib.kernel_launches()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    # ---- end-to-end region: pinned-host inputs in, tokens + losses out, every step ----------
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    host_tok = torch.empty(SONGS_PER_GPU, ROLLOUT_LEN + 1, 6, dtype=torch.int64).pin_memory()
    host_stats = torch.empty(4).pin_memory()
    for _ in range(args.steps):
        init = it.init_host.to(dev, non_blocking=True)
        tokens, stats = it.step(init)
        host_tok.copy_(tokens, non_blocking=True)
        host_stats.copy_(stats, non_blocking=True)
        torch.cuda.synchronize()                    # the caller reads the result every step
    it.flush()
    f1.record()
    torch.cuda.synchronize()
    ms_e2e = f0.elapsed_time(f1)
    # phase split (extra untimed step)
    it.phase_ms = {"rollout": 0.0, "update": 0.0}
    it.step(it.init_dev, time_phases=True)
    if OVERLAP_CRITIC:                                   # the deferred critic update of that step, timed alone
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        it.flush()
        c1.record()
        torch.cuda.synchronize()
        it.phase_ms["critic_update_alone"] = c0.elapsed_time(c1)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    # ---- cfg3 as BASELINE.json words it: 256 songs IN TOTAL, 256 / N per GPU (strong scaling) --------------------------
    strong = {"songs_total": SONGS_PER_GPU, "songs_per_gpu": SONGS_PER_GPU // world, "tokens_per_s": tokens_per_s_weak(ms, args.steps, 1),
              "ms_per_step": ms / args.steps, "note": "identical to the headline at 1 GPU"}
    if world > 1 and SONGS_PER_GPU % world == 0:
        it_s = PPOIteration(rank, world, dev, songs=SONGS_PER_GPU // world)
        it_s.attach_group()
        for _ in range(2):
            it_s.step(it_s.init_dev)
        it_s.flush()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_s = max(2, min(args.steps, 5))
        g0.record()
        for _ in range(n_s):
            it_s.step(it_s.init_dev)
        it_s.flush()
        g1.record()
        torch.cuda.synchronize()
        t = torch.tensor([g0.elapsed_time(g1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_s = float(t[0]) / n_s
        it_s.phase_ms = {"rollout": 0.0, "update": 0.0}
        it_s.step(it_s.init_dev, time_phases=True)
        strong = {"songs_total": SONGS_PER_GPU, "songs_per_gpu": SONGS_PER_GPU // world, "tokens_per_s": SONGS_PER_GPU * ROLLOUT_LEN / (ms_s * 1e-3),
                  "ms_per_step": ms_s, "steps": n_s, "phase_ms": {k: round(v, 3) for k, v in it_s.phase_ms.items()},
                  "minibatch": it_s.mb, "note": "256 songs x 1024 tokens in total, sharded over the GPUs; max over ranks"}
        del it_s
    if rank != 0:
        return
    tokens_step = SONGS_PER_GPU * ROLLOUT_LEN * world
    value = tokens_step * args.steps / (ms * 1e-3)
    e2e = tokens_step * args.steps / (ms_e2e * 1e-3)
    # ---- roofline of the kernel the metric names: chunked linear attention fwd+bwd ----------
    nf, tf = ktimes.get("linattn_fwd", (0, 0.0))
    nb, tb = ktimes.get("linattn_bwd", (0, 0.0))
    tok_call = MINIBATCH * ROLLOUT_LEN
    bytes_fwd, bytes_bwd = tok_call * 8 * 512, tok_call * 8 * 896            # SURVEY §8d per (token, head): 512 B fwd, 896 B bwd
    n_pairs = max(min(nf, nb), 1)
    ms_pair = (tf / max(nf, 1)) + (tb / max(nb, 1))
    achieved = (bytes_fwd + bytes_bwd) / (ms_pair * 1e-3) / 1e9 if ms_pair > 0 else 0.0
    roofline = {"kernel": f"linattn fwd+bwd ({cpmusic.ops.linattn_last_impl()})", "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": LINATTN_DRAM_BYTES_PER_PAIR.get((MINIBATCH, ROLLOUT_LEN), (None, None))[0],
                "traffic_source": "ncu dram__bytes_read+write per fwd+bwd launch group (%s)" % LINATTN_DRAM_BYTES_PER_PAIR.get((MINIBATCH, ROLLOUT_LEN), (None, "not captured for this shape"))[1],
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch_pair": bytes_fwd + bytes_bwd, "ms_fwd": tf / max(nf, 1), "ms_bwd": tb / max(nb, 1),
                "launch_pairs_timed": n_pairs, "share_of_step": (tf + tb) / ms,
                "tensor_frac_of_measured_bf16": (tok_call * 8 * 49152) / (ms_pair * 1e-3) / 1e12 / 1651.8 if ms_pair > 0 else 0.0}
    roofline_step = time_recurrent_step_kernel(dev, peak)
    if rank == 0:
        roofline["standalone"] = time_linattn_standalone(dev, peak)
    cpu = cpu_reference_sample() if world >= 1 and not args.no_cpu_baseline else None
    diag = {"rollout_us_per_token": round(it.phase_ms["rollout"] * 1e3 / ROLLOUT_LEN, 1), "rollout_mode": it.engine.mode,
            "rollout_layernorm_folded": bool(it.engine.fold), "under_torchrun": "TORCHELASTIC_RUN_ID" in os.environ,
            "cuda_device_max_connections": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"), "omp_num_threads": os.environ.get("OMP_NUM_THREADS"),
            "cpu_affinity": len(os.sched_getaffinity(0)), "state_base_mod_2MiB": int(it.engine.S.data_ptr() % (2 << 20)),
            "allocator_reserved_GB": round(torch.cuda.memory_reserved(dev) / 2**30, 2), "gpu_name": torch.cuda.get_device_name(dev),
            "critic_sms": it.critic_sms if OVERLAP_CRITIC else None,
            "host_ms_per_step": {k: round(v / max(it.host_ms["steps"], 1), 1) for k, v in it.host_ms.items() if k != "steps"}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": workload_config(world),
            "roofline": roofline, "roofline_recurrent_step": roofline_step, "cpu_baseline": cpu,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(it.init_host.numel() * 8),
                    "d2h_bytes_per_step": int(host_tok.numel() * 8 + 16), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks,
            "phase_ms": {k: round(v, 3) for k, v in it.phase_ms.items()},
            "rollout_kernels_per_token_step": it.engine.launches_per_step, "diag": diag,
            "strong_scaling": strong, "configs": None}
    if world == 1 and not args.no_configs:
        # BASELINE.json's other configurations on this GPU, a few timed iterations each (tools/bench_configs.py)
        del it
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_configs as bc
        line["configs"] = bc.run_cfg2(dev, 10) + bc.run_cfg4(dev, 5) + bc.run_cfg5(dev, 3)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cpmusic", choices=["cpmusic", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg2 / cfg4 / cfg5 timings appended at 1 GPU")
    ap.add_argument("--profile-step", nargs="?", const="all", default=None, choices=["all", "update", "rollout"],
                    help="after the warm-up run ONE iteration (or only its update phase / 8 token steps of its rollout) between "
                         "cudaProfilerStart/Stop and exit (for ncu --profile-from-start off)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_gpu(args, rank, world)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
