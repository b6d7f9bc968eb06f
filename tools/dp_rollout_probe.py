"""Two-rank probe of the rollout under torch.distributed: are simultaneous rollouts on the GPUs of one box slower than a rollout
next to an idle GPU?  (They are not: profiles/r01_summary.md section R.)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_rollout_probe.py"""
import os, sys, time
sys.path.insert(0, "/root/repo")
import torch, cpmusic
import torch.distributed as dist
VOCAB = [56, 135, 18, 87, 18, 25]
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
actor = cpmusic.LinearTransformer(VOCAB, dropout=0.1).to(dev)
init = torch.stack([torch.randint(0, n, (256,)) for n in VOCAB], -1).to(dev)
engines = {"1 step/graph": cpmusic.RolloutEngine(actor, 256, 512, greedy=False, seed=1),
           "8 steps/graph": cpmusic.GroupedRolloutEngine(actor, 256, 512, groups=1, steps_per_graph=8, greedy=False, seed=1)}
for name, eng in engines.items():
    eng.generate(init); torch.cuda.synchronize()
    for label, delay in (("lockstep", 0.0), ("staggered", 0.6 * rank), ("lockstep", 0.0)):
        dist.barrier(); torch.cuda.synchronize(); time.sleep(delay)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.generate(init); b.record(); torch.cuda.synchronize()
        print(f"rank {rank} {name} {label}: {a.elapsed_time(b) * 1e3 / 512:.1f} us/token", flush=True)
        time.sleep(0.8)
