#!/usr/bin/env bash
# A/B of the throw-away iteration before NCCL initialisation on 2 GPUs:
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_prewarm_ab.sh'
set -u
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port "$1" bench.py --gpus 2 --no-cpu-baseline; }
run 29511 > gpurun_out/bench_2gpu_default.json 2> gpurun_out/bench_2gpu_default.err
CPM_BENCH_PREWARM=1 run 29512 > gpurun_out/bench_2gpu_prewarm.json 2> gpurun_out/bench_2gpu_prewarm.err
for f in default prewarm; do python - "$f" <<'PY'
import json, sys
for line in open(f"gpurun_out/bench_2gpu_{sys.argv[1]}.json"):
    if line.startswith("{"):
        d = json.loads(line)
        print(sys.argv[1], round(d["value"]), "tokens/s", d.get("phase_ms"))
PY
done
