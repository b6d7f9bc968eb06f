#!/usr/bin/env bash
# One gpurun call that opens a round: full GPU suite, the queued attention experiment (DESIGN §7 "first GPU experiments"),
# the default bench line and its ncu launch list.  Everything lands in gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_round_start.sh'
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python tools/bench_linattn.py --impl 2 3 --shape "1024x8" --iters 20 > gpurun_out/linattn_impl2_vs_cp.jsonl 2>gpurun_out/linattn_impl2_vs_cp.err
cat gpurun_out/linattn_impl2_vs_cp.jsonl
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err && tail -c 1500 gpurun_out/bench_1gpu.json
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --launch-skip 105800 -c 9000 --csv \
    --log-file gpurun_out/launches_update_phase.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profile-step \
    > gpurun_out/ncu_launches.log 2>&1; echo "ncu exit $?"
# whole-bench A/B of the single-pass attention kernels (only meaningful if the micro-benchmark above favours impl 2)
CPM_LINATTN_IMPL=2 python bench.py --no-cpu-baseline > gpurun_out/bench_1gpu_impl2.json 2> gpurun_out/bench_1gpu_impl2.err && tail -c 600 gpurun_out/bench_1gpu_impl2.json
