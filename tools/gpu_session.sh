mkdir -p gpurun_out
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench exit $?"; tail -c 1500 gpurun_out/bench_1gpu.json; tail -5 gpurun_out/bench_1gpu.err
CPM_CHAIN_PDL=0 timeout 900 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/bench_1gpu_nopdl.json 2> gpurun_out/bench_1gpu_nopdl.err; echo "bench(nopdl) exit $?"; tail -c 400 gpurun_out/bench_1gpu_nopdl.json
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "full pytest exit $?" | tee -a gpurun_out/pytest_gpu.log; tail -40 gpurun_out/pytest_gpu.log
