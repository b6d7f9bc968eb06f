mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -q --timeout 300 -x > gpurun_out/gemm_tests.log 2>&1; echo "gemm pytest exit $?" | tee -a gpurun_out/gemm_tests.log; tail -5 gpurun_out/gemm_tests.log
timeout 300 python tools/bench_gemm.py > gpurun_out/bench_gemm.jsonl 2> gpurun_out/bench_gemm.err; echo "bench_gemm exit $?"; cat gpurun_out/bench_gemm.jsonl; tail -3 gpurun_out/bench_gemm.err
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x --deselect tests/test_gpu_gemm.py > gpurun_out/pytest_gpu.log 2>&1; echo "full pytest exit $?" | tee -a gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench exit $?"; tail -c 2500 gpurun_out/bench_1gpu.json; tail -5 gpurun_out/bench_1gpu.err
CPM_GEMM=lib timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_1gpu_libgemm.json 2> gpurun_out/bench_1gpu_libgemm.err; echo "bench(lib) exit $?"; tail -c 600 gpurun_out/bench_1gpu_libgemm.json
