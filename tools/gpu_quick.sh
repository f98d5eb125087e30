set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/q_pytest.log
timeout 300 python tools/k1_variants.py 4096 > gpurun_out/q_variants.log 2>&1; tail -6 gpurun_out/q_variants.log
timeout 300 python bench.py --steps 3 --warmup 3 --skip-e2e --no-cpu-baseline > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "rc=$?"; cat gpurun_out/q_bench.json; tail -3 gpurun_out/q_bench.err
