set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/k1f_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/k1f_pytest.log
SCHIC_K1_ALL=1 timeout 300 python tools/k1_variants.py 2048 > gpurun_out/k1f_variants.log 2>&1; tail -5 gpurun_out/k1f_variants.log
for mode in tile filter; do
  SCHIC_K1_MODE=$mode timeout 300 python bench.py --steps 3 --warmup 3 --skip-e2e --no-cpu-baseline > gpurun_out/k1f_bench_$mode.json 2> gpurun_out/k1f_bench_$mode.err; echo "$mode rc=$?"; cat gpurun_out/k1f_bench_$mode.json
done
timeout 600 python bench.py > gpurun_out/k1f_bench_full.json 2> gpurun_out/k1f_bench_full.err; echo "full rc=$?"; cat gpurun_out/k1f_bench_full.json
