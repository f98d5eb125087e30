set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/k1f_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/k1f_pytest.log
SCHIC_K1_MODE=filter timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "modulo or sweep or ragged or full_size" > gpurun_out/k1f_pytest2.log 2>&1; echo "pytest2 rc=$?"; tail -3 gpurun_out/k1f_pytest2.log
timeout 300 python tools/k1_variants.py 2048 > gpurun_out/k1f_variants.log 2>&1; tail -5 gpurun_out/k1f_variants.log
for mode in chunk filter; do
  SCHIC_K1_MODE=$mode timeout 300 python bench.py --steps 3 --warmup 3 --skip-e2e --no-cpu-baseline > gpurun_out/k1f_bench_$mode.json 2> gpurun_out/k1f_bench_$mode.err; echo "$mode rc=$?"; cat gpurun_out/k1f_bench_$mode.json
done
PROF="python bench.py --steps 1 --warmup 3 --skip-e2e --no-cpu-baseline"
SCHIC_K1_MODE=filter timeout 300 $PROF > gpurun_out/k1f_plain.log 2>&1 && \
SCHIC_K1_MODE=filter timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1_filter' -s 3 -c 1 -o gpurun_out/k1f_prof $PROF > gpurun_out/k1f_ncu.log 2>&1
echo "ncu rc=$?"
