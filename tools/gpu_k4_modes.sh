set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/k4m_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/k4m_pytest.log
for mode in hash window window64; do
  SCHIC_K4_MODE=$mode timeout 300 python bench.py --steps 3 --warmup 3 --skip-e2e --no-cpu-baseline > gpurun_out/k4m_bench_$mode.json 2> gpurun_out/k4m_bench_$mode.err; echo "$mode rc=$?"; cat gpurun_out/k4m_bench_$mode.json
done
PROF="python bench.py --steps 1 --warmup 3 --skip-e2e --no-cpu-baseline"
SCHIC_K4_MODE=window timeout 300 $PROF > gpurun_out/k4m_plain.log 2>&1 && \
SCHIC_K4_MODE=window timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k4_rerank' -s 3 -c 1 -o gpurun_out/k4m_prof_window $PROF > gpurun_out/k4m_ncu.log 2>&1
echo "ncu rc=$?"
