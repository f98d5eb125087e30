set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1e_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r1e_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1e_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r1e_smoke.log
timeout 600 python bench.py > gpurun_out/r1e_bench_n1.json 2> gpurun_out/r1e_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1e_bench_ref.json 2> gpurun_out/r1e_bench_ref.err; echo "ref rc=$?"
PROF="python bench.py --steps 2 --warmup 3 --skip-e2e --no-cpu-baseline"
timeout 300 $PROF > gpurun_out/r1e_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1e_launches.csv $PROF > gpurun_out/r1e_ncu_launch.log 2>&1
echo "ncu rc=$?"
