set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/q_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "rc=$?"; python -c "
import json; b=json.load(open('gpurun_out/q_bench.json')); print(b['value'], b['ms_per_step'], b['stage_ms']); print(b['e2e']['value'], b['e2e']['ms_per_step'])"
timeout 1500 python bench.py --workload s50 --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/s50_bench.json 2> gpurun_out/s50_bench.err; echo "rc=$?"; cat gpurun_out/s50_bench.json
