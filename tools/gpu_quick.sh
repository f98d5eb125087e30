mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/q_pytest.log
for m in plain packed; do
SCHIC_K3_MODE=$m timeout 300 python bench.py --steps 4 --warmup 3 --skip-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; b=json.loads(sys.stdin.read()); print('$m', b['ms_per_step'], b['stage_ms'])"
done
SCHIC_K3_MODE=packed timeout 1500 python bench.py --workload s50 --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e 2>/dev/null | python -c "
import json,sys; b=json.loads(sys.stdin.read()); print('s50 packed', b['ms_per_step'], b['stage_ms'])"
