mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/q_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/q_bench.err | python -c "
import json,sys; b=json.loads(sys.stdin.read()); print(b['value'], b['ms_per_step'], b['stage_ms']); print(b['e2e']['value'], b['e2e']['ms_per_step'], b['e2e']['h2d_bytes_per_step'], b['e2e']['stage_ms'])"; tail -3 gpurun_out/q_bench.err
