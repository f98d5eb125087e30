mkdir -p gpurun_out
export SCHIC_CUDA_LIB=$PWD/schicexplorer_b200/csrc/build/var/libschic_k4_pre.so
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; echo "pytest(pre lib) rc=$?"; tail -3 gpurun_out/q_pytest.log
for pre in 1 0; do
SCHIC_K4_PRE=$pre timeout 300 python bench.py --steps 4 --warmup 3 --skip-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; b=json.loads(sys.stdin.read()); print('win18 pre=$pre', b['ms_per_step'], b['stage_ms']['rerank'], b['stage_ms']['rerank_prep'])"
done
unset SCHIC_CUDA_LIB
timeout 300 python bench.py --steps 4 --warmup 3 --skip-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; b=json.loads(sys.stdin.read()); print('default win19', b['ms_per_step'], b['stage_ms']['rerank'], b['stage_ms']['rerank_prep'])"
