mkdir -p gpurun_out
for v in default 1024_19 768_19; do
  if [ $v = default ]; then unset SCHIC_CUDA_LIB; else export SCHIC_CUDA_LIB=$PWD/schicexplorer_b200/csrc/build/var/libschic_k4_$v.so; fi
  timeout 300 python bench.py --steps 4 --warmup 3 --skip-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; b=json.loads(sys.stdin.read()); print('$v', b['ms_per_step'], b['stage_ms']['rerank'], b['stage_ms']['rerank_prep'])"
done
