mkdir -p gpurun_out
for v in 0 1; do
  if [ $v = 1 ]; then export SCHIC_K4_VAL8=1; fi
  timeout 300 python bench.py --steps 4 --warmup 3 --skip-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; b=json.loads(sys.stdin.read()); print('val8=$v', b['ms_per_step'], b['stage_ms']['rerank'], b['stage_ms']['rerank_prep'])"
done
