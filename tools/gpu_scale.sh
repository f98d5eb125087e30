mkdir -p gpurun_out
VERIFY_CELLS=1500 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/verify_multi_gpu.py > gpurun_out/multi_verify_8.log 2>&1; echo "verify rc=$?"; tail -1 gpurun_out/multi_verify_8.log
for N in 1 2 4 8; do
  if [ $N = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/scale_bench_$N.json 2> gpurun_out/scale_bench_$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_bench_$N.json 2> gpurun_out/scale_bench_$N.err
  fi
  python -c "
import json
b=json.loads([l for l in open('gpurun_out/scale_bench_$N.json') if l.startswith('{')][-1]); print(b['n_gpus'], round(b['value']), round(b['ms_per_step'],3), {k: round(v,2) for k,v in b['stage_ms'].items() if v}, 'e2e', round(b['e2e']['value']), round(b['e2e']['ms_per_step'],2))"
done
