mkdir -p gpurun_out
N=${1:-2}
VERIFY_CELLS=${VERIFY_CELLS:-1500} timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/verify_multi_gpu.py > gpurun_out/multi_verify_$N.log 2>&1; echo "verify rc=$?"; grep PARITY gpurun_out/multi_verify_$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 > gpurun_out/multi_bench_$N.json 2> gpurun_out/multi_bench_$N.err; echo "bench rc=$?"
python -c "
import json,sys
b=json.loads([l for l in open('gpurun_out/multi_bench_$N.json') if l.startswith('{')][-1]); print(b['n_gpus'], round(b['value']), round(b['ms_per_step'],3), {k: round(v,2) for k,v in b['stage_ms'].items() if v}, 'e2e', round(b['e2e']['value']), round(b['e2e']['ms_per_step'],2))"
