set -x
mkdir -p gpurun_out
N=${1:-2}
VERIFY_CELLS=1500 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/verify_multi_gpu.py > gpurun_out/multi_verify_$N.log 2>&1; echo "verify rc=$?"; tail -1 gpurun_out/multi_verify_$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 tools/time_gather.py 2>/dev/null | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/multi_bench_$N.json 2> gpurun_out/multi_bench_$N.err; echo "bench rc=$?"
python -c "
import json,sys
b=json.loads([l for l in open('gpurun_out/multi_bench_$N.json') if l.startswith('{')][-1]); print(b['n_gpus'], b['value'], b['ms_per_step'], b['stage_ms']); print(b['e2e']['value'], b['e2e']['ms_per_step'], b['e2e'].get('host_call_ms'))"
