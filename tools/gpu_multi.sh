set -x
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L | head -8
VERIFY_CELLS=1500 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/verify_multi_gpu.py > gpurun_out/multi_verify_$N.log 2>&1; echo "verify rc=$?"; tail -4 gpurun_out/multi_verify_$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/multi_bench_$N.json 2> gpurun_out/multi_bench_$N.err; echo "bench rc=$?"; tail -3 gpurun_out/multi_bench_$N.err
python -c "
import json,sys
b=json.loads([l for l in open('gpurun_out/multi_bench_$N.json') if l.startswith('{')][-1]); print(b['n_gpus'], b['value'], b['ms_per_step'], b['stage_ms']); print(b['e2e']['value'], b['e2e']['ms_per_step'], b['e2e'].get('host_call_ms'))"
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/multi_bench_1.json 2>/dev/null; python -c "
import json
b=json.load(open('gpurun_out/multi_bench_1.json')); print(1, b['value'], b['ms_per_step'])"
