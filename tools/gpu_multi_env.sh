mkdir -p gpurun_out
N=8
run() {
  tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus $N --steps 6 --warmup 3 2>/dev/null | python -c "
import json,sys
b=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$tag', round(b['value']), round(b['ms_per_step'],3), {k: round(v,2) for k,v in b['stage_ms'].items() if v})"
}
run base A=1
run hiprio TORCH_NCCL_HIGH_PRIORITY=1
run ch32 NCCL_MIN_NCHANNELS=32
run hiprio_ch32 TORCH_NCCL_HIGH_PRIORITY=1 NCCL_MIN_NCHANNELS=32
run fast --version >/dev/null 2>&1
