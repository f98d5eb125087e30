# Multi-GPU check on N GPUs of one box: parity tool (sharded == single GPU == oracle), then bench lines.
# usage: bash tools/gpu_multi.sh N [TAG] [WORKLOADS="s10 s50"]
N=${1:-2}; TAG=${2:-r2}; WL=${3:-s10}
mkdir -p gpurun_out
VERIFY_CELLS=${VERIFY_CELLS:-1200} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/verify_multi_gpu.py > gpurun_out/${TAG}_verify_$N.log 2>&1; echo "verify rc=$?"; grep -i "identical\|PARITY" gpurun_out/${TAG}_verify_$N.log
for W in $WL; do
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --workload $W --steps ${STEPS:-5} --warmup 3 > gpurun_out/${TAG}_bench_${W}_$N.json 2> gpurun_out/${TAG}_bench_${W}_$N.err; echo "bench $W rc=$?"
  python - <<PY
import json
try:
    b = json.loads([l for l in open("gpurun_out/${TAG}_bench_${W}_$N.json") if l.startswith("{")][-1])
    print("$W", b["n_gpus"], round(b["value"]), "cells/s", round(b["ms_per_step"], 3), "ms", {k: round(v, 2) for k, v in b["stage_ms"].items() if v},
          "e2e", round(b["e2e"]["value"]), round(b["e2e"]["ms_per_step"], 2), b["parity"]["checksum"])
except Exception as e:
    print("unreadable", e); print(open("gpurun_out/${TAG}_bench_${W}_$N.err").read()[-1500:])
PY
done
