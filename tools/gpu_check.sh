# One GPU-box round trip: GPU test-suite, smoke, bench (ours + reference arm). Usage: bash tools/gpu_check.sh TAG [pytest-args]
TAG=${1:-r2}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv; nproc
timeout 1500 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/${TAG}_bench_n1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_n1.json", "gpurun_out/${TAG}_bench_ref.json"):
    try:
        b = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, round(b["value"]), "cells/s", round(b["ms_per_step"], 3), "ms")
    for k in ("stage_ms", "e2e", "e2e_class", "hash_width_other", "parity", "cpu_baseline"):
        if k in b: print("  ", k, json.dumps(b[k])[:700])
    if "roofline" in b: print("   roofline", json.dumps({k: v for k, v in b["roofline"].items() if k != "k1_signatures"})[:900])
PY
