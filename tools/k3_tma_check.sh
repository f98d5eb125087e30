timeout 600 python -m pytest tests -m gpu -x -q -k "k3 or collision or fixture or config or sharded or max_bin" 2>&1 | tail -3
SCHIC_K3_TMA=0 timeout 300 python -m pytest tests -m gpu -x -q -k "k3_packed or collision_blocks" 2>&1 | tail -2
for T in 1 0; do
  echo "== SCHIC_K3_TMA=$T s10"; SCHIC_K3_TMA=$T timeout 300 python bench.py --steps 5 --warmup 3 --skip-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
b = json.loads(sys.stdin.readline()); print(round(b['ms_per_step'], 3), {k: round(v, 3) for k, v in b['stage_ms'].items() if v}, b['parity']['checksum'])"
done
for T in 1 0; do
  echo "== SCHIC_K3_TMA=$T s50 (20k cells)"; SCHIC_K3_TMA=$T timeout 600 python bench.py --workload s50 --cells 20000 --steps 3 --warmup 3 --skip-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
b = json.loads(sys.stdin.readline()); print(round(b['ms_per_step'], 3), {k: round(v, 3) for k, v in b['stage_ms'].items() if v}, b['parity']['checksum'])"
done
