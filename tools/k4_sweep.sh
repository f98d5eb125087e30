# K4 experiments on one GPU: window-group size / lanes-per-slice sweep of the counts kernel (device-resident leg only),
# optional ncu capture.  usage: [MBS="0 40 160"] [LPSS="8 4"] [NCU_MB=40 NCU_LPS=8] [TAG=r2c] bash tools/k4_sweep.sh
mkdir -p gpurun_out
PROF="python bench.py --steps 5 --warmup 3 --skip-e2e --no-cpu-baseline"
for LPS in ${LPSS:-8}; do
for MB in ${MBS:-0 20 40 80 160}; do
  echo "== SCHIC_K4_LPS=$LPS SCHIC_K4_GROUP_MB=$MB"; SCHIC_K4_LPS=$LPS SCHIC_K4_GROUP_MB=$MB timeout 300 $PROF 2>&1 | tail -1 | python -c "
import sys, json
b = json.loads(sys.stdin.readline()); print(round(b['ms_per_step'], 3), {k: round(v, 3) for k, v in b['stage_ms'].items() if v}, b['rerank_kernel'], b['parity']['checksum'])"
done
done
if [ -n "$NCU_MB" ]; then
  SCHIC_K4_LPS=${NCU_LPS:-8} SCHIC_K4_GROUP_MB=$NCU_MB timeout 900 ncu --set full --clock-control none --import-source on -k regex:k4c_rerank_kernel -c 1 -o gpurun_out/${TAG:-r2c}_k4c_prof -f python bench.py --steps 1 --warmup 3 --skip-e2e --no-cpu-baseline > gpurun_out/${TAG:-r2c}_ncu.log 2>&1; echo "ncu rc=$?"
fi
