# Launch list + ncu --set full capture of one device-resident build (profiles/ summaries are made from these).
# usage: bash tools/gpu_profile.sh TAG
TAG=${1:-r2}
mkdir -p gpurun_out
PROF="python bench.py --steps 2 --warmup 3 --skip-e2e --no-cpu-baseline"
timeout 300 $PROF > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $PROF > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
# full capture of the last build only: skip the launches of the 3 warm-up + first timed step
N=$(python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_launches.csv")) if len(r)>5]
print(max(0, (len(rows)-1)//5*4))
PY
)
timeout 1200 ncu --set full --clock-control none --import-source on -s $N -c 60 -o gpurun_out/${TAG}_prof -f $PROF > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$? (skipped $N launches)"
