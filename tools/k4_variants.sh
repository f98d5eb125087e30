# Experiment builds of the counts kernel (schicexplorer_b200/build.py build_cuda(variant=..., extra_flags=...)) timed on the
# device-resident leg. usage: VARIANTS="nb1:4 t512c2nb2:8" bash tools/k4_variants.sh   (variant[:lanes per slice])
PROF="python bench.py --steps 5 --warmup 3 --skip-e2e --no-cpu-baseline"
for VL in ${VARIANTS}; do
  V=${VL%%:*}; L=${VL##*:}; [ "$L" = "$V" ] && L=4
  echo "== $V lanes/slice $L"; SCHIC_CUDA_LIB=$PWD/schicexplorer_b200/csrc/build_$V/libschic_mh_$V.so SCHIC_K4_LPS=$L SCHIC_K4_GROUP_MB=${MB:-80} timeout 300 $PROF 2>&1 | tail -1 | python -c "
import sys, json
b = json.loads(sys.stdin.readline()); print(round(b['ms_per_step'], 3), 'rerank', round(b['stage_ms']['rerank'], 3), b['parity']['checksum'])"
done
