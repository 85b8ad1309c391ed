# A/B of variant libraries built by scripts/build_variant.sh on the bench configuration:
#   gpurun -- '[BENCH_ARGS="--variant evimo2"] bash scripts/run_variants.sh NAME1 NAME2 ...'
# (prints step / knn / lut_backward ms and the work list)
cp motionpriorcmax_b200/_lib/libcmax_b200.so /tmp/default.so
for v in default "$@"; do
  if [ $v = default ]; then cp /tmp/default.so motionpriorcmax_b200/_lib/libcmax_b200.so; else cp _variants/$v/libcmax_b200.so motionpriorcmax_b200/_lib/libcmax_b200.so; fi
  python bench.py $BENCH_ARGS --steps 10 --warmup 3 --no-cpu --no-train --no-e2e --no-packed > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err || tail -3 gpurun_out/var_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/var_$v.json")); s=d["roofline"]["stage_ms_per_launch"]
    print("$v", "step %.3f"%d["ms_per_step"], "knn %.3f lutbwd %.3f"%(s["knn_select"], s["lut_backward"]), "worklist", d["roofline"]["knn_worklist_cells"]["total"],
          "| sum of stages %.3f |"%sum(s.values()), " ".join("%s %.3f"%(k[:9], x) for k, x in s.items() if k not in ("knn_select", "lut_backward")))
except Exception as e: print("$v failed", e)
PY
done
cp /tmp/default.so motionpriorcmax_b200/_lib/libcmax_b200.so
