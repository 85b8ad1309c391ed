python bench.py --steps 20 --warmup 3 --no-cpu --no-train > gpurun_out/r2_b10.json 2>gpurun_out/r2_b10.err; tail -3 gpurun_out/r2_b10.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_b10.json"))
print("value ms", d["ms_per_step"], "e2e", {k:v for k,v in d["e2e"].items() if k!='h2d_note'}, "\nref-layout", {k:v for k,v in d["e2e_reference_layout"].items() if k!='note'})
PY
