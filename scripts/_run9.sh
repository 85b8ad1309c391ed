python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_b9.json 2>gpurun_out/r2_b9.err; tail -3 gpurun_out/r2_b9.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_b9.json"))
print("value ms", d["ms_per_step"], "e2e", d["e2e"], "\nref-layout", d["e2e_reference_layout"], "\ntrain", d.get("train_step"))
print(d["roofline"]["stage_ms_per_launch"]); print("packed", d.get("packed_layout",{}).get("ms_per_step"), d.get("packed_layout",{}).get("stage_ms_per_launch")); print(d.get("cpu_baseline"))
PY
