python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/atomic_bench.py > gpurun_out/r2_atomic_microbench.json 2>gpurun_out/r2_atomic.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_atomic_microbench.json"))
for k,v in d.items():
    if "34MB" in k or "1.2MB" in k: print(k, round(v["Gops_per_s"],1), round(v.get("Grequests_per_s",0),1))
PY
