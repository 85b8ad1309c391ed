python -m pytest tests -m gpu -x -q 2>&1 | tail -6
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-train"
P='import json,sys; d=json.load(open(sys.argv[1])); s=d["packed_layout"]["stage_ms_per_launch"]; print(sys.argv[1], round(d["ms_per_step"],3), "packed", round(d["packed_layout"]["ms_per_step"],3), "ev fwd/bwd", round(s["event_forward"],4), round(s["event_backward"],4))'
$B > gpurun_out/r2_b15.json 2>gpurun_out/r2_b15.err; python -c "$P" gpurun_out/r2_b15.json
python scripts/sweep.py --steps 5 --batch 14 --events 1e6,1e7,5e7 --layouts packed 2>&1 | grep '^{"n_gpus"' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['B'], d['events_per_window'], d['layout'], round(d['ms_per_step'],3), round(d['events_per_s']/1e9,2), 'fwd', round(d['event_forward_ms'],3), 'bwd', round(d['event_backward_ms'],3), 'hbm', round(d['hbm_frac'],3))"
