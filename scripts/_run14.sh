python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-train"
P='import json,sys; d=json.load(open(sys.argv[1])); s=d["roofline"]["stage_ms_per_launch"]; print(sys.argv[1], round(d["ms_per_step"],3), "knn", round(s["knn_select"],3), "bwd", round(s["lut_backward"],3), "bin", round(s["bin_points"],3), d["roofline"]["knn_worklist_cells"]["total"], "packed", d["packed_layout"]["ms_per_step"])'
$B > gpurun_out/r2_b14.json 2>gpurun_out/r2_b14.err; python -c "$P" gpurun_out/r2_b14.json
