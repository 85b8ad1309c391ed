python bench.py --steps 20 --warmup 5 --no-cpu --no-train > gpurun_out/r2_b19.json 2>gpurun_out/r2_b19.err; tail -3 gpurun_out/r2_b19.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_b19.json"))
e=d["e2e"]
print("value %.2f G %.3f ms | e2e %.2f G %.3f ms (blocking %.3f) bytes %d copy %.2f ms %.1f GB/s | cg %.3f | compact12 %.3f ms | ref %.3f | packed %.3f"%(d["value"]/1e9,d["ms_per_step"],e["value"]/1e9,e["ms_per_step"],e["blocking_item"]["ms_per_step_rank0"],e["h2d_bytes_per_step"],e["copy_alone_ms"],e["h2d_GBps_per_rank_copy_alone"],e["with_coeff_grid_from_host"]["ms_per_step_rank0"],e["compact_12B_layout"]["ms_per_step_rank0"],d["e2e_reference_layout"]["ms_per_step"],d["packed_layout"]["ms_per_step"]))
PY
