python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_b11_n2.json 2>gpurun_out/r2_b11.err; tail -3 gpurun_out/r2_b11.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_b11_n2.json"))
print("N", d["n_gpus"], "value", d["value"]/1e9, "ms", d["ms_per_step"], "\ne2e", {k:v for k,v in d["e2e"].items() if k!='h2d_note'}, "\nref-layout", {k:v for k,v in d["e2e_reference_layout"].items() if k!='note'}, "\ntrain", {k:v for k,v in d["train_step"].items() if k!='what'})
PY
