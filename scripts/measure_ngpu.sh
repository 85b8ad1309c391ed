# Round measurement at N GPUs (usage: gpurun --gpus N -- "bash scripts/measure_ngpu.sh N"): bench, bench without NUMA binding, sweep.
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -2 gpurun_out/r02_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 --no-numa --no-train --no-packed > gpurun_out/r02_bench_n${N}_nonuma.json 2> gpurun_out/r02_bench_n${N}_nonuma.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 scripts/sweep.py --steps 5 --batch 14 --events 1e6,1e7 > gpurun_out/r02_sweep_n$N.json 2> gpurun_out/r02_sweep_n$N.err
python - <<PY
import json
for f in ("gpurun_out/r02_bench_n$N.json","gpurun_out/r02_bench_n${N}_nonuma.json"):
    d=json.load(open(f))
    print(f, "N", d["n_gpus"], "value %.2f G %.3f ms | e2e %.2f G %.3f ms copy %.2f ms %.1f GB/s numa %s"%(d["value"]/1e9,d["ms_per_step"],d["e2e"]["value"]/1e9,d["e2e"]["ms_per_step"],d["e2e"]["copy_alone_ms"],d["e2e"]["h2d_GBps_per_rank_copy_alone"],d["e2e"].get("host_numa_binding_rank0")), "| ref layout %.3f ms"%d["e2e_reference_layout"]["ms_per_step"], "| train", d.get("train_step",{}).get("ms_per_step"))
PY
