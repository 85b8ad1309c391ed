# Round measurement on one GPU (gpurun -- "bash scripts/measure_n1.sh"): bench, reference arm, variants, sweep, ncu launch list.
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/r02_bench_reference_n1.err
for v in evimo2 evimo2_tref10 dsec_tref5; do python bench.py --variant $v --steps 10 --warmup 3 --no-cpu --no-train > gpurun_out/r02_bench_$v.json 2> gpurun_out/r02_bench_$v.err; done
python bench.py --variant k3_det --steps 10 --warmup 3 --no-cpu --no-train > gpurun_out/r02_bench_k3_det_b1.json 2> gpurun_out/r02_bench_k3_det_b1.err
python bench.py --variant k3_det --batch 14 --steps 10 --warmup 3 --no-cpu --no-train > gpurun_out/r02_bench_k3_det_b14.json 2> gpurun_out/r02_bench_k3_det_b14.err
python bench.py --dist edges --steps 10 --warmup 3 --no-cpu --no-train > gpurun_out/r02_bench_edges.json 2> gpurun_out/r02_bench_edges.err
python scripts/sweep.py --steps 5 > gpurun_out/r02_sweep_n1.json 2> gpurun_out/r02_sweep_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 60 --csv --log-file gpurun_out/r02_ncu_launch_list.csv python bench.py --steps 2 --prof-warmup 1 --no-e2e --no-cpu --no-train > /dev/null 2>&1
ls -la gpurun_out/r02_*
