#!/bin/bash
# one GPU session: tests, bench, ncu launch list, ncu full capture of the hot kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --prof-warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
ncu --set full --clock-control none --import-source on -k regex:"knn_select|event_forward|event_backward|lut_backward|image_" -s 11 -c 7 \
    -o gpurun_out/prof_full -f python bench.py --steps 1 --prof-warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
