#!/bin/bash
# one GPU session: tests, bench, ncu launch list, ncu full capture of every kernel of one step
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; tail -c 600 gpurun_out/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --prof-warmup 1 --no-cpu --no-e2e --no-packed > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:"bin_points|knn_|lut_|event_|image_|smooth_|finalize|traj_" -s 28 -c 28 \
    -o gpurun_out/prof_full -f python bench.py --steps 1 --prof-warmup 1 --no-cpu --no-e2e --no-packed > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
# packed layout: the tile kernels and the device packer
ncu --set full --clock-control none --import-source on -k regex:"tile_kernel|pack_" -s 5 -c 5 \
    -o gpurun_out/prof_packed -f python bench.py --steps 1 --prof-warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_packed.log 2>&1
tail -1 gpurun_out/ncu_packed.log
ls -la gpurun_out | head -30
