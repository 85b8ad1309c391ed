bash scripts/_measN.sh 2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 scripts/sharded_check.py > gpurun_out/r02_event_sharded_n2.json 2> gpurun_out/r02_event_sharded_n2.err; tail -2 gpurun_out/r02_event_sharded_n2.err; cat gpurun_out/r02_event_sharded_n2.json | tail -1 | cut -c1-1500
