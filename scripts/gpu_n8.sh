N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 4 > gpurun_out/bench_r2c_n$N.json 2> gpurun_out/bench_r2c_n$N.err
python -c "import json;d=json.load(open('gpurun_out/bench_r2c_n$N.json'));print('N$N:',round(d['ms_per_step'],2),'e2e',round(d['e2e']['ms_per_step'],2),d['phase_ms_per_step'],d['phase_ms_per_step_max_over_ranks'],d['sharded_result_bitwise_equal_to_one_gpu'],d['config']['ring_blocks'],d['replicas']['ms_per_step'])"
