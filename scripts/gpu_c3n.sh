N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --config 3 --gpus $N --steps 10 --warmup 4 > gpurun_out/bench_r2c_c3_n$N.json 2> gpurun_out/bench_r2c_c3_n$N.err
tail -2 gpurun_out/bench_r2c_c3_n$N.err
python -c "import json;d=json.load(open('gpurun_out/bench_r2c_c3_n$N.json'));print('cfg3 N$N:',round(d['ms_per_step'],2),'e2e',round(d['e2e']['ms_per_step'],2),d['phase_ms_per_step'],d['sharded_result_bitwise_equal_to_one_gpu'],d['config']['ring_blocks'],d['replicas']['ms_per_step'])"
