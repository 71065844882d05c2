python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 4 > gpurun_out/bench_r2c_n2.json 2> gpurun_out/bench_r2c_n2.err
tail -3 gpurun_out/bench_r2c_n2.err
python -c "import json;d=json.load(open('gpurun_out/bench_r2c_n2.json'));print('N2:',round(d['ms_per_step'],2),'e2e',round(d['e2e']['ms_per_step'],2),d['phase_ms_per_step'],d['phase_ms_per_step_max_over_ranks'],d['sharded_result_bitwise_equal_to_one_gpu'],d['config']['ring_blocks'],d['replicas']['ms_per_step'])"
