set -x
ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 3 -c 1 -f -o gpurun_out/tile_c3 python bench.py --config 3 --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:zcont_kernel -s 3 -c 1 -f -o gpurun_out/zcont_c2 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu_zc.log 2>&1
python scripts/ring_block_probe.py 2 0 57 3 > gpurun_out/probe_0_57.log 2>&1
python scripts/ring_block_probe.py 2 221 269 3 > gpurun_out/probe_221_269.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_rb0.csv python scripts/ring_block_probe.py 2 0 57 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_rb7.csv python scripts/ring_block_probe.py 2 221 269 2 > /dev/null 2>&1
cat gpurun_out/probe_0_57.log gpurun_out/probe_221_269.log
