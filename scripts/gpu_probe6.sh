ncu --set full --clock-control none --import-source on -k regex:chan_kernel -s 3 -c 1 -f -o gpurun_out/chan_c3 python bench.py --config 3 --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu_chan.log 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k same_bits 2>&1 | tail -3
