python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for v in default c16 c24; do
  L=A=1; if [ "$v" != "default" ]; then L=RADLITE_B200_LIB=$PWD/radlite_b200/libradlite_b200_$v.so; fi
  echo "== $v"; env $L python scripts/time_cfg.py 1 3
done
ncu --set full --clock-control none --import-source on -k regex:chan_kernel -s 3 -c 1 -f -o gpurun_out/chan_c3 python bench.py --config 3 --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu_chan.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c1.csv python bench.py --config 1 --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
