python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for v in default noco; do
  L=A=1; if [ "$v" != "default" ]; then L=RADLITE_B200_LIB=$PWD/radlite_b200/libradlite_b200_$v.so; fi
  echo "== $v"
  env $L python scripts/ring_block_probe.py 2 0 57 3 | tail -1
  env $L python scripts/ring_block_probe.py 2 112 140 3 | tail -1
  env $L python scripts/ring_block_probe.py 2 228 269 3 | tail -1
  env $L python scripts/time_cfg.py 2 5
done
