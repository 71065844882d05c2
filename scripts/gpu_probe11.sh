python -m pytest tests -m gpu -x -q 2>&1 | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c5.csv python bench.py --config 5 --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu_c5.log 2>&1
python scripts/launch_shares.py gpurun_out/launches_c5.csv | head -14
