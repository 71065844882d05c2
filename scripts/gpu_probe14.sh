ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv --log-file gpurun_out/launches_c5b.csv python scripts/time_cfg.py 5 > gpurun_out/b_ncu_c5b.log 2>&1
python scripts/launch_shares.py gpurun_out/launches_c5b.csv | head -16
python scripts/time_cfg.py 5 5
