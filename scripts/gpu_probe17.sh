python scripts/time_cfg.py 2 5 3 4
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
