python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python scripts/time_cfg.py 2 5 3
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_p12.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/bench_p12.json'));print('cfg 2:',round(d['ms_per_step'],2),'e2e',round(d['e2e']['ms_per_step'],2),{k:round(v,2) for k,v in d['phase_ms_per_step'].items()},round(d['roofline']['frac'],4))"
