python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/time_cfg.py 3
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_p16.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/bench_p16.json'));print('cfg 2:',round(d['ms_per_step'],2),'e2e',round(d['e2e']['ms_per_step'],2),{k:round(v,2) for k,v in d['phase_ms_per_step'].items()},round(d['roofline']['frac'],4))"
python bench.py --config 3 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_p16c3.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/bench_p16c3.json'));print('cfg 3:',round(d['ms_per_step'],2),'e2e',round(d['e2e']['ms_per_step'],2),{k:round(v,2) for k,v in d['phase_ms_per_step'].items()},round(d['roofline']['frac'],4))"
