python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/time_cfg.py 1 3
for c in 1 3; do python bench.py --config $c --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c${c}_chan.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/bench_c${c}_chan.json'));print('cfg $c:',round(d['ms_per_step'],2),'e2e',round(d['e2e']['ms_per_step'],2),{k:round(v,2) for k,v in d['phase_ms_per_step'].items()},round(d['roofline']['frac'],3))"; done
