for ms in 100 200 1000 100000; do
python bench.py --steps 10 --warmup 3 --no-cpu --clock-ms $ms > gpurun_out/bench_clk$ms.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/bench_clk$ms.json'));print('CLK $ms',round(d['ms_per_step'],2),round(d['e2e']['ms_per_step'],2),{k:round(v,2) for k,v in d['phase_ms_per_step'].items()},d['clocks'])"
done
python scripts/time_cfg.py 2 1 3
