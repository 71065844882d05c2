for spec in "3 100 3" "3 100000 3" "3 200 10" "1 200 5" "1 100000 5" "1 200 40"; do
set -- $spec
python bench.py --config $1 --steps $3 --warmup 3 --no-cpu --clock-ms $2 > gpurun_out/bench_p4.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/bench_p4.json'));print('cfg $1 clk $2 steps $3:',round(d['ms_per_step'],2),'wall',round(d['wall_ms_per_step'],2),'e2e',round(d['e2e']['ms_per_step'],2),{k:round(v,2) for k,v in d['phase_ms_per_step'].items()},d['clocks'].get('samples'))"
done
