# Round-end evidence on one GPU: tests, bench lines of every config, reference arm, launch lists, full ncu captures,
# memcheck.  usage: bash scripts/gpu_final.sh <tag>   (outputs under gpurun_out/)
tag=${1:-fin}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${tag}_c2.json 2> gpurun_out/bench_${tag}_c2.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err
for c in 1 3 4 5; do python bench.py --config $c --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${tag}_c$c.json 2>/dev/null; done
for c in 2 1 3 4 5; do python -c "import json;d=json.load(open('gpurun_out/bench_${tag}_c$c.json'));print('cfg $c:',round(d['ms_per_step'],2),'e2e',round(d['e2e']['ms_per_step'],2),{k:round(v,2) for k,v in d['phase_ms_per_step'].items()},round(d['roofline']['frac'],4), d.get('cpu_baseline') and d['cpu_baseline']['value'])"; done
tail -c 600 gpurun_out/bench_${tag}_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${tag}_c5.csv python bench.py --config 5 --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${tag}_c3.csv python bench.py --config 3 --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ztile_kernel -s 3 -c 1 -f -o gpurun_out/ztile_${tag} python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:chan_kernel -s 3 -c 1 -f -o gpurun_out/chan_${tag} python bench.py --config 3 --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu_${tag}_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ztile_kernel -s 3 -c 1 -f -o gpurun_out/ztile_c5_${tag} python bench.py --config 5 --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu_${tag}_c5.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python scripts/san_case.py > gpurun_out/san_memcheck_${tag}.log 2>&1; tail -2 gpurun_out/san_memcheck_${tag}.log
ls -la gpurun_out | tail -12
