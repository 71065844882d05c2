# One GPU visit: parity tests, bench line, ncu launch list, one full ncu capture of the dominant kernel.
# usage: bash scripts/gpu_round.sh <tag>     (outputs under gpurun_out/)
tag=${1:-cur}
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 2500 gpurun_out/bench_$tag.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; tail -c 1500 gpurun_out/bench_ref_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ztile_kernel -s 3 -c 1 -f -o gpurun_out/ztile_$tag python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu2.log 2>&1
ls -la gpurun_out | tail -8
