# GPU visit for the ztile kernel: smoke under compute-sanitizer, parity tests, bench of the variants.
# usage: bash scripts/gpu_z.sh <tag> [variant names...]
tag=${1:-z}; shift
set -x
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/san_$tag.log 2>&1; tail -15 gpurun_out/san_$tag.log
timeout 1500 python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -40
run() { # label, env...
  label=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_$label.json 2>gpurun_out/bench_$label.err
  python -c "import json;d=json.load(open('gpurun_out/bench_$label.json'));print('RESULT $label',round(d['ms_per_step'],2),{k:round(v,2) for k,v in d['phase_ms_per_step'].items()},round(d['roofline']['frac'],4))" || tail -5 gpurun_out/bench_$label.err
}
run ${tag}_default A=1
run ${tag}_lw8 RL_ZLW=8
run ${tag}_lw32 RL_ZLW=32
run ${tag}_old RL_KERNEL=tile
for v in "$@"; do
  run ${tag}_$v RADLITE_B200_LIB=$PWD/radlite_b200/libradlite_b200_$v.so
done
