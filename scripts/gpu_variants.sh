# usage: bash scripts/gpu_variants.sh [notest] name1 name2 ...   (default library first; name "t64" = default lib with 64-thread tiles)
set -x
if [ "$1" = "notest" ]; then shift; else python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
run() { # label, env...
  label=$1; shift
  env "$@" python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_$label.json 2>gpurun_out/bench_$label.err
  python -c "import json;d=json.load(open('gpurun_out/bench_$label.json'));print('RESULT $label',round(d['ms_per_step'],2),{k:round(v,2) for k,v in d['phase_ms_per_step'].items()},round(d['roofline']['frac'],4))" || tail -5 gpurun_out/bench_$label.err
}
run default A=1
for v in "$@"; do
  if [ "$v" = "t64" ]; then run t64 RL_TILE_THREADS=64; else run $v RADLITE_B200_LIB=$PWD/radlite_b200/libradlite_b200_$v.so; fi
done
