# usage: bash scripts/gpu_v.sh <tag> "label:VAR=val,VAR=val" ...   (variant lib via LIB=<name>)
tag=$1; shift
for spec in "$@"; do
  label=${spec%%:*}; envs=${spec#*:}
  args=()
  IFS=',' read -ra kv <<< "$envs"
  for e in "${kv[@]}"; do
    case $e in LIB=*) args+=("RADLITE_B200_LIB=$PWD/radlite_b200/libradlite_b200_${e#LIB=}.so");; *) args+=("$e");; esac
  done
  env "${args[@]}" timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
  env "${args[@]}" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_${tag}_$label.json 2>gpurun_out/bench_${tag}_$label.err
  python -c "import json;d=json.load(open('gpurun_out/bench_${tag}_$label.json'));print('RESULT $label',round(d['ms_per_step'],2),{k:round(v,2) for k,v in d['phase_ms_per_step'].items()},round(d['roofline']['frac'],4))" || tail -5 gpurun_out/bench_${tag}_$label.err
done
