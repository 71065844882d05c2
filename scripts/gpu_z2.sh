# usage: bash scripts/gpu_z2.sh <tag> <variant-for-ncu or "default"> [bench variants...]
tag=$1; nv=$2; shift 2
set -x
timeout 1500 python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -12
run() { # label, env...
  label=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_$label.json 2>gpurun_out/bench_$label.err
  python -c "import json;d=json.load(open('gpurun_out/bench_$label.json'));print('RESULT $label',round(d['ms_per_step'],2),{k:round(v,2) for k,v in d['phase_ms_per_step'].items()},round(d['roofline']['frac'],4))" || tail -5 gpurun_out/bench_$label.err
}
run ${tag}_default A=1
for v in "$@"; do
  run ${tag}_$v RADLITE_B200_LIB=$PWD/radlite_b200/libradlite_b200_$v.so
done
if [ "$nv" != "none" ]; then
  L=A=1; if [ "$nv" != "default" ]; then L=RADLITE_B200_LIB=$PWD/radlite_b200/libradlite_b200_$nv.so; fi
  env $L timeout 900 ncu --set full --clock-control none --import-source on -k regex:ztile_kernel -s 3 -c 1 -f -o gpurun_out/ztile_$tag python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu2.log 2>&1
  ls -la gpurun_out/ztile_$tag.ncu-rep
fi
