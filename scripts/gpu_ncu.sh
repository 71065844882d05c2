# usage: bash scripts/gpu_ncu.sh <tag> <variant or "default"> [kernel regex]
tag=$1; nv=$2; k=${3:-ztile_kernel}
L=A=1; if [ "$nv" != "default" ]; then L=RADLITE_B200_LIB=$PWD/radlite_b200/libradlite_b200_$nv.so; fi
env $L timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/${k}_$tag python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu_$tag.log 2>&1
ls -la gpurun_out/${k}_$tag.ncu-rep
