set -x
for cfg in "f16x3 3" "f16 3"; do
  set -- $cfg
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_curve_kernel -s 3 -c 1 -o gpurun_out/r02b_tc_${1}_c$2 \
     python bench.py --config $2 --steps 10 --warmup 3 --precision $1 --no-cpu --no-other > gpurun_out/r02b_ncu_${1}_c$2.log 2>&1
done
