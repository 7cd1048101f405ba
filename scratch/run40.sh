set -x
for cfg in "f16x3f 3" "f16x3f 5" "f16 5"; do
  set -- $cfg
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_curve_kernel -s 3 -c 1 -f -o gpurun_out/r02_tc_${1}_c$2 \
     python bench.py --config $2 --steps 10 --warmup 3 --precision $1 --no-cpu --no-other > gpurun_out/r02_ncu_${1}_c$2.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 10 --warmup 3 --no-cpu --no-other > gpurun_out/r02_launches_bench.log 2>&1
