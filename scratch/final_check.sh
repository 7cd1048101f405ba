set -x
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > gpurun_out/r2u_pytest_full.log
timeout 400 python bench.py > gpurun_out/r2u_bench_default.json 2> gpurun_out/r2u_bench_default.err
for cfg in "f16x3f 3" "f16 3"; do
  set -- $cfg
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_curve_kernel -s 3 -c 1 -f -o gpurun_out/r02_tc_${1}_c$2 \
     python bench.py --config $2 --steps 10 --warmup 3 --precision $1 --no-cpu --no-other > gpurun_out/r02_ncu_${1}_c$2.log 2>&1
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 10 --warmup 3 --no-cpu --no-other > gpurun_out/r02_launches_bench.log 2>&1
timeout 200 python bench.py --config 5 --no-cpu --no-other > gpurun_out/r2u_bench_c5.json 2> gpurun_out/r2u_bench_c5.err
