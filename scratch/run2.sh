set -x
for v in base cvt maskh hint1k hint10k; do
  export VLG_B200_LIB=$PWD/scratch/variants_build/$v.so
  python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "track_reference or deterministic or edge_shapes" 2>&1 | tail -3 > gpurun_out/r2b_pytest_$v.log
  python bench.py --steps 10 --warmup 3 --precision f16 --no-cpu > gpurun_out/r2b_bench_f16_$v.json 2> gpurun_out/r2b_bench_f16_$v.err
  python bench.py --steps 10 --warmup 3 --precision f16x3 --no-cpu > gpurun_out/r2b_bench_f16x3_$v.json 2> gpurun_out/r2b_bench_f16x3_$v.err
done
unset VLG_B200_LIB
python scratch/prof_tc.py 148 2 f16 > gpurun_out/r2b_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_curve_kernel -s 2 -c 1 -o gpurun_out/r2b_tc_f16 python scratch/prof_tc.py 148 2 f16 > gpurun_out/r2b_ncu.log 2>&1
